/* gcdyn_b200.h — C ABI of libgcdyn_b200.so, the B200 (sm_100a) implementation of
 * gcdyn.jl's multitype branching-process tree log-likelihood.
 *
 * The reference (thanasibakis/gcdyn.jl) has no FFI boundary: the path is ordinary Julia
 * multiple dispatch.  These entry points are what a Julia `ccall` binding for that path
 * needs; each one names the reference interface it replaces (paths relative to the
 * reference checkout).  INTEGRATION.md shows the Julia-side stubs.
 *
 * Conventions
 *   - plain pointers and sizes only; every function returns GCDYN_OK (0) or a negative
 *     GCDYN_E* code and never throws, exits or longjmps across the boundary;
 *   - caller owns every input/output buffer; inputs are copied before the call returns
 *     and never retained; opaque handles own device memory;
 *   - the library calls cudaSetDevice itself, installs no signal handlers, and a
 *     gcdyn_ctx may be used from any host thread (one call at a time per ctx);
 *   - STREAMS: a ctx, a forest handle and a communicator own scratch memory that is reused
 *     from call to call (table, coefficient rows, partial sums, flags; the forest's moments
 *     and calibrated grid degree).  Calls that share any of these handles must be issued on
 *     ONE stream - the ctx's own, or the one passed to the *_resident entry points - or be
 *     separated by a synchronisation of that stream; results of an unsynchronised resident
 *     call are ordered on its stream like any kernel's.  A forest caches its moments for one
 *     present_time: alternating present times on one handle is correct but rebuilds them;
 *   - there is NO CPU fallback: without a CUDA device every compute call fails with
 *     GCDYN_ENODEVICE.
 *   - "density evaluates to zero" cases of the reference (`@warn` + `return -Inf`,
 *     src/multitype_branching_process.jl:154-158,181-185,211-214,238-241) come back as
 *     -Inf values with a non-zero status word, not as error codes.
 */
#ifndef GCDYN_B200_H
#define GCDYN_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GCDYN_B200_VERSION 130  /* major*100 + minor; 1.10 adds gcdyn_comm_*, gcdyn_loglik_batch_sharded, gcdyn_mh_*;
                                   1.20 adds GCDYN_FLAG_TOTALS, gcdyn_last_cells, gcdyn_loglik_batch_resident_sharded,
                                   gcdyn_loglik_grad_batch; 1.30 adds gcdyn_comm_rows_*, gcdyn_comm_error, gcdyn_last_traj_sharded */

/* return codes */
#define GCDYN_OK            0
#define GCDYN_EINVAL       (-1)  /* bad argument / malformed forest or parameters          */
#define GCDYN_ECUDA        (-2)  /* a CUDA runtime call failed (see gcdyn_last_error)       */
#define GCDYN_ENOMEM       (-3)  /* host or device allocation failed                        */
#define GCDYN_EUNSUPPORTED (-4)  /* e.g. n_types above GCDYN_MAX_TYPES                      */
#define GCDYN_ENODEVICE    (-5)  /* no usable CUDA device                                   */

#define GCDYN_MAX_TYPES 128

/* which likelihood: replaces the four reference entry points
 *   ODE            StatsAPI.loglikelihood(model, tree(s), present_time; reltol, abstol, maxiters)
 *                  src/multitype_branching_process.jl:110-117,250-257
 *   NAIVE          gcdyn.naive_loglikelihood(model, tree)                         src/extra_likelihoods.jl:10
 *   STADLER        gcdyn.stadler_appx_loglikelihood(model, tree, present_time)    src/extra_likelihoods.jl:45
 *   STADLER_UNCOND gcdyn.stadler_appx_unconditioned_loglikelihood(...)            src/extra_likelihoods.jl:115 */
#define GCDYN_METHOD_ODE            0
#define GCDYN_METHOD_NAIVE          1
#define GCDYN_METHOD_STADLER        2
#define GCDYN_METHOD_STADLER_UNCOND 3

/* how birth rates are given: replaces λ(model, type), src/multitype_branching_process.jl:27-43 */
#define GCDYN_LAMBDA_TABLE    0  /* lambda[P*K], pset-major (DiscreteBranchingProcess, or pre-evaluated) */
#define GCDYN_LAMBDA_CONSTANT 1  /* lambda[P]                (ConstantBranchingProcess)                   */
#define GCDYN_LAMBDA_SIGMOID  2  /* xscale,xshift,yscale,yshift[P] + type_space[K], evaluated on device
                                    (SigmoidalBranchingProcess; sigmoid() at :4-7)                          */

/* per-(tree,pset) status words */
#define GCDYN_STATUS_OK         0
#define GCDYN_STATUS_SELF_LOOP  1  /* type_change with type == up.type        -> -Inf (mbp.jl:181-185)      */
#define GCDYN_STATUS_SOLVER     2  /* p left [0,1] or went non-finite          -> -Inf (mbp.jl:154,211,238)  */
#define GCDYN_STATUS_DOMAIN     4  /* log/sqrt of a negative number: Julia would throw DomainError; value NaN */

typedef struct gcdyn_ctx gcdyn_ctx;          /* one CUDA device + stream + workspace            */
typedef struct gcdyn_forest gcdyn_forest;    /* a flattened forest resident in HBM              */
typedef struct gcdyn_dparams gcdyn_dparams;  /* a batch of parameter sets resident in HBM       */

/* Flattened forest = what the host-side flattener produces from `TreeNode` trees
 * (src/types.jl:23-37): per tree the nodes of PostOrderTraversal(tree.children[1])
 * (src/treenode.jl:159-175, mbp.jl:166); the root is per-tree metadata.  Event codes are
 * 0-based positions in EVENTS (src/types.jl:6).  Type indices are 0-based positions in
 * model.type_space (type_space_index, mbp.jl:16-18). */
typedef struct {
    int64_t n_trees;
    int64_t n_nodes;
    int32_t n_types;
    int32_t reserved;
    const int64_t* tree_off;       /* [n_trees+1] CSR over nodes                                          */
    const uint8_t* event;          /* [n_nodes]                                                            */
    const int32_t* type_idx;       /* [n_nodes] node's own type (new type of a type_change); may be -1     */
    const int32_t* btype_idx;      /* [n_nodes] type of the branch above the node = idx(up.type)           */
    const double*  t_start;        /* [n_nodes] up.time                                                    */
    const double*  t_end;          /* [n_nodes] time                                                       */
    const int32_t* parent;         /* [n_nodes] position of `up` inside the tree, -1 if `up` is the root   */
    const int64_t* child_off;      /* [n_nodes+1] CSR over children                                        */
    const int32_t* child_idx;      /* [child_off[n_nodes]] positions inside the tree, `children` order     */
    const uint8_t* live;           /* [n_nodes] or NULL (= all 1): 0 marks subtrees the ODE likelihood
                                      computes but never uses (3rd+ child of a birth, mbp.jl:173-179)      */
    const int32_t* root_type_idx;  /* [n_trees] idx(tree.type)                                             */
    const double*  root_time;      /* [n_trees] tree.time                                                  */
    const uint8_t* self_loop;      /* [n_trees] or NULL: 1 = the reference returns -Inf for this tree      */
} gcdyn_forest_desc;

/* A batch of P parameter sets = P model objects (src/types.jl:65-207) in canonical form. */
typedef struct {
    int32_t n_psets;
    int32_t n_types;
    int32_t response;              /* GCDYN_LAMBDA_*                                                       */
    int32_t gamma_pset_stride;     /* 0: one Γ shared by all psets; n_types*n_types: one Γ per pset        */
    const double* lambda;          /* TABLE: [P*K]; CONSTANT: [P]; SIGMOID: NULL                           */
    const double* xscale;          /* SIGMOID: [P] each, else NULL                                         */
    const double* xshift;
    const double* yscale;
    const double* yshift;
    const double* type_space;      /* [K] (SIGMOID only, else may be NULL)                                 */
    const double* mu;              /* [P]                                                                  */
    const double* delta;           /* [P]                                                                  */
    const double* rho;             /* [P]                                                                  */
    const double* sigma;           /* [P]                                                                  */
    const double* Gamma;           /* K×K COLUMN-major (Julia Matrix{Float64}), 1 or P matrices            */
} gcdyn_params;

/* Replaces the solver keywords reltol/abstol/maxiters (mbp.jl:114-116).  The kernels use a fixed-order
 * Taylor-series integrator whose cell widths adapt to an error indicator; all zero = defaults. */
typedef struct {
    int32_t steps;   /* 0 = adaptive cells (default); S > 0 = exactly S uniform cells on [0, present_time],
                        no adaptivity: a pset whose indicator then exceeds tol is a solver failure (-Inf)  */
    int32_t order;   /* M, Taylor order (8, 12, 16, 20 or 24); 0 = 16                                  */
    double  tol;     /* accepted size of the last Taylor terms (error indicator); 0 = 1e-11, which
                        keeps the relative error of a log-likelihood below ~1e-12        */
    int32_t flags;   /* GCDYN_FLAG_*                                                                    */
    int32_t max_cells; /* the reference's `maxiters` (mbp.jl:116): a parameter set whose trajectory needs more
                        adaptive cells than this is a solver failure - status bit 2 and -Inf, where the reference
                        warns and returns -Inf (mbp.jl:154-158); 0 = bounded by the table capacity only          */
} gcdyn_opts;

#define GCDYN_FLAG_SWEEP_ONLY 4 /* per-tree sums by the item sweep only (skip the Chebyshev-moment GEMM) */
#define GCDYN_FLAG_TOTALS     8 /* only out_pset is wanted (what loglikelihood(model, trees, T) returns, mbp.jl:250-259):
                                   the per-pset sums come straight from the trajectory kernel as the product of the pset's
                                   coefficient row with the forest's COLUMN SUMS of the moment matrix - no per-tree values are
                                   formed, out_tree_pset and status must be NULL.  Same value as the sum of the per-tree
                                   path up to the order of summation (~1e-14 relative).  A forest containing a self-loop
                                   gives -Inf, as the plain sum would. */

int  gcdyn_version(void);
int  gcdyn_device_count(void);

int  gcdyn_ctx_create(int device, gcdyn_ctx** out);
void gcdyn_ctx_destroy(gcdyn_ctx* ctx);
/* Last error message of this ctx (or of the calling thread when ctx == NULL). Never NULL. */
const char* gcdyn_last_error(const gcdyn_ctx* ctx);

/* Validate, copy to the device and index a flattened forest.  Upload once per dataset and
 * reuse across many likelihood evaluations. */
int  gcdyn_forest_create(gcdyn_ctx* ctx, const gcdyn_forest_desc* desc, gcdyn_forest** out);
void gcdyn_forest_destroy(gcdyn_forest* forest);
int  gcdyn_forest_info(const gcdyn_forest* forest, int64_t* n_trees, int64_t* n_nodes,
                       int64_t* n_items, int32_t* n_types);
/* Per-tree histogram of event codes, computed on the device: out[n_trees*7], tree-major. */
int  gcdyn_forest_event_counts(gcdyn_ctx* ctx, const gcdyn_forest* forest, int64_t* out);

/* Batched likelihood, HOST buffers (the call a Julia `ccall` makes).
 *   out_pset      [P]        Σ over trees, per parameter set  (= loglikelihood(model, trees, T), mbp.jl:250-259)
 *   out_tree_pset [P*T] or NULL   per tree, pset-major: out_tree_pset[p*T + t]  (a Julia T×P Matrix)
 *   status        [P*T] or NULL   GCDYN_STATUS_* bits, same layout
 * present_time is ignored by GCDYN_METHOD_NAIVE. */
int  gcdyn_loglik_batch(gcdyn_ctx* ctx, const gcdyn_forest* forest, const gcdyn_params* params,
                        double present_time, int method, const gcdyn_opts* opts,
                        double* out_pset, double* out_tree_pset, int32_t* status);

/* Device-resident variant: parameters uploaded once, outputs are DEVICE pointers, work is
 * enqueued on `stream` (a cudaStream_t passed as void*; NULL = the ctx's own stream) and the
 * call returns without synchronising.  opts->steps == 0 uses the step count derived from the
 * rates at upload time; no refinement is attempted. */
int  gcdyn_params_upload(gcdyn_ctx* ctx, const gcdyn_params* params, gcdyn_dparams** out);
void gcdyn_dparams_destroy(gcdyn_dparams* dparams);
int  gcdyn_loglik_batch_resident(gcdyn_ctx* ctx, const gcdyn_forest* forest, const gcdyn_dparams* dparams,
                                 double present_time, int method, const gcdyn_opts* opts,
                                 double* d_out_pset, double* d_out_tree_pset, int32_t* d_status,
                                 void* stream);
/* Exact gradient of the per-parameter-set log-likelihood (ODE method) by forward sensitivities of the trajectory: what
 * the reference obtains by pushing ForwardDiff dual numbers through loglikelihood - its model structs are parametric in
 * T<:Real for exactly that (src/types.jl:92, src/multitype_branching_process.jl:118-122).  Directions = the autodiff-able
 * fields of the model:
 *     GCDYN_LAMBDA_SIGMOID   (xscale, xshift, yscale, yshift, mu, delta)    n_dirs = 6
 *     GCDYN_LAMBDA_CONSTANT  (lambda, mu, delta)                            n_dirs = 3
 * (GCDYN_LAMBDA_TABLE has K + 2 directions and n_types > 32 two integrating warps per vector: GCDYN_EUNSUPPORTED.)
 *   out_pset [P]          sum over trees of the log-likelihood (same value as gcdyn_loglik_batch with GCDYN_FLAG_TOTALS)
 *   out_grad [P*n_dirs]   d out_pset[p] / d direction j at out_grad[p*n_dirs + j]; NaN for a parameter set whose value is
 *                         -Inf or that the forest's Chebyshev grid does not resolve (its out_pset is still exact)
 * Cost: about one plain evaluation plus one sensitivity integration (all directions advance in lockstep with p). */
int  gcdyn_loglik_grad_batch(gcdyn_ctx* ctx, const gcdyn_forest* forest, const gcdyn_params* params, double present_time,
                             const gcdyn_opts* opts, double* out_pset, double* out_grad, int32_t* n_dirs);

/* Diagnostics of the last ODE evaluation on this ctx (host copies; synchronises the stream):
 * steps and order used, number of kernels launched, and per-pset error indicators [P] (may be NULL). */
int  gcdyn_last_diag(gcdyn_ctx* ctx, int32_t* steps, int32_t* order, int32_t* n_launches,
                     double* err_indicator, int32_t n_psets);
/* Taylor cells each of the first n_psets parameter sets of the last ODE evaluation took (host copy; synchronises).
 * opts.max_cells (the reference's `maxiters`) caps them; a parameter set that would exceed the cap or the table capacity
 * comes back as a solver failure (-Inf), as mbp.jl:154-158 does. */
int  gcdyn_last_cells(gcdyn_ctx* ctx, int32_t* cells, int32_t n_psets);

/* Native forward simulator (host code, OpenMP): draws n_trees trees from the law of the reference's
 * rand_tree(model, present_time, init_type, n; reject_stubs, min_leaves, max_rejections)
 * (src/multitype_branching_process.jl:318-402, sample_child! :432-461, mutate! :411-420, pruning with the
 * child re-ordering of delete!, src/treenode.jl:38-48) and emits them already flattened.  `model` holds exactly
 * one parameter set.  Tree i of the call uses the RNG stream (seed, first_tree_index + i), so forests are
 * reproducible piecewise.  Two-phase export: ask the sizes, allocate, export, destroy. */
typedef struct gcdyn_sim gcdyn_sim;
int  gcdyn_sim_forest(const gcdyn_params* model, double present_time, int32_t init_type_idx, int64_t n_trees,
                      uint64_t seed, int64_t first_tree_index, int32_t reject_stubs, int32_t min_leaves,
                      int32_t max_rejections, gcdyn_sim** out);
int  gcdyn_sim_sizes(const gcdyn_sim* sim, int64_t* n_trees, int64_t* n_nodes, int64_t* n_child_entries,
                     int64_t* n_rejections);
int  gcdyn_sim_export(const gcdyn_sim* sim, int64_t* tree_off, uint8_t* event, int32_t* type_idx, int32_t* btype_idx,
                      double* t_start, double* t_end, int32_t* parent, int64_t* child_off, int32_t* child_idx,
                      int32_t* root_type_idx, double* root_time);
void gcdyn_sim_destroy(gcdyn_sim* sim);

/* Chebyshev/GEMM stage of the last ODE evaluation: interpolation degree, effective degree used by the
 * GEMM (max over psets), and how many psets fell back to the item sweep.  All 0 if the stage did not run. */
int  gcdyn_last_cheb_degree(gcdyn_ctx* ctx, int32_t* interp_degree, int32_t* effective_degree, int32_t* n_fallback);

/* Per-kernel timing for the roofline report: when on, every evaluation brackets its kernels with CUDA
 * events on the launching stream; gcdyn_last_kernel_ms waits for the last evaluation and returns
 * ms[6] = {setup, trajectory incl. its Chebyshev tail (or Stadler conditioning), tree_const (only when Gamma
 * changed), gemm, sweep (or alternate), finish (or reduce)}; an interval whose kernel did not run holds only the
 * event gap (~3 us). */
int  gcdyn_set_profiling(gcdyn_ctx* ctx, int on);
int  gcdyn_last_kernel_ms(gcdyn_ctx* ctx, float* ms);

/* FP64 FMA peak probe (roofline denominator; not part of the likelihood path): runs a
 * register-resident DFMA loop on every SM and returns achieved TFLOP/s. */
int  gcdyn_probe_fp64_peak(gcdyn_ctx* ctx, double* tflops);

/* ---- multi-GPU: peer all-reduce of the [P] per-pset sums ---------------------------------
 * With the trees sharded over GPUs (one process per GPU), the plain sum over trees of
 * src/multitype_branching_process.jl:258 becomes a sum over ranks of [P] doubles.  These entry points do
 * it with ONE kernel per rank over NVLink peer memory (P2P stores into every peer's window + sequence
 * flags) instead of a library collective; every rank ends with bit-identical sums.
 *   gcdyn_comm_create   allocates this rank's window (max_psets doubles per slot) and returns its 64-byte
 *                       CUDA IPC handle in `handle_out`;
 *   gcdyn_comm_connect  takes the `world` handles (64 bytes each, rank order; exchanged by the host, e.g.
 *                       torch.distributed / MPI all-gather) and maps the peers' windows;
 *   gcdyn_comm_allreduce enqueues the kernel on `stream`: `d_inout[P]` (device) is summed over ranks in
 *                       place.  Collective: every rank must call it the same number of times. */
typedef struct gcdyn_comm gcdyn_comm;
int  gcdyn_comm_create(gcdyn_ctx* ctx, int32_t rank, int32_t world, int32_t max_psets, void* handle_out,
                       gcdyn_comm** out);
int  gcdyn_comm_connect(gcdyn_comm* comm, const void* handles);
int  gcdyn_comm_allreduce(gcdyn_comm* comm, double* d_inout, int32_t n_psets, void* stream);
/* gcdyn_loglik_batch for a forest that is this rank's SHARD of the trees: `out_pset` receives the sums over all
 * ranks' trees (partial sums -> peer all-reduce -> pinned host memory, one synchronisation); `out_tree_pset` and
 * `status` stay per-shard.  Collective. */
int  gcdyn_loglik_batch_sharded(gcdyn_ctx* ctx, const gcdyn_forest* forest, const gcdyn_params* params,
                                double present_time, int method, const gcdyn_opts* opts, gcdyn_comm* comm,
                                double* out_pset, double* out_tree_pset, int32_t* status);
/* The device-resident form: like gcdyn_loglik_batch_resident (device outputs, caller's stream, no synchronisation), with
 * d_out_pset receiving the sums over all ranks' trees.  On the contraction / totals paths the all-reduce is FUSED into the
 * kernel that produces the sums: every CTA stores its parameter set's sum into every rank's window over NVLink, the last
 * one publishes the flags, waits (bounded) for the peers and adds the ranks in order - no collective is launched.
 * Collective: every rank must make the same sequence of sharded calls. */
int  gcdyn_loglik_batch_resident_sharded(gcdyn_ctx* ctx, const gcdyn_forest* forest, const gcdyn_dparams* dparams,
                                         double present_time, int method, const gcdyn_opts* opts, gcdyn_comm* comm,
                                         double* d_out_pset, double* d_out_tree_pset, int32_t* d_status, void* stream);
/* Trajectory sharding (strong scaling of the per-parameter-set work).  The trajectory integration and the coefficient
 * rows it produces do not depend on the trees, so with the trees sharded every rank would repeat them for all P sets.
 * A communicator that owns a ROW WINDOW splits them instead: on a sharded ODE evaluation of at least 512 parameter sets
 * (GCDYN_TRAJ_SHARD_MIN_PSETS overrides) without GCDYN_FLAG_TOTALS, rank r integrates the sets [P*r/world, P*(r+1)/world),
 * the rows are all-gathered by one kernel of P2P stores into every peer's window, and each rank contracts all P rows with
 * its own trees.  Values are bit-identical to the unsplit evaluation.
 *   gcdyn_comm_rows_create   allocates the window (`bytes` >= 2 * P * (Jst*8 + 32), Jst = row length at the calibrated
 *                            degree; a call whose rows do not fit fails with GCDYN_ENOMEM and names the size) and returns
 *                            its 64-byte CUDA IPC handle;
 *   gcdyn_comm_rows_connect  takes the `world` handles in rank order.
 * Every rank must hold the same parameter sets and make the same calls.  gcdyn_comm_error returns and clears the
 * communicator's error word after the caller synchronised a resident call: 0 none, 1 a peer did not arrive in time,
 * 2 the ranks disagreed on the grid degree (or a peer could not take the contraction path) - the sums are NaN / -Inf.
 * gcdyn_last_traj_sharded: 1 when the last evaluation on `ctx` integrated only this rank's slice. */
int  gcdyn_comm_rows_create(gcdyn_comm* comm, int64_t bytes, void* handle_out);
int  gcdyn_comm_rows_connect(gcdyn_comm* comm, const void* handles);
int  gcdyn_comm_error(gcdyn_comm* comm);
int  gcdyn_last_traj_sharded(gcdyn_ctx* ctx);
void gcdyn_comm_destroy(gcdyn_comm* comm);

/* ---- device-resident random-walk Metropolis (BASELINE config 4) --------------------------
 * The reference ships no sampler (README.md:3); this is the consumer of the batched likelihood.  C chains over
 * theta = (log xscale, xshift, log yscale, log yshift, log mu, log delta) of a SigmoidalBranchingProcess
 * (src/types.jl:65-110) advance in lock-step: per iteration one proposal kernel, ONE batched likelihood of C
 * parameter sets on `forest`, the peer all-reduce when `comm` is given (trees sharded over ranks), and one
 * accept/reject kernel - no host round trip inside an iteration.  Flat prior on the box [lower, upper];
 * upper[2..5] must be finite (they bound the rates that size the integrator's tables).  Random numbers are
 * a hash of (seed, iteration, chain, slot): every rank draws identical proposals. */
typedef struct {
    int32_t n_chains;
    int32_t n_types;
    const double* type_space;      /* [K]                                   */
    const double* Gamma;           /* K x K column-major, shared            */
    double rho, sigma;
    const double* theta0;          /* [n_chains*6]                          */
    const double* lower;           /* [6]                                   */
    const double* upper;           /* [6]                                   */
    double step;                   /* proposal standard deviation           */
    uint64_t seed;
} gcdyn_mh_desc;
typedef struct gcdyn_mh gcdyn_mh;
int  gcdyn_mh_create(gcdyn_ctx* ctx, const gcdyn_forest* forest, const gcdyn_mh_desc* desc, double present_time,
                     gcdyn_comm* comm, gcdyn_mh** out);
/* Runs `iterations` Metropolis steps and waits for them.  theta_hist [iterations*n_chains*6] and
 * loglik_hist [iterations*n_chains] (host, may be NULL) receive the state after every step. */
int  gcdyn_mh_run(gcdyn_mh* mh, int64_t iterations, double* theta_hist, double* loglik_hist);
/* Current state: theta [n_chains*6], loglik [n_chains], accepted [n_chains] (host; any may be NULL). */
int  gcdyn_mh_state(gcdyn_mh* mh, double* theta, double* loglik, int64_t* accepted, int64_t* iteration);
/* Restore a checkpoint written from gcdyn_mh_state: the chain continues with the random numbers of iteration+1, so a
 * restored run reproduces an uninterrupted one. */
int  gcdyn_mh_restore(gcdyn_mh* mh, const double* theta, const double* loglik, const int64_t* accepted, int64_t iteration);
void gcdyn_mh_destroy(gcdyn_mh* mh);
/* The sampler's counter-based uniform in (0,1) for (seed, iteration, chain, slot), computed on the host with the same
 * code the kernels run (slots 2j, 2j+1: Box-Muller pair of coordinate j; slot 12: the accept draw).  No device needed. */
double gcdyn_mh_uniform(uint64_t seed, uint64_t iteration, uint64_t chain, uint64_t slot);

#ifdef __cplusplus
}
#endif
#endif /* GCDYN_B200_H */
