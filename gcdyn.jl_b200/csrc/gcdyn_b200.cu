// libgcdyn_b200 — host side of the C ABI declared in include/gcdyn_b200.h.
// Handles, validation, packing, workspace management and kernel launches.  No C++ exception
// leaves this file: every extern "C" entry point catches and converts to a GCDYN_E* code.
#include "../../include/gcdyn_b200.h"
#include "kernels.cuh"
#include "cheb_gemm.cuh"
#include "simulator.h"
#include "peer_allreduce.cuh"
#include "metropolis.cuh"
#include "grad.cuh"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <chrono>
#include <cstring>
#include <new>
#include <string>
#include <vector>

using namespace gcdyn;

namespace {

thread_local std::string g_tls_error;

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

struct PinBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMallocHost(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

}  // namespace

struct gcdyn_ctx {
    int device = 0;
    int sm_count = 148;
    int smem_optin = 232448;
    cudaStream_t stream = nullptr;
    std::string err;
    // workspace (grow-only)
    DevBuf params_blob, pack, table, out_tree, out_pset, status, probe, gemm_a, flags, partial, grad_out;
    PinBuf pin_in, pin_out;
    // host-buffer calls: the shared Γ last uploaded into params_blob (and the blob's size), so that later calls with the same Γ
    // need no host→device copy at all — the per-pset scalars are then read by the kernels straight from pinned host memory
    std::vector<double> blob_gamma;
    size_t blob_doubles = 0;
    int params_mapped = 0;                // $GCDYN_PARAMS_MAPPED: 1 = read per-pset parameters from pinned host memory
    // diagnostics of the last evaluation
    int last_steps = 0, last_order = 0, last_launches = 0, last_P = 0;
    PsetPack last_pack{};
    // optional per-kernel timing (CUDA events on the launching stream)
    bool profile = false;
    cudaEvent_t ev[7] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    int last_cheb_D = 0;
    int max_cells = 0;                    // opts.max_cells of the call being enqueued (0 = the table capacity)
    bool last_traj_sharded = false;       // the last evaluation integrated only this rank's slice of the parameter sets
    const int* last_feedback = nullptr;   // pinned feedback words of the forest the last evaluation ran on
    bool gemm_attr_set = false;
    bool ev_valid = false;
};

struct gcdyn_forest {
    int device = 0;
    ForestView v{};
    std::vector<void*> allocs;
    int64_t n_chunks = 0;
    int64_t* d_chunk_off = nullptr;
    int64_t max_tree_items = 0;
    // nodal Chebyshev moments of the forest for one (present_time, grid degree) (built lazily, cached)
    mutable MomentView mom{};
    mutable double mom_T = -1.0;
    mutable double* d_Mt = nullptr;
    mutable size_t Mt_cap = 0;            // doubles
    mutable double* d_colsum = nullptr;
    mutable size_t colsum_cap = 0;        // doubles
    mutable bool mom_refused = false;     // the moments did not fit (or do not pay): this forest takes the item sweep
    // grid-degree control: the calibrated degree (0 = not calibrated yet) and the present_time it belongs to; feedback
    // words {degree of a grid some pset was NOT resolved on, probe: sufficient degree, probe: grid it was measured on}
    mutable int D_cur = 0;
    mutable double D_T = -1.0;
    mutable int* h_feedback = nullptr;    // pinned, written by finish_kernel / degree_probe_kernel
    // shared-Γ constant Σ count·log Γ_xy per tree, cached for the last Γ seen
    mutable double* d_tree_const = nullptr;   // [T] constants, then one more slot: their sum (GCDYN_FLAG_TOTALS)
    mutable std::vector<double> tc_gamma;
    bool any_self_loop = false;
};

struct gcdyn_comm {
    gcdyn_ctx* ctx = nullptr;
    int rank = 0, world = 1, cap = 0;
    void* window = nullptr;               // this rank's allocation: data then flags
    void* peer[gcdyn::PEER_MAX_WORLD] = {};   // peers' windows as mapped here (peer[rank] == window)
    bool connected = false;
    unsigned long long timeout_ns = 20000000000ULL;   // bounded wait for the peers' flags ($GCDYN_PEER_TIMEOUT_MS)
    int* h_err = nullptr;                 // pinned: raised by the kernel when a peer did not arrive in time
    // trajectory sharding: a second peer-mapped window that holds the coefficient rows of the whole batch
    //   [half 0 | half 1 | flag words [world] | CTA counter]   (peer_allreduce.cuh, row_exchange_kernel)
    void* rows = nullptr;
    void* rows_peer[gcdyn::PEER_MAX_WORLD] = {};
    size_t rows_half = 0;                 // bytes per half
    bool rows_connected = false;
    unsigned long long rows_seq = 0;      // exchanges enqueued so far (host-side: the exchange is never graph-captured)
    bool push_from_traj = true;           // $GCDYN_ROW_PUSH=kernel: the rows are copied by row_exchange_kernel instead
    int shard_min_psets = 512;            // batches smaller than this are integrated by every rank ($GCDYN_TRAJ_SHARD_MIN_PSETS)
};

struct gcdyn_sim {
    gcdyn_simu::FlatOut out;
};

struct gcdyn_dparams {
    int device = 0;
    void* blob = nullptr;
    ParamsView v{};
    double rate_max = 0.0;   // max over psets and types of λ_i + μ + γ_i, from the host copy
    std::vector<double> host_gamma;   // the shared Γ (column-major), empty when every pset has its own
};

namespace {

int fail(gcdyn_ctx* ctx, int code, const std::string& msg) {
    if (ctx) ctx->err = msg;
    g_tls_error = msg;
    return code;
}

#define CU_TRY(ctx, expr)                                                                         \
    do {                                                                                          \
        cudaError_t e__ = (expr);                                                                 \
        if (e__ != cudaSuccess)                                                                   \
            return fail(ctx, e__ == cudaErrorMemoryAllocation ? GCDYN_ENOMEM : GCDYN_ECUDA,       \
                        std::string(#expr) + ": " + cudaGetErrorString(e__));                     \
    } while (0)

// ---------------------------------------------------------------------------------------------
// parameter packing
// ---------------------------------------------------------------------------------------------
struct ParamLayout {
    size_t n_doubles = 0;
    size_t o_lambda = 0, o_xs = 0, o_xsh = 0, o_ys = 0, o_ysh = 0, o_ts = 0, o_mu = 0, o_delta = 0, o_rho = 0,
           o_sigma = 0, o_gamma = 0;
    size_t n_lambda = 0, n_gamma = 0;
};

int check_params(gcdyn_ctx* ctx, const gcdyn_params* pr) {
    if (!pr) return fail(ctx, GCDYN_EINVAL, "params is NULL");
    if (pr->n_psets <= 0) return fail(ctx, GCDYN_EINVAL, "n_psets must be positive");
    if (pr->n_types <= 0) return fail(ctx, GCDYN_EINVAL, "n_types must be positive");
    if (pr->n_types > GCDYN_MAX_TYPES) return fail(ctx, GCDYN_EUNSUPPORTED, "n_types exceeds GCDYN_MAX_TYPES");
    if (pr->response < 0 || pr->response > 2) return fail(ctx, GCDYN_EINVAL, "unknown response kind");
    if (pr->gamma_pset_stride != 0 && pr->gamma_pset_stride != pr->n_types * pr->n_types)
        return fail(ctx, GCDYN_EINVAL, "gamma_pset_stride must be 0 or n_types*n_types");
    if (!pr->mu || !pr->delta || !pr->rho || !pr->sigma || !pr->Gamma)
        return fail(ctx, GCDYN_EINVAL, "mu, delta, rho, sigma and Gamma are required");
    if (pr->response == GCDYN_LAMBDA_SIGMOID) {
        if (!pr->xscale || !pr->xshift || !pr->yscale || !pr->yshift || !pr->type_space)
            return fail(ctx, GCDYN_EINVAL, "sigmoid response needs xscale, xshift, yscale, yshift and type_space");
    } else if (!pr->lambda) {
        return fail(ctx, GCDYN_EINVAL, "lambda is required for table/constant response");
    }
    for (int p = 0; p < pr->n_psets; ++p) {   // types.jl:78-81
        if (!(pr->rho[p] >= 0.0 && pr->rho[p] <= 1.0)) return fail(ctx, GCDYN_EINVAL, "rho must be between 0 and 1");
        if (!(pr->sigma[p] >= 0.0 && pr->sigma[p] <= 1.0)) return fail(ctx, GCDYN_EINVAL, "sigma must be between 0 and 1");
    }
    return GCDYN_OK;
}

ParamLayout layout_of(const gcdyn_params* pr) {
    ParamLayout L;
    const size_t P = pr->n_psets, K = pr->n_types;
    size_t o = 0;
    auto take = [&](size_t n) { size_t r = o; o += n; return r; };
    L.n_lambda = pr->response == 0 ? P * K : (pr->response == 1 ? P : 0);
    L.o_lambda = take(L.n_lambda);
    if (pr->response == 2) { L.o_xs = take(P); L.o_xsh = take(P); L.o_ys = take(P); L.o_ysh = take(P); L.o_ts = take(K); }
    L.o_mu = take(P); L.o_delta = take(P); L.o_rho = take(P); L.o_sigma = take(P);
    L.n_gamma = pr->gamma_pset_stride ? P * K * K : K * K;
    L.o_gamma = take(L.n_gamma);
    L.n_doubles = o;
    return L;
}

void pack_host(const gcdyn_params* pr, const ParamLayout& L, double* dst) {
    const size_t P = pr->n_psets, K = pr->n_types;
    if (L.n_lambda) std::memcpy(dst + L.o_lambda, pr->lambda, L.n_lambda * 8);
    if (pr->response == 2) {
        std::memcpy(dst + L.o_xs, pr->xscale, P * 8);
        std::memcpy(dst + L.o_xsh, pr->xshift, P * 8);
        std::memcpy(dst + L.o_ys, pr->yscale, P * 8);
        std::memcpy(dst + L.o_ysh, pr->yshift, P * 8);
        std::memcpy(dst + L.o_ts, pr->type_space, K * 8);
    }
    std::memcpy(dst + L.o_mu, pr->mu, P * 8);
    std::memcpy(dst + L.o_delta, pr->delta, P * 8);
    std::memcpy(dst + L.o_rho, pr->rho, P * 8);
    std::memcpy(dst + L.o_sigma, pr->sigma, P * 8);
    std::memcpy(dst + L.o_gamma, pr->Gamma, L.n_gamma * 8);
}

ParamsView view_of(const gcdyn_params* pr, const ParamLayout& L, const double* d) {
    ParamsView v{};
    v.P = pr->n_psets; v.K = pr->n_types; v.response = pr->response; v.gamma_stride = pr->gamma_pset_stride;
    v.lambda = L.n_lambda ? d + L.o_lambda : nullptr;
    if (pr->response == 2) {
        v.xscale = d + L.o_xs; v.xshift = d + L.o_xsh; v.yscale = d + L.o_ys; v.yshift = d + L.o_ysh;
        v.type_space = d + L.o_ts;
    }
    v.mu = d + L.o_mu; v.delta = d + L.o_delta; v.rho = d + L.o_rho; v.sigma = d + L.o_sigma;
    v.Gamma = d + L.o_gamma;
    return v;
}

// max over psets and types of Λ_i = λ_i + μ + γ_i (host copy; picks the Taylor step count)
double rate_max_of(const gcdyn_params* pr) {
    const int P = pr->n_psets, K = pr->n_types;
    double best = 0.0;
    for (int p = 0; p < P; ++p) {
        const double* G = pr->Gamma + (size_t)p * pr->gamma_pset_stride;
        for (int i = 0; i < K; ++i) {
            double lam;
            if (pr->response == 0) lam = pr->lambda[(size_t)p * K + i];
            else if (pr->response == 1) lam = pr->lambda[p];
            else lam = pr->yscale[p] * (1.0 / (1.0 + std::exp(-(pr->xscale[p] * (pr->type_space[i] - pr->xshift[p]))))) + pr->yshift[p];
            // off-diagonal mass also bounds the spectrum of δΓ; use |λ| + |μ| + 2|γ_i|
            const double g = std::fabs(pr->delta[p] * G[i + (size_t)i * K]);
            const double r = std::fabs(lam) + std::fabs(pr->mu[p]) + 2.0 * g;
            if (std::isfinite(r) && r > best) best = r;   // a pset with non-finite rates fails on its own; it must not size the others' tables
        }
    }
    return best;
}

// ---------------------------------------------------------------------------------------------
// step-count rule: h·rate_max ≤ target(M), from the numerical study in tests/algo_model.py
// ---------------------------------------------------------------------------------------------
constexpr int S_CAP = 8192;
constexpr int FLAG_FORCE_MOMENTS = 0x100;   // internal: take the moment path whatever the batch size (gradients need it)

int normalise_order(int order) {
    if (order == 0) return 16;
    const int allowed[] = {8, 12, 16, 20, 24};
    for (int a : allowed) if (order == a) return a;
    return -1;
}

int steps_for(double present_time, double rate_max, double target) {
    if (!(rate_max == rate_max) || !(present_time > 0.0)) return 4;
    double s = std::ceil(present_time * rate_max / target);
    if (s < 4.0) s = 4.0;
    if (s > (double)S_CAP) s = (double)S_CAP;
    return (int)s;
}

// Capacity (in cells) of the per-pset table: 1.5× the step count the calibrated worst case needs.
int cap_steps(double present_time, double rate_max, int M) {
    const int safe = steps_for(present_time, rate_max, hrate_target_safe(M));
    return std::min(S_CAP, ladder_ceil((long long)safe + safe / 2));
}

// Cells the fastest pset is expected to need (the kernel's own first guess, one rung up): sizes the
// shared-memory staging buffer of the sweep.
int expected_steps(double present_time, double rate_max, int M) {
    return ladder_next(ladder_ceil(steps_for(present_time, rate_max, hrate_target_start(M))));
}

// ---------------------------------------------------------------------------------------------
// workspace carving for PsetPack
// ---------------------------------------------------------------------------------------------
int carve_pack(gcdyn_ctx* ctx, int P, int K, PsetPack* pk) {
    const size_t ntw = (size_t)K + (size_t)K * K + 2;
    const size_t PK = (size_t)P * K;
    const size_t n_doubles = PK * 6 + (size_t)P * ntw + (size_t)P;
    const size_t bytes = n_doubles * 8 + (size_t)P * 8 + 64;
    CU_TRY(ctx, ctx->pack.ensure(bytes));
    double* d = ctx->pack.as<double>();
    pk->ntw = (int)ntw;
    pk->lam = d; d += PK;
    pk->Lam = d; d += PK;
    pk->logcond = d; d += PK;
    pk->st_c = d; d += PK;
    pk->st_A = d; d += PK;
    pk->st_B = d; d += PK;
    pk->nt = d; d += (size_t)P * ntw;
    pk->perr = d; d += P;
    pk->pstatus = reinterpret_cast<int*>(d);
    pk->S = pk->pstatus + P;
    return GCDYN_OK;
}

// Dynamic shared memory of traj_kernel: [d_n broadcast buffers | δΓ (K > 64 only)] for the integrating warps, then
// TrajSmem: [cos table | node values | check | ring meta | ring of A_n | two row buffers].  Mirrors the kernel.
size_t traj_smem_bytes(int K, int D, int M) {
    const size_t KSB = K <= 32 ? (((size_t)K + 3) & ~(size_t)3) : (K <= 64 ? (((size_t)K + 7) & ~(size_t)7) : (((size_t)K + 3) & ~(size_t)3));
    size_t n = 2 * KSB + (K <= 64 ? 0 : (((size_t)K * K + 1) & ~(size_t)1));
    const size_t KS = ((size_t)K + 1) & ~(size_t)1;
    if (D > 0) n += 2 * (size_t)D + (((size_t)D + 3) & ~(size_t)1) + ((((size_t)D + 1) * K + 1) & ~(size_t)1) + 24 * (size_t)K;
    const size_t slots = K <= 64 ? RING_SLOTS : RING_SLOTS_WIDE;      // ring_slots(KPL): the K > 64 layout keeps two cells in flight
    n += slots * 4 + slots * (M + 1) * KS;
    if (D > 0) n += 2 * (size_t)K * (M + 2);
    return n * sizeof(double);
}

template <int M, int KP, int KPL, int NPW = 1>
int launch_traj_inst(gcdyn_ctx* ctx, const ParamsView& pv, const PsetPack& pk, const TrajCfg& cfg, const ChebCfg& cc,
                     double* tab, cudaStream_t st) {
    const size_t smem = traj_smem_bytes(pv.K, cc.D, M);
    if (smem > 48 * 1024)
        CU_TRY(ctx, cudaFuncSetAttribute(traj_kernel<M, KP, KPL, NPW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    traj_kernel<M, KP, KPL, NPW><<<cfg.n_ctas > 0 ? cfg.n_ctas : pv.P, TRAJ_THREADS, smem, st>>>(pv, pk, cfg, cc, tab);
    CU_TRY(ctx, cudaGetLastError());
    return GCDYN_OK;
}

template <int M>
int launch_traj_m(gcdyn_ctx* ctx, const ParamsView& pv, const PsetPack& pk, const TrajCfg& cfg, const ChebCfg& cc,
                  double* tab, cudaStream_t st) {
    const int K = pv.K;
    if (K <= 4) return launch_traj_inst<M, 4, 1>(ctx, pv, pk, cfg, cc, tab, st);
    if (K <= 8) return launch_traj_inst<M, 8, 1>(ctx, pv, pk, cfg, cc, tab, st);
    if (K <= 12) return launch_traj_inst<M, 12, 1>(ctx, pv, pk, cfg, cc, tab, st);
    if (K <= 16) return launch_traj_inst<M, 16, 1>(ctx, pv, pk, cfg, cc, tab, st);
    if (K <= 20) return launch_traj_inst<M, 20, 1>(ctx, pv, pk, cfg, cc, tab, st);
    if (K <= 24) return launch_traj_inst<M, 24, 1>(ctx, pv, pk, cfg, cc, tab, st);
    if (K <= 28) return launch_traj_inst<M, 28, 1>(ctx, pv, pk, cfg, cc, tab, st);
    if (K <= 32) return launch_traj_inst<M, 32, 1>(ctx, pv, pk, cfg, cc, tab, st);
    // 32 < K ≤ 64: two integrating warps, one type per thread, δΓ row in registers
    if (K <= 40) return launch_traj_inst<M, 40, 1, 2>(ctx, pv, pk, cfg, cc, tab, st);
    if (K <= 48) return launch_traj_inst<M, 48, 1, 2>(ctx, pv, pk, cfg, cc, tab, st);
    if (K <= 56) return launch_traj_inst<M, 56, 1, 2>(ctx, pv, pk, cfg, cc, tab, st);
    if (K <= 64) return launch_traj_inst<M, 64, 1, 2>(ctx, pv, pk, cfg, cc, tab, st);
    return launch_traj_inst<M, 0, 4>(ctx, pv, pk, cfg, cc, tab, st);
}

int launch_traj(gcdyn_ctx* ctx, int M, const ParamsView& pv, const PsetPack& pk, const TrajCfg& cfg, const ChebCfg& cc,
                double* tab, cudaStream_t st) {
    switch (M) {
        case 8: return launch_traj_m<8>(ctx, pv, pk, cfg, cc, tab, st);
        case 12: return launch_traj_m<12>(ctx, pv, pk, cfg, cc, tab, st);
        case 16: return launch_traj_m<16>(ctx, pv, pk, cfg, cc, tab, st);
        case 20: return launch_traj_m<20>(ctx, pv, pk, cfg, cc, tab, st);
        case 24: return launch_traj_m<24>(ctx, pv, pk, cfg, cc, tab, st);
    }
    return fail(ctx, GCDYN_EINVAL, "unsupported Taylor order");
}

template <int M>
int launch_sweep_m(gcdyn_ctx* ctx, const gcdyn_forest* f, const PsetPack& pk, const double* tab, long long tab_stride,
                   int S_cap, int grid_doubles, double T, int P, size_t want_smem, const int* only_flagged,
                   double* out_tree, int32_t* status, cudaStream_t st) {
    // dynamic shared memory = staging buffer for one pset's table (bulk-copied); psets whose table is
    // larger than the buffer are read from global memory by the same kernel
    size_t smem = std::min(std::max(want_smem, (size_t)64 * 1024), (size_t)ctx->smem_optin - 2048);
    smem &= ~(size_t)127;
    if (smem > 48 * 1024)
        CU_TRY(ctx, cudaFuncSetAttribute(sweep_kernel<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 1;
    CU_TRY(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, sweep_kernel<M>, SWEEP_THREADS, smem));
    if (per_sm < 1) per_sm = 1;
    const long long units = (long long)P * f->n_chunks;
    const long long grid = std::max<long long>(1, std::min<long long>(units, (long long)ctx->sm_count * per_sm));
    sweep_kernel<M><<<(unsigned)grid, SWEEP_THREADS, smem, st>>>(f->v, pk, tab, tab_stride, S_cap, grid_doubles,
                                                      f->d_chunk_off, (int)f->n_chunks, P, T,
                                                      (int)(smem / 8), only_flagged, out_tree, status);
    CU_TRY(ctx, cudaGetLastError());
    return GCDYN_OK;
}

int launch_sweep(gcdyn_ctx* ctx, int M, const gcdyn_forest* f, const PsetPack& pk, const double* tab,
                 long long tab_stride, int S_cap, int grid_doubles, double T, int P, size_t want_smem,
                 const int* only_flagged, double* out_tree, int32_t* status, cudaStream_t st) {
    switch (M) {
        case 8: return launch_sweep_m<8>(ctx, f, pk, tab, tab_stride, S_cap, grid_doubles, T, P, want_smem, only_flagged, out_tree, status, st);
        case 12: return launch_sweep_m<12>(ctx, f, pk, tab, tab_stride, S_cap, grid_doubles, T, P, want_smem, only_flagged, out_tree, status, st);
        case 16: return launch_sweep_m<16>(ctx, f, pk, tab, tab_stride, S_cap, grid_doubles, T, P, want_smem, only_flagged, out_tree, status, st);
        case 20: return launch_sweep_m<20>(ctx, f, pk, tab, tab_stride, S_cap, grid_doubles, T, P, want_smem, only_flagged, out_tree, status, st);
        case 24: return launch_sweep_m<24>(ctx, f, pk, tab, tab_stride, S_cap, grid_doubles, T, P, want_smem, only_flagged, out_tree, status, st);
    }
    return fail(ctx, GCDYN_EINVAL, "unsupported Taylor order");
}

// Degrees the nodal Chebyshev grid may take.
// every even degree up to 64 (a column block of K doubles per degree: no reason to round further), then coarser rungs
int degree_ladder_ceil(int d) {
    if (d <= 8) return 8;
    if (d <= 64) return (d + 1) & ~1;
    const int coarse[] = {72, 80, 96, 112, 128};
    for (int v : coarse) if (v >= d) return v;
    return 128;
}

// Degree of the Chebyshev–Gauss–Lobatto grid.  The FIRST evaluation on a forest (per present_time) calibrates it: it runs
// on a generous grid sized from the rates, a probe measures the degree the batch's interpolants really need (full
// Chebyshev spectrum of a few psets), the call synchronises ONCE, adopts that degree and repeats itself on it.  From
// then on the degree is a property of the forest handle — every later evaluation uses it, so results are bit-identical
// from call to call.  Psets the grid does not resolve are summed exactly (slower, still correct); only when a quarter of
// a batch is in that state — the parameters have moved to another regime — does finish_kernel report back through a
// pinned word, and the next call takes the next rung.
size_t fit_degree_smem(int K, int D) { return (2 * (size_t)D + ((size_t)D + 1) * K) * 8 + 1024; }
// The nodal stage needs node values, cosine table and check scratch next to the ring: for the widest layouts it does not fit
// at any degree, and the evaluation takes the exact item sweep (which needs none of it).
bool nodal_stage_fits(int K, int M, int smem_optin) {
    return traj_smem_bytes(K, 8, M) + 1024 <= (size_t)smem_optin && fit_degree_smem(K, 8) <= (size_t)smem_optin;
}
int clamp_degree(int D, int K, int M, int smem_optin) {
    while (D > 8 && (traj_smem_bytes(K, D, M) + 1024 > (size_t)smem_optin || fit_degree_smem(K, D) > (size_t)smem_optin)) D -= 2;
    return D;
}
int initial_degree(double T, double rate, int K, int M, int smem_optin) {
    double want = 2.0 * T * rate + 16.0;
    if (!(want == want) || want > 128.0) want = 128.0;
    return clamp_degree(degree_ladder_ceil((int)std::ceil(want)), K, M, smem_optin);
}

// Build (once per forest, present_time and grid degree) the per-tree nodal Chebyshev moments and count columns.
// Returns GCDYN_OK with f->mom.Mt == nullptr when the moments are refused (they do not fit or do not pay): the caller
// then takes the exact item sweep, which needs none of this memory.
int ensure_moments(gcdyn_ctx* ctx, const gcdyn_forest* f, double T, int D, cudaStream_t st, int* launches) {
    if (f->mom.Mt && f->mom_T == T && f->mom.D == D) return GCDYN_OK;
    f->mom.Mt = nullptr;
    if (f->mom_refused) return GCDYN_OK;
    const int K = f->v.K;
    const int C1 = 2 * K + 4;
    long long J1 = C1 + (long long)(D + 1) * K;
    J1 = (J1 + 3) & ~3LL;
    long long Jtc = J1;
    long long Jst = Jtc + (long long)K * K;
    Jst = (Jst + 3) & ~3LL;
    const int64_t Tn = f->v.n_trees;
    // cost model: the contraction does 2·Jst flops per (tree, pset) on the tensor pipe, the sweep ≈ 40 flops per item at
    // a far lower rate; with fewer than ≈ Jst/100 items per tree the sweep wins.  Memory: never take more than half of
    // what is free.
    const size_t need = std::max<size_t>((size_t)Tn * (size_t)Jst, 1);
    size_t free_b = 0, total_b = 0;
    cudaMemGetInfo(&free_b, &total_b);
    const bool pays = (double)Tn * (double)Jst <= 100.0 * (double)std::max<int64_t>(f->v.n_items, 1);
    const bool fits = need <= f->Mt_cap || need * 8 <= free_b / 2;
    if (!pays || !fits) { f->mom_refused = true; return GCDYN_OK; }
    if (need > f->Mt_cap || (size_t)Jst > f->colsum_cap) {
        if (f->d_Mt) cudaFree(f->d_Mt);
        if (f->d_colsum) cudaFree(f->d_colsum);
        f->d_Mt = f->d_colsum = nullptr; f->Mt_cap = f->colsum_cap = 0;
        cudaError_t e1 = cudaMalloc(&f->d_Mt, need * 8);
        cudaError_t e2 = e1 == cudaSuccess ? cudaMalloc(&f->d_colsum, (size_t)Jst * 8) : e1;
        if (e1 != cudaSuccess || e2 != cudaSuccess) {     // not an error of the evaluation: the sweep path needs no moments
            cudaGetLastError();
            if (f->d_Mt) cudaFree(f->d_Mt);
            f->d_Mt = f->d_colsum = nullptr;
            f->mom_refused = true;
            return GCDYN_OK;
        }
        f->Mt_cap = need; f->colsum_cap = (size_t)Jst;
    }
    if (!f->h_feedback) {
        CU_TRY(ctx, cudaHostAlloc(&f->h_feedback, 4 * sizeof(int), cudaHostAllocMapped));
        f->h_feedback[0] = f->h_feedback[1] = f->h_feedback[2] = f->h_feedback[3] = 0;
    }
    const size_t smem = ((size_t)C1 + (size_t)K * K) * 4;
    if (smem > 48 * 1024)
        CU_TRY(ctx, cudaFuncSetAttribute(counts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (Tn > 0) {
        counts_kernel<<<(unsigned)Tn, 128, smem, st>>>(f->v, D, Jtc, Jst, f->d_Mt);
        CU_TRY(ctx, cudaGetLastError());
        const long long warps = (long long)Tn * K;
        moments_kernel<<<(unsigned)((warps + 7) / 8), 256, 0, st>>>(f->v, T, D, Jst, C1, f->d_Mt);
        CU_TRY(ctx, cudaGetLastError());
        const size_t nsm = (2 * (size_t)D + ((size_t)D + 1) * K) * 8;
        if (nsm > 48 * 1024)
            CU_TRY(ctx, cudaFuncSetAttribute(nodal_transform_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)nsm));
        nodal_transform_kernel<<<(unsigned)Tn, 256, nsm, st>>>(K, D, C1, Jst, f->d_Mt);
        CU_TRY(ctx, cudaGetLastError());
        *launches += 3;
    }
    colsum_kernel<<<(unsigned)((Jst + 31) / 32), 256, 0, st>>>(f->d_Mt, Tn, Jst, f->d_colsum);
    CU_TRY(ctx, cudaGetLastError());
    *launches += 1;
    f->mom.K = K; f->mom.C1 = C1; f->mom.D = D; f->mom.J1 = J1; f->mom.Jtc = Jtc; f->mom.Jst = Jst;
    f->mom.Mt = f->d_Mt; f->mom.colsum = f->d_colsum;
    f->mom_T = T;
    f->tc_gamma.clear();
    return GCDYN_OK;
}

// Launch configuration with programmatic stream serialization: the kernel may become resident while its predecessor
// in the stream is still running; it calls pdl_wait() before touching the predecessor's output.
thread_local cudaLaunchAttribute g_pdl_attr[1];
cudaLaunchConfig_t pdl_config(dim3 grid, dim3 block, size_t smem, cudaStream_t st) {
    g_pdl_attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    g_pdl_attr[0].val.programmaticStreamSerializationAllowed = 1;
    cudaLaunchConfig_t lc{};
    lc.gridDim = grid;
    lc.blockDim = block;
    lc.dynamicSmemBytes = smem;
    lc.stream = st;
    lc.attrs = g_pdl_attr;
    lc.numAttrs = 1;
    return lc;
}

template <int M>
int launch_finish_m(gcdyn_ctx* ctx, const gcdyn_forest* f, const ParamsView& pv, const PsetPack& pk, const double* tab, const TrajCfg& cfg, double T,
                    const double* partial, int splits, const double* tree_const, const int* need_sweep, int P,
                    double* out_tree, int32_t* status, double* out_pset, cudaStream_t st, int totals_only,
                    const gcdyn::FusedAllReduce& fc, int mh_on, const gcdyn::MhView& mh) {
    cudaLaunchConfig_t lc = pdl_config(dim3((unsigned)P), dim3(256), 0, st);
    CU_TRY(ctx, cudaLaunchKernelEx(&lc, finish_kernel<M>, f->v, pv, pk, tab, cfg.tab_stride, cfg.S_cap, cfg.grid_doubles, T, partial,
                                   splits, tree_const, need_sweep, P, f->mom.D, const_cast<int*>(need_sweep) + P, f->h_feedback, totals_only, out_tree, status, out_pset, fc, mh_on, mh));
    return GCDYN_OK;
}

int launch_finish(gcdyn_ctx* ctx, int M, const gcdyn_forest* f, const ParamsView& pv, const PsetPack& pk, const double* tab, const TrajCfg& cfg,
                  double T, const double* partial, int splits, const double* tree_const, const int* need_sweep, int P,
                  double* out_tree, int32_t* status, double* out_pset, cudaStream_t st, int totals_only,
                  const gcdyn::FusedAllReduce& fc, int mh_on = 0, const gcdyn::MhView& mh = gcdyn::MhView{}) {
    switch (M) {
        case 8: return launch_finish_m<8>(ctx, f, pv, pk, tab, cfg, T, partial, splits, tree_const, need_sweep, P, out_tree, status, out_pset, st, totals_only, fc, mh_on, mh);
        case 12: return launch_finish_m<12>(ctx, f, pv, pk, tab, cfg, T, partial, splits, tree_const, need_sweep, P, out_tree, status, out_pset, st, totals_only, fc, mh_on, mh);
        case 16: return launch_finish_m<16>(ctx, f, pv, pk, tab, cfg, T, partial, splits, tree_const, need_sweep, P, out_tree, status, out_pset, st, totals_only, fc, mh_on, mh);
        case 20: return launch_finish_m<20>(ctx, f, pv, pk, tab, cfg, T, partial, splits, tree_const, need_sweep, P, out_tree, status, out_pset, st, totals_only, fc, mh_on, mh);
        case 24: return launch_finish_m<24>(ctx, f, pv, pk, tab, cfg, T, partial, splits, tree_const, need_sweep, P, out_tree, status, out_pset, st, totals_only, fc, mh_on, mh);
    }
    return fail(ctx, GCDYN_EINVAL, "unsupported Taylor order");
}

int comm_enqueue(gcdyn_comm* c, const double* d_in, double* out, int n_psets, cudaStream_t st);   // peer all-reduce (below)
int comm_fused_view(gcdyn_comm* c, int n_psets, gcdyn::FusedAllReduce* out);                      // its form fused into finish_kernel
void comm_poison(gcdyn_comm* c, int n_psets, cudaStream_t st);                                     // NaN contribution of a rank that failed
int comm_row_exchange(gcdyn_comm* c, unsigned long long seq, int P, long long Jst, long long Juse, int p0, int p1, int D,
                      const PsetPack& pk, int* need_sweep, cudaStream_t st);                                             // all-gather of the coefficient rows

// Enqueue one evaluation on `st`.  All pointers are device pointers; d_out_tree/d_status may be null.
// `comm`: the forest is this rank's shard — d_out_pset receives the sums over all ranks (fused into finish_kernel on the
// contraction / totals paths, a separate one-CTA kernel behind the other paths).
// S_fixed > 0 pins the number of Taylor cells; otherwise every pset adapts it inside traj_kernel.
int enqueue_eval(gcdyn_ctx* ctx, const gcdyn_forest* f, const ParamsView& pv, double rate_max, double T, int method,
                 int M, int S_fixed, double tol, int flags, const double* host_gamma_shared, double* d_out_pset,
                 double* d_out_tree, int32_t* d_status, cudaStream_t st, PsetPack* pk_out, double rate_typical = 0.0,
                 gcdyn_comm* comm = nullptr, const gcdyn::MhView* mh = nullptr, bool* mh_fused = nullptr) {
    // rate_max bounds every pset's rates (table capacity); rate_typical (> 0: the sampler's current rates) only picks
    // the Chebyshev interpolation degree — a pset it is too small for is flagged and summed exactly
    const int P = pv.P, K = pv.K;
    if (P > 65535) return fail(ctx, GCDYN_EUNSUPPORTED, "more than 65535 parameter sets per call");
    if (comm && comm->world <= 1) comm = nullptr;         // a world of one rank: nothing to combine
    if (mh_fused) *mh_fused = false;
    PsetPack pk{};
    int rc = carve_pack(ctx, P, K, &pk);
    if (rc) return rc;
    const int64_t Tn = f->v.n_trees;
    if (!d_out_tree) {
        CU_TRY(ctx, ctx->out_tree.ensure((size_t)P * Tn * 8));
        d_out_tree = ctx->out_tree.as<double>();
    }
    int launches = 0;
    const bool prof = ctx->profile && ctx->ev[0];
    ctx->ev_valid = false;
    if (prof) cudaEventRecord(ctx->ev[0], st);
    if (method == GCDYN_METHOD_ODE) {
        if (prof) cudaEventRecord(ctx->ev[1], st);
        TrajCfg cfg{};
        cfg.T = T;
        cfg.tol = tol;
        cfg.S_fixed = S_fixed;
        cfg.S_cap = S_fixed > 0 ? S_fixed : cap_steps(T, rate_max, M);
        if (S_fixed <= 0 && ctx->max_cells > 0) cfg.S_cap = std::min(cfg.S_cap, ctx->max_cells);   // the caller's `maxiters`
        const long long L = M + 2;
        cfg.grid_doubles = 2 * cfg.S_cap + 2;
        cfg.tab_stride = cfg.grid_doubles + (long long)cfg.S_cap * K * L;
        const size_t tab_bytes = (size_t)P * (size_t)cfg.tab_stride * 8;
        CU_TRY(ctx, ctx->table.ensure(tab_bytes));
        // ---- per-tree sums: Chebyshev-moment GEMM (DMMA), with the item sweep as exact fallback.  The Chebyshev
        //      stage (node values = coefficient row, resolution check, fallback flag) runs inside traj_kernel.
        if (traj_smem_bytes(K, 0, M) + 1024 > (size_t)ctx->smem_optin)
            return fail(ctx, GCDYN_EUNSUPPORTED, "this many types at this Taylor order do not fit the trajectory kernel's shared memory");
        bool use_gemm = !(flags & GCDYN_FLAG_SWEEP_ONLY) && Tn > 0 && !f->mom_refused && nodal_stage_fits(K, M, ctx->smem_optin) &&
                        (P >= 8 || (f->mom.Mt && f->mom_T == T) || (flags & FLAG_FORCE_MOMENTS));
        const int* only_flagged = nullptr;
        ChebCfg cc{};
        int* need_sweep = nullptr;
        bool calibrating = false;
        if (use_gemm) {
            int D;
            if (f->D_cur == 0 || f->D_T != T) {
                D = initial_degree(T, rate_typical > 0.0 ? std::min(rate_typical, rate_max) : rate_max, K, M, ctx->smem_optin);
                calibrating = true;
            } else {
                if (f->h_feedback && f->h_feedback[0] == f->D_cur && f->D_cur < 128)      // some pset was not resolved on it
                    f->D_cur = clamp_degree(degree_ladder_ceil(f->D_cur + 1), K, M, ctx->smem_optin);
                D = f->D_cur;
            }
            rc = ensure_moments(ctx, f, T, D, st, &launches);
            if (rc) return rc;
            if (!f->mom.Mt) use_gemm = false;              // refused: too large for the free memory, or the sweep is cheaper
        }
        // trajectory sharding (trees sharded over ranks AND a batch of several waves): every rank integrates a slice of the
        // parameter sets and the coefficient rows are all-gathered over peer memory.  Whether the exchange takes place is a
        // function of (P, world, flags) alone — every rank decides alike and makes the same number of exchanges.
        const bool shard_eligible = comm && comm->rows_connected && !mh && !(flags & (GCDYN_FLAG_TOTALS | GCDYN_FLAG_SWEEP_ONLY)) &&
                                    P >= comm->shard_min_psets && P >= comm->world;
        bool shard_traj = false;
        unsigned long long row_seq = 0;
        int own_lo = 0, own_hi = P;
        if (shard_eligible) {
            row_seq = ++comm->rows_seq;
            const size_t need_bytes = use_gemm ? (size_t)P * (size_t)f->mom.Jst * 8 + (size_t)P * sizeof(gcdyn::RowMeta) : 0;
            if (!use_gemm || need_bytes > comm->rows_half) {
                // this rank cannot take part (no contraction here, or the window is too small for this degree): tell the
                // peers (degree 0 in the flag word) so that nobody contracts rows that never arrive; every rank reports
                rc = comm_row_exchange(comm, row_seq, 0, 0, 0, 0, 0, 0, pk, nullptr, st);
                if (rc) return rc;
                ++launches;
                if (use_gemm) return fail(ctx, GCDYN_ENOMEM, "row window too small for this batch: gcdyn_comm_rows_create needs " +
                                                             std::to_string(2 * need_bytes) + " bytes");
            } else {
                shard_traj = true;
                own_lo = (int)((long long)P * comm->rank / comm->world);
                own_hi = (int)((long long)P * (comm->rank + 1) / comm->world);
            }
        }
        if (use_gemm) {
            if (!shard_traj) CU_TRY(ctx, ctx->gemm_a.ensure((size_t)P * (size_t)f->mom.Jst * 8));
            CU_TRY(ctx, ctx->flags.ensure((size_t)(P + 2) * 4));
            need_sweep = ctx->flags.as<int>();
            cc.T = T;
            cc.tol = 1e-11;
            cc.D = f->mom.D;
            cc.compact_tc = pv.gamma_stride == 0 ? 1 : 0;
            cc.mv = f->mom;
            cc.A = shard_traj ? reinterpret_cast<double*>(static_cast<char*>(comm->rows) + (size_t)(row_seq & 1ULL) * comm->rows_half)
                              : ctx->gemm_a.as<double>();
            cc.need_sweep = need_sweep;
            cc.nflag = need_sweep + P;
        }
        const double* tree_const = nullptr;
        gcdyn::FusedAllReduce fc{};
        fc.pv.world = 1;
        // (the calibrating first call takes the full path once; its repetition on the adopted grid is the totals call)
        const bool totals = use_gemm && (flags & GCDYN_FLAG_TOTALS) && !calibrating;
        if (use_gemm) {
            if (cc.compact_tc) {
                // shared Γ: Σ count·log Γ_xy is the same for every pset — one constant per tree, recomputed only
                // when Γ (or the moments it is built from) changed since the last call on this forest; done BEFORE the
                // trajectory kernel so that nothing stands between it and the contraction
                const size_t kk = (size_t)K * K;
                if (!f->d_tree_const) CU_TRY(ctx, cudaMalloc(&f->d_tree_const, ((size_t)Tn + 1) * 8));
                const bool same = host_gamma_shared && f->tc_gamma.size() == kk &&
                                  std::memcmp(f->tc_gamma.data(), host_gamma_shared, kk * 8) == 0;
                if (!same) {
                    tc_const_kernel<<<(unsigned)((Tn + 7) / 8), 256, 0, st>>>(f->mom, Tn, pv.Gamma, f->d_tree_const);
                    CU_TRY(ctx, cudaGetLastError());
                    reduce_kernel<<<1, 256, 0, st>>>(f->d_tree_const, Tn, f->d_tree_const + Tn);   // their sum, for the totals path
                    CU_TRY(ctx, cudaGetLastError());
                    launches += 2;
                    if (host_gamma_shared) f->tc_gamma.assign(host_gamma_shared, host_gamma_shared + kk);
                    else f->tc_gamma.clear();
                }
                tree_const = f->d_tree_const;
            }
            if (comm) {                                    // fused all-reduce: this rank's sums never leave the kernel chain
                rc = comm_fused_view(comm, P, &fc);
                if (rc) return rc;
                if (totals) {
                    CU_TRY(ctx, ctx->out_pset.ensure((size_t)P * 8));
                    fc.totals_in = ctx->out_pset.as<double>();
                }
            }
            if (totals && mh) {                            // fused Metropolis iteration (see traj_kernel / finish_kernel)
                cc.mh_on = 1;
                cc.mh_sharded = comm ? 1 : 0;
                cc.mh = *mh;
                if (mh_fused) *mh_fused = true;
            }
            if (totals) {
                cc.totals = comm ? ctx->out_pset.as<double>() : d_out_pset;
                cc.tc_total = tree_const ? tree_const + Tn : nullptr;
                cc.any_self_loop = f->any_self_loop ? 1 : 0;
            }
        }
        if (shard_traj) {
            cfg.p_base = own_lo; cfg.n_ctas = own_hi - own_lo; cfg.own_lo = own_lo; cfg.own_hi = own_hi;
            // the rows travel from the CTAs that finish them (GCDYN_ROW_PUSH=kernel: from row_exchange_kernel instead)
            if (comm->push_from_traj) {
                cc.push.w2 = (int)((cc.compact_tc ? f->mom.J1 : f->mom.Jst) / 2);
                for (int r = 0; r < comm->world; ++r)
                    if (r != comm->rank)
                        cc.push.dst[cc.push.n++] = reinterpret_cast<double*>(static_cast<char*>(comm->rows_peer[r]) + (size_t)(row_seq & 1ULL) * comm->rows_half);
            }
        }
        rc = launch_traj(ctx, M, pv, pk, cfg, cc, ctx->table.as<double>(), st);
        if (rc) return rc;
        ++launches;
        if (shard_traj) {
            const long long juse = cc.push.n > 0 ? 0 : (cc.compact_tc ? f->mom.J1 : f->mom.Jst);
            cc.push.n = 0;
            rc = comm_row_exchange(comm, row_seq, P, f->mom.Jst, juse, own_lo, own_hi, cc.D, pk, need_sweep, st);
            if (rc) return rc;
            // psets of other ranks that were flagged for the exact fallback: their Taylor tables are needed here
            cfg.p_base = 0; cfg.n_ctas = 0; cfg.only_missing = 1;
            rc = launch_traj(ctx, M, pv, pk, cfg, cc, ctx->table.as<double>(), st);
            if (rc) return rc;
            cfg.only_missing = 0;
            launches += 2;
        }
        ctx->last_traj_sharded = shard_traj;
        if (prof) cudaEventRecord(ctx->ev[2], st);
        if (totals) {
            // the per-pset sums are already in d_out_pset; finish_kernel only redoes the psets the grid did not resolve
            if (prof) { cudaEventRecord(ctx->ev[3], st); cudaEventRecord(ctx->ev[4], st); cudaEventRecord(ctx->ev[5], st); }
            rc = launch_finish(ctx, M, f, pv, pk, ctx->table.as<double>(), cfg, T, nullptr, 0, nullptr, need_sweep, P,
                               d_out_tree, d_status, d_out_pset, st, 1, fc, cc.mh_on, cc.mh);
            if (rc) return rc;
            ++launches;
            if (prof) { cudaEventRecord(ctx->ev[6], st); ctx->ev_valid = true; }
            ctx->last_cheb_D = cc.D; ctx->last_feedback = f->h_feedback; ctx->last_steps = cfg.S_cap; ctx->last_order = M;
            ctx->last_launches = launches; ctx->last_P = P; ctx->last_pack = pk;
            if (pk_out) *pk_out = pk;
            return GCDYN_OK;
        }
        if (use_gemm) {
            const MomentView& mv = f->mom;
            if (prof) cudaEventRecord(ctx->ev[3], st);
            // one linear tile index, pset tiles fastest: the CTAs resident together share a few tree panels of Mt and all of
            // A (which always fits in L2), so Mt streams from HBM once per launch — with tree tiles fastest a forest whose
            // moments exceed L2 re-read all of Mt for every row of pset tiles (config 5: 5.7 GB of DRAM reads for a 0.2 GB
            // operand, profiles/r2_ncu_c5_traj_gemm_summary.csv)
            const long long n_tt = (Tn + GT_TILE - 1) / GT_TILE, n_pt = (P + GT_TILE - 1) / GT_TILE;
            if (n_tt * n_pt > 0x7fffffffLL) return fail(ctx, GCDYN_EUNSUPPORTED, "too many (tree, pset) tiles for one launch");
            dim3 grid((unsigned)(n_tt * n_pt), 1);
            if (!ctx->gemm_attr_set) {
                CU_TRY(ctx, cudaFuncSetAttribute(gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GT_SMEM_BYTES));
                ctx->gemm_attr_set = true;
            }
            // split the column range when the tile grid alone cannot fill the machine (deterministic: fixed order in finish)
            const long long tiles = n_tt * n_pt;
            // (measured at config-3 size, 256 tiles: 1 → 24.9 µs, 2 → 25.8, 3 → 24.7, 4 → 22.9, 6 → 22.8 with a slower finish)
            const int splits = tiles >= 4LL * ctx->sm_count ? 1 : (tiles >= 2LL * ctx->sm_count ? 2 : 4);
            grid.z = (unsigned)splits;
            CU_TRY(ctx, ctx->partial.ensure((size_t)splits * P * (size_t)Tn * 8));
            {
                cudaLaunchConfig_t lc = pdl_config(grid, dim3(128), GT_SMEM_BYTES, st);
                CU_TRY(ctx, cudaLaunchKernelEx(&lc, gemm_kernel, (const double*)cc.A, mv, cc.compact_tc ? 0 : 1,
                                               P, (int64_t)Tn, ctx->partial.as<double>()));
            }
            ++launches;
            if (prof) { cudaEventRecord(ctx->ev[4], st); cudaEventRecord(ctx->ev[5], st); }
            rc = launch_finish(ctx, M, f, pv, pk, ctx->table.as<double>(), cfg, T, ctx->partial.as<double>(), splits, tree_const,
                               need_sweep, P, d_out_tree, d_status, d_out_pset, st, 0, fc);
            if (rc) return rc;
            ++launches;
            if (prof) { cudaEventRecord(ctx->ev[6], st); ctx->ev_valid = true; }
            if (calibrating) {
                // first evaluation on this forest: measure the degree the batch really needs, adopt it, and repeat
                int adopt = cc.D;
                const size_t psm = (2 * (size_t)cc.D + 2 * ((size_t)cc.D + 1) * K) * 8;
                if (psm + 1024 <= (size_t)ctx->smem_optin) {
                    if (psm > 48 * 1024)
                        CU_TRY(ctx, cudaFuncSetAttribute(degree_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)psm));
                    CU_TRY(ctx, ctx->probe.ensure(64));
                    CU_TRY(ctx, cudaMemsetAsync(ctx->probe.p, 0, 8, st));
                    degree_probe_kernel<<<P, 256, psm, st>>>((const double*)cc.A, mv, (const int*)need_sweep,
                                                             (const int*)pk.pstatus, cc.tol, ctx->probe.as<int>());
                    CU_TRY(ctx, cudaGetLastError());
                    CU_TRY(ctx, cudaMemcpyAsync(f->h_feedback + 1, ctx->probe.p, 8, cudaMemcpyDeviceToHost, st));
                    CU_TRY(ctx, cudaStreamSynchronize(st));
                    // +3: the last three coefficients of the adopted grid must lie in the resolved tail; a sampler (rate_typical
                    // given) gets six more degrees of head-room, because its parameter sets will move away from this batch
                    const int margin = 3 + (rate_typical > 0.0 ? 6 : 0);
                    if (f->h_feedback[2] == cc.D && f->h_feedback[1] > 4)
                        adopt = std::min(cc.D, clamp_degree(degree_ladder_ceil(f->h_feedback[1] + margin), K, M, ctx->smem_optin));
                }
                f->D_cur = adopt;
                f->D_T = T;
                // (sharded: every rank repeats, whatever it adopted — the ranks must make the same number of collective calls)
                if (adopt != cc.D || comm)
                    return enqueue_eval(ctx, f, pv, rate_max, T, method, M, S_fixed, tol, flags, host_gamma_shared, d_out_pset,
                                        d_out_tree == ctx->out_tree.as<double>() ? nullptr : d_out_tree, d_status, st, pk_out, rate_typical, comm, mh, mh_fused);
            }
            ctx->last_cheb_D = cc.D;
            ctx->last_feedback = f->h_feedback;
            ctx->last_steps = cfg.S_cap;
            ctx->last_order = M;
            ctx->last_launches = launches;
            ctx->last_P = P;
            ctx->last_pack = pk;
            if (pk_out) *pk_out = pk;
            return GCDYN_OK;
        } else {
            ctx->last_cheb_D = 0;
        }
        if (prof) { cudaEventRecord(ctx->ev[3], st); cudaEventRecord(ctx->ev[4], st); }
        const int S_stage = S_fixed > 0 ? S_fixed : std::min(cfg.S_cap, expected_steps(T, rate_max, M));
        const size_t want_smem = ((size_t)cfg.grid_doubles + (size_t)S_stage * K * L) * 8;
        rc = launch_sweep(ctx, M, f, pk, ctx->table.as<double>(), cfg.tab_stride, cfg.S_cap, cfg.grid_doubles, T, P,
                          want_smem, only_flagged,
                          d_out_tree, d_status, st);
        if (rc) return rc;
        ++launches;
        ctx->last_steps = cfg.S_cap;
    } else {
        prep_kernel<<<P, 128, 0, st>>>(pv, pk, method);
        CU_TRY(ctx, cudaGetLastError());
        ++launches;
        if (prof) cudaEventRecord(ctx->ev[1], st);
        if (method == GCDYN_METHOD_STADLER) {
            stadler_cond_kernel<<<P, 128, 0, st>>>(pv, pk, T);
            CU_TRY(ctx, cudaGetLastError());
            ++launches;
        }
        if (prof) { cudaEventRecord(ctx->ev[2], st); cudaEventRecord(ctx->ev[3], st); cudaEventRecord(ctx->ev[4], st); }
        dim3 grid((unsigned)f->n_chunks, (unsigned)P);
        alt_kernel<<<grid, 256, 0, st>>>(f->v, pv, pk, f->d_chunk_off, T, method, d_out_tree, d_status);
        CU_TRY(ctx, cudaGetLastError());
        ++launches;
        ctx->last_steps = 0;
    }
    if (prof) cudaEventRecord(ctx->ev[5], st);
    if (comm) {                                            // sweep / alternate paths: partial sums, then the one-CTA all-reduce
        CU_TRY(ctx, ctx->out_pset.ensure((size_t)P * 8));
        reduce_kernel<<<P, 256, 0, st>>>(d_out_tree, Tn, ctx->out_pset.as<double>());
        CU_TRY(ctx, cudaGetLastError());
        rc = comm_enqueue(comm, ctx->out_pset.as<double>(), d_out_pset, P, st);
        if (rc) return rc;
        ++launches;
    } else {
        reduce_kernel<<<P, 256, 0, st>>>(d_out_tree, Tn, d_out_pset);
        CU_TRY(ctx, cudaGetLastError());
    }
    ++launches;
    if (prof) { cudaEventRecord(ctx->ev[6], st); ctx->ev_valid = true; }
    ctx->last_order = method == GCDYN_METHOD_ODE ? M : 0;
    ctx->last_launches = launches;
    ctx->last_P = P;
    ctx->last_pack = pk;
    if (pk_out) *pk_out = pk;
    return GCDYN_OK;
}

template <class T>
int upload(gcdyn_ctx* ctx, gcdyn_forest* f, const T* src, size_t n, const T** dst) {
    void* d = nullptr;
    CU_TRY(ctx, cudaMalloc(&d, std::max<size_t>(n, 1) * sizeof(T)));
    f->allocs.push_back(d);
    if (n) CU_TRY(ctx, cudaMemcpy(d, src, n * sizeof(T), cudaMemcpyHostToDevice));
    *dst = reinterpret_cast<const T*>(d);
    return GCDYN_OK;
}

void free_forest(gcdyn_forest* f) {
    if (!f) return;
    int prev = 0;
    cudaGetDevice(&prev);
    cudaSetDevice(f->device);
    for (void* p : f->allocs) cudaFree(p);
    if (f->d_Mt) cudaFree(f->d_Mt);
    if (f->d_colsum) cudaFree(f->d_colsum);
    if (f->d_tree_const) cudaFree(f->d_tree_const);
    if (f->h_feedback) cudaFreeHost(f->h_feedback);
    cudaSetDevice(prev);
    delete f;
}

int forest_create_impl(gcdyn_ctx* ctx, const gcdyn_forest_desc* d, gcdyn_forest** out) {
    if (!ctx || !d || !out) return fail(ctx, GCDYN_EINVAL, "NULL argument");
    *out = nullptr;
    const int64_t T = d->n_trees, N = d->n_nodes;
    const int K = d->n_types;
    if (T < 0 || N < 0) return fail(ctx, GCDYN_EINVAL, "negative sizes");
    if (K <= 0) return fail(ctx, GCDYN_EINVAL, "n_types must be positive");
    if (K > GCDYN_MAX_TYPES) return fail(ctx, GCDYN_EUNSUPPORTED, "n_types exceeds GCDYN_MAX_TYPES");
    if (!d->tree_off || !d->root_type_idx || !d->root_time) return fail(ctx, GCDYN_EINVAL, "tree_off/root arrays are required");
    if (N > 0 && (!d->event || !d->type_idx || !d->btype_idx || !d->t_start || !d->t_end))
        return fail(ctx, GCDYN_EINVAL, "per-node arrays are required");
    if (d->tree_off[0] != 0 || d->tree_off[T] != N) return fail(ctx, GCDYN_EINVAL, "tree_off must start at 0 and end at n_nodes");
    for (int64_t t = 0; t < T; ++t)
        if (d->tree_off[t + 1] < d->tree_off[t]) return fail(ctx, GCDYN_EINVAL, "tree_off must be non-decreasing");
    // A tree the reference gives -Inf for (self-loop, mbp.jl:181-185) was not examined past the offending node by the
    // flattener either: its remaining records may carry placeholder indices.  Such trees are never looked at by the
    // kernels beyond their self_loop flag, so their records are exempt from the index checks (and yield no lookup items).
    for (int64_t t = 0; t < T; ++t) {
        const bool looped = d->self_loop && d->self_loop[t];
        for (int64_t i = d->tree_off[t]; i < d->tree_off[t + 1]; ++i) {
            if (d->event[i] > 6) return fail(ctx, GCDYN_EINVAL, "event code out of range");
            if (!(d->t_start[i] == d->t_start[i]) || !(d->t_end[i] == d->t_end[i])) return fail(ctx, GCDYN_EINVAL, "NaN node time");
            if (looped) continue;
            if (d->btype_idx[i] < 0 || d->btype_idx[i] >= K) return fail(ctx, GCDYN_EINVAL, "btype_idx out of range");
            if (d->type_idx[i] < -1 || d->type_idx[i] >= K) return fail(ctx, GCDYN_EINVAL, "type_idx out of range");
            if (d->event[i] == EV_TC && d->type_idx[i] < 0) return fail(ctx, GCDYN_EINVAL, "type_change to a type outside the type space");
        }
    }
    for (int64_t t = 0; t < T; ++t)
        if (d->root_type_idx[t] < -1 || d->root_type_idx[t] >= K) return fail(ctx, GCDYN_EINVAL, "root_type_idx out of range");
    if (d->parent && d->child_off && d->child_idx) {   // structural cross-check of the CSR the flattener built
        for (int64_t t = 0; t < T; ++t) {
            const int64_t lo = d->tree_off[t], n = d->tree_off[t + 1] - lo;
            for (int64_t i = 0; i < n; ++i) {
                const int32_t par = d->parent[lo + i];
                if (par < -1 || par >= n || (par >= 0 && par <= i)) return fail(ctx, GCDYN_EINVAL, "parent is not a later post-order position");
                for (int64_t c = d->child_off[lo + i]; c < d->child_off[lo + i + 1]; ++c) {
                    const int32_t ch = d->child_idx[c];
                    if (ch < 0 || ch >= i || d->parent[lo + ch] != (int32_t)i) return fail(ctx, GCDYN_EINVAL, "child_idx inconsistent with parent");
                }
            }
        }
    }

    // ---- lookup items of the regrouped ODE sum
    std::vector<int64_t> item_off(T + 1, 0);
    std::vector<double> it_time;
    std::vector<int2> it_meta;
    std::vector<int32_t> n_death(T, 0), n_surv(T, 0);
    std::vector<uint8_t> live_all, loop_all;
    it_time.reserve(N + T); it_meta.reserve(N + T);
    auto push = [&](double tm, int type, int w, int aux) {
        it_time.push_back(tm);
        it_meta.push_back(make_int2((int32_t)((uint32_t)type | ((uint32_t)(uint8_t)(int8_t)w << 24)), aux));
    };
    int64_t max_items = 0;
    for (int64_t t = 0; t < T; ++t) {
        const bool looped = d->self_loop && d->self_loop[t];
        for (int64_t i = d->tree_off[t]; i < d->tree_off[t + 1] && !looped; ++i) {
            if (d->live && !d->live[i]) continue;
            const int ev = d->event[i], x = d->btype_idx[i], y = d->type_idx[i];
            const double te = d->t_end[i];
            if (ev == EV_BIRTH) {
                if (y == x || y < 0) push(te, x, +1, x);
                else { push(te, x, -1, x); push(te, y, +2, -1); }
            } else if (ev == EV_TC) {
                push(te, x, -1, K + x * K + y);
                push(te, y, +1, -1);
            } else if (ev == EV_SDEATH) {
                push(te, x, -1, -1);
                ++n_death[t];
            } else if (ev == EV_SSURV) {
                ++n_surv[t];
            }
        }
        const int rt = d->root_type_idx[t] < 0 ? 0 : d->root_type_idx[t];
        push(d->root_time[t], rt, +1, -1);
        item_off[t + 1] = (int64_t)it_time.size();
        max_items = std::max(max_items, item_off[t + 1] - item_off[t]);
    }
    // ---- chunks of trees with roughly equal work
    const int64_t n_items = (int64_t)it_time.size();
    const int64_t target = 4096;
    std::vector<int64_t> chunk_off{0};
    int64_t acc = 0;
    for (int64_t t = 0; t < T; ++t) {
        const int64_t wgt = std::max<int64_t>(item_off[t + 1] - item_off[t], d->tree_off[t + 1] - d->tree_off[t]);
        acc += wgt;
        if (acc >= target) { chunk_off.push_back(t + 1); acc = 0; }
    }
    if (chunk_off.back() != T) chunk_off.push_back(T);
    if (chunk_off.size() == 1) chunk_off.push_back(0);   // empty forest: one empty chunk
    if ((int64_t)chunk_off.size() - 1 > 2147483647LL) return fail(ctx, GCDYN_EUNSUPPORTED, "too many chunks");

    gcdyn_forest* f = new (std::nothrow) gcdyn_forest();
    if (!f) return fail(ctx, GCDYN_ENOMEM, "host allocation failed");
    f->device = ctx->device;
    f->max_tree_items = max_items;
    if (d->self_loop) for (int64_t t = 0; t < T; ++t) if (d->self_loop[t]) { f->any_self_loop = true; break; }
    ForestView& v = f->v;
    v.n_trees = T; v.n_nodes = N; v.n_items = n_items; v.K = K;
    if (!d->live) live_all.assign((size_t)std::max<int64_t>(N, 1), 1);
    if (!d->self_loop) loop_all.assign((size_t)std::max<int64_t>(T, 1), 0);
    int rc = GCDYN_OK;
#define UP(field, src, n) if (!rc) rc = upload(ctx, f, src, (size_t)(n), &v.field)
    UP(tree_off, d->tree_off, T + 1);
    UP(event, d->event, N);
    UP(type_idx, d->type_idx, N);
    UP(btype_idx, d->btype_idx, N);
    UP(t_start, d->t_start, N);
    UP(t_end, d->t_end, N);
    UP(live, d->live ? d->live : live_all.data(), N);
    UP(root_type, d->root_type_idx, T);
    UP(root_time, d->root_time, T);
    UP(self_loop, d->self_loop ? d->self_loop : loop_all.data(), T);
    UP(item_off, item_off.data(), T + 1);
    UP(it_time, it_time.data(), n_items);
    UP(it_meta, it_meta.data(), n_items);
    UP(n_death, n_death.data(), T);
    UP(n_surv, n_surv.data(), T);
#undef UP
    const int64_t* dco = nullptr;
    if (!rc) rc = upload(ctx, f, chunk_off.data(), chunk_off.size(), &dco);
    if (rc) { free_forest(f); return rc; }
    f->d_chunk_off = const_cast<int64_t*>(dco);
    f->n_chunks = (int64_t)chunk_off.size() - 1;
    *out = f;
    return GCDYN_OK;
}


int batch_impl(gcdyn_ctx* ctx, const gcdyn_forest* f, const gcdyn_params* pr, double T, int method,
               const gcdyn_opts* opts, double* out_pset, double* out_tree_pset, int32_t* status, gcdyn_comm* comm = nullptr) {
    if (!ctx || !f || !out_pset) return fail(ctx, GCDYN_EINVAL, "NULL argument");
    // $GCDYN_E2E_TRACE: host-side timeline of the call (microseconds between the stamps), averaged over 10 calls, on stderr
    static const bool e2e_trace = std::getenv("GCDYN_E2E_TRACE") != nullptr;
    static double e2e_acc[5] = {0, 0, 0, 0, 0};
    static int e2e_n = 0;
    auto now = [] { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double ts0 = e2e_trace ? now() : 0.0;
    int rc = check_params(ctx, pr);
    if (rc) return rc;
    if (method < 0 || method > 3) return fail(ctx, GCDYN_EINVAL, "unknown method");
    if (pr->n_types != f->v.K) return fail(ctx, GCDYN_EINVAL, "params and forest disagree on n_types");
    if (f->device != ctx->device) return fail(ctx, GCDYN_EINVAL, "forest lives on another device");
    if (method != GCDYN_METHOD_NAIVE && !(T > 0.0) ) return fail(ctx, GCDYN_EINVAL, "present_time must be positive");
    gcdyn_opts o{};
    if (opts) o = *opts;
    const int M = normalise_order(o.order);
    if (M < 0) return fail(ctx, GCDYN_EINVAL, "order must be 0, 8, 12, 16, 20 or 24");
    if (o.steps < 0 || o.steps > S_CAP) return fail(ctx, GCDYN_EINVAL, "steps out of range");
    const double tol = o.tol > 0.0 ? o.tol : 1e-11;
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const int P = pr->n_psets;
    const int64_t Tn = f->v.n_trees;

    // parameters: pack into pinned staging, one H2D copy
    const ParamLayout L = layout_of(pr);
    CU_TRY(ctx, ctx->pin_in.ensure(L.n_doubles * 8));
    pack_host(pr, L, reinterpret_cast<double*>(ctx->pin_in.p));
    CU_TRY(ctx, ctx->params_blob.ensure(L.n_doubles * 8));
    const double ts1 = e2e_trace ? now() : 0.0;
    ParamsView pv = view_of(pr, L, ctx->params_blob.as<double>());
    {
        const size_t kk = (size_t)pr->n_types * pr->n_types;
        const bool same_gamma = ctx->params_mapped && pr->gamma_pset_stride == 0 && ctx->blob_doubles == L.n_doubles &&
                                ctx->blob_gamma.size() == kk && std::memcmp(ctx->blob_gamma.data(), pr->Gamma, kk * 8) == 0;
        if (same_gamma) {
            // no copy node: the kernels read this call's per-pset parameters over PCIe from the pinned staging buffer
            // (a few hundred bytes per CTA); Γ — K² doubles that EVERY CTA reads — stays in the device blob
            const double* gamma_dev = pv.Gamma;
            pv = view_of(pr, L, reinterpret_cast<const double*>(ctx->pin_in.p));
            pv.Gamma = gamma_dev;
        } else {
            CU_TRY(ctx, cudaMemcpyAsync(ctx->params_blob.p, ctx->pin_in.p, L.n_doubles * 8, cudaMemcpyHostToDevice, st));
            if (pr->gamma_pset_stride == 0) { ctx->blob_gamma.assign(pr->Gamma, pr->Gamma + kk); ctx->blob_doubles = L.n_doubles; }
            else { ctx->blob_gamma.clear(); ctx->blob_doubles = 0; }
        }
    }
    const double ts2 = e2e_trace ? now() : 0.0;

    const bool want_rows = out_tree_pset || status;
    if ((o.flags & GCDYN_FLAG_TOTALS) && want_rows) return fail(ctx, GCDYN_EINVAL, "GCDYN_FLAG_TOTALS produces no per-tree values: out_tree_pset and status must be NULL");
    if (status) CU_TRY(ctx, ctx->status.ensure((size_t)P * Tn * 4));
    if (comm) CU_TRY(ctx, ctx->out_pset.ensure((size_t)P * 8));
    // the per-pset sums are written by the last kernel straight into pinned host memory (device-accessible under
    // unified addressing): no device→host copy node for the 8·P bytes a sampler waits on
    CU_TRY(ctx, ctx->pin_out.ensure((size_t)P * 8 + 64));
    double* h_out = reinterpret_cast<double*>(ctx->pin_out.p);

    // one pass: every pset adapts its own step count inside traj_kernel (or uses opts->steps as given);
    // a pset that cannot be resolved within the table capacity comes back as a solver failure (-Inf)
    ctx->max_cells = o.max_cells > 0 ? o.max_cells : 0;
    rc = enqueue_eval(ctx, f, pv, rate_max_of(pr), T, method, M, o.steps, tol, o.flags,
                      pr->gamma_pset_stride == 0 ? pr->Gamma : nullptr, h_out, nullptr,
                      status ? ctx->status.as<int32_t>() : nullptr, st, nullptr, 0.0, comm);
    if (rc && comm) {
        // a rank-local failure must still release the peers: contribute NaN (every rank then sees NaN sums, and this rank
        // returns its error) instead of leaving them to wait for a flag that never comes
        const std::string why = ctx->err;
        cudaGetLastError();
        comm_poison(comm, P, st);
        cudaStreamSynchronize(st);
        return fail(ctx, rc, why);
    }
    if (rc) return rc;
    if (want_rows) {
        if (out_tree_pset)
            CU_TRY(ctx, cudaMemcpyAsync(out_tree_pset, ctx->out_tree.p, (size_t)P * Tn * 8, cudaMemcpyDeviceToHost, st));
        if (status)
            CU_TRY(ctx, cudaMemcpyAsync(status, ctx->status.p, (size_t)P * Tn * 4, cudaMemcpyDeviceToHost, st));
    }
    const double ts3 = e2e_trace ? now() : 0.0;
    CU_TRY(ctx, cudaStreamSynchronize(st));
    const double ts4 = e2e_trace ? now() : 0.0;
    std::memcpy(out_pset, h_out, (size_t)P * 8);
    if (e2e_trace) {
        const double ts5 = now();
        e2e_acc[0] += ts1 - ts0; e2e_acc[1] += ts2 - ts1; e2e_acc[2] += ts3 - ts2; e2e_acc[3] += ts4 - ts3; e2e_acc[4] += ts5 - ts4;
        if (++e2e_n == 10) {
            std::fprintf(stderr, "[gcdyn e2e] validate+pack %.1f | H2D enqueue %.1f | enqueue kernels %.1f | synchronize %.1f | copy out %.1f us\n",
                         e2e_acc[0] / 10, e2e_acc[1] / 10, e2e_acc[2] / 10, e2e_acc[3] / 10, e2e_acc[4] / 10);
            e2e_n = 0;
            for (double& a : e2e_acc) a = 0.0;
        }
    }
    if (comm && comm->h_err && *comm->h_err) {
        const int why = *comm->h_err;
        *comm->h_err = 0;
        return fail(ctx, GCDYN_ECUDA, (why & 2) ? "row exchange: the ranks disagree on the grid degree, or a rank could not take the contraction path"
                                                : "peer all-reduce timed out: a rank did not arrive (its evaluation failed or it is gone)");
    }
    return GCDYN_OK;
}

}  // namespace

// =================================================================================================
// extern "C" entry points
// =================================================================================================
#define GUARD_BEGIN try {
#define GUARD_END(ctx)                                                                        \
    } catch (const std::bad_alloc&) { return fail(ctx, GCDYN_ENOMEM, "host allocation failed"); } \
    catch (const std::exception& e) { return fail(ctx, GCDYN_EINVAL, e.what()); }             \
    catch (...) { return fail(ctx, GCDYN_EINVAL, "unknown C++ exception"); }

namespace {
template <int M, int KP>
int launch_grad_inst(gcdyn_ctx* ctx, const ParamsView& pv, const gcdyn::GradCfg& gc, cudaStream_t st) {
    constexpr int NV = 1 + gcdyn::GRAD_MAX_DIRS;
    const size_t doubles = (size_t)2 * NV * KP + (size_t)NV * (M + 1) * KP + (((size_t)gc.D + 3) & ~(size_t)1) + (size_t)NV * KP;
    const size_t smem = doubles * 8;
    if (smem > 48 * 1024)
        CU_TRY(ctx, cudaFuncSetAttribute(gcdyn::traj_grad_kernel<M, KP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    gcdyn::traj_grad_kernel<M, KP><<<pv.P, gcdyn::GRAD_THREADS, smem, st>>>(pv, gc);
    CU_TRY(ctx, cudaGetLastError());
    return GCDYN_OK;
}
template <int M>
int launch_grad_m(gcdyn_ctx* ctx, const ParamsView& pv, const gcdyn::GradCfg& gc, cudaStream_t st) {
    const int K = pv.K;
    if (K <= 4) return launch_grad_inst<M, 4>(ctx, pv, gc, st);
    if (K <= 8) return launch_grad_inst<M, 8>(ctx, pv, gc, st);
    if (K <= 12) return launch_grad_inst<M, 12>(ctx, pv, gc, st);
    if (K <= 16) return launch_grad_inst<M, 16>(ctx, pv, gc, st);
    if (K <= 20) return launch_grad_inst<M, 20>(ctx, pv, gc, st);
    if (K <= 24) return launch_grad_inst<M, 24>(ctx, pv, gc, st);
    if (K <= 28) return launch_grad_inst<M, 28>(ctx, pv, gc, st);
    return launch_grad_inst<M, 32>(ctx, pv, gc, st);
}
}  // namespace

extern "C" {

int gcdyn_version(void) { return GCDYN_B200_VERSION; }

int gcdyn_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

const char* gcdyn_last_error(const gcdyn_ctx* ctx) {
    if (ctx) return ctx->err.c_str();
    return g_tls_error.c_str();
}

int gcdyn_ctx_create(int device, gcdyn_ctx** out) {
    gcdyn_ctx* none = nullptr;
    GUARD_BEGIN
    if (!out) return fail(none, GCDYN_EINVAL, "out is NULL");
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) {
        cudaGetLastError();
        return fail(none, GCDYN_ENODEVICE, "no CUDA device available (libgcdyn_b200 has no CPU fallback)");
    }
    if (device < 0 || device >= n) return fail(none, GCDYN_EINVAL, "device index out of range");
    CU_TRY(none, cudaSetDevice(device));
    cudaDeviceProp prop{};
    CU_TRY(none, cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10)
        return fail(none, GCDYN_ENODEVICE, std::string("device is ") + prop.name + " (sm_" + std::to_string(prop.major) +
                                               std::to_string(prop.minor) + "); this library is built for sm_100a only");
    gcdyn_ctx* ctx = new gcdyn_ctx();
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    ctx->smem_optin = (int)prop.sharedMemPerBlockOptin;
    if (const char* ev = std::getenv("GCDYN_PARAMS_MAPPED")) ctx->params_mapped = std::atoi(ev);
    cudaError_t e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) { delete ctx; return fail(none, GCDYN_ECUDA, cudaGetErrorString(e)); }
    *out = ctx;
    return GCDYN_OK;
    GUARD_END(none)
}

void gcdyn_ctx_destroy(gcdyn_ctx* ctx) {
    if (!ctx) return;
    int prev = 0;
    cudaGetDevice(&prev);
    cudaSetDevice(ctx->device);
    if (ctx->stream) { cudaStreamSynchronize(ctx->stream); cudaStreamDestroy(ctx->stream); }
    for (auto& e : ctx->ev) if (e) cudaEventDestroy(e);
    ctx->params_blob.release(); ctx->pack.release(); ctx->table.release(); ctx->out_tree.release();
    ctx->out_pset.release(); ctx->status.release(); ctx->probe.release(); ctx->gemm_a.release(); ctx->flags.release(); ctx->partial.release(); ctx->grad_out.release();
    ctx->pin_in.release(); ctx->pin_out.release();
    cudaSetDevice(prev);
    delete ctx;
}

int gcdyn_forest_create(gcdyn_ctx* ctx, const gcdyn_forest_desc* desc, gcdyn_forest** out) {
    GUARD_BEGIN
    if (!ctx) return fail(ctx, GCDYN_EINVAL, "ctx is NULL");
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    return forest_create_impl(ctx, desc, out);
    GUARD_END(ctx)
}

void gcdyn_forest_destroy(gcdyn_forest* forest) { free_forest(forest); }

int gcdyn_forest_info(const gcdyn_forest* f, int64_t* n_trees, int64_t* n_nodes, int64_t* n_items, int32_t* n_types) {
    if (!f) return fail(nullptr, GCDYN_EINVAL, "forest is NULL");
    if (n_trees) *n_trees = f->v.n_trees;
    if (n_nodes) *n_nodes = f->v.n_nodes;
    if (n_items) *n_items = f->v.n_items;
    if (n_types) *n_types = f->v.K;
    return GCDYN_OK;
}

int gcdyn_forest_event_counts(gcdyn_ctx* ctx, const gcdyn_forest* f, int64_t* out) {
    GUARD_BEGIN
    if (!ctx || !f || !out) return fail(ctx, GCDYN_EINVAL, "NULL argument");
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    const int64_t T = f->v.n_trees;
    if (T == 0) return GCDYN_OK;
    CU_TRY(ctx, ctx->status.ensure((size_t)T * 7 * 8));
    int64_t* d = ctx->status.as<int64_t>();
    const int wpb = 8;
    event_count_kernel<<<(unsigned)((T + wpb - 1) / wpb), wpb * 32, 0, ctx->stream>>>(f->v, d);
    CU_TRY(ctx, cudaGetLastError());
    CU_TRY(ctx, cudaMemcpyAsync(out, d, (size_t)T * 7 * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return GCDYN_OK;
    GUARD_END(ctx)
}

int gcdyn_loglik_batch(gcdyn_ctx* ctx, const gcdyn_forest* forest, const gcdyn_params* params, double present_time,
                       int method, const gcdyn_opts* opts, double* out_pset, double* out_tree_pset, int32_t* status) {
    GUARD_BEGIN
    return batch_impl(ctx, forest, params, present_time, method, opts, out_pset, out_tree_pset, status);
    GUARD_END(ctx)
}

int gcdyn_params_upload(gcdyn_ctx* ctx, const gcdyn_params* pr, gcdyn_dparams** out) {
    GUARD_BEGIN
    if (!ctx || !out) return fail(ctx, GCDYN_EINVAL, "NULL argument");
    *out = nullptr;
    int rc = check_params(ctx, pr);
    if (rc) return rc;
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    const ParamLayout L = layout_of(pr);
    std::vector<double> host(L.n_doubles);
    pack_host(pr, L, host.data());
    gcdyn_dparams* dp = new gcdyn_dparams();
    dp->device = ctx->device;
    cudaError_t e = cudaMalloc(&dp->blob, L.n_doubles * 8);
    if (e == cudaSuccess) e = cudaMemcpy(dp->blob, host.data(), L.n_doubles * 8, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        if (dp->blob) cudaFree(dp->blob);
        delete dp;
        return fail(ctx, GCDYN_ECUDA, cudaGetErrorString(e));
    }
    dp->v = view_of(pr, L, reinterpret_cast<const double*>(dp->blob));
    dp->rate_max = rate_max_of(pr);
    if (pr->gamma_pset_stride == 0) dp->host_gamma.assign(pr->Gamma, pr->Gamma + (size_t)pr->n_types * pr->n_types);
    *out = dp;
    return GCDYN_OK;
    GUARD_END(ctx)
}

void gcdyn_dparams_destroy(gcdyn_dparams* dp) {
    if (!dp) return;
    int prev = 0;
    cudaGetDevice(&prev);
    cudaSetDevice(dp->device);
    if (dp->blob) cudaFree(dp->blob);
    cudaSetDevice(prev);
    delete dp;
}

int gcdyn_loglik_batch_resident(gcdyn_ctx* ctx, const gcdyn_forest* f, const gcdyn_dparams* dp, double present_time,
                                int method, const gcdyn_opts* opts, double* d_out_pset, double* d_out_tree_pset,
                                int32_t* d_status, void* stream) {
    GUARD_BEGIN
    if (!ctx || !f || !dp || !d_out_pset) return fail(ctx, GCDYN_EINVAL, "NULL argument");
    if (method < 0 || method > 3) return fail(ctx, GCDYN_EINVAL, "unknown method");
    if (dp->v.K != f->v.K) return fail(ctx, GCDYN_EINVAL, "params and forest disagree on n_types");
    if (f->device != ctx->device || dp->device != ctx->device) return fail(ctx, GCDYN_EINVAL, "handles live on another device");
    if (method != GCDYN_METHOD_NAIVE && !(present_time > 0.0)) return fail(ctx, GCDYN_EINVAL, "present_time must be positive");
    gcdyn_opts o{};
    if (opts) o = *opts;
    const int M = normalise_order(o.order);
    if (M < 0) return fail(ctx, GCDYN_EINVAL, "order must be 0, 8, 12, 16, 20 or 24");
    if (o.steps < 0 || o.steps > S_CAP) return fail(ctx, GCDYN_EINVAL, "steps out of range");
    if ((o.flags & GCDYN_FLAG_TOTALS) && (d_out_tree_pset || d_status))
        return fail(ctx, GCDYN_EINVAL, "GCDYN_FLAG_TOTALS produces no per-tree values: d_out_tree_pset and d_status must be NULL");
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = stream ? reinterpret_cast<cudaStream_t>(stream) : ctx->stream;
    ctx->max_cells = o.max_cells > 0 ? o.max_cells : 0;
    return enqueue_eval(ctx, f, dp->v, dp->rate_max, present_time, method, M, o.steps, o.tol > 0.0 ? o.tol : 1e-11,
                        o.flags, dp->host_gamma.empty() ? nullptr : dp->host_gamma.data(), d_out_pset, d_out_tree_pset,
                        d_status, st, nullptr);
    GUARD_END(ctx)
}

int gcdyn_loglik_batch_resident_sharded(gcdyn_ctx* ctx, const gcdyn_forest* f, const gcdyn_dparams* dp, double present_time,
                                        int method, const gcdyn_opts* opts, gcdyn_comm* comm, double* d_out_pset,
                                        double* d_out_tree_pset, int32_t* d_status, void* stream) {
    GUARD_BEGIN
    if (!ctx || !f || !dp || !d_out_pset || !comm) return fail(ctx, GCDYN_EINVAL, "NULL argument");
    if (comm->ctx != ctx) return fail(ctx, GCDYN_EINVAL, "comm belongs to another context");
    if (method < 0 || method > 3) return fail(ctx, GCDYN_EINVAL, "unknown method");
    if (dp->v.K != f->v.K) return fail(ctx, GCDYN_EINVAL, "params and forest disagree on n_types");
    if (f->device != ctx->device || dp->device != ctx->device) return fail(ctx, GCDYN_EINVAL, "handles live on another device");
    if (method != GCDYN_METHOD_NAIVE && !(present_time > 0.0)) return fail(ctx, GCDYN_EINVAL, "present_time must be positive");
    gcdyn_opts o{};
    if (opts) o = *opts;
    const int M = normalise_order(o.order);
    if (M < 0) return fail(ctx, GCDYN_EINVAL, "order must be 0, 8, 12, 16, 20 or 24");
    if (o.steps < 0 || o.steps > S_CAP) return fail(ctx, GCDYN_EINVAL, "steps out of range");
    if ((o.flags & GCDYN_FLAG_TOTALS) && (d_out_tree_pset || d_status))
        return fail(ctx, GCDYN_EINVAL, "GCDYN_FLAG_TOTALS produces no per-tree values: d_out_tree_pset and d_status must be NULL");
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = stream ? reinterpret_cast<cudaStream_t>(stream) : ctx->stream;
    ctx->max_cells = o.max_cells > 0 ? o.max_cells : 0;
    const int rc = enqueue_eval(ctx, f, dp->v, dp->rate_max, present_time, method, M, o.steps, o.tol > 0.0 ? o.tol : 1e-11,
                                o.flags, dp->host_gamma.empty() ? nullptr : dp->host_gamma.data(), d_out_pset, d_out_tree_pset,
                                d_status, st, nullptr, 0.0, comm);
    if (rc) {                                              // release the peers: NaN contribution instead of a missing one
        const std::string why = ctx->err;
        cudaGetLastError();
        comm_poison(comm, dp->v.P, st);
        return fail(ctx, rc, why);
    }
    return GCDYN_OK;
    GUARD_END(ctx)
}

int gcdyn_last_diag(gcdyn_ctx* ctx, int32_t* steps, int32_t* order, int32_t* n_launches, double* err_indicator,
                    int32_t n_psets) {
    GUARD_BEGIN
    if (!ctx) return fail(ctx, GCDYN_EINVAL, "ctx is NULL");
    if (steps) {
        *steps = ctx->last_steps;
        if (ctx->last_order > 0 && ctx->last_pack.S && ctx->last_P > 0) {   // largest per-pset step count actually used
            std::vector<int> hs(ctx->last_P);
            CU_TRY(ctx, cudaSetDevice(ctx->device));
            CU_TRY(ctx, cudaDeviceSynchronize());
            CU_TRY(ctx, cudaMemcpy(hs.data(), ctx->last_pack.S, (size_t)ctx->last_P * 4, cudaMemcpyDeviceToHost));
            *steps = *std::max_element(hs.begin(), hs.end());
        }
    }
    if (order) *order = ctx->last_order;
    if (n_launches) *n_launches = ctx->last_launches;
    if (err_indicator && n_psets > 0) {
        if (n_psets > ctx->last_P || !ctx->last_pack.perr) return fail(ctx, GCDYN_EINVAL, "no evaluation with that many psets to report on");
        CU_TRY(ctx, cudaSetDevice(ctx->device));
        CU_TRY(ctx, cudaDeviceSynchronize());
        CU_TRY(ctx, cudaMemcpy(err_indicator, ctx->last_pack.perr, (size_t)n_psets * 8, cudaMemcpyDeviceToHost));
    }
    return GCDYN_OK;
    GUARD_END(ctx)
}

int gcdyn_loglik_grad_batch(gcdyn_ctx* ctx, const gcdyn_forest* f, const gcdyn_params* pr, double T, const gcdyn_opts* opts,
                            double* out_pset, double* out_grad, int32_t* n_dirs) {
    GUARD_BEGIN
    if (!ctx || !f || !out_pset || !out_grad) return fail(ctx, GCDYN_EINVAL, "NULL argument");
    int rc = check_params(ctx, pr);
    if (rc) return rc;
    if (pr->n_types != f->v.K) return fail(ctx, GCDYN_EINVAL, "params and forest disagree on n_types");
    if (f->device != ctx->device) return fail(ctx, GCDYN_EINVAL, "forest lives on another device");
    if (!(T > 0.0)) return fail(ctx, GCDYN_EINVAL, "present_time must be positive");
    if (pr->response == GCDYN_LAMBDA_TABLE)
        return fail(ctx, GCDYN_EUNSUPPORTED, "gradients need a parametric birth response (sigmoid or constant): a table has K + 2 directions");
    if (pr->n_types > 32) return fail(ctx, GCDYN_EUNSUPPORTED, "gradients are built for n_types <= 32");
    gcdyn_opts o{};
    if (opts) o = *opts;
    const int M = normalise_order(o.order);
    if (M < 0 || M > 16) return fail(ctx, GCDYN_EINVAL, "gradient order must be 0, 8, 12 or 16");
    const double tol = o.tol > 0.0 ? o.tol : 1e-11;
    const int NS = pr->response == GCDYN_LAMBDA_SIGMOID ? 6 : 3;
    if (n_dirs) *n_dirs = NS;
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const int P = pr->n_psets;
    // parameters → device; a plain totals evaluation first: it builds and calibrates the forest's nodal moments, gives the
    // exact value of parameter sets the grid does not resolve, and tells us which those are
    const ParamLayout L = layout_of(pr);
    CU_TRY(ctx, ctx->pin_in.ensure(L.n_doubles * 8));
    pack_host(pr, L, reinterpret_cast<double*>(ctx->pin_in.p));
    CU_TRY(ctx, ctx->params_blob.ensure(L.n_doubles * 8));
    CU_TRY(ctx, cudaMemcpyAsync(ctx->params_blob.p, ctx->pin_in.p, L.n_doubles * 8, cudaMemcpyHostToDevice, st));
    ctx->blob_doubles = 0;                                 // (the blob no longer holds what gcdyn_loglik_batch cached there)
    const ParamsView pv = view_of(pr, L, ctx->params_blob.as<double>());
    CU_TRY(ctx, ctx->grad_out.ensure(((size_t)P * (NS + 2)) * 8 + (size_t)P * 4));
    double* d_tot = ctx->grad_out.as<double>();
    double* d_ll = d_tot + P;
    double* d_grad = d_ll + P;
    int* d_st = reinterpret_cast<int*>(d_grad + (size_t)P * NS);
    ctx->max_cells = 0;
    rc = enqueue_eval(ctx, f, pv, rate_max_of(pr), T, GCDYN_METHOD_ODE, M, 0, tol, GCDYN_FLAG_TOTALS | FLAG_FORCE_MOMENTS,
                      pr->gamma_pset_stride == 0 ? pr->Gamma : nullptr, d_tot, nullptr, nullptr, st, nullptr);
    if (rc) return rc;
    if (!f->mom.Mt || f->mom_T != T)
        return fail(ctx, GCDYN_EUNSUPPORTED, "gradients need the forest's moment matrix (refused for this forest: too large or too few items per tree)");
    gcdyn::GradCfg gc{};
    gc.T = T; gc.tol = tol; gc.D = f->mom.D; gc.NS = NS; gc.compact_tc = pr->gamma_pset_stride == 0 ? 1 : 0;
    gc.any_self_loop = f->any_self_loop ? 1 : 0;
    gc.mv = f->mom;
    gc.tc_total = gc.compact_tc && f->d_tree_const ? f->d_tree_const + f->v.n_trees : nullptr;
    gc.out_ll = d_ll; gc.out_grad = d_grad; gc.out_status = d_st;
    switch (M) {
        case 8: rc = launch_grad_m<8>(ctx, pv, gc, st); break;
        case 12: rc = launch_grad_m<12>(ctx, pv, gc, st); break;
        default: rc = launch_grad_m<16>(ctx, pv, gc, st); break;
    }
    if (rc) return rc;
    std::vector<double> h_tot(P), h_ll(P);
    std::vector<int> h_flag(P);
    CU_TRY(ctx, cudaMemcpyAsync(h_tot.data(), d_tot, (size_t)P * 8, cudaMemcpyDeviceToHost, st));
    CU_TRY(ctx, cudaMemcpyAsync(h_ll.data(), d_ll, (size_t)P * 8, cudaMemcpyDeviceToHost, st));
    CU_TRY(ctx, cudaMemcpyAsync(out_grad, d_grad, (size_t)P * NS * 8, cudaMemcpyDeviceToHost, st));
    CU_TRY(ctx, cudaMemcpyAsync(h_flag.data(), ctx->flags.p, (size_t)P * 4, cudaMemcpyDeviceToHost, st));
    CU_TRY(ctx, cudaStreamSynchronize(st));
    const double nan = std::nan("");
    for (int p = 0; p < P; ++p) {
        if (h_flag[p]) {                                   // summed exactly by the plain path; no gradient on this grid
            out_pset[p] = h_tot[p];
            for (int j = 0; j < NS; ++j) out_grad[(size_t)p * NS + j] = nan;
        } else {
            out_pset[p] = h_ll[p];
            if (!std::isfinite(h_ll[p])) for (int j = 0; j < NS; ++j) out_grad[(size_t)p * NS + j] = nan;
        }
    }
    return GCDYN_OK;
    GUARD_END(ctx)
}

int gcdyn_last_cells(gcdyn_ctx* ctx, int32_t* cells, int32_t n_psets) {
    GUARD_BEGIN
    if (!ctx || !cells || n_psets <= 0) return fail(ctx, GCDYN_EINVAL, "bad argument");
    if (n_psets > ctx->last_P || !ctx->last_pack.S || ctx->last_order <= 0)
        return fail(ctx, GCDYN_EINVAL, "no ODE evaluation with that many psets to report on");
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    CU_TRY(ctx, cudaDeviceSynchronize());
    CU_TRY(ctx, cudaMemcpy(cells, ctx->last_pack.S, (size_t)n_psets * 4, cudaMemcpyDeviceToHost));
    return GCDYN_OK;
    GUARD_END(ctx)
}

int gcdyn_last_cheb_degree(gcdyn_ctx* ctx, int32_t* interp_degree, int32_t* effective_degree, int32_t* n_fallback) {
    GUARD_BEGIN
    if (!ctx) return fail(ctx, GCDYN_EINVAL, "ctx is NULL");
    if (interp_degree) *interp_degree = ctx->last_cheb_D;
    if (effective_degree) *effective_degree = 0;
    if (n_fallback) *n_fallback = 0;
    if ((effective_degree || n_fallback) && ctx->last_cheb_D > 0 && ctx->flags.p && ctx->last_P > 0) {
        std::vector<int> h(ctx->last_P);
        CU_TRY(ctx, cudaSetDevice(ctx->device));
        CU_TRY(ctx, cudaDeviceSynchronize());
        CU_TRY(ctx, cudaMemcpy(h.data(), ctx->flags.p, (size_t)ctx->last_P * 4, cudaMemcpyDeviceToHost));
        if (n_fallback) { int n = 0; for (int p = 0; p < ctx->last_P; ++p) n += h[p] != 0; *n_fallback = n; }
        // the degree the calibration probe found sufficient for its batch
        if (effective_degree && ctx->last_feedback) *effective_degree = ctx->last_feedback[1];   // measured at calibration
    }
    return GCDYN_OK;
    GUARD_END(ctx)
}

// ---- native simulator (host code; no device needed) ------------------------------------------------------
int gcdyn_sim_forest(const gcdyn_params* model, double present_time, int32_t init_type_idx, int64_t n_trees,
                     uint64_t seed, int64_t first_tree_index, int32_t reject_stubs, int32_t min_leaves,
                     int32_t max_rejections, gcdyn_sim** out) {
    gcdyn_ctx* none = nullptr;
    GUARD_BEGIN
    if (!out) return fail(none, GCDYN_EINVAL, "out is NULL");
    *out = nullptr;
    int rc = check_params(none, model);
    if (rc) return rc;
    if (model->n_psets != 1) return fail(none, GCDYN_EINVAL, "the simulator takes exactly one parameter set");
    const int K = model->n_types;
    if (init_type_idx < 0 || init_type_idx >= K) return fail(none, GCDYN_EINVAL, "init_type_idx out of range");
    if (n_trees < 0 || !(present_time > 0.0)) return fail(none, GCDYN_EINVAL, "n_trees must be >= 0 and present_time > 0");
    std::vector<double> lam(K), G((size_t)K * K);
    for (int i = 0; i < K; ++i) {
        if (model->response == 0) lam[i] = model->lambda[i];
        else if (model->response == 1) lam[i] = model->lambda[0];
        else lam[i] = model->yscale[0] * (1.0 / (1.0 + std::exp(-(model->xscale[0] * (model->type_space[i] - model->xshift[0]))))) + model->yshift[0];
        for (int j = 0; j < K; ++j) G[(size_t)i * K + j] = model->Gamma[i + (size_t)j * K];   // column-major in, row-major here
    }
    for (int i = 0; i < K; ++i) {
        const double tot = lam[i] + model->mu[0] + model->delta[0] * -G[(size_t)i * K + i];
        if (!(tot > 0.0) || !(lam[i] >= 0.0) || !(model->mu[0] >= 0.0)) return fail(none, GCDYN_EINVAL, "rates must be non-negative with a positive total");
    }
    const gcdyn_simu::Model m{K, lam.data(), model->mu[0], model->delta[0], model->rho[0], model->sigma[0], G.data()};
    // blocks of trees are simulated in parallel; each tree has its own RNG stream, so the result is independent of threading
    const int64_t block = 64;
    const int64_t n_blocks = (n_trees + block - 1) / block;
    std::vector<gcdyn_simu::FlatOut> parts((size_t)n_blocks);
    int failed = 0;
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t b = 0; b < n_blocks; ++b) {
        std::vector<gcdyn_simu::Node> nodes;
        std::vector<double> wbuf((size_t)K);
        const int64_t lo = b * block, hi = std::min(n_trees, lo + block);
        for (int64_t t = lo; t < hi; ++t) {
            if (!gcdyn_simu::simulate_tree(m, present_time, init_type_idx, seed, (uint64_t)(first_tree_index + t), reject_stubs != 0,
                                          min_leaves, max_rejections, parts[(size_t)b], nodes, wbuf)) {
#pragma omp atomic write
                failed = 1;
                break;
            }
        }
    }
    if (failed) return fail(none, GCDYN_EINVAL, "max_rejections trees rejected");
    gcdyn_sim* sim = new gcdyn_sim();
    gcdyn_simu::FlatOut& o = sim->out;
    for (auto& part : parts) {
        const int64_t nb = (int64_t)o.event.size(), cb = (int64_t)o.child_idx.size();
        for (size_t k = 1; k < part.tree_off.size(); ++k) o.tree_off.push_back(part.tree_off[k] + nb);
        for (size_t k = 1; k < part.child_off.size(); ++k) o.child_off.push_back(part.child_off[k] + cb);
        o.event.insert(o.event.end(), part.event.begin(), part.event.end());
        o.type_idx.insert(o.type_idx.end(), part.type_idx.begin(), part.type_idx.end());
        o.btype_idx.insert(o.btype_idx.end(), part.btype_idx.begin(), part.btype_idx.end());
        o.t_start.insert(o.t_start.end(), part.t_start.begin(), part.t_start.end());
        o.t_end.insert(o.t_end.end(), part.t_end.begin(), part.t_end.end());
        o.parent.insert(o.parent.end(), part.parent.begin(), part.parent.end());
        o.child_idx.insert(o.child_idx.end(), part.child_idx.begin(), part.child_idx.end());
        o.root_type_idx.insert(o.root_type_idx.end(), part.root_type_idx.begin(), part.root_type_idx.end());
        o.root_time.insert(o.root_time.end(), part.root_time.begin(), part.root_time.end());
        o.rejections += part.rejections;
        part = gcdyn_simu::FlatOut();
    }
    *out = sim;
    return GCDYN_OK;
    GUARD_END(none)
}

int gcdyn_sim_sizes(const gcdyn_sim* sim, int64_t* n_trees, int64_t* n_nodes, int64_t* n_child_entries, int64_t* n_rejections) {
    if (!sim) return fail(nullptr, GCDYN_EINVAL, "sim is NULL");
    if (n_trees) *n_trees = (int64_t)sim->out.tree_off.size() - 1;
    if (n_nodes) *n_nodes = (int64_t)sim->out.event.size();
    if (n_child_entries) *n_child_entries = (int64_t)sim->out.child_idx.size();
    if (n_rejections) *n_rejections = sim->out.rejections;
    return GCDYN_OK;
}

int gcdyn_sim_export(const gcdyn_sim* sim, int64_t* tree_off, uint8_t* event, int32_t* type_idx, int32_t* btype_idx,
                     double* t_start, double* t_end, int32_t* parent, int64_t* child_off, int32_t* child_idx,
                     int32_t* root_type_idx, double* root_time) {
    gcdyn_ctx* none = nullptr;
    GUARD_BEGIN
    if (!sim) return fail(none, GCDYN_EINVAL, "sim is NULL");
    const gcdyn_simu::FlatOut& o = sim->out;
    auto cp = [](auto* dst, const auto& v) { if (dst && !v.empty()) std::memcpy(dst, v.data(), v.size() * sizeof(v[0])); };
    cp(tree_off, o.tree_off); cp(event, o.event); cp(type_idx, o.type_idx); cp(btype_idx, o.btype_idx);
    cp(t_start, o.t_start); cp(t_end, o.t_end); cp(parent, o.parent); cp(child_off, o.child_off);
    cp(child_idx, o.child_idx); cp(root_type_idx, o.root_type_idx); cp(root_time, o.root_time);
    return GCDYN_OK;
    GUARD_END(none)
}

void gcdyn_sim_destroy(gcdyn_sim* sim) { delete sim; }

int gcdyn_set_profiling(gcdyn_ctx* ctx, int on) {
    GUARD_BEGIN
    if (!ctx) return fail(ctx, GCDYN_EINVAL, "ctx is NULL");
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    if (on && !ctx->ev[0])
        for (auto& e : ctx->ev) CU_TRY(ctx, cudaEventCreate(&e));
    ctx->profile = on != 0;
    ctx->ev_valid = false;
    return GCDYN_OK;
    GUARD_END(ctx)
}

int gcdyn_last_kernel_ms(gcdyn_ctx* ctx, float* ms) {
    GUARD_BEGIN
    if (!ctx || !ms) return fail(ctx, GCDYN_EINVAL, "NULL argument");
    if (!ctx->ev_valid) return fail(ctx, GCDYN_EINVAL, "no profiled evaluation to report on");
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    CU_TRY(ctx, cudaEventSynchronize(ctx->ev[6]));
    for (int i = 0; i < 6; ++i) CU_TRY(ctx, cudaEventElapsedTime(&ms[i], ctx->ev[i], ctx->ev[i + 1]));
    return GCDYN_OK;
    GUARD_END(ctx)
}

int gcdyn_probe_fp64_peak(gcdyn_ctx* ctx, double* tflops) {
    GUARD_BEGIN
    if (!ctx || !tflops) return fail(ctx, GCDYN_EINVAL, "NULL argument");
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    CU_TRY(ctx, ctx->probe.ensure(64));
    const int iters = 1 << 15;
    const int blocks = ctx->sm_count * 8;
    cudaEvent_t a, b;
    CU_TRY(ctx, cudaEventCreate(&a));
    CU_TRY(ctx, cudaEventCreate(&b));
    double best = 0.0;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(a, ctx->stream);
        dfma_probe_kernel<<<blocks, 256, 0, ctx->stream>>>(ctx->probe.as<double>(), iters, 1.0 + rep);
        cudaEventRecord(b, ctx->stream);
        cudaError_t e = cudaEventSynchronize(b);
        if (e != cudaSuccess) { cudaEventDestroy(a); cudaEventDestroy(b); return fail(ctx, GCDYN_ECUDA, cudaGetErrorString(e)); }
        float ms = 0.f;
        cudaEventElapsedTime(&ms, a, b);
        const double flops = 2.0 * 8.0 * (double)iters * 256.0 * (double)blocks;
        if (rep > 0) best = std::max(best, flops / (ms * 1e-3) / 1e12);
    }
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    *tflops = best;
    return GCDYN_OK;
    GUARD_END(ctx)
}

}  // extern "C"

// ---- peer all-reduce (multi-GPU) ---------------------------------------------------------------------------
namespace {
size_t comm_data_bytes(int world, int cap) { return (size_t)2 * world * cap * sizeof(double); }
// byte offset of the 16-byte entries of the fused form: behind data | flags | call counter | CTA counter, 16-byte aligned
size_t comm_ll_offset(int world, int cap) { return (comm_data_bytes(world, cap) + ((size_t)world + 2) * 8 + 15) & ~(size_t)15; }
}

extern "C" int gcdyn_comm_create(gcdyn_ctx* ctx, int32_t rank, int32_t world, int32_t max_psets, void* handle_out,
                                 gcdyn_comm** out) {
    GUARD_BEGIN
    if (!ctx || !out || !handle_out) return fail(ctx, GCDYN_EINVAL, "NULL argument");
    *out = nullptr;
    if (world < 1 || world > gcdyn::PEER_MAX_WORLD || rank < 0 || rank >= world || max_psets < 1)
        return fail(ctx, GCDYN_EINVAL, "bad rank/world/max_psets (world <= 16)");
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    gcdyn_comm* c = new gcdyn_comm();
    c->ctx = ctx; c->rank = rank; c->world = world; c->cap = (max_psets + 1) & ~1;
    if (const char* ev = std::getenv("GCDYN_PEER_TIMEOUT_MS")) {
        const double ms = std::atof(ev);
        if (ms > 0.0) c->timeout_ns = (unsigned long long)(ms * 1e6);
    }
    if (cudaHostAlloc(&c->h_err, sizeof(int), cudaHostAllocMapped) != cudaSuccess) { cudaGetLastError(); c->h_err = nullptr; }
    else *c->h_err = 0;
    // data | flags | call counter | fused-kernel CTA counter | self-validating entries of the fused form (16 B each)
    const size_t bytes = comm_ll_offset(world, c->cap) + comm_data_bytes(world, c->cap) * 2;
    cudaError_t e = cudaMalloc(&c->window, bytes);
    if (e == cudaSuccess) e = cudaMemset(c->window, 0, bytes);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    cudaIpcMemHandle_t h{};
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, c->window);
    if (e != cudaSuccess) {
        if (c->window) cudaFree(c->window);
        delete c;
        return fail(ctx, GCDYN_ECUDA, cudaGetErrorString(e));
    }
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    std::memcpy(handle_out, &h, 64);
    c->peer[rank] = c->window;
    *out = c;
    return GCDYN_OK;
    GUARD_END(ctx)
}

extern "C" int gcdyn_comm_connect(gcdyn_comm* c, const void* handles) {
    gcdyn_ctx* ctx = c ? c->ctx : nullptr;
    GUARD_BEGIN
    if (!c || !handles) return fail(ctx, GCDYN_EINVAL, "NULL argument");
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    for (int r = 0; r < c->world; ++r) {
        if (r == c->rank || c->peer[r]) continue;
        cudaIpcMemHandle_t h;
        std::memcpy(&h, static_cast<const char*>(handles) + (size_t)r * 64, 64);
        CU_TRY(ctx, cudaIpcOpenMemHandle(&c->peer[r], h, cudaIpcMemLazyEnablePeerAccess));
    }
    c->connected = true;
    return GCDYN_OK;
    GUARD_END(ctx)
}

namespace {
int comm_enqueue(gcdyn_comm* c, const double* d_in, double* out, int n_psets, cudaStream_t st) {
    gcdyn_ctx* ctx = c->ctx;
    if (!c->connected && c->world > 1) return fail(ctx, GCDYN_EINVAL, "gcdyn_comm_connect has not been called");
    if (n_psets < 1 || n_psets > c->cap) return fail(ctx, GCDYN_EINVAL, "n_psets exceeds the window capacity");
    gcdyn::PeerView pv{};
    pv.rank = c->rank; pv.world = c->world; pv.cap = c->cap;
    for (int r = 0; r < c->world; ++r) {
        pv.win[r].data = static_cast<double*>(c->peer[r]);
        pv.win[r].flag = reinterpret_cast<unsigned long long*>(static_cast<char*>(c->peer[r]) + comm_data_bytes(c->world, c->cap));
    }
    pv.seq_dev = pv.win[c->rank].flag + c->world;
    cudaLaunchConfig_t lc = pdl_config(dim3(1), dim3(1024), 0, st);
    CU_TRY(ctx, cudaLaunchKernelEx(&lc, gcdyn::peer_allreduce_kernel, pv, d_in, out, n_psets, c->timeout_ns, c->h_err));
    return GCDYN_OK;
}
}  // namespace

namespace {
int comm_fused_view(gcdyn_comm* c, int n_psets, gcdyn::FusedAllReduce* out) {
    gcdyn_ctx* ctx = c->ctx;
    if (!c->connected && c->world > 1) return fail(ctx, GCDYN_EINVAL, "gcdyn_comm_connect has not been called");
    if (n_psets < 1 || n_psets > c->cap) return fail(ctx, GCDYN_EINVAL, "n_psets exceeds the window capacity");
    gcdyn::FusedAllReduce fc{};
    fc.pv.rank = c->rank; fc.pv.world = c->world; fc.pv.cap = c->cap;
    for (int r = 0; r < c->world; ++r) {
        fc.pv.win[r].data = static_cast<double*>(c->peer[r]);
        fc.pv.win[r].flag = reinterpret_cast<unsigned long long*>(static_cast<char*>(c->peer[r]) + comm_data_bytes(c->world, c->cap));
        fc.pv.win[r].ll = reinterpret_cast<uint4*>(static_cast<char*>(c->peer[r]) + comm_ll_offset(c->world, c->cap));
    }
    fc.pv.seq_dev = fc.pv.win[c->rank].flag + c->world;
    fc.done = reinterpret_cast<int*>(fc.pv.win[c->rank].flag + c->world + 1);
    fc.timeout_ns = c->timeout_ns;
    fc.err = c->h_err;
    fc.totals_in = nullptr;
    *out = fc;
    return GCDYN_OK;
}
int comm_row_exchange(gcdyn_comm* c, unsigned long long seq, int P, long long Jst, long long Juse, int p0, int p1, int D,
                      const PsetPack& pk, int* need_sweep, cudaStream_t st) {
    gcdyn_ctx* ctx = c->ctx;
    gcdyn::RowExchange rx{};
    const size_t flag_off = 2 * c->rows_half;
    for (int r = 0; r < c->world; ++r) {
        char* w = static_cast<char*>(c->rows_peer[r]);
        rx.rows[r] = reinterpret_cast<double*>(w + (size_t)(seq & 1ULL) * c->rows_half);
        rx.flag[r] = reinterpret_cast<unsigned long long*>(w + flag_off);
    }
    rx.done = reinterpret_cast<int*>(static_cast<char*>(c->rows) + flag_off + (size_t)c->world * 8);
    rx.rank = c->rank; rx.world = c->world; rx.seq = seq; rx.timeout_ns = c->timeout_ns; rx.err = c->h_err;
    // enough CTAs to keep the NVLink stores of a several-MB slice in flight; one is plenty for a refusal (P == 0)
    const long long doubles = (long long)(p1 - p0) * Juse;
    const int grid = (int)std::max<long long>(1, std::min<long long>(2LL * ctx->sm_count, (doubles / 2 + 255) / 256));
    gcdyn::row_exchange_kernel<<<grid, 256, 0, st>>>(rx, P, Jst, Juse, p0, p1, D, pk.pstatus, need_sweep, pk.S, pk.perr);
    CU_TRY(ctx, cudaGetLastError());
    return GCDYN_OK;
}

void comm_poison(gcdyn_comm* c, int n_psets, cudaStream_t st) {
    gcdyn::FusedAllReduce fc{};
    if (c->world <= 1 || n_psets < 1 || n_psets > c->cap || comm_fused_view(c, n_psets, &fc) != GCDYN_OK) return;
    gcdyn::peer_poison_kernel<<<1, 256, 0, st>>>(fc.pv, n_psets);
    cudaGetLastError();
}
}  // namespace

extern "C" int gcdyn_comm_rows_create(gcdyn_comm* c, int64_t bytes, void* handle_out) {
    gcdyn_ctx* ctx = c ? c->ctx : nullptr;
    GUARD_BEGIN
    if (!c || !handle_out || bytes < 2) return fail(ctx, GCDYN_EINVAL, "bad argument");
    if (c->rows) return fail(ctx, GCDYN_EINVAL, "this communicator already has a row window");
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    const size_t half = (((size_t)bytes + 1) / 2 + 255) & ~(size_t)255;
    const size_t total = 2 * half + ((size_t)c->world + 2) * 8;
    cudaError_t e = cudaMalloc(&c->rows, total);
    if (e == cudaSuccess) e = cudaMemset(static_cast<char*>(c->rows) + 2 * half, 0, total - 2 * half);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    cudaIpcMemHandle_t h{};
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, c->rows);
    if (e != cudaSuccess) {
        if (c->rows) cudaFree(c->rows);
        c->rows = nullptr;
        return fail(ctx, e == cudaErrorMemoryAllocation ? GCDYN_ENOMEM : GCDYN_ECUDA, cudaGetErrorString(e));
    }
    std::memcpy(handle_out, &h, 64);
    c->rows_half = half;
    c->rows_peer[c->rank] = c->rows;
    if (const char* ev = std::getenv("GCDYN_TRAJ_SHARD_MIN_PSETS")) c->shard_min_psets = std::max(1, std::atoi(ev));
    if (const char* ev = std::getenv("GCDYN_ROW_PUSH")) c->push_from_traj = std::strcmp(ev, "kernel") != 0;
    if (c->world == 1) c->rows_connected = true;
    return GCDYN_OK;
    GUARD_END(ctx)
}

extern "C" int gcdyn_comm_rows_connect(gcdyn_comm* c, const void* handles) {
    gcdyn_ctx* ctx = c ? c->ctx : nullptr;
    GUARD_BEGIN
    if (!c || !handles) return fail(ctx, GCDYN_EINVAL, "NULL argument");
    if (!c->rows) return fail(ctx, GCDYN_EINVAL, "gcdyn_comm_rows_create has not been called");
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    for (int r = 0; r < c->world; ++r) {
        if (r == c->rank || c->rows_peer[r]) continue;
        cudaIpcMemHandle_t h;
        std::memcpy(&h, static_cast<const char*>(handles) + (size_t)r * 64, 64);
        CU_TRY(ctx, cudaIpcOpenMemHandle(&c->rows_peer[r], h, cudaIpcMemLazyEnablePeerAccess));
    }
    c->rows_connected = true;
    return GCDYN_OK;
    GUARD_END(ctx)
}

extern "C" int gcdyn_comm_error(gcdyn_comm* c) {
    if (!c || !c->h_err) return 0;
    const int e = *c->h_err;
    *c->h_err = 0;
    return e;
}

extern "C" int gcdyn_last_traj_sharded(gcdyn_ctx* ctx) { return ctx && ctx->last_traj_sharded ? 1 : 0; }

extern "C" int gcdyn_comm_allreduce(gcdyn_comm* c, double* d_inout, int32_t n_psets, void* stream) {
    gcdyn_ctx* ctx = c ? c->ctx : nullptr;
    GUARD_BEGIN
    if (!c || !d_inout) return fail(ctx, GCDYN_EINVAL, "NULL argument");
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = stream ? static_cast<cudaStream_t>(stream) : ctx->stream;
    return comm_enqueue(c, d_inout, d_inout, n_psets, st);
    GUARD_END(ctx)
}

extern "C" int gcdyn_loglik_batch_sharded(gcdyn_ctx* ctx, const gcdyn_forest* forest, const gcdyn_params* params,
                                          double present_time, int method, const gcdyn_opts* opts, gcdyn_comm* comm,
                                          double* out_pset, double* out_tree_pset, int32_t* status) {
    GUARD_BEGIN
    if (!comm) return fail(ctx, GCDYN_EINVAL, "comm is NULL");
    if (comm->ctx != ctx) return fail(ctx, GCDYN_EINVAL, "comm belongs to another context");
    return batch_impl(ctx, forest, params, present_time, method, opts, out_pset, out_tree_pset, status, comm);
    GUARD_END(ctx)
}

extern "C" void gcdyn_comm_destroy(gcdyn_comm* c) {
    if (!c) return;
    int prev = 0;
    cudaGetDevice(&prev);
    cudaSetDevice(c->ctx->device);
    cudaDeviceSynchronize();
    for (int r = 0; r < c->world; ++r)
        if (r != c->rank && c->peer[r]) cudaIpcCloseMemHandle(c->peer[r]);
    for (int r = 0; r < c->world; ++r)
        if (r != c->rank && c->rows_peer[r]) cudaIpcCloseMemHandle(c->rows_peer[r]);
    if (c->rows) cudaFree(c->rows);
    if (c->window) cudaFree(c->window);
    if (c->h_err) cudaFreeHost(c->h_err);
    cudaSetDevice(prev);
    delete c;
}

// ---- device-resident Metropolis ----------------------------------------------------------------------------
struct gcdyn_mh {
    gcdyn_ctx* ctx = nullptr;
    const gcdyn_forest* forest = nullptr;
    gcdyn_comm* comm = nullptr;
    int C = 0, K = 0;
    double T = 0.0, rate_bound = 0.0, rate_typical = 0.0;
    void* blob = nullptr;              // parameters in the likelihood's layout (SIGMOID response) + sampler state
    ParamsView pv{};
    gcdyn::MhView v{};
    std::vector<double> host_gamma;
    int64_t iteration = 0;
    bool have_ll = false;
    void* hist = nullptr;              // device history buffer (grow-only)
    size_t hist_cap = 0;
    // one Metropolis iteration (propose → trajectory → finish → [all-reduce] → accept → tick) captured as a CUDA graph
    cudaGraphExec_t graph = nullptr, graph_many = nullptr;   // one iteration / MH_UNROLL iterations
    int graph_D = 0;                   // nodal grid degree the captured iteration was built for
    int64_t uncaptured_iters = 0;      // iterations run launch by launch (diagnostics)
};

extern "C" int gcdyn_mh_create(gcdyn_ctx* ctx, const gcdyn_forest* f, const gcdyn_mh_desc* d, double present_time,
                               gcdyn_comm* comm, gcdyn_mh** out) {
    GUARD_BEGIN
    if (!ctx || !f || !d || !out) return fail(ctx, GCDYN_EINVAL, "NULL argument");
    *out = nullptr;
    if (d->n_chains < 1 || d->n_chains > 65535) return fail(ctx, GCDYN_EINVAL, "n_chains out of range");
    if (d->n_types != f->v.K) return fail(ctx, GCDYN_EINVAL, "desc and forest disagree on n_types");
    if (!d->type_space || !d->Gamma || !d->theta0 || !d->lower || !d->upper) return fail(ctx, GCDYN_EINVAL, "NULL array in desc");
    if (!(d->rho >= 0.0 && d->rho <= 1.0) || !(d->sigma >= 0.0 && d->sigma <= 1.0)) return fail(ctx, GCDYN_EINVAL, "rho, sigma must be between 0 and 1");
    if (!(present_time > 0.0) || !(d->step > 0.0)) return fail(ctx, GCDYN_EINVAL, "present_time and step must be positive");
    for (int j = 2; j < gcdyn::MH_DIM; ++j)
        if (!std::isfinite(d->upper[j])) return fail(ctx, GCDYN_EINVAL, "upper[2..5] must be finite: they bound the rates");
    if (comm && comm->cap < d->n_chains) return fail(ctx, GCDYN_EINVAL, "comm window smaller than n_chains");
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    const int C = d->n_chains, K = d->n_types;
    gcdyn_mh* m = new gcdyn_mh();
    m->ctx = ctx; m->forest = f; m->comm = comm; m->C = C; m->K = K; m->T = present_time;
    m->host_gamma.assign(d->Gamma, d->Gamma + (size_t)K * K);
    // rates are bounded on the prior box: lambda <= yscale + yshift, |gamma_i| <= delta * max|Gamma_ii|
    double gmax = 0.0;
    for (int i = 0; i < K; ++i) gmax = std::max(gmax, std::fabs(d->Gamma[i + (size_t)i * K]));
    const auto up = [&](int j) { return std::exp(std::min(d->upper[j], 50.0)); };
    m->rate_bound = up(2) + up(3) + up(4) + 2.0 * up(5) * gmax;
    // the interpolation degree follows the rates at the starting points (with a margin), not the worst corner of the box
    for (int c = 0; c < C; ++c) {
        const double* t = d->theta0 + (size_t)c * gcdyn::MH_DIM;
        const double r = std::exp(t[2]) + std::exp(t[3]) + std::exp(t[4]) + 2.0 * std::exp(t[5]) * gmax;
        if (std::isfinite(r)) m->rate_typical = std::max(m->rate_typical, 1.5 * r);
    }
    // blob: xscale xshift yscale yshift mu delta rho sigma [C each] | type_space [K] | Gamma [K*K] | theta prop [6C each]
    //       | ll ll_prop [C each] | accepted [C i64] | inside [C i32]
    const size_t nd = (size_t)8 * C + K + (size_t)K * K + (size_t)2 * gcdyn::MH_DIM * C + 2 * (size_t)C + C + (C + 1) / 2 + 8 + 8 + 2;
    cudaError_t e = cudaMalloc(&m->blob, nd * 8);
    if (e != cudaSuccess) { delete m; return fail(ctx, GCDYN_ECUDA, cudaGetErrorString(e)); }
    std::vector<double> h(nd, 0.0);
    double* base = static_cast<double*>(m->blob);
    size_t o = 0;
    auto take = [&](size_t n) { size_t r = o; o += n; return r; };
    const size_t o_xs = take(C), o_xsh = take(C), o_ys = take(C), o_ysh = take(C), o_mu = take(C), o_de = take(C);
    const size_t o_rho = take(C), o_sig = take(C), o_ts = take(K), o_G = take((size_t)K * K);
    const size_t o_th = take((size_t)gcdyn::MH_DIM * C), o_pr = take((size_t)gcdyn::MH_DIM * C);
    const size_t o_ll = take(C), o_llp = take(C), o_acc = take(C), o_in = take((C + 1) / 2 + 1);
    const size_t o_it = take(1), o_hist = take(4), o_tick = take(1);
    for (int c = 0; c < C; ++c) { h[o_rho + c] = d->rho; h[o_sig + c] = d->sigma; }
    std::memcpy(&h[o_ts], d->type_space, (size_t)K * 8);
    std::memcpy(&h[o_G], d->Gamma, (size_t)K * K * 8);
    std::memcpy(&h[o_th], d->theta0, (size_t)gcdyn::MH_DIM * C * 8);
    e = cudaMemcpy(m->blob, h.data(), nd * 8, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { cudaFree(m->blob); delete m; return fail(ctx, GCDYN_ECUDA, cudaGetErrorString(e)); }
    ParamsView& pv = m->pv;
    pv.P = C; pv.K = K; pv.response = GCDYN_LAMBDA_SIGMOID; pv.gamma_stride = 0; pv.lambda = nullptr;
    pv.xscale = base + o_xs; pv.xshift = base + o_xsh; pv.yscale = base + o_ys; pv.yshift = base + o_ysh;
    pv.type_space = base + o_ts; pv.mu = base + o_mu; pv.delta = base + o_de; pv.rho = base + o_rho; pv.sigma = base + o_sig;
    pv.Gamma = base + o_G;
    gcdyn::MhView& v = m->v;
    v.C = C; v.K = K;
    v.theta = base + o_th; v.prop = base + o_pr; v.ll = base + o_ll; v.ll_prop = base + o_llp;
    v.accepted = reinterpret_cast<long long*>(base + o_acc); v.inside = reinterpret_cast<int*>(base + o_in);
    v.xscale = base + o_xs; v.xshift = base + o_xsh; v.yscale = base + o_ys; v.yshift = base + o_ysh;
    v.mu = base + o_mu; v.delta = base + o_de;
    for (int j = 0; j < gcdyn::MH_DIM; ++j) { v.lower[j] = d->lower[j]; v.upper[j] = d->upper[j]; }
    v.step = d->step; v.seed = d->seed;
    v.it_dev = reinterpret_cast<unsigned long long*>(base + o_it);
    v.hist = reinterpret_cast<gcdyn::MhView::Hist*>(base + o_hist);
    v.tick_count = reinterpret_cast<int*>(base + o_tick);
    *out = m;
    return GCDYN_OK;
    GUARD_END(ctx)
}

namespace {
// one batched likelihood of the C parameter sets currently in the blob → dst[C] (all-reduced when sharded)
// `hook`: fuse this iteration's proposal, accept/reject and counter into the likelihood's kernels; *fused tells whether that
// happened (it cannot when the forest takes the item sweep)
int mh_loglik(gcdyn_mh* m, double* dst, cudaStream_t st, bool hook = false, bool* fused = nullptr) {
    gcdyn_ctx* ctx = m->ctx;
    ctx->max_cells = 0;
    int rc = enqueue_eval(ctx, m->forest, m->pv, m->rate_bound, m->T, GCDYN_METHOD_ODE, 16, 0, 1e-11,
                          GCDYN_FLAG_TOTALS | FLAG_FORCE_MOMENTS, m->host_gamma.data(),
                          dst, nullptr, nullptr, st, nullptr, m->rate_typical, m->comm, hook ? &m->v : nullptr, fused);
    if (rc && m->comm) {              // release the peers with NaN sums (see batch_impl), then report the local error
        const std::string why = ctx->err;
        cudaGetLastError();
        comm_poison(m->comm, m->C, st);
        return fail(ctx, rc, why);
    }
    return rc;
}
}  // namespace

extern "C" int gcdyn_mh_run(gcdyn_mh* m, int64_t iterations, double* theta_hist, double* loglik_hist) {
    gcdyn_ctx* ctx = m ? m->ctx : nullptr;
    GUARD_BEGIN
    if (!m || iterations < 0) return fail(ctx, GCDYN_EINVAL, "bad argument");
    if ((theta_hist == nullptr) != (loglik_hist == nullptr)) return fail(ctx, GCDYN_EINVAL, "give both history buffers or neither");
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const int C = m->C, threads = 128, blocks = (C + threads - 1) / threads;
    double *d_ht = nullptr, *d_hl = nullptr;
    if (theta_hist && iterations > 0) {
        const size_t need = (size_t)iterations * C * (gcdyn::MH_DIM + 1) * 8;
        if (need > m->hist_cap) {
            if (m->hist) cudaFree(m->hist);
            m->hist = nullptr; m->hist_cap = 0;
            CU_TRY(ctx, cudaMalloc(&m->hist, need));
            m->hist_cap = need;
        }
        d_ht = static_cast<double*>(m->hist);
        d_hl = d_ht + (size_t)iterations * C * gcdyn::MH_DIM;
    }
    // where this run's history rows go (read by mh_accept_kernel through device memory, so the captured graph is
    // the same for every run) and the iteration the device counter stands at
    {
        gcdyn::MhView::Hist h{d_ht, d_hl, (long long)m->iteration, (long long)iterations};
        const unsigned long long it0 = (unsigned long long)m->iteration;
        CU_TRY(ctx, cudaMemcpyAsync(m->v.hist, &h, sizeof(h), cudaMemcpyHostToDevice, st));
        CU_TRY(ctx, cudaMemcpyAsync(m->v.it_dev, &it0, sizeof(it0), cudaMemcpyHostToDevice, st));
        CU_TRY(ctx, cudaStreamSynchronize(st));            // h, it0 are stack variables
    }
    if (!m->have_ll) {
        gcdyn::mh_state_params_kernel<<<blocks, threads, 0, st>>>(m->v);
        CU_TRY(ctx, cudaGetLastError());
        int rc = mh_loglik(m, m->v.ll, st);                 // also calibrates the forest's nodal grid (synchronises once)
        if (rc) return rc;
        m->have_ll = true;
    }
    const gcdyn_forest* f = m->forest;
    auto enqueue_iteration = [&]() -> int {
        // fused form: the likelihood's own kernels draw the proposals, decide accept/reject and bump the counter — an
        // iteration is traj_kernel + finish_kernel.  Forests that take the item sweep use the stand-alone kernels.
        if (!f->mom_refused && f->v.n_trees > 0 && std::getenv("GCDYN_MH_UNFUSED") == nullptr) {
            bool fused = false;
            const int rc = mh_loglik(m, m->v.ll_prop, st, true, &fused);
            if (rc) return rc;
            if (fused) return GCDYN_OK;
            return fail(ctx, GCDYN_ECUDA, "internal: the fused Metropolis iteration was not taken");
        }
        gcdyn::mh_propose_kernel<<<blocks, threads, 0, st>>>(m->v);
        CU_TRY(ctx, cudaGetLastError());
        int rc = mh_loglik(m, m->v.ll_prop, st);
        if (rc) return rc;
        gcdyn::mh_accept_kernel<<<blocks, threads, 0, st>>>(m->v);
        CU_TRY(ctx, cudaGetLastError());
        gcdyn::mh_tick_kernel<<<1, 1, 0, st>>>(m->v);
        CU_TRY(ctx, cudaGetLastError());
        return GCDYN_OK;
    };
    const bool use_graph = std::getenv("GCDYN_MH_NO_GRAPH") == nullptr;
    // Iterations are replayed from CUDA graphs: one holding MH_UNROLL consecutive iterations (the kernels take the iteration
    // number, the history row and the all-reduce sequence from device memory, so every replay is the same graph) and one
    // holding a single iteration for the remainder.  A graph is valid as long as the forest's grid degree stands; a pending
    // widening request (a quarter of a batch unresolved) is served by launch-by-launch iterations, then re-captured.
    constexpr int MH_UNROLL = 16;
    auto capture = [&](int n, cudaGraphExec_t* out) -> int {
        cudaGraph_t g = nullptr;
        CU_TRY(ctx, cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
        int rc = GCDYN_OK;
        for (int k = 0; k < n && !rc; ++k) rc = enqueue_iteration();
        const cudaError_t ce = cudaStreamEndCapture(st, &g);
        if (rc || ce != cudaSuccess || !g) {
            if (g) cudaGraphDestroy(g);
            cudaGetLastError();
            return rc ? rc : fail(ctx, GCDYN_ECUDA, std::string("graph capture of a Metropolis iteration failed: ") + cudaGetErrorString(ce));
        }
        const cudaError_t ie = cudaGraphInstantiate(out, g, 0);
        cudaGraphDestroy(g);
        if (ie != cudaSuccess) { *out = nullptr; return fail(ctx, GCDYN_ECUDA, std::string("cudaGraphInstantiate: ") + cudaGetErrorString(ie)); }
        return GCDYN_OK;
    };
    for (int64_t i = 0; i < iterations;) {
        const bool widen_pending = f->h_feedback && f->D_cur > 0 && f->h_feedback[0] == f->D_cur && f->D_cur < 128;
        if (m->graph && (m->graph_D != f->D_cur || widen_pending || !use_graph)) {
            cudaGraphExecDestroy(m->graph);
            if (m->graph_many) cudaGraphExecDestroy(m->graph_many);
            m->graph = m->graph_many = nullptr;
        }
        if (use_graph && !m->graph && !widen_pending && m->uncaptured_iters > 0) {
            // every workspace already has its size (an uncaptured iteration ran before)
            int rc = capture(1, &m->graph);
            if (!rc) rc = capture(MH_UNROLL, &m->graph_many);
            if (rc) return rc;
            m->graph_D = f->D_cur;
        }
        if (use_graph && m->graph) {
            if (iterations - i >= MH_UNROLL) { CU_TRY(ctx, cudaGraphLaunch(m->graph_many, st)); i += MH_UNROLL; m->iteration += MH_UNROLL; }
            else { CU_TRY(ctx, cudaGraphLaunch(m->graph, st)); ++i; ++m->iteration; }
        } else {
            const int rc = enqueue_iteration();
            if (rc) return rc;
            ++m->uncaptured_iters;
            ++i;
            ++m->iteration;
        }
    }
    if (d_ht) {
        CU_TRY(ctx, cudaMemcpyAsync(theta_hist, d_ht, (size_t)iterations * C * gcdyn::MH_DIM * 8, cudaMemcpyDeviceToHost, st));
        CU_TRY(ctx, cudaMemcpyAsync(loglik_hist, d_hl, (size_t)iterations * C * 8, cudaMemcpyDeviceToHost, st));
    }
    CU_TRY(ctx, cudaStreamSynchronize(st));
    if (m->comm && m->comm->h_err && *m->comm->h_err) {
        *m->comm->h_err = 0;
        return fail(ctx, GCDYN_ECUDA, "peer all-reduce timed out: a rank did not arrive (its evaluation failed or it is gone)");
    }
    return GCDYN_OK;
    GUARD_END(ctx)
}

extern "C" int gcdyn_mh_state(gcdyn_mh* m, double* theta, double* loglik, int64_t* accepted, int64_t* iteration) {
    gcdyn_ctx* ctx = m ? m->ctx : nullptr;
    GUARD_BEGIN
    if (!m) return fail(ctx, GCDYN_EINVAL, "mh is NULL");
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    if (theta) CU_TRY(ctx, cudaMemcpy(theta, m->v.theta, (size_t)m->C * gcdyn::MH_DIM * 8, cudaMemcpyDeviceToHost));
    if (loglik) CU_TRY(ctx, cudaMemcpy(loglik, m->v.ll, (size_t)m->C * 8, cudaMemcpyDeviceToHost));
    if (accepted) CU_TRY(ctx, cudaMemcpy(accepted, m->v.accepted, (size_t)m->C * 8, cudaMemcpyDeviceToHost));
    if (iteration) *iteration = m->iteration;
    return GCDYN_OK;
    GUARD_END(ctx)
}

extern "C" int gcdyn_mh_restore(gcdyn_mh* m, const double* theta, const double* loglik, const int64_t* accepted, int64_t iteration) {
    gcdyn_ctx* ctx = m ? m->ctx : nullptr;
    GUARD_BEGIN
    if (!m || !theta || !loglik || !accepted || iteration < 0) return fail(ctx, GCDYN_EINVAL, "bad argument");
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    CU_TRY(ctx, cudaMemcpy(m->v.theta, theta, (size_t)m->C * gcdyn::MH_DIM * 8, cudaMemcpyHostToDevice));
    CU_TRY(ctx, cudaMemcpy(m->v.ll, loglik, (size_t)m->C * 8, cudaMemcpyHostToDevice));
    CU_TRY(ctx, cudaMemcpy(m->v.accepted, accepted, (size_t)m->C * 8, cudaMemcpyHostToDevice));
    m->iteration = iteration;
    m->have_ll = true;
    return GCDYN_OK;
    GUARD_END(ctx)
}

extern "C" double gcdyn_mh_uniform(uint64_t seed, uint64_t iteration, uint64_t chain, uint64_t slot) {
    return gcdyn::mh_uniform(seed, iteration, chain, slot);
}

extern "C" void gcdyn_mh_destroy(gcdyn_mh* m) {
    if (!m) return;
    int prev = 0;
    cudaGetDevice(&prev);
    cudaSetDevice(m->ctx->device);
    cudaStreamSynchronize(m->ctx->stream);
    if (m->graph) cudaGraphExecDestroy(m->graph);
    if (m->graph_many) cudaGraphExecDestroy(m->graph_many);
    if (m->blob) cudaFree(m->blob);
    if (m->hist) cudaFree(m->hist);
    cudaSetDevice(prev);
    delete m;
}

#ifdef GCDYN_TRAJ_TRACE
extern "C" int gcdyn_trace_dump(long long* out) {
    return (int)cudaMemcpyFromSymbol(out, gcdyn::g_trace_out, sizeof(long long) * 1024 * 24);
}
#endif
