// Peer all-reduce of the per-pset sums over NVLink peer memory (the multi-GPU form of the plain
// `sum` over trees at src/multitype_branching_process.jl:258 when the trees are sharded over GPUs).
//
// Every rank owns a window in its own HBM, mapped into every peer through CUDA IPC:
//     data[2 parities][world slots][cap doubles]   flag[world] (u64 sequence numbers)
// One kernel per evaluation and rank: (1) store this rank's [P] sums into slot `rank` of EVERY peer's window
// (P2P stores over NVLink/NVSwitch), fence, then publish the call's sequence number in every peer's flag[rank];
// (2) wait until all `world` flags of the local window carry that sequence number; (3) sum the slots in rank
// order.  Every rank adds the same numbers in the same order, so all ranks hold bit-identical results.
// The parity alternates per call: a rank can be at most one call ahead of its slowest peer (it needs that
// peer's flag for the current call before it can start the next), so the slots of call n+2 never overwrite
// slots of call n that are still being read.
// The wait is BOUNDED: a peer that never arrives (it failed before enqueueing, or died) must not leave this GPU spinning
// inside a stream synchronisation.  After `timeout_ns` of %globaltimer the kernel gives up, writes NaN to every output
// and raises `*err` (pinned host memory; the host-buffer entry points turn it into an error code).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace gcdyn {

struct PeerWindow {
    double* data;                  // [2][world][cap]
    unsigned long long* flag;      // [world]
    uint4* ll;                     // [2][world][cap] self-validating entries {lo32, seq, hi32, seq} of the fused form
};

constexpr int PEER_MAX_WORLD = 16;

struct PeerView {
    PeerWindow win[PEER_MAX_WORLD];    // win[r] = rank r's window as mapped in this process (win[rank] is local)
    int rank, world, cap;
    unsigned long long* seq_dev;       // this rank's call counter, in device memory: the kernel takes the next sequence
                                       // number from it, so a captured CUDA graph can replay the all-reduce unchanged
};

// The same all-reduce fused into the kernel that produces the sums (finish_kernel, cheb_gemm.cuh): world == 1 = off.
// Its entries are SELF-VALIDATING (the low-latency protocol of collective libraries): a double travels as two aligned
// 8-byte words {low half, seq} and {high half, seq}; an 8-byte store is single-copy atomic, so a reader that finds the
// call's sequence number in both words has the value — no fence, no separate flag, no ordering between stores is needed,
// and the cost of the collective is one NVLink store latency.  Parities alternate per call as above.
struct FusedAllReduce {
    PeerView pv;
    unsigned long long timeout_ns;
    int* err;                          // pinned host word raised on a timeout
    int* done;                         // device counter of the CTAs that have stored their sum (reset by the last one)
    const double* totals_in;           // GCDYN_FLAG_TOTALS: this rank's sums as traj_kernel wrote them
};

__device__ __forceinline__ void ll_store(uint4* slot, double v, unsigned seq) {
    const unsigned long long bits = (unsigned long long)__double_as_longlong(v);
    const uint2 w0 = make_uint2((unsigned)bits, seq), w1 = make_uint2((unsigned)(bits >> 32), seq);
    asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(slot), "r"(w0.x), "r"(w0.y) : "memory");
    asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(reinterpret_cast<char*>(slot) + 8), "r"(w1.x), "r"(w1.y) : "memory");
}
__device__ __forceinline__ bool ll_load(const uint4* slot, unsigned seq, double& v) {
    uint4 w;
    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(w.x), "=r"(w.y), "=r"(w.z), "=r"(w.w) : "l"(slot) : "memory");
    if (w.y != seq || w.w != seq) return false;
    v = __longlong_as_double((long long)(((unsigned long long)w.z << 32) | (unsigned long long)w.x));
    return true;
}

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// One CTA.  `in` = this rank's [P] partial sums, `out` = the all-reduced sums (may alias `in`; may be pinned host memory).
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

__global__ void __launch_bounds__(1024) peer_allreduce_kernel(PeerView pv, const double* in, double* out, int P,
                                                              unsigned long long timeout_ns, int* err) {
    __shared__ int s_timeout;
    if (threadIdx.x == 0) s_timeout = 0;
    asm volatile("griddepcontrol.wait;" ::: "memory");      // launched behind finish_kernel with programmatic serialization
    const unsigned long long seq = *pv.seq_dev + 1;         // every rank makes the same number of calls
    const int par = (int)(seq & 1ULL);
    const size_t slot = ((size_t)par * pv.world + pv.rank) * pv.cap;
    for (int i = threadIdx.x; i < P * pv.world; i += blockDim.x) {
        const int r = i / P, p = i - r * P;
        pv.win[r].data[slot + p] = in[p];
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x < pv.world) st_release_sys(pv.win[threadIdx.x].flag + pv.rank, seq);
    if (threadIdx.x < pv.world) {
        const unsigned long long* f = pv.win[pv.rank].flag + threadIdx.x;
        const unsigned long long t0 = global_ns();
        while (ld_acquire_sys(f) < seq) {
            __nanosleep(100);
            if (global_ns() - t0 > timeout_ns) { atomicExch(&s_timeout, 1); break; }
        }
    }
    __syncthreads();
    if (s_timeout) {                                        // CTA-uniform after the barrier
        for (int p = threadIdx.x; p < P; p += blockDim.x) out[p] = __longlong_as_double(0x7ff8000000000000LL);
        if (threadIdx.x == 0 && err) *err = 1;
        if (threadIdx.x == 0) *pv.seq_dev = seq;
        return;
    }
    const double* local = pv.win[pv.rank].data + (size_t)par * pv.world * pv.cap;
    for (int p = threadIdx.x; p < P; p += blockDim.x) {
        double s = 0.0;
        for (int r = 0; r < pv.world; ++r) s += __ldcg(local + (size_t)r * pv.cap + p);
        out[p] = s;
    }
    __syncthreads();                                        // everybody has read seq
    if (threadIdx.x == 0) *pv.seq_dev = seq;
}

// A rank whose evaluation failed locally still owes its peers a contribution, or they would wait for the timeout: NaN
// in both forms (slots + flag, self-validating entries) for the next sequence number, without waiting for anybody.
__global__ void __launch_bounds__(256) peer_poison_kernel(PeerView pv, int P) {
    const unsigned long long seq = *pv.seq_dev + 1;
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    const size_t base = ((size_t)(seq & 1ULL) * pv.world + pv.rank) * pv.cap;
    for (int i = threadIdx.x; i < P * pv.world; i += blockDim.x) {
        const int r = i / P, p = i - r * P;
        pv.win[r].data[base + p] = nan;
        ll_store(pv.win[r].ll + base + p, nan, (unsigned)seq);
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x < pv.world) st_release_sys(pv.win[threadIdx.x].flag + pv.rank, seq);
    if (threadIdx.x == 0) *pv.seq_dev = seq;
}

// ---- trajectory sharding: all-gather of the coefficient rows -------------------------------------------------------
// The per-pset work (trajectory integration + node values, traj_kernel) does not depend on the trees, so with the trees
// sharded over ranks every rank would repeat it for all P parameter sets.  With a ROW WINDOW (gcdyn_comm_rows_create)
// each rank integrates only its slice [p0, p1) of the parameter sets, writes those coefficient rows A[p][0..Jst) into its
// own window and this kernel pushes them — plus 32 bytes of per-pset metadata (status, fallback flag, cell count, error
// indicator) — into every peer's window with 16-byte P2P stores; the last CTA publishes {sequence, grid degree} in every
// peer's flag word, waits (bounded) for the peers' flags and copies the peers' metadata into the local arrays.  When the
// kernel completes, the local window holds all P rows and the contraction runs as usual.  Two halves alternate per call:
// a rank that is one call ahead (it cannot be further: the LL all-reduce of the call in between is a barrier) writes the
// other half.  The flag carries the grid degree: ranks that disagree (or a rank that could not take the contraction path,
// degree 0) raise `*err = 2` instead of contracting rows of the wrong shape.
struct RowMeta { int pstatus, need_sweep, S, D; double perr, pad; };
static_assert(sizeof(RowMeta) == 32, "RowMeta is two 16-byte stores");

struct RowExchange {
    double* rows[PEER_MAX_WORLD];            // rank r's window, the half of this call: [P*Jst doubles | RowMeta[P]]
    unsigned long long* flag[PEER_MAX_WORLD];// rank r's flag words [world]
    int* done;                               // local CTA counter (reset by the last CTA)
    int rank, world;
    unsigned long long seq, timeout_ns;
    int* err;
};

__global__ void __launch_bounds__(256) row_exchange_kernel(RowExchange rx, int P, long long Jst, long long Juse, int p0, int p1, int D,
                                                           int* __restrict__ pstatus, int* __restrict__ need_sweep,
                                                           int* __restrict__ S, double* __restrict__ perr) {
    __shared__ int s_last, s_bad;
    double* mine = rx.rows[rx.rank];
    RowMeta* meta_mine = reinterpret_cast<RowMeta*>(mine + (size_t)P * Jst);
    // this rank's metadata first (its own copy too: the import below reads every pset from the window)
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
    for (int p = p0 + tid; p < p1; p += nth) {
        RowMeta m;
        m.pstatus = pstatus[p]; m.need_sweep = need_sweep[p]; m.S = S[p]; m.D = D; m.perr = perr[p]; m.pad = 0.0;
        for (int r = 0; r < rx.world; ++r) {
            RowMeta* dst = reinterpret_cast<RowMeta*>(rx.rows[r] + (size_t)P * Jst) + p;
            reinterpret_cast<int4*>(dst)[0] = reinterpret_cast<const int4*>(&m)[0];
            reinterpret_cast<int4*>(dst)[1] = reinterpret_cast<const int4*>(&m)[1];
        }
    }
    // the rows of the slice, the first Juse columns of each (shared Γ: the K² type-change columns are not used): one load,
    // world − 1 peer stores (Jst and Juse are multiples of 4 doubles: 16-byte aligned)
    const unsigned w2 = (unsigned)(Juse / 2);                                     // double2 per row
    const long long n2 = (long long)(p1 - p0) * w2;
    const double2* src = reinterpret_cast<const double2*>(mine);
    for (long long i = tid; i < n2; i += nth) {
        const long long row = i / w2;
        const long long at = ((long long)p0 + row) * (Jst / 2) + (i - row * w2);
        const double2 v = __ldcg(src + at);
        for (int r = 0; r < rx.world; ++r)
            if (r != rx.rank) reinterpret_cast<double2*>(rx.rows[r])[at] = v;
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(rx.done, 1) == (int)gridDim.x - 1;
    __syncthreads();
    if (!s_last) return;
    if (threadIdx.x == 0) { *rx.done = 0; s_bad = 0; }
    __threadfence_system();
    __syncthreads();
    const unsigned long long word = (rx.seq << 8) | (unsigned long long)(D & 0xff);
    if (threadIdx.x < rx.world) st_release_sys(rx.flag[threadIdx.x] + rx.rank, word);
    if (threadIdx.x < rx.world) {
        const unsigned long long* f = rx.flag[rx.rank] + threadIdx.x;
        const unsigned long long t0 = global_ns();
        unsigned long long w;
        while (((w = ld_acquire_sys(f)) >> 8) < rx.seq) {
            __nanosleep(100);
            if (global_ns() - t0 > rx.timeout_ns) { atomicOr(&s_bad, 1); break; }
        }
        if ((w >> 8) >= rx.seq && (int)(w & 0xff) != (D & 0xff)) atomicOr(&s_bad, 2);
    }
    __syncthreads();
    if (s_bad) {
        if (threadIdx.x == 0 && rx.err) *rx.err = s_bad;
        // nothing of the peers can be trusted: every pset they own is marked as a solver failure (−Inf/NaN downstream)
        for (int p = threadIdx.x; p < P; p += blockDim.x)
            if (p < p0 || p >= p1) { pstatus[p] = 2 /* ST_SOLVER */; need_sweep[p] = 0; S[p] = 0; perr[p] = __longlong_as_double(0x7ff8000000000000LL); }
        return;
    }
    for (int p = threadIdx.x; p < P; p += blockDim.x) {
        if (p >= p0 && p < p1) continue;
        const int4 a = __ldcg(reinterpret_cast<const int4*>(meta_mine + p));
        const double2 b = __ldcg(reinterpret_cast<const double2*>(meta_mine + p) + 1);
        pstatus[p] = a.x; need_sweep[p] = a.y; S[p] = a.z; perr[p] = b.x;
    }
}

}  // namespace gcdyn
