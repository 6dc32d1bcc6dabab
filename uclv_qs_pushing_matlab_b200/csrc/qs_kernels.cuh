// qs_kernels.cuh — sm_100a kernels of the batched pusher-slider NMPC engine.
//
// Data layout (HBM): every per-problem quantity is a structure-of-arrays slab
//     slab[(stage * DIM + comp) * Bp + problem],   Bp = batch rounded up to 32,
// so a warp that walks problems reads/writes one 256-byte line per instruction, both in the
// (problem, stage)-parallel linearisation kernel and in the problem-per-thread QP kernel.
// The per-object spline tables (pp-form, 10 KB each) are staged once per CTA into shared memory
// by a single bulk TMA copy (cp.async.bulk + mbarrier).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "qs_solver.cuh"

namespace qs {

// ------------------------------------------------------------------------------------------------
// model staging: one bulk TMA copy global -> shared per CTA, completion on an mbarrier
// ------------------------------------------------------------------------------------------------
// Two halves so that a kernel can put model-independent work between the issue of the bulk copy and the wait for it.
__device__ __forceinline__ void stage_models_issue(const double* __restrict__ gmodels, int nmodels) {
    extern __shared__ __align__(128) unsigned char qs_smem[];
    uint64_t* mbar = reinterpret_cast<uint64_t*>(qs_smem);
    double* dst = reinterpret_cast<double*>(qs_smem + 128);
    const uint32_t bytes = (uint32_t)nmodels * MODEL_DOUBLES * 8u;
    const uint32_t mbar_s = (uint32_t)__cvta_generic_to_shared(mbar);
    const uint32_t dst_s = (uint32_t)__cvta_generic_to_shared(dst);
    const bool leader = (threadIdx.x | threadIdx.y | threadIdx.z) == 0;     // one thread of the CTA (2-D blocks: k_linesearch)
    if (leader) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar_s) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (leader) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar_s), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(dst_s), "l"(gmodels), "r"(bytes), "r"(mbar_s) : "memory");
    }
}
__device__ __forceinline__ const double* stage_models_wait() {
    extern __shared__ __align__(128) unsigned char qs_smem[];
    const uint32_t mbar_s = (uint32_t)__cvta_generic_to_shared(qs_smem);
    uint32_t done = 0;
    while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(mbar_s), "r"(0u) : "memory");
    }
    return reinterpret_cast<const double*>(qs_smem + 128);
}
__device__ __forceinline__ const double* stage_models(const double* __restrict__ gmodels, int nmodels) {
    stage_models_issue(gmodels, nmodels);
    return stage_models_wait();
}
inline size_t model_smem_bytes(int nmodels) { return 128 + (size_t)nmodels * MODEL_DOUBLES * 8; }

// ------------------------------------------------------------------------------------------------
// AoS <-> SoA transposition (set/get): user buffer [nb][R]  <->  slab rows [row0 .. row0+R) x Bp
// ------------------------------------------------------------------------------------------------
__global__ void k_aos_to_soa(const double* __restrict__ src, double* __restrict__ dst, int nb, int R, int row0, int lo, int Bp) {
    __shared__ double tile[32][33];
    const int bx = blockIdx.x * 32, ry = blockIdx.y * 32;
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {          // read: r fastest
        const int i = bx + j, r = ry + threadIdx.x;
        if (i < nb && r < R) tile[j][threadIdx.x] = src[(size_t)i * R + r];
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {          // write: problem fastest
        const int r = ry + j, i = bx + threadIdx.x;
        if (i < nb && r < R) dst[(size_t)(row0 + r) * Bp + lo + i] = tile[threadIdx.x][j];
    }
}
__global__ void k_soa_to_aos(const double* __restrict__ src, double* __restrict__ dst, int nb, int R, int row0, int lo, int Bp) {
    __shared__ double tile[32][33];
    const int bx = blockIdx.x * 32, ry = blockIdx.y * 32;
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int r = ry + j, i = bx + threadIdx.x;
        if (i < nb && r < R) tile[j][threadIdx.x] = src[(size_t)(row0 + r) * Bp + lo + i];
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int i = bx + j, r = ry + threadIdx.x;
        if (i < nb && r < R) dst[(size_t)i * R + r] = tile[threadIdx.x][j];
    }
}
__global__ void k_fill_int(int* p, int n, int v) { int i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n) p[i] = v; }

// ------------------------------------------------------------------------------------------------
// K1 / K2 stateless evaluation kernels (config 2 entry points), AoS [cnt][dim] like the C-ABI
// ------------------------------------------------------------------------------------------------
__global__ void k_eval_spline(const double* __restrict__ gmodel, int cnt, const double* __restrict__ s, int wrap, int single,
                              double* C, double* Cd, double* Cdd, double* tv, double* nv, double* kappa) {
    const double* M = stage_models(gmodel, 1);
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= cnt) return;
    double sg = s[i];
    if (wrap == 1) sg = matlab_mod(sg, M[1], single != 0);
    else if (wrap == 2) sg = wrap_dyn(sg, M[1]);
    Curve c; curve_eval(M, sg, c);
    if (C) reinterpret_cast<double2*>(C)[i] = make_double2(c.cx, c.cy);
    if (Cd) reinterpret_cast<double2*>(Cd)[i] = make_double2(c.dx, c.dy);
    if (Cdd) { double ex, ey; curve_dd(M, sg, ex, ey); reinterpret_cast<double2*>(Cdd)[i] = make_double2(ex, ey); }
    if (tv || nv) {
        const double nrm = sqrt(c.dx * c.dx + c.dy * c.dy);
        const double tx = c.dx / nrm, ty = c.dy / nrm;
        if (tv) reinterpret_cast<double2*>(tv)[i] = make_double2(tx, ty);
        if (nv) reinterpret_cast<double2*>(nv)[i] = make_double2(ty, -tx);
    }
    if (kappa) kappa[i] = (c.dx * c.hy - c.dy * c.hx) / (c.dx * c.dx + c.dy * c.dy);
}

__global__ void k_eval_dynamics(const double* __restrict__ gmodel, int cnt, const double* __restrict__ x, const double* __restrict__ u,
                                double* f, double* Jx, double* Ju) {
    const double* M = stage_models(gmodel, 1);
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= cnt) return;
    const double2 xa = reinterpret_cast<const double2*>(x)[2 * i + 1];   // theta, s
    const double2 ua = reinterpret_cast<const double2*>(u)[i];
    Dyn d;
    if (Jx || Ju) dyn_eval<true>(M, xa.x, xa.y, ua.x, ua.y, d); else dyn_eval<false>(M, xa.x, xa.y, ua.x, ua.y, d);
    double2* fo = reinterpret_cast<double2*>(f) + 2 * i;
    fo[0] = make_double2(d.f[0], d.f[1]); fo[1] = make_double2(d.f[2], d.f[3]);
    if (Jx) {
        double2* o = reinterpret_cast<double2*>(Jx) + 8 * i;
        const double jt[4] = {-d.f[1], d.f[0], 0.0, 0.0};
#pragma unroll
        for (int r = 0; r < 4; ++r) { o[2 * r] = make_double2(0.0, 0.0); o[2 * r + 1] = make_double2(jt[r], d.fs[r]); }
    }
    if (Ju) {
        double2* o = reinterpret_cast<double2*>(Ju) + 4 * i;
#pragma unroll
        for (int r = 0; r < 4; ++r) o[r] = make_double2(d.fun[r], d.fut[r]);
    }
}

// 3 CTAs of 128 threads per SM (168 registers): measured +24 % over the unconstrained 182-register build
__global__ void __launch_bounds__(128, 3)
k_eval_erk4(const double* __restrict__ gmodel, int cnt, const double* __restrict__ x, const double* __restrict__ u, double dt,
            double* Phi, double* A, double* Bo) {
    const double* M = stage_models(gmodel, 1);
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= cnt) return;
    const double2 x01 = reinterpret_cast<const double2*>(x)[2 * i], x23 = reinterpret_cast<const double2*>(x)[2 * i + 1];
    const double2 ua = reinterpret_cast<const double2*>(u)[i];
    const double xv[4] = {x01.x, x01.y, x23.x, x23.y};
    double P4[4], Sm[16];
    erk4_sens(M, xv, ua.x, ua.y, dt, P4, Sm);
    double2* po = reinterpret_cast<double2*>(Phi) + 2 * i;
    po[0] = make_double2(P4[0], P4[1]); po[1] = make_double2(P4[2], P4[3]);
    if (A) {
        double2* o = reinterpret_cast<double2*>(A) + 8 * i;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            o[2 * r] = make_double2(r == 0 ? 1.0 : 0.0, r == 1 ? 1.0 : 0.0);
            o[2 * r + 1] = make_double2(Sm[4 * r + 0], Sm[4 * r + 1]);
        }
    }
    if (Bo) {
        double2* o = reinterpret_cast<double2*>(Bo) + 4 * i;
#pragma unroll
        for (int r = 0; r < 4; ++r) o[r] = make_double2(Sm[4 * r + 2], Sm[4 * r + 3]);
    }
}

__global__ void k_eval_vbound(const double* __restrict__ gmodel, int cnt, const double* __restrict__ s, CtrlDev cp,
                              double* vb, double* ta) {
    const double* M = stage_models(gmodel, 1);
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= cnt) return;
    double t;
    const double v = v_bound_of(M, s[i], cp.v_alpha, cp.d_v_bound, cp.t_angle0, cp.u_t_ub, cp.single != 0, &t);
    if (vb) vb[i] = v;
    if (ta) ta[i] = t;
}

// ------------------------------------------------------------------------------------------------
// solver kernels: thin launch wrappers around the per-thread bodies in qs_solver.cuh
// ------------------------------------------------------------------------------------------------
// one thread per (problem, stage); consecutive threads walk problems -> coalesced slab accesses
__global__ void __launch_bounds__(128, 3) k_linearise(SolverDev S) {
    const double* Mall = stage_models(S.models, S.nmodels);
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int k = (int)(tid / S.Bp), b = (int)(tid % S.Bp);
    if (k > S.N || b >= S.B) return;
    if (S.done && S.done[b]) return;
    linearise_one(S, Mall, k, b);
}

// one problem per thread, ppw problems packed per warp (one warp per CTA)
__global__ void __launch_bounds__(32) k_qp(SolverDev S, IpmOpts o, int ppw, int apply) {
    const int lane = threadIdx.x & 31;
    if (lane >= ppw) return;
    const int b = blockIdx.x * ppw + lane;
    if (b >= S.B) return;
    if (S.done && S.done[b]) return;
    qp_one(S, o, b, apply);
}

// K4 v2: one or two problems per warp (SEG = 32 / 16 lanes each), persistent CTAs of up to QW_MAX_WARPS warps (one CTA
// per SM); finished warps pull the next problem(s) from a global work queue (S.ndone[1], order S.order).  The IPM state of each problem lives in its warp's slice of
// the dynamic shared memory; the read-only linearisation lives in tensor memory: the CTA allocates all 512 TMEM
// columns, warp w owns TMEM lanes 32*(w%4).. (the quarter the hardware lets it address) and columns 256*(w/4)..
constexpr int QW_MAX_WARPS = 8;
// Work-queue order of the warp QP kernel: problems sorted by DESCENDING IPM iteration count of their previous solve
// (counting sort, one CTA).  Long problems first shortens the tail of the queue, and — with two problems per warp —
// neighbours in the order need about the same number of iterations, so a segment rarely idles while its partner
// finishes (random pairs lose ~10 % of the segment-iterations).  In closed loop the previous count is a good
// predictor; the order never changes a result (every problem's arithmetic is independent of its partner).
constexpr int QO_BINS = 64;          // qp_order_cta scans the bins with one warp, two per lane
static_assert(QO_BINS == 64, "qp_order_cta: two bins per lane");
// counting sort of the problems by descending previous IPM iteration count, by ONE CTA.  The batch has a handful of distinct
// counts, so plain shared-memory atomics serialise (4096 increments on ~10 addresses: 22 us); lanes with equal keys are grouped
// with match.any, one atomic per group (1 us).
__device__ __forceinline__ void qp_order_cta(const SolverDev& S, int* __restrict__ order) {
    __shared__ int cnt[QO_BINS], pos[QO_BINS];
    const int lane = threadIdx.x & 31;
    if (threadIdx.x < QO_BINS) cnt[threadIdx.x] = 0;
    __syncthreads();
    for (int b0 = 0; b0 < S.B; b0 += blockDim.x) {               // uniform trip count: match.any is warp-collective
        const int b = b0 + threadIdx.x;
        const int k = b < S.B ? S.qp_last[b] : 0;
        const int key = b < S.B ? QO_BINS - 1 - (k < 0 ? 0 : (k > QO_BINS - 1 ? QO_BINS - 1 : k)) : QO_BINS;
        const unsigned peers = __match_any_sync(0xffffffffu, key);
        if (key < QO_BINS && lane == __ffs(peers) - 1) atomicAdd(&cnt[key], __popc(peers));
    }
    __syncthreads();
    if (threadIdx.x < 32) {                                      // exclusive prefix sum of the 64 bins by one warp
        const int c0 = cnt[2 * lane], c1 = cnt[2 * lane + 1];
        int a = c0 + c1;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, a, o); if (lane >= o) a += t; }
        pos[2 * lane] = a - c0 - c1; pos[2 * lane + 1] = a - c1;
    }
    __syncthreads();
    for (int b0 = 0; b0 < S.B; b0 += blockDim.x) {
        const int b = b0 + threadIdx.x;
        const int k = b < S.B ? S.qp_last[b] : 0;
        const int key = b < S.B ? QO_BINS - 1 - (k < 0 ? 0 : (k > QO_BINS - 1 ? QO_BINS - 1 : k)) : QO_BINS;
        const unsigned peers = __match_any_sync(0xffffffffu, key);
        const int leader = __ffs(peers) - 1;
        int base = 0;
        if (key < QO_BINS && lane == leader) base = atomicAdd(&pos[key], __popc(peers));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (key < QO_BINS) order[base + __popc(peers & ((1u << lane) - 1u))] = b;
    }
}
__global__ void __launch_bounds__(1024) k_qp_order(SolverDev S, int* __restrict__ order) { qp_order_cta(S, order); }
template <int C, int HV, int SEG>
__global__ void __launch_bounds__(32 * QW_MAX_WARPS, 1) k_qp_warp(SolverDev S, IpmOpts o, int apply, int per_problem_doubles, int lockstep) {
    extern __shared__ __align__(16) double qw_smem[];
    __shared__ unsigned tmem_base;
    const int wid = threadIdx.x >> 5;
    if (wid == 0) {
        const unsigned dst = (unsigned)__cvta_generic_to_shared(&tmem_base);
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" :: "r"(dst) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tbase = tmem_base;
    WarpCtxDev w{(int)(threadIdx.x & 31), tbase + (((unsigned)(wid & 3) * 32u) << 16) + (unsigned)(wid >> 2) * 256u, lockstep};
    // work queue: one atomic per warp fetches 32 / SEG consecutive problems, segment `seg` takes the seg-th of them
    auto next = [&](int seg) -> int {
        constexpr int PPW = 32 / SEG;
        int b = -1;
        for (;;) {
            if (w.lane() == 0) b = atomicAdd(S.ndone + 1, PPW);
            b = w.bcast_int(b);
            if (b >= S.B) return -1;
            const int slot = b + seg;
            const int mine = slot < S.B ? (S.order ? S.order[slot] : slot) : S.B;
            const bool ok = mine < S.B && !(S.done && S.done[mine]);      // full SQP: skip problems that already converged
            if (w.wany(ok ? 1 : 0)) return ok ? mine : -1;
        }
    };
    qp_warp_persistent<WarpCtxDev, C, HV, SEG>(w, qw_smem + (size_t)wid * (32 / SEG) * per_problem_doubles, per_problem_doubles, S, o, apply, next);
    // the warps leave the loop independently or through the CTA-wide vote (w.lockstep()); once all of them are here no TMEM access is in flight
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (wid == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" :: "r"(tbase) : "memory");
}

// SQP-level kernels: CTA = 32 problems (threadIdx.x, coalesced slab accesses) x LS_CHUNKS stage chunks (threadIdx.y).
// Line search: the
// N ERK4 evaluations of every merit-function trial are spread over the chunk threads and summed through shared
// memory, instead of one serial loop per problem (which was 59 % of a full-SQP solve at N = 100).
constexpr int LS_CHUNKS = 8;
struct ChunkCta {
    int y, ny;
    double* red;                                   // [ny][32]
    __device__ __forceinline__ double sum(double v) const {
        red[y * 32 + threadIdx.x] = v;
        __syncthreads();
        double t = 0.0;
        for (int i = 0; i < ny; ++i) t += red[i * 32 + threadIdx.x];      // same order in every chunk thread: identical totals
        __syncthreads();
        return t;
    }
    __device__ __forceinline__ double max(double v) const {
        red[y * 32 + threadIdx.x] = v;
        __syncthreads();
        double t = red[threadIdx.x];
        for (int i = 1; i < ny; ++i) t = fmax(t, red[i * 32 + threadIdx.x]);
        __syncthreads();
        return t;
    }
    __device__ __forceinline__ bool any(bool p) const { return __syncthreads_or(p ? 1 : 0) != 0; }
};
// NLP residuals + convergence test: same CTA shape (the serial per-problem loop was 7 % of a full-SQP solve)
__global__ void __launch_bounds__(32 * LS_CHUNKS) k_nlp_res(SolverDev S, SqpOpts o, int it) {
    __shared__ double red[LS_CHUNKS * 32];
    const int b = blockIdx.x * 32 + threadIdx.x;
    const bool live = b < S.B && !S.done[b];
    ChunkCta ch{(int)threadIdx.y, LS_CHUNKS, red};
    if (nlp_res_one(S, o, it, live ? b : 0, ch, live)) atomicAdd(S.ndone, 1);
}
__global__ void __launch_bounds__(32 * LS_CHUNKS) k_linesearch(SolverDev S, SqpOpts o, int it) {
    __shared__ double red[LS_CHUNKS * 32];
    const double* Mall = stage_models(S.models, S.nmodels);
    const int b = blockIdx.x * 32 + threadIdx.x;
    const bool live = b < S.B && !S.done[b];
    ChunkCta ch{(int)threadIdx.y, LS_CHUNKS, red};
    if (linesearch_one(S, o, Mall, it, live ? b : 0, ch, live)) atomicAdd(S.ndone, 1);
}

__global__ void k_cost(SolverDev S) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= S.B) return;
    cost_one(S, b);
}

__global__ void k_shift(SolverDev S) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= S.B) return;
    shift_one(S, blockIdx.y, b);
}

// forward-Euler plant step on caller arrays [B][4], [B][2] (helper.m:294, 307)
__global__ void k_plant_step(SolverDev S, double* x, const double* u) {
    const double* Mall = stage_models(S.models, S.nmodels);
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= S.B) return;
    const double* M = Mall + (size_t)S.objid[b] * MODEL_DOUBLES;
    double2* xp = reinterpret_cast<double2*>(x) + 2 * b;
    const double2 x01 = xp[0], x23 = xp[1];
    const double2 ua = reinterpret_cast<const double2*>(u)[b];
    Dyn d;
    dyn_eval<false>(M, x23.x, x23.y, ua.x, ua.y, d);
    xp[0] = make_double2(fma(S.dt, d.f[0], x01.x), fma(S.dt, d.f[1], x01.y));
    xp[1] = make_double2(fma(S.dt, d.f[2], x23.x), fma(S.dt, d.f[3], x23.y));
}

// ---- device-resident closed loop (qspush_closed_loop): thin wrappers, one thread per problem / per (stage, problem)
__global__ void k_loop_window(SolverDev S, LoopDev L, int idx) {
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int k = (int)(tid / S.Bp), b = (int)(tid % S.Bp);
    if (k >= S.N || b >= S.B) return;
    loop_window_one(S, L, idx, k, b);
}
__global__ void __launch_bounds__(128) k_loop_state(SolverDev S, LoopDev L, int step, double* xs, double* log_x) {
    const double* Mall = stage_models(S.models, S.nmodels);
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= S.B) return;
    loop_state_one(S, L, Mall, step, xs, log_x, b);
}
__global__ void __launch_bounds__(128) k_loop_post(SolverDev S, LoopDev L, int step, double* xs, double* log_u, int* log_status) {
    const double* Mall = stage_models(S.models, S.nmodels);
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= S.B) return;
    loop_post_one(S, L, Mall, step, xs, log_u, log_status, b);
}


// ---- FP64 roofline denominator, measured (MEASURED_PEAKS.json has no FP64 entry): 16 independent
// DFMA chains per thread, 4 x 256-thread CTAs per SM resident, no memory traffic.  qspush_measure_fp64_peak times it.
__global__ void __launch_bounds__(256) k_fp64_peak(double* __restrict__ out, int iters, double a, double b) {
    double v[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = (double)(threadIdx.x + i) * 1e-3;
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = fma(v[i], a, b);
    }
    double t = 0.0;
#pragma unroll
    for (int i = 0; i < 16; ++i) t += v[i];
    if (t == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = t;     // never true: keeps the chains alive
}

// ---- controller-level step (qspush_step): the whole NMPC_controller.solve pre-processing in ONE launch and the whole
// post-processing in one.  k_prepare (step form): constr_x0 from the caller's [B][4] array, reference window of period *idx
// (NMPC_controller.m:307-313, 343-348; skipped when no trajectory is set), then prepare_one (x0 wrap, cold start, v_bound
// clip, Euler rollout).  k_step_out: u0 = get('u', 0) and status to [B][2] / [B] arrays; the last CTA also sorts the
// work queue of the NEXT solve by this solve's IPM iteration counts (what k_qp_order does as a launch of its own).
// K6.  ONE kernel for both entry points (two instantiations of prepare_one were contracted differently by the compiler and
// disagreed in the last bit): qspush_prepare passes x0 = nullptr / L.traj = nullptr and the kernel works on the slabs as set;
// qspush_step passes the caller's [B][4] state and the device-resident reference trajectory.
// 256 threads per 128 problems: threads 0..127 run the (serial, latency-bound) rollout of one problem each, threads 128..255 write
// the reference window of the same problems meanwhile (r02: the window loop was 18 % and the wait for the model tables 10 % of
// the kernel when one thread did everything in sequence).
constexpr int PREP_PROBLEMS = 128;
__global__ void __launch_bounds__(2 * PREP_PROBLEMS) k_prepare(SolverDev S, CtrlDev cp, LoopDev L, const int* __restrict__ idx, const double* __restrict__ x0) {
    stage_models_issue(S.models, S.nmodels);
    const int b = blockIdx.x * PREP_PROBLEMS + (threadIdx.x & (PREP_PROBLEMS - 1));
    if (threadIdx.x >= PREP_PROBLEMS) {                         // window threads: no model needed
        if (b < S.B && L.traj) {
            const int i0 = *idx;
            for (int k = 0; k < S.N; ++k) loop_window_one(S, L, i0, k, b);
        }
        return;
    }
    if (b < S.B && x0) {
        const double2 a = reinterpret_cast<const double2*>(x0)[2 * b], c = reinterpret_cast<const double2*>(x0)[2 * b + 1];
        QS_EL(S.x0, 0, b) = a.x; QS_EL(S.x0, 1, b) = a.y; QS_EL(S.x0, 2, b) = c.x; QS_EL(S.x0, 3, b) = c.y;
    }
    const double* Mall = stage_models_wait();
    if (b >= S.B) return;
    prepare_one(S, cp, Mall, b);
}
__global__ void __launch_bounds__(1024) k_step_out(SolverDev S, double* __restrict__ u0, int* __restrict__ status, int* __restrict__ order) {
    if (blockIdx.x + 1 == gridDim.x) { if (order) qp_order_cta(S, order); return; }   // last CTA: counting sort, descending previous iteration count
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= S.B) return;
    reinterpret_cast<double2*>(u0)[b] = make_double2(QS_EL(S.u, 0, b), QS_EL(S.u, 1, b));
    status[b] = S.status[b];
}

}  // namespace qs
