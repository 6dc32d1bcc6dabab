// qs_qp_warp.cuh — K4, second generation: one WARP per problem, parallel in time.
//
// Same QP and the same Mehrotra predictor-corrector IPM in Newton-step form as qs_qp.cuh (true residuals
// every iteration, centering floor, stall exit), but the horizon is spread over the 32 lanes of a warp:
// lane l owns the C consecutive stages l*C .. l*C+C-1 (C = ceil((N+1)/32)); the whole IPM state of the
// problem lives on chip — shared memory records plus a lane-private block of tensor memory (no HBM traffic
// inside the iteration loop) — and every "sweep" of the serial algorithm becomes either lane-local work or a
// warp scan (operands exchanged through shared memory, reductions through shuffles):
//
//   * Riccati matrices P_k: suffix scan of conditional value functions V_{i->j}(x_i, x_j), each an element
//     (A, C, J) with  V = max_lam 1/2 x'Jx - 1/2 lam'C lam + lam'(x_j - A x_i)  [vector parts omitted]
//     (Sarkka & Garcia-Fernandez, "Temporal parallelization of dynamic programming and LQ control").
//     The combination is evaluated in a symmetric, Cholesky-based form that keeps C and J positive
//     semi-definite by construction:  J2 = L2 L2',  T = C1 L2,  W = I + L2' T = Lw Lw',
//     Z = Lw^-1 L2' A1,  Y = Lw^-1 T'   =>   A = A2 (A1 - Y'Z),  C = A2 (C1 - Y'Y) A2' + C2,  J = Z'Z + J1.
//   * with P at the right boundary of its chunk, each lane runs the ordinary Riccati stage (qs_qp.cuh)
//     over its own C stages: gains K_k, Cholesky factors, P_k;
//   * the vector recursions p_k = Abar_k' p_{k+1} + d_k (backward) and dx_{k+1} = Abar_k dx_k + bbar_k
//     (forward), Abar = A - B K, are suffix / prefix scans of affine maps — twice per iteration
//     (predictor and corrector run through ONE copy of the solve code: instruction-cache footprint);
//   * residual norms, step lengths and complementarity sums are warp reductions.
//
// Constraint rows: the default set h = [s; u_n; u_t] (selection rows) or, template flag HV = 1, the reference's
// parked set h = [u_n; u_t - v_bound(s); u_t + v_bound(s)] whose rows couple ds and du_t (barrier cross term).
//
// The code is written against a tiny warp context (lane id, reductions, TMEM block, CTA vote) so that
// tests/hostsim can run the identical source on the CPU with a fiber-based warp emulator.
#pragma once
#include "qs_qp.cuh"

namespace qs {

#if defined(__CUDACC__)
// tcgen05.ld / tcgen05.st of n doubles (2n 32-bit columns) of this thread's TMEM lane.  The instructions are
// warp-collective (.sync.aligned): callers keep them outside lane-divergent code, the __syncwarp() in front
// re-converges lanes that left the previous loop iteration through different paths.
#define QW_TM_REGS8 "t0,t1,t2,t3,t4,t5,t6,t7"
#define QW_TM_REGS16 QW_TM_REGS8 ",t8,t9,t10,t11,t12,t13,t14,t15"
#define QW_TM_REGS32 QW_TM_REGS16 ",t16,t17,t18,t19,t20,t21,t22,t23,t24,t25,t26,t27,t28,t29,t30,t31"
struct WarpCtxDev {
    int lane_;
    unsigned tm_;        // TMEM address of this warp's private block: (32 * (warp % 4)) << 16 | first column
    int ls_;             // 1: the warps of this CTA meet at a vote once per IPM iteration (chosen by launch_qp)
    // The 32-bit destination registers are asm OUTPUTS (not block-local temporaries copied out with mov.b64): the register
    // allocator then places the LDTM destination block where the consumers read it, each aligned pair being one double —
    // with block-local temporaries every load was followed by one MOV per 32-bit register (6 % of the kernel's instructions).
    template <int n> __device__ __forceinline__ void tm_ld(int off, double* v) const {
        static_assert(n == 4 || n == 8 || n == 16, "tm_ld: 4, 8 or 16 doubles");
        const unsigned a = tm_ + 2u * (unsigned)off;
        unsigned r[2 * n];
        __syncwarp();
        if constexpr (n == 4) {
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n\t"
                         "tcgen05.wait::ld.sync.aligned;"
                         : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(a) : "memory");
        } else if constexpr (n == 8) {
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n\t"
                         "tcgen05.wait::ld.sync.aligned;"
                         : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                           "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(a) : "memory");
        } else {
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
                         "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n\t"
                         "tcgen05.wait::ld.sync.aligned;"
                         : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                           "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                           "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                           "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]) : "r"(a) : "memory");
        }
#pragma unroll
        for (int i = 0; i < n; ++i) v[i] = __hiloint2double((int)r[2 * i + 1], (int)r[2 * i]);
    }
    template <int n> __device__ __forceinline__ void tm_st(int off, const double* v) const {
        static_assert(n == 4 || n == 8, "tm_st: 4 or 8 doubles (16: tm_st16)");
        const unsigned a = tm_ + 2u * (unsigned)off;
        __syncwarp();
        if constexpr (n == 4) {
            asm volatile("{\n\t.reg .b32 t<8>;\n\t"
                         "mov.b64 {t0,t1}, %1;\n\tmov.b64 {t2,t3}, %2;\n\tmov.b64 {t4,t5}, %3;\n\tmov.b64 {t6,t7}, %4;\n\t"
                         "tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {" QW_TM_REGS8 "};\n\t"
                         "tcgen05.wait::st.sync.aligned;\n\t}"
                         :: "r"(a), "d"(v[0]), "d"(v[1]), "d"(v[2]), "d"(v[3]) : "memory");
        } else {
            asm volatile("{\n\t.reg .b32 t<16>;\n\t"
                         "mov.b64 {t0,t1}, %1;\n\tmov.b64 {t2,t3}, %2;\n\tmov.b64 {t4,t5}, %3;\n\tmov.b64 {t6,t7}, %4;\n\t"
                         "mov.b64 {t8,t9}, %5;\n\tmov.b64 {t10,t11}, %6;\n\tmov.b64 {t12,t13}, %7;\n\tmov.b64 {t14,t15}, %8;\n\t"
                         "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {" QW_TM_REGS16 "};\n\t"
                         "tcgen05.wait::st.sync.aligned;\n\t}"
                         :: "r"(a), "d"(v[0]), "d"(v[1]), "d"(v[2]), "d"(v[3]), "d"(v[4]), "d"(v[5]), "d"(v[6]), "d"(v[7]) : "memory");
        }
    }
    // 16 doubles -> 32 columns
    __device__ __forceinline__ void tm_st16(int off, const double* v) const {
        const unsigned a = tm_ + 2u * (unsigned)off;
        __syncwarp();
        asm volatile("{\n\t.reg .b32 t<32>;\n\t"
                     "mov.b64 {t0,t1}, %1;\n\tmov.b64 {t2,t3}, %2;\n\tmov.b64 {t4,t5}, %3;\n\tmov.b64 {t6,t7}, %4;\n\t"
                     "mov.b64 {t8,t9}, %5;\n\tmov.b64 {t10,t11}, %6;\n\tmov.b64 {t12,t13}, %7;\n\tmov.b64 {t14,t15}, %8;\n\t"
                     "mov.b64 {t16,t17}, %9;\n\tmov.b64 {t18,t19}, %10;\n\tmov.b64 {t20,t21}, %11;\n\tmov.b64 {t22,t23}, %12;\n\t"
                     "mov.b64 {t24,t25}, %13;\n\tmov.b64 {t26,t27}, %14;\n\tmov.b64 {t28,t29}, %15;\n\tmov.b64 {t30,t31}, %16;\n\t"
                     "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {" QW_TM_REGS32 "};\n\t"
                     "tcgen05.wait::st.sync.aligned;\n\t}"
                     :: "r"(a), "d"(v[0]), "d"(v[1]), "d"(v[2]), "d"(v[3]), "d"(v[4]), "d"(v[5]), "d"(v[6]), "d"(v[7]),
                        "d"(v[8]), "d"(v[9]), "d"(v[10]), "d"(v[11]), "d"(v[12]), "d"(v[13]), "d"(v[14]), "d"(v[15])
                     : "memory");
    }
    __device__ __forceinline__ int lane() const { return lane_; }
    __device__ __forceinline__ double shfl(double v, int src) const { return __shfl_sync(0xffffffffu, v, src & 31); }
    // reductions over the SEG lanes of this lane's segment (SEG = 32: the warp; 16: two problems per warp, short horizons)
    template <int SEG = 32> __device__ __forceinline__ double wmax(double v) const {
#pragma unroll
        for (int o = SEG / 2; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
        return v;
    }
    template <int SEG = 32> __device__ __forceinline__ double wmin(double v) const {
#pragma unroll
        for (int o = SEG / 2; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
        return v;
    }
    template <int SEG = 32> __device__ __forceinline__ double wsum(double v) const {
#pragma unroll
        for (int o = SEG / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        return v;
    }
    template <int SEG = 32> __device__ __forceinline__ int wany(int p) const {
        const unsigned m = __ballot_sync(0xffffffffu, p);
        if constexpr (SEG == 32) return m != 0u;
        else return (m & (((1u << SEG) - 1u) << (lane_ & ~(SEG - 1)))) != 0u;
    }
    __device__ __forceinline__ void sync() const { __syncwarp(); }
    // CTA-wide vote of the CTAs with more than one warp per scheduler (one barrier per IPM iteration so that the warps share
    // instruction fetches; every CTA did this up to r02 v14); returns true when every warp of the CTA has finished
    __device__ __forceinline__ bool lockstep() const { return ls_ != 0; }
    __device__ __forceinline__ bool cta_all(bool pred) const { return __syncthreads_and(pred ? 1 : 0) != 0; }
    __device__ __forceinline__ int bcast_int(int v) const { return __shfl_sync(0xffffffffu, v, 0); }
};
#endif

// optional per-phase cycle counters (build with -DQW_PROFILE; development aid, tools/gpu_phase_profile.py)
#if defined(QW_PROFILE) && defined(__CUDACC__)
__device__ unsigned long long qw_prof[16];
#endif
#if defined(QW_PROFILE) && defined(__CUDA_ARCH__)
#define QW_T0() long long qw_t__ = clock64()
#define QW_TICK(slot) do { const long long n__ = clock64(); if (w.lane() == 0) atomicAdd(&qw_prof[slot], (unsigned long long)(n__ - qw_t__)); qw_t__ = n__; } while (0)
#else
#define QW_T0() do { } while (0)
#define QW_TICK(slot) do { } while (0)
#endif

// Per-problem state.  Two on-chip homes:
//   * shared memory: one RECORD of R_ROWS doubles per (local stage j, lane l) at sm[(j*L + l)*R_ROWS + row]: a row
//     access is a compile-time immediate offset from the record pointer (no integer address arithmetic), and the odd
//     record stride makes the 64-bit accesses of a half-warp conflict-free;
//   * tensor memory (TMEM): the read-only linearisation of the stage (A, B, b, h, g: 29 doubles, padded to
//     32) sits in the lane's own TMEM lane, written once per problem and
//     read back with tcgen05.ld.  TMEM is otherwise idle in this kernel; moving these rows out of the records brings
//     the record from 101 to 73 doubles, i.e. 8 instead of 6 resident problems per SM at N = 40 (shared memory is
//     what bounds the number of resident warps).
enum : int {
    R_Z = 0, R_PIK = 6, R_LAM = 10, R_T = 16,                                      // point (22): z, pi_k, lam, t
    R_RG = 22, R_RB = 28,                                                          // residuals
    R_K = 32, R_LI = 40, R_DZA = 43, R_PV = 46, R_KFF = 50,                        // factor, affine step, p_k, k_ff: 20 rows that are dead during the element scan
    R_GT = 52,                                                                     // rhs / step (aliased)
    R_P = 58, R_PB = 68,                                                           // P_k, P_{k+1} r_b (long horizons: in TMEM instead)
    R_ROWS = 73, R_ROWS_LONG = 59
};
// Long horizons (C >= 3: N >= 64) are the most shared-memory-starved configurations (3 resident problems per SM at
// N = 100), and with at most 4 warps per CTA each warp owns a full TMEM lane quarter (512 columns): P_k and
// P_{k+1} r_b move to TMEM as well, the record shrinks to 59 rows (4 problems per SM at N = 100).
QS_HD constexpr int qw_rows(int C) { return C >= 3 ? R_ROWS_LONG : R_ROWS; }
// offsets (doubles) inside the TMEM block of one local stage
// (A, B 0..15, h 16..19 and b, g 32..43 are written once per problem; 20..31 are rewritten every IPM iteration in phase (2):
// the slack reciprocals 1/t_l, 1/t_u and the barrier diagonal D = lam_l/t_l + lam_u/t_u, reused by phases (4)-(8) — 30 FP64
// divisions per stage and iteration become 6; 48..63, long horizons only: P_k (10) and P_{k+1} r_b (4))
// r02: [h | 1/t | D] are one 16-double group (one tcgen05.ld / st where two or three were issued) and [b | g] another
enum : int { QW_TM_AB = 0, QW_TM_HH = 16, QW_TM_IT = 20, QW_TM_D = 28, QW_TM_BV = 32, QW_TM_G = 36, QW_TM_P = 48, QW_TM_STAGE = 48, QW_TM_STAGE_LONG = 64 };
QS_HD constexpr int qw_tm_stage(int C) { return C >= 3 ? QW_TM_STAGE_LONG : QW_TM_STAGE; }
// Mapping of a horizon onto a warp.  seg = 16: two problems per warp, one per 16-lane segment, each lane owning
// C = ceil((N+1)/16) stages; seg = 32: one problem per warp, C = ceil((N+1)/32).  Measured on B200 (tools/gpu_half_sweep.py):
// the two-problem mapping wins for every horizon it fits (N = 10: +65 %, 20: +31 %, 31: +20 %, 40: +8 %, 47: +15 %, 55: +5 %;
// 56..63: +7..15 % once the work queue pairs problems of equal iteration count, k_qp_order): the same 8 problems are resident per SM either way (shared memory), but one instruction stream serves
// two problems with 85-100 % instead of 35-65 % of the lanes busy and the scans are one step shorter.
#ifndef QW_PLAN_MAX
#define QW_PLAN_MAX 63            // longest horizon that runs two problems per warp (C = ceil((N+1)/16) <= 4)
#endif
#ifndef QW_PLAN8_MAX
#define QW_PLAN8_MAX (-1)         // longest horizon that runs four problems per warp (8-lane segments); -1: never
#endif
struct QwPlan { int C, seg; };
QS_HD constexpr QwPlan qp_warp_plan(int N) {
    return (N <= QW_PLAN8_MAX) ? QwPlan{(N + 1 + 7) / 8, 8}
         : (N <= QW_PLAN_MAX) ? QwPlan{(N + 1 + 15) / 16, 16} : QwPlan{(N + 1 + 31) / 32, 32};
}
QS_HD constexpr int qp_warp_chunk(int N) { return qp_warp_plan(N).C; }
QS_HD constexpr int qp_warp_lanes(int N, int C) { return (N + 1 + C - 1) / C; }
// Exchange areas of the warp scans.  Affine maps (M 16, d 4): a dedicated area behind the records, odd stride 21 per
// lane.  Scan elements (A 16, C 10, J 10): for C >= 2 they travel through rows R_K.. of the j = 0 (A) and j = 1
// (C, J) records, which are dead during the element scan; for C == 1 through a dedicated area of stride 37.
constexpr int QW_XA = 21, QW_XE = 37;
QS_HD constexpr size_t qp_warp_smem_doubles(int N, int C_plan = 0) {   // per problem; C_plan = 0: chunk of qp_warp_plan(N)
    const int C = C_plan ? C_plan : qp_warp_chunk(N), L = qp_warp_lanes(N, C);
    return (size_t)qw_rows(C) * C * L + (size_t)QW_XA * L + (C >= 2 ? 0 : (size_t)QW_XE * L);
}

// ---- small dense helpers --------------------------------------------------------------------------
// Cholesky of a packed symmetric positive SEMI-definite 4x4 (lower, LT indexing); non-positive pivots give
// a zero column.  id[i] = 1/L[i][i] (0 for dropped pivots).
QS_HD void chol4_psd(const double A[10], double Lo[10], double id[4]) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        double d = A[LT(j, j)];
#pragma unroll
        for (int p = 0; p < j; ++p) d = fma(-Lo[LT(j, p)], Lo[LT(j, p)], d);
        const bool pos = d > 1e-14 * fabs(A[LT(j, j)]) && d > 0.0;
        const double inv = pos ? qs_rsqrt(d) : 0.0;
        Lo[LT(j, j)] = pos ? d * inv : 0.0;
        id[j] = inv;
#pragma unroll
        for (int i = j + 1; i < 4; ++i) {
            double a = A[LT(i, j)];
#pragma unroll
            for (int p = 0; p < j; ++p) a = fma(-Lo[LT(i, p)], Lo[LT(j, p)], a);
            Lo[LT(i, j)] = a * inv;
        }
    }
}

struct Elem { double A[16]; double C[10]; double J[10]; };   // A row-major, C and J packed lower

// Exchange through shared memory instead of 64-bit shuffles (2 instructions per value instead of ~8):
// every lane deposits n values in its column of the exchange area, the partner column is read after a
// warp sync.  Lanes outside [0, L) neither write nor read.
template <class Ctx>
QS_HD void xch_put(const Ctx& w, double* __restrict__ slot, bool on, const double* v, int n) {
    if (on) {
#pragma unroll
        for (int i = 0; i < n; ++i) slot[i] = v[i];
    }
    (void)w;
}

QS_HD void elem_identity(Elem& e) {
#pragma unroll
    for (int i = 0; i < 16; ++i) e.A[i] = (i % 5 == 0) ? 1.0 : 0.0;
#pragma unroll
    for (int i = 0; i < 10; ++i) { e.C[i] = 0.0; e.J[i] = 0.0; }
}

// e1 <- e1 (x) e2 : e1 covers the earlier interval (i -> j), (A2, C2, J2) the later one (j -> k).
// JONLY: only J of the result is formed (last step of the suffix scan: A and C of the aggregate are never read again;
// A2 and C2 are not referenced)
template <bool JONLY = false>
QS_HD void elem_combine(Elem& e1, const double A2[16], const double C2[10], const double J2[10]) {
    double L2[10], i2[4];
    chol4_psd(J2, L2, i2);
    // T = C1 L2 (4x4), L2 lower triangular: T[i][j] = sum_{p>=j} C1[i][p] L2[p][j]
    double T[16];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            double a = 0.0;
#pragma unroll
            for (int p = j; p < 4; ++p) a = fma(e1.C[LT(i, p)], L2[LT(p, j)], a);
            T[4 * i + j] = a;
        }
    // W = I + L2' T (symmetric positive definite, eigenvalues >= 1)
    double W[10], Lw[10], iw[4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j <= i; ++j) {
            double a = (i == j) ? 1.0 : 0.0;
#pragma unroll
            for (int p = i; p < 4; ++p) a = fma(L2[LT(p, i)], T[4 * p + j], a);
            W[LT(i, j)] = a;
        }
    chol4_psd(W, Lw, iw);
    // Z = Lw^-1 (L2' A1)   and   Y = Lw^-1 T'   (forward substitutions, column by column)
    double Z[16], Y[16];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        double g[4], t[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            double a = 0.0;
#pragma unroll
            for (int p = i; p < 4; ++p) a = fma(L2[LT(p, i)], e1.A[4 * p + c], a);
            g[i] = a;
            t[i] = T[4 * c + i];                               // column c of T' = row c of T
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            double a = g[i], b = t[i];
#pragma unroll
            for (int p = 0; p < i; ++p) { a = fma(-Lw[LT(i, p)], Z[4 * p + c], a); if constexpr (!JONLY) b = fma(-Lw[LT(i, p)], Y[4 * p + c], b); }
            Z[4 * i + c] = a * iw[i];
            Y[4 * i + c] = b * iw[i];
        }
    }
    if constexpr (JONLY) {                                     // J = Z'Z + J1 (same operation order as below)
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j <= i; ++j) {
                double b = e1.J[LT(i, j)];
#pragma unroll
                for (int p = 0; p < 4; ++p) b = fma(Z[4 * p + i], Z[4 * p + j], b);
                e1.J[LT(i, j)] = b;
            }
        return;
    }
    // FA = A1 - Y'Z ; FC = C1 - Y'Y (sym) ; J = Z'Z + J1
    double FA[16], FC[10];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            double a = e1.A[4 * i + j];
#pragma unroll
            for (int p = 0; p < 4; ++p) a = fma(-Y[4 * p + i], Z[4 * p + j], a);
            FA[4 * i + j] = a;
        }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j <= i; ++j) {
            double a = e1.C[LT(i, j)], b = e1.J[LT(i, j)];
#pragma unroll
            for (int p = 0; p < 4; ++p) { a = fma(-Y[4 * p + i], Y[4 * p + j], a); b = fma(Z[4 * p + i], Z[4 * p + j], b); }
            FC[LT(i, j)] = a;
            e1.J[LT(i, j)] = b;
        }
    // A = A2 FA ; C = A2 FC A2' + C2
    double AF[16];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            double a = 0.0, b = 0.0;
#pragma unroll
            for (int p = 0; p < 4; ++p) { a = fma(A2[4 * i + p], FA[4 * p + j], a); b = fma(A2[4 * i + p], FC[LT(p, j)], b); }
            e1.A[4 * i + j] = a;
            AF[4 * i + j] = b;                                 // A2 FC
        }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j <= i; ++j) {
            double a = C2[LT(i, j)];
#pragma unroll
            for (int p = 0; p < 4; ++p) a = fma(AF[4 * i + p], A2[4 * j + p], a);
            e1.C[LT(i, j)] = a;
        }
}

// e <- e (x) E for a SINGLE-STAGE element e on the left: its C is the rank-2 matrix Bh Bh' (Bh = B Lr^-T, columns bh0, bh1),
// so (I + C1 J2)^-1 = I - Bh (I2 + Bh' J2 Bh)^-1 Bh' J2 (Woodbury) needs one 2x2 Cholesky instead of the two 4x4 ones of the
// general combine.  With G = J2 Bh, S = I2 + Bh'G = Ls Ls', Y = Ls^-1 G', V = Ls^-1 Bh':
//   A = A2 (A1 - V'(Y A1)),   C = (A2 V')(A2 V')' + C2,   J = A1' (J2 - Y'Y) A1 + J1.
// S >= I in exact arithmetic; C and J stay symmetric PSD by construction like in elem_combine.
// 372 instead of 620 fused multiply-adds and 2 instead of 8 reciprocal square roots on the dependent chain.
// In: e.A = A1, e.J = J1 (e.C is not read).  Out: e = the aggregate.
QS_HD void elem_stage_combine(Elem& e, const double bh0[4], const double bh1[4], const double A2[16], const double C2[10], const double J2[10]) {
    double G0[4], G1[4];
    sym4_mul(J2, bh0, G0); sym4_mul(J2, bh1, G1);
    const double s00 = 1.0 + dot4(bh0, G0), s10 = dot4(bh1, G0), s11 = 1.0 + dot4(bh1, G1);
    // (J2 is PSD up to rounding; with the coupled rows of h_variant 1 its (s, s) entry is the difference of two barrier terms of
    // ~1e24 and can come out as negative noise: a non-positive pivot is dropped like in chol4_psd)
    const bool p0 = s00 > 0.0;
    const double i00 = p0 ? qs_rsqrt(s00) : 0.0, l10 = s10 * i00, d11 = s11 - l10 * l10;
    const double i11 = d11 > 0.0 ? qs_rsqrt(d11) : 0.0;
    double y0[4], y1[4], v0[4], v1[4], t0[4], t1[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        y0[i] = G0[i] * i00; y1[i] = (G1[i] - l10 * y0[i]) * i11;
        v0[i] = bh0[i] * i00; v1[i] = (bh1[i] - l10 * v0[i]) * i11;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        double a = 0.0, b = 0.0;
#pragma unroll
        for (int p = 0; p < 4; ++p) { a = fma(y0[p], e.A[4 * p + j], a); b = fma(y1[p], e.A[4 * p + j], b); }
        t0[j] = a; t1[j] = b;
    }
    // X = J2 - Y'Y ;  J = A1' X A1 + J1   (A1 is still needed: J first)
    double X[10], XA[16];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j <= i; ++j) X[LT(i, j)] = J2[LT(i, j)] - fma(y0[i], y0[j], y1[i] * y1[j]);
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            double a = 0.0;
#pragma unroll
            for (int p = 0; p < 4; ++p) a = fma(X[LT(i, p)], e.A[4 * p + j], a);
            XA[4 * i + j] = a;
        }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j <= i; ++j) {
            double a = e.J[LT(i, j)];
#pragma unroll
            for (int p = 0; p < 4; ++p) a = fma(e.A[4 * p + i], XA[4 * p + j], a);
            e.J[LT(i, j)] = a;
        }
    // FA = A1 - V'(Y A1) ;  A = A2 FA
    double FA[16];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) FA[4 * i + j] = e.A[4 * i + j] - fma(v0[i], t0[j], v1[i] * t1[j]);
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            double a = 0.0;
#pragma unroll
            for (int p = 0; p < 4; ++p) a = fma(A2[4 * i + p], FA[4 * p + j], a);
            e.A[4 * i + j] = a;
        }
    // C = (A2 V')(A2 V')' + C2
    double w0[4], w1[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        double a = 0.0, b = 0.0;
#pragma unroll
        for (int p = 0; p < 4; ++p) { a = fma(A2[4 * i + p], v0[p], a); b = fma(A2[4 * i + p], v1[p], b); }
        w0[i] = a; w1[i] = b;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j <= i; ++j) e.C[LT(i, j)] = C2[LT(i, j)] + fma(w0[i], w0[j], w1[i] * w1[j]);
}

// affine maps v -> M v + d ; own <- own o partner  (own is applied AFTER partner)
QS_HD void aff_compose(double M[16], double d[4], const double Mp[16], const double dp[4]) {
    double Mn[16], dn[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        double a = d[i];
#pragma unroll
        for (int p = 0; p < 4; ++p) a = fma(M[4 * i + p], dp[p], a);
        dn[i] = a;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            double b = 0.0;
#pragma unroll
            for (int p = 0; p < 4; ++p) b = fma(M[4 * i + p], Mp[4 * p + j], b);
            Mn[4 * i + j] = b;
        }
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) M[i] = Mn[i];
#pragma unroll
    for (int i = 0; i < 4; ++i) d[i] = dn[i];
}

// vector part of aff_compose only: d <- M dp + d (same operation order)
QS_HD void aff_apply(const double M[16], double d[4], const double dp[4]) {
    double dn[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        double a = d[i];
#pragma unroll
        for (int p = 0; p < 4; ++p) a = fma(M[4 * i + p], dp[p], a);
        dn[i] = a;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) d[i] = dn[i];
}

// M <- M Mp (4x4, row-major)
QS_HD void mat4_mul(double M[16], const double Mp[16]) {
    double Mn[16];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            double b = 0.0;
#pragma unroll
            for (int p = 0; p < 4; ++p) b = fma(M[4 * i + p], Mp[4 * p + j], b);
            Mn[4 * i + j] = b;
        }
#pragma unroll
    for (int i = 0; i < 16; ++i) M[i] = Mn[i];
}

// closed-loop transition Abar = A - B K of a stage (row-major), A = [e1 e2 a3 a4]
QS_HD void closed_loop(const StageLin& L, const double K0[4], const double K1[4], double Ab[16]) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const double a = (j == 0) ? (i == 0 ? 1.0 : 0.0) : (j == 1) ? (i == 1 ? 1.0 : 0.0) : (j == 2 ? L.a3[i] : L.a4[i]);
            Ab[4 * i + j] = a - fma(L.b1[i], K0[j], L.b2[i] * K1[j]);
        }
}

// Ratio test without divisions: keep the pair (num, den) with the smallest num/den among the candidates
// v / (-dv) with dv < 0 (cross-multiplication; one division per lane at the end instead of one per candidate).
QS_HD void ratio_min(double v, double dv, double& num, double& den) {
    const double dn = -dv;
    const bool better = (dn > 0.0) && (v * den < num * dn);
    num = better ? v : num;
    den = better ? dn : den;
}

// Newton step of one two-sided bound pair (lower: t = v - dl, upper: t = du - v), Mehrotra corrector included.
struct IneqStep { double dtl, dtu, dll, dlu; };
QS_HD IneqStep ineq_step(double v, double dva, double dv, double ll, double lu, double tl, double tu, double itl, double itu, double dl, double du, double smu, double t_min) {
    const double rdl = v - dl - tl, rdu = du - v - tu;
    const double dtal = dva + rdl, dtau = -dva + rdu;
    const double cl = (-ll - ll * dtal * itl) * dtal, cu = (-lu - lu * dtau * itu) * dtau;
    IneqStep s;
    s.dtl = dv + rdl; s.dtu = -dv + rdu;
    s.dll = -(ll * tl - fmax(smu, ll * t_min) + cl + ll * s.dtl) * itl;      // centering target of a pair: max(sigma mu, lam t_min)
    s.dlu = -(lu * tu - fmax(smu, lu * t_min) + cu + lu * s.dtu) * itu;
    return s;
}

#define QW_SM(row, j) sm[((size_t)(j) * Lw_ + lane) * qw_rows(C) + (row)]

// A_k (columns 3, 4) and B_k of the local stage whose TMEM block starts at `base` (warp-collective: every lane calls it)
template <class Ctx>
QS_HD void qw_ld_lin(const Ctx& w, int base, StageLin& L) {
    double v[16];
    w.template tm_ld<16>(base + QW_TM_AB, v);
#pragma unroll
    for (int i = 0; i < 4; ++i) { L.a3[i] = v[i]; L.a4[i] = v[4 + i]; L.b1[i] = v[8 + i]; L.b2[i] = v[12 + i]; }
}
// 1/t_l (0..2) and 1/t_u (4..6) of local stage j, as left by phase (2) of the current iteration
template <class Ctx>
QS_HD void qw_ld_it(const Ctx& w, int base, double* it8) { w.template tm_ld<8>(base + QW_TM_IT, it8); }
// h_k (3 values) at the linearisation point and beta_k = v_bound'(s_k) (0 unless h_variant 1): 4 doubles
template <class Ctx>
QS_HD void qw_ld_h(const Ctx& w, int base, double* h4) { w.template tm_ld<4>(base + QW_TM_HH, h4); }
// h_k / beta_k (4) and the slack reciprocals (8) with ONE load of the [h | 1/t | D] group
template <class Ctx>
QS_HD void qw_ld_hit(const Ctx& w, int base, double* h4, double* it8) {
    double v[16];
    w.template tm_ld<16>(base + QW_TM_HH, v);
#pragma unroll
    for (int i = 0; i < 4; ++i) h4[i] = v[i];
#pragma unroll
    for (int i = 0; i < 8; ++i) it8[i] = v[4 + i];
}
// One Newton solve with the current factorisation: rhs gt (R_GT rows) and r_b (R_RB) ->
// step dz (R_GT rows, aliased), costate offsets p_k (R_PV), feed-forward k_ff (R_KFF).
template <class Ctx, int C, int SEG>
QS_HD void qp_warp_solve(const Ctx& w, double* __restrict__ sm, int N, int Lw_, bool live, int pass) {
    const int lane = w.lane() & (SEG - 1);
    const bool act = lane < Lw_ && live;
    // The matrix parts of the two affine recursions depend on the factorisation only, not on the right-hand side: with TMEM room
    // for one more 4x4 per lane (C <= 3) the chunk's composed backward matrix M = Abar_k0' ... Abar_k1' (stages k < N) is built by
    // the predictor pass (pass 0) only and kept there; the corrector reloads it, and the forward chunk matrix of either pass is
    // its transpose.  The vector parts are rolled through the stages directly (Abar' p = A'p - K'(B'p), forward_stage).  The
    // terminal stage leaves M alone: a chunk matrix that reaches the terminal stage is never applied to a vector by the scan.
    constexpr bool KEEP = (C <= 3);
    constexpr int TM_M = C * qw_tm_stage(C);
    // ---- (a) local: d_k, kff0_k and the chunk's composed backward map  p_start = M p_end + d
    double M[16], d[4];
    bool acc_identity = true;
#pragma unroll
    for (int i = 0; i < 16; ++i) M[i] = (i % 5 == 0) ? 1.0 : 0.0;
#pragma unroll
    for (int i = 0; i < 4; ++i) d[i] = 0.0;
    {
#pragma unroll 1
        for (int j = C - 1; j >= 0; --j) {
            const int k = lane * C + j;
            StageLin L;
            double Pst[16];
            qw_ld_lin(w, j * qw_tm_stage(C), L);
            if constexpr (C >= 3) w.template tm_ld<16>(j * qw_tm_stage(C) + QW_TM_P, Pst);
            if (!act || k > N) continue;
            if (k == N) {                                       // terminal: p_N = rg_N (constant map)
                if constexpr (!KEEP) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) M[i] = 0.0;
                    acc_identity = false;
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) { d[i] = QW_SM(R_RG + 2 + i, j); QW_SM(R_PV + i, j) = d[i]; }
                continue;
            }
            double gt[6], Pb[4], K0[4], K1[4], Li[3], m[6];
#pragma unroll
            for (int i = 0; i < 6; ++i) gt[i] = QW_SM(R_GT + i, j);
#pragma unroll
            for (int i = 0; i < 4; ++i) { if constexpr (C >= 3) Pb[i] = Pst[10 + i]; else Pb[i] = QW_SM(R_PB + i, j); K0[i] = QW_SM(R_K + i, j); K1[i] = QW_SM(R_K + 4 + i, j); }
#pragma unroll
            for (int i = 0; i < 3; ++i) Li[i] = QW_SM(R_LI + i, j);
            lin_T_mul_add(L, Pb, gt, m);
            const double y0 = m[0] * Li[0], y1 = (m[1] - Li[1] * y0) * Li[2];
            const double k1 = y1 * Li[2], k0 = (y0 - Li[1] * k1) * Li[0];
            QW_SM(R_KFF, j) = k0; QW_SM(R_KFF + 1, j) = k1;
            double dk[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) { dk[i] = m[2 + i] - fma(K0[i], m[0], K1[i] * m[1]); QW_SM(R_PV + i, j) = dk[i]; }
            if constexpr (KEEP) {
                // vector part of acc <- f_k o acc:  d <- Abar_k' d + d_k
                {
                    const double m0 = dot4(L.b1, d), m1 = dot4(L.b2, d);
                    double dn[4];
                    dn[0] = dk[0] + d[0] - fma(K0[0], m0, K1[0] * m1);
                    dn[1] = dk[1] + d[1] - fma(K0[1], m0, K1[1] * m1);
                    dn[2] = dk[2] + dot4(L.a3, d) - fma(K0[2], m0, K1[2] * m1);
                    dn[3] = dk[3] + dot4(L.a4, d) - fma(K0[3], m0, K1[3] * m1);
#pragma unroll
                    for (int i = 0; i < 4; ++i) d[i] = dn[i];
                }
                if (pass == 0) {                                // matrix part: M <- Abar_k' M
                    double Ab[16], At[16];
                    closed_loop(L, K0, K1, Ab);
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int q = 0; q < 4; ++q) At[4 * i + q] = Ab[4 * q + i];
                    if (!acc_identity) mat4_mul(At, M);
                    acc_identity = false;
#pragma unroll
                    for (int i = 0; i < 16; ++i) M[i] = At[i];
                }
            } else {
                double Ab[16], At[16];
                closed_loop(L, K0, K1, Ab);
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int q = 0; q < 4; ++q) At[4 * i + q] = Ab[4 * q + i];
                // acc <- f_k o acc
                if (!acc_identity) aff_compose(At, dk, M, d);
                acc_identity = false;
#pragma unroll
                for (int i = 0; i < 16; ++i) M[i] = At[i];
#pragma unroll
                for (int i = 0; i < 4; ++i) d[i] = dk[i];
            }
        }
    }
    if constexpr (KEEP) {                                       // warp-collective TMEM accesses: every lane, outside divergent code
        if (pass == 0) w.tm_st16(TM_M, M);
        else w.template tm_ld<16>(TM_M, M);
    }
    // ---- (b) suffix scan over lanes (exchange through shared memory); afterwards d = p at the first stage of the chunk
    double* xa = sm + (size_t)qw_rows(C) * C * Lw_ + (size_t)lane * QW_XA;      // this lane's affine exchange slot
#pragma unroll 1
    for (int dl = 1; dl < Lw_; dl <<= 1) {
        const bool last = (dl << 1) >= Lw_;                    // after the last step only d is read
        w.sync();
        if (!last) xch_put(w, xa, act, M, 16);
        xch_put(w, xa + 16, act, d, 4);
        w.sync();
        if (act && lane + dl < Lw_) {
            double Mp[16], dp[4];
            const double* xp = xa + (size_t)dl * QW_XA;
#pragma unroll
            for (int i = 0; i < 4; ++i) dp[i] = xp[16 + i];
            if (last) {
                aff_apply(M, d, dp);
            } else {
#pragma unroll
                for (int i = 0; i < 16; ++i) Mp[i] = xp[i];
                aff_compose(M, d, Mp, dp);
            }
        }
    }
    w.sync();
    xch_put(w, xa + 16, act, d, 4);
    w.sync();
    double pe[4] = {0, 0, 0, 0};                                // p at the right boundary of the chunk
    if (act && lane + 1 < Lw_) {
#pragma unroll
        for (int i = 0; i < 4; ++i) pe[i] = xa[QW_XA + 16 + i];
    }
    // ---- (c) local back-substitution: k_ff and p_k
    {
#pragma unroll 1
        for (int j = C - 1; j >= 0; --j) {
            const int k = lane * C + j;
            StageLin L;
            qw_ld_lin(w, j * qw_tm_stage(C), L);
            if (!act || k > N) continue;
            if (k == N) {
#pragma unroll
                for (int i = 0; i < 4; ++i) pe[i] = QW_SM(R_PV + i, j);
                continue;
            }
            double K0[4], K1[4], Li[3];
#pragma unroll
            for (int i = 0; i < 4; ++i) { K0[i] = QW_SM(R_K + i, j); K1[i] = QW_SM(R_K + 4 + i, j); }
#pragma unroll
            for (int i = 0; i < 3; ++i) Li[i] = QW_SM(R_LI + i, j);
            const double m0 = dot4(L.b1, pe), m1 = dot4(L.b2, pe);
            const double y0 = m0 * Li[0], y1 = (m1 - Li[1] * y0) * Li[2];
            const double k1 = y1 * Li[2], k0 = (y0 - Li[1] * k1) * Li[0];
            QW_SM(R_KFF, j) += k0; QW_SM(R_KFF + 1, j) += k1;
            // p_k = Abar' p_{k+1} + d_k = A'p - K'(B'p) + d_k
            double pk[4];
            pk[0] = QW_SM(R_PV + 0, j) + pe[0] - fma(K0[0], m0, K1[0] * m1);
            pk[1] = QW_SM(R_PV + 1, j) + pe[1] - fma(K0[1], m0, K1[1] * m1);
            pk[2] = QW_SM(R_PV + 2, j) + dot4(L.a3, pe) - fma(K0[2], m0, K1[2] * m1);
            pk[3] = QW_SM(R_PV + 3, j) + dot4(L.a4, pe) - fma(K0[3], m0, K1[3] * m1);
#pragma unroll
            for (int i = 0; i < 4; ++i) { QW_SM(R_PV + i, j) = pk[i]; pe[i] = pk[i]; }
        }
    }
    // ---- (d) forward: chunk's composed map dx_end = M dx_start + d, prefix scan, local rollout
    acc_identity = true;
#pragma unroll
    for (int i = 0; i < 4; ++i) d[i] = 0.0;
    if constexpr (KEEP) {
        // forward chunk matrix Abar_k1 ... Abar_k0 = (backward chunk matrix)'
        double Mb[16];
        w.template tm_ld<16>(TM_M, Mb);
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int q = 0; q < 4; ++q) M[4 * i + q] = Mb[4 * q + i];
#pragma unroll 1
        for (int j = 0; j < C; ++j) {
            const int k = lane * C + j;
            StageLin L;
            qw_ld_lin(w, j * qw_tm_stage(C), L);
            if (!act || k >= N) continue;
            double K0[4], K1[4], bk[4], kff[2], u[2];
#pragma unroll
            for (int i = 0; i < 4; ++i) { K0[i] = QW_SM(R_K + i, j); K1[i] = QW_SM(R_K + 4 + i, j); bk[i] = QW_SM(R_RB + i, j); }
            kff[0] = QW_SM(R_KFF, j); kff[1] = QW_SM(R_KFF + 1, j);
            forward_stage(L, bk, K0, K1, kff, d, u);            // d <- Abar_k d + (r_b - B k_ff)
        }
    } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) M[i] = (i % 5 == 0) ? 1.0 : 0.0;
#pragma unroll 1
        for (int j = 0; j < C; ++j) {
            const int k = lane * C + j;
            StageLin L;
            qw_ld_lin(w, j * qw_tm_stage(C), L);
            if (!act || k >= N) continue;
            double K0[4], K1[4], Ab[16], bb[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) { K0[i] = QW_SM(R_K + i, j); K1[i] = QW_SM(R_K + 4 + i, j); }
            const double f0 = QW_SM(R_KFF, j), f1 = QW_SM(R_KFF + 1, j);
#pragma unroll
            for (int i = 0; i < 4; ++i) bb[i] = QW_SM(R_RB + i, j) - fma(L.b1[i], f0, L.b2[i] * f1);
            closed_loop(L, K0, K1, Ab);
            if (!acc_identity) aff_compose(Ab, bb, M, d);
            acc_identity = false;
#pragma unroll
            for (int i = 0; i < 16; ++i) M[i] = Ab[i];
#pragma unroll
            for (int i = 0; i < 4; ++i) d[i] = bb[i];
        }
    }
#pragma unroll 1
    for (int dl = 1; dl < Lw_; dl <<= 1) {
        const bool last = (dl << 1) >= Lw_;                    // after the last step only d is read
        w.sync();
        if (!last) xch_put(w, xa, act, M, 16);
        xch_put(w, xa + 16, act, d, 4);
        w.sync();
        if (act && lane - dl >= 0) {
            double Mp[16], dp[4];
            const double* xp = xa - (size_t)dl * QW_XA;
#pragma unroll
            for (int i = 0; i < 4; ++i) dp[i] = xp[16 + i];
            if (last) {
                aff_apply(M, d, dp);
            } else {
#pragma unroll
                for (int i = 0; i < 16; ++i) Mp[i] = xp[i];
                aff_compose(M, d, Mp, dp);
            }
        }
    }
    w.sync();
    xch_put(w, xa + 16, act, d, 4);
    w.sync();
    double x[4] = {0, 0, 0, 0};                                // dx at the first stage of the chunk (dx_0 = 0)
    if (act && lane >= 1) {
#pragma unroll
        for (int i = 0; i < 4; ++i) x[i] = xa[16 + i - QW_XA];
    }
    {
#pragma unroll 1
        for (int j = 0; j < C; ++j) {
            const int k = lane * C + j;
            StageLin L;
            qw_ld_lin(w, j * qw_tm_stage(C), L);
            if (!act || k > N) continue;
            if (k == N) {
                QW_SM(R_GT + 0, j) = 0.0; QW_SM(R_GT + 1, j) = 0.0;
#pragma unroll
                for (int i = 0; i < 4; ++i) QW_SM(R_GT + 2 + i, j) = x[i];
                continue;
            }
            double K0[4], K1[4], bk[4], kff[2], u[2];
#pragma unroll
            for (int i = 0; i < 4; ++i) { K0[i] = QW_SM(R_K + i, j); K1[i] = QW_SM(R_K + 4 + i, j); bk[i] = QW_SM(R_RB + i, j); }
            kff[0] = QW_SM(R_KFF, j); kff[1] = QW_SM(R_KFF + 1, j);
            const double xk[4] = {x[0], x[1], x[2], x[3]};
            forward_stage(L, bk, K0, K1, kff, x, u);
            QW_SM(R_GT + 0, j) = u[0]; QW_SM(R_GT + 1, j) = u[1];
#pragma unroll
            for (int i = 0; i < 4; ++i) QW_SM(R_GT + 2 + i, j) = xk[i];
        }
    }
}

// Per-problem IPM state kept in registers across iterations (everything else lives in shared memory).
struct QwState {
    bool fin;                   // this segment's problem is finished (or the segment is idle): its lanes only take part in collectives
    int it, stall, status;
    double rmax_prev, r_stat, r_eq, r_in, r_cp;
    double qN[4];
};

// ---- load the linearisation of problem V into the warp's shared memory, initial point
template <class Ctx, int C, int HV, int SEG>
QS_HD void qw_init(const Ctx& w, double* __restrict__ sm, const QpConst& Q, const QpView& V, QwState& st, bool live) {
    const int N = Q.N;
    const int lane = w.lane() & (SEG - 1);
    const int Lw_ = qp_warp_lanes(N, C);
    const bool act = lane < Lw_ && live;
    // ---------------- load the linearisation (into the lane's TMEM block), initial point
#pragma unroll 1
    for (int j = 0; j < C; ++j) {
        const int k = lane * C + j;
        const bool on = act && k < N;
        double v[48];
#pragma unroll
        for (int i = 0; i < 48; ++i) v[i] = 0.0;
        if (act && k <= N) {
#pragma unroll
            for (int i = 0; i < 6; ++i) QW_SM(R_Z + i, j) = 0.0;
#pragma unroll
            for (int i = 0; i < 4; ++i) QW_SM(R_PIK + i, j) = 0.0;
        }
        if (on) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                v[QW_TM_AB + i] = QS_AT(V.A, k, 8, i); v[QW_TM_AB + 4 + i] = QS_AT(V.A, k, 8, 4 + i);
                v[QW_TM_AB + 8 + i] = QS_AT(V.B, k, 8, i); v[QW_TM_AB + 12 + i] = QS_AT(V.B, k, 8, 4 + i);
                v[QW_TM_BV + i] = QS_AT(V.b, k, 4, i);
            }
#pragma unroll
            for (int i = 0; i < 6; ++i) v[QW_TM_G + i] = QS_AT(V.g, k, 6, i);
#pragma unroll
            for (int i = 0; i < 4; ++i) v[QW_TM_HH + i] = QS_AT(V.hv, k, 4, i);     // h_k (3) and v_bound'(s_k)
        }
        w.tm_st16(j * qw_tm_stage(C), v);                      // warp-collective: every lane stores (zeros when idle)
        w.tm_st16(j * qw_tm_stage(C) + 16, v + 16);
        w.tm_st16(j * qw_tm_stage(C) + 32, v + 32);
        if (on) {
            const double* h = v + QW_TM_HH;
            if (k == 0) {
#pragma unroll
                for (int i = 0; i < 4; ++i) QW_SM(R_Z + 2 + i, j) = V.dx0[i * V.stride];
            }
            double z6[6];
#pragma unroll
            for (int i = 0; i < 6; ++i) z6[i] = QW_SM(R_Z + i, j);
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const bool bon = h_on(HV, k, c);
                const double z = qw_row(HV, c, h[3], z6);
                double tl = fmax(z - (Q.lh[c] - h[c]), Q.thr), tu = fmax((Q.uh[c] - h[c]) - z, Q.thr);
                double ll = Q.mu0 / tl, lu = Q.mu0 / tu;
                if (!bon) { tl = 1.0; tu = 1.0; ll = 0.0; lu = 0.0; }
                QW_SM(R_T + c, j) = tl; QW_SM(R_T + 3 + c, j) = tu;
                QW_SM(R_LAM + c, j) = ll; QW_SM(R_LAM + 3 + c, j) = lu;
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) st.qN[i] = live ? V.qN[i * V.stride] : 0.0;
    st.fin = !live;
    st.it = 0; st.stall = 0; st.status = 1; st.rmax_prev = 1e300;
    st.r_stat = st.r_eq = st.r_in = st.r_cp = 0.0;
    w.sync();

}

// ---- one IPM iteration: true residuals + stopping tests, factorisation (parallel-in-time), predictor,
// corrector, step.  Returns 0 to continue, 1 when the problem is finished (st.status set).
// With SEG = 16 a warp carries two problems (lanes 0..15 and 16..31, horizons N <= 15): every reduction is taken over
// the segment, a segment whose problem has finished keeps executing the warp-collective instructions with all its
// stores masked (act = false), and the function returns 1 when every segment of the warp is finished.
template <class Ctx, int C, int HV, int SEG>
QS_HD int qw_iterate(const Ctx& w, double* __restrict__ sm, const QpConst& Q, QwState& st) {
    const int N = Q.N;
    const int lane = w.lane() & (SEG - 1);
    const int Lw_ = qp_warp_lanes(N, C);
    bool act = lane < Lw_ && !st.fin;
    constexpr int hvar = HV;                            // constraint set: compile-time, the default set pays nothing for the coupled rows
    const int m_on = hvar ? 6 * N : 6 * N - 2;
    const double t4 = 4.0 * Q.t_min;
    int& status = st.status; int& it = st.it; int& stall = st.stall;
    double& rmax_prev = st.rmax_prev; double& r_stat = st.r_stat; double& r_eq = st.r_eq; double& r_in = st.r_in; double& r_cp = st.r_cp;
    const double* qN = st.qN;
    {
    QW_T0();
    // ================= (1) true residuals =================
    double nx[4] = {0, 0, 0, 0}, npi[4] = {0, 0, 0, 0};    // x and pi of the stage right of the chunk (neighbour lane, j = 0)
    w.sync();
    if (act && lane + 1 < Lw_) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            nx[i] = sm[((size_t)lane + 1) * qw_rows(C) + R_Z + 2 + i];      // record (j = 0, lane + 1)
            npi[i] = sm[((size_t)lane + 1) * qw_rows(C) + R_PIK + i];
        }
    }
    double l_stat = 0.0, l_eq = 0.0, l_in = 0.0, l_cp = 0.0, l_mu = 0.0;
    bool l_nan = false;
    {
#pragma unroll 1
        for (int j = C - 1; j >= 0; --j) {
            const int k = lane * C + j;
            StageLin L;
            double gk[8], hk[4], bv[4];
            qw_ld_lin(w, j * qw_tm_stage(C), L);
            {
                double bg[16];                                  // [b | g] group
                w.template tm_ld<16>(j * qw_tm_stage(C) + QW_TM_BV, bg);
#pragma unroll
                for (int i = 0; i < 4; ++i) bv[i] = bg[i];
#pragma unroll
                for (int i = 0; i < 8; ++i) gk[i] = bg[4 + i];
            }
            w.template tm_ld<4>(j * qw_tm_stage(C) + QW_TM_HH, hk);
            if (!act || k > N) continue;
            double z6[6], pik[4];
#pragma unroll
            for (int i = 0; i < 6; ++i) z6[i] = QW_SM(R_Z + i, j);
#pragma unroll
            for (int i = 0; i < 4; ++i) pik[i] = QW_SM(R_PIK + i, j);
            if (k == N) {
                double xN[4] = {z6[2], z6[3], z6[4], z6[5]}, rg[4];
                sym4_mul(Q.QN, xN, rg);
#pragma unroll
                for (int i = 0; i < 4; ++i) { rg[i] += qN[i] - pik[i]; QW_SM(R_RG + 2 + i, j) = rg[i]; l_stat = fmax(l_stat, fabs(rg[i])); l_nan = l_nan || !(rg[i] == rg[i]); }
            } else {
                const double* Hk = Q.H + (size_t)k * 21;
                double gh[6], rg[6];
#pragma unroll
                for (int i = 0; i < 6; ++i) {
                    double a = gk[i];
#pragma unroll
                    for (int q = 0; q < 6; ++q) a = fma(Hk[LT(i, q)], z6[q], a);
                    gh[i] = a;
                }
                lin_T_mul_add(L, npi, gh, rg);
#pragma unroll
                for (int i = 0; i < 4; ++i) rg[2 + i] -= pik[i];
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    if (!h_on(hvar, k, c)) continue;
                    const double h = hk[c], v = qw_row(hvar, c, hk[3], z6);
                    const double ll = QW_SM(R_LAM + c, j), lu = QW_SM(R_LAM + 3 + c, j), tl = QW_SM(R_T + c, j), tu = QW_SM(R_T + 3 + c, j);
                    qw_row_add(hvar, c, hk[3], lu - ll, rg);
                    const double rdl = v - (Q.lh[c] - h) - tl, rdu = (Q.uh[c] - h) - v - tu;
                    l_in = fmax(l_in, fmax(fabs(rdl), fabs(rdu)));
                    const double pl = tl > t4 ? ll * tl : 0.0, pu = tu > t4 ? lu * tu : 0.0;   // converged active pairs (slack at its floor) leave mu
                    l_cp = fmax(l_cp, fmax(pl, pu));
                    l_mu += pl + pu;
                }
                if (k == 0) { rg[2] = 0.0; rg[3] = 0.0; rg[4] = 0.0; rg[5] = 0.0; }
#pragma unroll
                for (int i = 0; i < 6; ++i) { QW_SM(R_RG + i, j) = rg[i]; l_stat = fmax(l_stat, fabs(rg[i])); l_nan = l_nan || !(rg[i] == rg[i]); }
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    double a = bv[i] + (i < 2 ? z6[2 + i] : 0.0) - nx[i];
                    a = fma(L.a3[i], z6[4], a); a = fma(L.a4[i], z6[5], a);
                    a = fma(L.b1[i], z6[0], a); a = fma(L.b2[i], z6[1], a);
                    QW_SM(R_RB + i, j) = a; l_eq = fmax(l_eq, fabs(a)); l_nan = l_nan || !(a == a);
                }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) { nx[i] = z6[2 + i]; npi[i] = pik[i]; }
        }
    }
    const double mu_sum = w.template wsum<SEG>(l_mu);
    const double mu = mu_sum / (double)m_on;
    {
        const double q_stat = w.template wmax<SEG>(l_stat), q_eq = w.template wmax<SEG>(l_eq), q_in = w.template wmax<SEG>(l_in), q_cp = w.template wmax<SEG>(l_cp);
        const bool q_nan = w.template wany<SEG>(l_nan ? 1 : 0) || !(mu == mu);
        if (!st.fin) {                                       // stopping tests of this segment's problem
            r_stat = q_stat; r_eq = q_eq; r_in = q_in; r_cp = q_cp;
            if (q_nan) { status = 2; st.fin = true; }
            else {
                const int fin_ = qp_stop_test(Q, r_stat, r_eq, r_in, r_cp, it, rmax_prev, stall);
                if (fin_ >= 0) { status = fin_; st.fin = true; }
            }
        }
    }
    if (!w.wany(st.fin ? 0 : 1)) return 1;                  // every segment of the warp is finished
    act = act && !st.fin;
    QW_TICK(1);
    // ================= (2) barrier terms, affine rhs, stage elements, chunk aggregate =================
    Elem E; elem_identity(E);
    bool E_is_identity = true;                              // combining with the identity is a copy: skip the arithmetic
    {
#pragma unroll 1
        for (int j = C - 1; j >= 0; --j) {
            const int k = lane * C + j;
            StageLin L;
            double hk[4];
            qw_ld_lin(w, j * qw_tm_stage(C), L);
            qw_ld_h(w, j * qw_tm_stage(C), hk);
            // barrier terms first: slack reciprocals and D go to the TMEM block (warp-collective store), affine rhs to R_GT
            double it8[8], D[4], Dc[3] = {0.0, 0.0, 0.0};
#pragma unroll
            for (int i = 0; i < 8; ++i) it8[i] = 0.0;
#pragma unroll
            for (int i = 0; i < 4; ++i) D[i] = 0.0;
            if (act && k < N) {
                double gt[6], z6[6];
#pragma unroll
                for (int i = 0; i < 6; ++i) { gt[i] = QW_SM(R_RG + i, j); z6[i] = QW_SM(R_Z + i, j); }
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const double h = hk[c], v = qw_row(hvar, c, hk[3], z6);
                    const double ll = QW_SM(R_LAM + c, j), lu = QW_SM(R_LAM + 3 + c, j), tl = QW_SM(R_T + c, j), tu = QW_SM(R_T + 3 + c, j);
                    const double itl = 1.0 / tl, itu = 1.0 / tu;
                    const bool on = h_on(hvar, k, c);
                    const double rdl = on ? v - (Q.lh[c] - h) - tl : 0.0, rdu = on ? (Q.uh[c] - h) - v - tu : 0.0;
                    it8[c] = itl; it8[4 + c] = itu;
                    Dc[c] = ll * itl + lu * itu;
                    qw_row_add(hvar, c, hk[3], (ll + ll * rdl * itl) - (lu + lu * rdu * itu), gt);
                }
                // barrier Hessian sum_c D_c a_c a_c' as (D_ss, D_unun, D_utut, D_s,ut)
                if (hvar) { const double be = hk[3]; D[0] = be * be * (Dc[1] + Dc[2]); D[1] = Dc[0]; D[2] = Dc[1] + Dc[2]; D[3] = be * (Dc[2] - Dc[1]); }
                else { D[0] = Dc[0]; D[1] = Dc[1]; D[2] = Dc[2]; D[3] = 0.0; }
#pragma unroll
                for (int i = 0; i < 6; ++i) QW_SM(R_GT + i, j) = gt[i];
            }
            {
                double hid[16];                                 // [h | 1/t | D] group in one store (h unchanged)
#pragma unroll
                for (int i = 0; i < 4; ++i) { hid[i] = hk[i]; hid[12 + i] = D[i]; }
#pragma unroll
                for (int i = 0; i < 8; ++i) hid[4 + i] = it8[i];
                w.tm_st16(j * qw_tm_stage(C) + QW_TM_HH, hid);
            }
            if (!act || k > N) continue;
            if (k == N) {                                   // terminal value function: J = Q_N, A = 0, C = 0
#pragma unroll
                for (int i = 0; i < 16; ++i) E.A[i] = 0.0;
#pragma unroll
                for (int i = 0; i < 10; ++i) { E.C[i] = 0.0; E.J[i] = Q.QN[i]; }
                E_is_identity = false;
                continue;
            }
            const double* Hk = Q.H + (size_t)k * 21;
            // element of stage k: eliminate u.  Rt = H_uu + D_u, S = H_ux, Qt = H_xx + D_s
            const double r00 = Hk[LT(0, 0)] + D[1], r10 = Hk[LT(1, 0)], r11 = Hk[LT(1, 1)] + D[2];
            const double i00 = qs_rsqrt(r00), l10 = r10 * i00, i11 = qs_rsqrt(r11 - l10 * l10);
            // Bh = B Lr^-T  (4x2):  columns of B R^-1 B' = Bh Bh'
            double bh0[4], bh1[4], sh0[4], sh1[4];             // Sh = Lr^-1 S (2x4)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                bh0[i] = L.b1[i] * i00; bh1[i] = (L.b2[i] - l10 * bh0[i]) * i11;
                const double s0 = Hk[LT(2 + i, 0)], s1 = Hk[LT(2 + i, 1)] + (i == 3 ? D[3] : 0.0);
                sh0[i] = s0 * i00; sh1[i] = (s1 - l10 * sh0[i]) * i11;
            }
            Elem e;
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const double a = (q == 0) ? (i == 0 ? 1.0 : 0.0) : (q == 1) ? (i == 1 ? 1.0 : 0.0) : (q == 2 ? L.a3[i] : L.a4[i]);
                    e.A[4 * i + q] = a - fma(bh0[i], sh0[q], bh1[i] * sh1[q]);
                }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int q = 0; q <= i; ++q) e.J[LT(i, q)] = Hk[LT(2 + i, 2 + q)] - fma(sh0[i], sh0[q], sh1[i] * sh1[q]);
            if constexpr (hvar) {
                // coupled rows: D_ss - sh1_s^2 = beta^2 (D1 + D2) - (sigma + beta (D2 - D1))^2 / (rho + D1 + D2) is a difference of two
                // barrier terms of up to ~1e24 whose exact value is O(rho); expanded with (D1 + D2)^2 - (D2 - D1)^2 = 4 D1 D2 every large
                // term is positive:  [beta^2 (rho (D1 + D2) + 4 D1 D2) - sigma (sigma + 2 beta (D2 - D1))] / (rho + D1 + D2)
                const double be = hk[3], rho = Hk[LT(1, 1)] - l10 * l10, sig = Hk[LT(5, 1)] - l10 * sh0[3];
                const double num = be * be * fma(rho, Dc[1] + Dc[2], 4.0 * Dc[1] * Dc[2]) - sig * fma(2.0 * be, Dc[2] - Dc[1], sig);
                e.J[LT(3, 3)] = Hk[LT(5, 5)] - sh0[3] * sh0[3] + num * (i11 * i11);
            } else {
                e.J[LT(3, 3)] += D[0];
            }
            // (stage 0 needs no special case: dx_0 = 0 is imposed by the forward scan, the element of stage 0 only feeds lanes left of it)
            if (!E_is_identity) elem_stage_combine(e, bh0, bh1, E.A, E.C, E.J);   // E <- e (x) E, rank-2 form
            else {
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int q = 0; q <= i; ++q) e.C[LT(i, q)] = fma(bh0[i], bh0[q], bh1[i] * bh1[q]);
            }
            E = e;
            E_is_identity = false;
        }
    }
    QW_TICK(2);
    // ================= (3) suffix scan of the chunk aggregates =================
    // exchange slots of this lane: A (16) and C, J (10 + 10)
    double* xeA = (C >= 2) ? &QW_SM(R_K, 0) : sm + (size_t)qw_rows(C) * C * Lw_ + (size_t)QW_XA * Lw_ + (size_t)lane * QW_XE;
    double* xeCJ = (C >= 2) ? &QW_SM(R_K, (C >= 2 ? 1 : 0)) : xeA + 16;
    const size_t xstride = (C >= 2) ? (size_t)qw_rows(C) : (size_t)QW_XE;     // distance between neighbouring lanes' slots
#pragma unroll 1
    for (int dl = 1; dl < Lw_; dl <<= 1) {
        const bool last = (dl << 1) >= Lw_;                    // after the last step only J is read (P at the chunk boundaries)
        w.sync();
        if (!last) { xch_put(w, xeA, act, E.A, 16); xch_put(w, xeCJ, act, E.C, 10); }
        xch_put(w, xeCJ + 10, act, E.J, 10);
        w.sync();
        if (act && lane + dl < Lw_) {
            double A2[16], C2[10], J2[10];
            const double* pA = xeA + (size_t)dl * xstride;
            const double* pCJ = xeCJ + (size_t)dl * xstride;
#pragma unroll
            for (int i = 0; i < 10; ++i) J2[i] = pCJ[10 + i];
            if (last) {
                elem_combine<true>(E, A2, C2, J2);
            } else {
#pragma unroll
                for (int i = 0; i < 16; ++i) A2[i] = pA[i];
#pragma unroll
                for (int i = 0; i < 10; ++i) C2[i] = pCJ[i];
                elem_combine<false>(E, A2, C2, J2);
            }
        }
    }
    w.sync();
    xch_put(w, xeCJ + 10, act, E.J, 10);
    w.sync();
    double P[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};          // P at the right boundary of the chunk
    if (act && lane + 1 < Lw_) {
#pragma unroll
        for (int i = 0; i < 10; ++i) P[i] = xeCJ[xstride + 10 + i];
    }
    QW_TICK(3);
    // ================= (4) local Riccati over the chunk: K_k, Cholesky, P_k, P_{k+1} r_b =================
    bool ok = true;
    {
#pragma unroll 1
        for (int j = C - 1; j >= 0; --j) {
            const int k = lane * C + j;
            StageLin L;
            double D[4];
            qw_ld_lin(w, j * qw_tm_stage(C), L);
            w.template tm_ld<4>(j * qw_tm_stage(C) + QW_TM_D, D);
            double Pst[16];                                 // P_k (10), P_{k+1} r_b (4): to shared memory, or to TMEM on long horizons
#pragma unroll
            for (int i = 0; i < 16; ++i) Pst[i] = 0.0;
            if (act && k <= N) {
                if (k == N) {
#pragma unroll
                    for (int i = 0; i < 10; ++i) P[i] = Q.QN[i];
                } else {
                    double rb[4], Pb[4], K0[4], K1[4], Li[3];
#pragma unroll
                    for (int i = 0; i < 4; ++i) rb[i] = QW_SM(R_RB + i, j);
                    sym4_mul(P, rb, Pb);
                    ok = riccati_factor_stage(L, Q.H + (size_t)k * 21, D, P, K0, K1, Li, D[3]) && ok;
#pragma unroll
                    for (int i = 0; i < 4; ++i) { QW_SM(R_K + i, j) = K0[i]; QW_SM(R_K + 4 + i, j) = K1[i]; Pst[10 + i] = Pb[i]; }
#pragma unroll
                    for (int i = 0; i < 3; ++i) QW_SM(R_LI + i, j) = Li[i];
                }
#pragma unroll
                for (int i = 0; i < 10; ++i) Pst[i] = P[i];
                if constexpr (C < 3) {
#pragma unroll
                    for (int i = 0; i < 10; ++i) QW_SM(R_P + i, j) = Pst[i];
                    if (k < N) {
#pragma unroll
                        for (int i = 0; i < 4; ++i) QW_SM(R_PB + i, j) = Pst[10 + i];
                    }
                }
            }
            if constexpr (C >= 3) w.tm_st16(j * qw_tm_stage(C) + QW_TM_P, Pst);     // warp-collective
        }
    }
    if (w.template wany<SEG>(ok ? 0 : 1) && !st.fin) { status = 2; st.fin = true; }
    if (!w.wany(st.fin ? 0 : 1)) return 1;
    act = act && !st.fin;
    QW_TICK(4);
    // ================= (5)-(7) predictor and corrector share ONE copy of the solve code =================
    double smu = 0.0;
#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass) {
        if (pass == 1) {
            // corrector rhs
#pragma unroll 1
            for (int j = 0; j < C; ++j) {
                const int k = lane * C + j;
                double hk[4], it8[8];
                qw_ld_hit(w, j * qw_tm_stage(C), hk, it8);
                if (!act || k >= N) continue;
                double gt[6], z6[6];
#pragma unroll
                for (int i = 0; i < 6; ++i) { gt[i] = QW_SM(R_RG + i, j); z6[i] = QW_SM(R_Z + i, j); }
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    if (!h_on(hvar, k, c)) continue;
                    const double h = hk[c], v = qw_row(hvar, c, hk[3], z6), dva = QW_SM(R_DZA + c, j);
                    const double ll = QW_SM(R_LAM + c, j), lu = QW_SM(R_LAM + 3 + c, j), tl = QW_SM(R_T + c, j), tu = QW_SM(R_T + 3 + c, j);
                    const double rdl = v - (Q.lh[c] - h) - tl, rdu = (Q.uh[c] - h) - v - tu;
                    const double dtl = dva + rdl, dtu = -dva + rdu;
                    const double itl = it8[c], itu = it8[4 + c];
                    const double cl = (-ll - ll * dtl * itl) * dtl, cu = (-lu - lu * dtu * itu) * dtu;
                    qw_row_add(hvar, c, hk[3], (ll * tl - fmax(smu, ll * Q.t_min) + cl + ll * rdl) * itl - (lu * tu - fmax(smu, lu * Q.t_min) + cu + lu * rdu) * itu, gt);
                }
#pragma unroll
                for (int i = 0; i < 6; ++i) QW_SM(R_GT + i, j) = gt[i];
            }
        }
        QW_TICK(5);
        qp_warp_solve<Ctx, C, SEG>(w, sm, N, Lw_, !st.fin, pass);
        QW_TICK(6);
        if (pass == 0) {
            // step to the boundary of the affine step, mu_aff, centering parameter
            double a_aff = 1.0, a_num = 1.0, a_den = 1.0, S1 = 0.0, S2 = 0.0;
            {
#pragma unroll 1
                for (int j = 0; j < C; ++j) {
                    const int k = lane * C + j;
                    double hk[4], it8[8];
                    qw_ld_hit(w, j * qw_tm_stage(C), hk, it8);
                    if (!act || k >= N) continue;
                    double z6[6], dz6[6];
#pragma unroll
                    for (int i = 0; i < 6; ++i) { z6[i] = QW_SM(R_Z + i, j); dz6[i] = QW_SM(R_GT + i, j); }
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        const double dva = qw_row(hvar, c, hk[3], dz6);
                        QW_SM(R_DZA + c, j) = dva;
                        if (!h_on(hvar, k, c)) continue;
                        const double h = hk[c], v = qw_row(hvar, c, hk[3], z6);
                        const double ll = QW_SM(R_LAM + c, j), lu = QW_SM(R_LAM + 3 + c, j), tl = QW_SM(R_T + c, j), tu = QW_SM(R_T + 3 + c, j);
                        const double dtl = dva + (v - (Q.lh[c] - h) - tl), dtu = -dva + ((Q.uh[c] - h) - v - tu);
                        const double dll = -ll - ll * dtl * it8[c], dlu = -lu - lu * dtu * it8[4 + c];
                        ratio_min(tl, dtl, a_num, a_den); ratio_min(tu, dtu, a_num, a_den);
                        ratio_min(ll, dll, a_num, a_den); ratio_min(lu, dlu, a_num, a_den);
                        const double wl = tl > t4 ? 1.0 : 0.0, wu = tu > t4 ? 1.0 : 0.0;
                        S1 += wl * (ll * dtl + tl * dll) + wu * (lu * dtu + tu * dlu);
                        S2 += wl * (dll * dtl) + wu * (dlu * dtu);
                    }
                }
            }
            a_aff = w.template wmin<SEG>(fmin(a_aff, a_num / a_den)); S1 = w.template wsum<SEG>(S1); S2 = w.template wsum<SEG>(S2);
            const double mu_aff = (mu_sum + a_aff * (S1 + a_aff * S2)) / (double)m_on;
            double sigma = (mu > 0.0) ? mu_aff / mu : 0.0;
            sigma = sigma * sigma * sigma;
            smu = fmax(sigma * mu, 0.1 * Q.tol_cp);
        }
    }
    QW_TICK(7);
    // ================= (8) step length and update =================
    double m_num = 1.0, m_den = 1.0, l_num = 1.0, l_den = 1.0;         // ratio tests: slacks (primal step) / multipliers (dual step)
    {
#pragma unroll 1
        for (int j = 0; j < C; ++j) {
            const int k = lane * C + j;
            double hk[4], it8[8];
            qw_ld_hit(w, j * qw_tm_stage(C), hk, it8);
            if (!act || k >= N) continue;
            double z6[6], dz6[6];
#pragma unroll
            for (int i = 0; i < 6; ++i) { z6[i] = QW_SM(R_Z + i, j); dz6[i] = QW_SM(R_GT + i, j); }
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                if (!h_on(hvar, k, c)) continue;
                IneqStep s_ = ineq_step(qw_row(hvar, c, hk[3], z6), QW_SM(R_DZA + c, j), qw_row(hvar, c, hk[3], dz6),
                                        QW_SM(R_LAM + c, j), QW_SM(R_LAM + 3 + c, j), QW_SM(R_T + c, j), QW_SM(R_T + 3 + c, j),
                                        it8[c], it8[4 + c], Q.lh[c] - hk[c], Q.uh[c] - hk[c], smu, Q.t_min);
                ratio_min(QW_SM(R_T + c, j), s_.dtl, m_num, m_den); ratio_min(QW_SM(R_T + 3 + c, j), s_.dtu, m_num, m_den);
                ratio_min(QW_SM(R_LAM + c, j), s_.dll, l_num, l_den); ratio_min(QW_SM(R_LAM + 3 + c, j), s_.dlu, l_num, l_den);
            }
        }
    }
    const double a_p = w.template wmin<SEG>(fmin(1.0, m_num / m_den)), a_d = w.template wmin<SEG>(fmin(1.0, l_num / l_den));
    double alpha, alpha_d;
    qp_step_lengths(Q, a_p, a_d, smu, mu, alpha, alpha_d);
    if ((!(alpha == alpha) || !(alpha_d == alpha_d)) && !st.fin) { status = 2; st.fin = true; }
    if (!w.wany(st.fin ? 0 : 1)) return 1;
    act = act && !st.fin;
    {
#pragma unroll 1
        for (int j = 0; j < C; ++j) {
            const int k = lane * C + j;
            double hk[4], it8[8], Pst[16];
            qw_ld_hit(w, j * qw_tm_stage(C), hk, it8);
            if constexpr (C >= 3) w.template tm_ld<16>(j * qw_tm_stage(C) + QW_TM_P, Pst);
            if (!act || k > N) continue;
            double dz[6];
#pragma unroll
            for (int i = 0; i < 6; ++i) dz[i] = QW_SM(R_GT + i, j);
            if (k >= 1) {                                   // dpi_k = P_k dx_k + p_k
                double Pk[10], dxk[4] = {dz[2], dz[3], dz[4], dz[5]}, dp[4];
#pragma unroll
                for (int i = 0; i < 10; ++i) { if constexpr (C >= 3) Pk[i] = Pst[i]; else Pk[i] = QW_SM(R_P + i, j); }
                sym4_mul(Pk, dxk, dp);
#pragma unroll
                for (int i = 0; i < 4; ++i) QW_SM(R_PIK + i, j) = fma(alpha_d, dp[i] + QW_SM(R_PV + i, j), QW_SM(R_PIK + i, j));
            }
            if (k < N) {
                double z6[6];
#pragma unroll
                for (int i = 0; i < 6; ++i) z6[i] = QW_SM(R_Z + i, j);
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    if (!h_on(hvar, k, c)) continue;
                    IneqStep s_ = ineq_step(qw_row(hvar, c, hk[3], z6), QW_SM(R_DZA + c, j), qw_row(hvar, c, hk[3], dz),
                                            QW_SM(R_LAM + c, j), QW_SM(R_LAM + 3 + c, j), QW_SM(R_T + c, j), QW_SM(R_T + 3 + c, j),
                                            it8[c], it8[4 + c], Q.lh[c] - hk[c], Q.uh[c] - hk[c], smu, Q.t_min);
                    QW_SM(R_T + c, j) = fma(alpha, s_.dtl, QW_SM(R_T + c, j));
                    QW_SM(R_T + 3 + c, j) = fma(alpha, s_.dtu, QW_SM(R_T + 3 + c, j));
                    QW_SM(R_LAM + c, j) = fma(alpha_d, s_.dll, QW_SM(R_LAM + c, j));
                    QW_SM(R_LAM + 3 + c, j) = fma(alpha_d, s_.dlu, QW_SM(R_LAM + 3 + c, j));
                }
            }
#pragma unroll
            for (int i = 0; i < 6; ++i) QW_SM(R_Z + i, j) = fma(alpha, dz[i], QW_SM(R_Z + i, j));
        }
    }
    w.sync();
    QW_TICK(8);
    if (!st.fin) ++it;
    return 0;
    }
}

// ---- write the point back to the slabs: V.z (du, dx), V.pi (pi[k] = pi_{k+1}), V.lam, V.t
template <class Ctx, int C, int SEG>
QS_HD void qw_writeback(const Ctx& w, double* __restrict__ sm, const QpConst& Q, const QpView& V, bool live, bool step_and_slacks = true) {
    const int N = Q.N;
    const int lane = w.lane() & (SEG - 1);
    const int Lw_ = qp_warp_lanes(N, C);
    const bool act = lane < Lw_ && live;
    // ---------------- write the point back to the slabs
    if (act) {
#pragma unroll 1
        for (int j = 0; j < C; ++j) {
            const int k = lane * C + j;
            if (k > N) continue;
            if (step_and_slacks) {                          // (du, dx): read by the SQP-level kernels only; the RTI epilogue applies it from shared memory
#pragma unroll
                for (int i = 0; i < 6; ++i) QS_AT(V.z, k, 6, i) = QW_SM(R_Z + i, j);
            }
            if (k >= 1) {
#pragma unroll
                for (int i = 0; i < 4; ++i) QS_AT(V.pi, k - 1, 4, i) = QW_SM(R_PIK + i, j);
            }
            if (k < N) {
#pragma unroll
                for (int i = 0; i < 6; ++i) { QS_AT(V.lam, k, 6, i) = QW_SM(R_LAM + i, j); if (step_and_slacks) QS_AT(V.t, k, 6, i) = QW_SM(R_T + i, j); }
            }
        }
    }
}

}  // namespace qs
