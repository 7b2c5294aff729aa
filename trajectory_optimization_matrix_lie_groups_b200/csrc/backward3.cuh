// Backward Riccati sweep for the 12-dimensional families (SE3, drone): one CTA of 2 warps per group
// of 32 problems (lane = problem, warp = role), 4 CTAs per SM so that all 512 groups of a 16k batch
// are resident at once.  Same recursion as backward.cuh (traopt_controller.py:2178-2321 / 2912-3068,
// gradient norms :2323-2349 / :3070-3093) with the work of one stage split by 3x3 column blocks of
// V_xx(i) = Q_xx - Y^T Y, two column blocks per warp, in three phases separated by CTA barriers:
//
//            phase 1                         | phase 2                               | phase 3
//   warp 0   X0 = V A[:,0] -> Y0, K0, V(0,0) | X2 -> Y2, K2, V(2,2)                  | V(0,2), V(1,2), V(2,3)
//   warp 1   X1 = V A[:,1] -> Y1, K1, V(1,1) | V(0,1);  X3 -> Y3, K3, V(3,3)         | V(0,3), V(1,3); SS: adjoint p
//
// Both warps first repeat the small serial part (Q_uu, its regularised Cholesky factor, y, k): it
// costs ~350 of the ~3500 FP64 operations of a stage and saves a broadcast.  An off-diagonal block
// V(r,c) = l_xx + A[:,r]^T V A[:,c] - Y_r^T Y_c is formed by the warp that owns X_c = V A[:,c]
// (V(2,3) by the owner of X_2, as X_2^T A[:,3]); it needs the other column's Y, which is why
// Y_0, Y_1, Y_3 go through shared memory.  V is updated IN PLACE: every new block is held in
// registers until the second barrier, after which nobody reads V(i+1) any more.
//
// Shared memory per problem (doubles, one column of 32 lanes each, conflict-free):
//   V    78   packed upper triangle of V_xx(i+1), overwritten in place after the second barrier
//   Vx   12   V_x(i+1), ditto
//   Y    NU*9 columns 0..5 and 9..11 of Y = L^-1 Q_ux
//   REC  A_LEN+NX+NU   the stage's record prefix (blocks of A = f_x, defect d, l_u), brought in by ONE
//        TMA bulk copy (cp.async.bulk, mbarrier completion) issued as soon as the previous stage
//        released the buffer; it lands while the warps factorise Q_uu.
// = 219 doubles (SE3) -> 54.8 KB per CTA, 4 CTAs per SM.  l_x, l_xx (used once each) are read
// straight from HBM/L2; the record two stages ahead is prefetched into L2.
#pragma once
#include "backward.cuh"

namespace trajopt {

// Section marks of a stage.  They are NAMED BARRIERS among the warps of a CTA that run the same code (B3_FENCE: the warps of
// one role across the CTA's groups; B3_FENCE_ALL: every warp, in the replicated serial part).  Two effects, measured on
// B200 (16384 x 955, ncu): (1) ptxas does not move instructions across a BAR, so each section is scheduled on its own and
// the live ranges stay short: 580 bytes of spills per thread -> 0, long-scoreboard stalls 1.25 -> 0.50 cycles per issued
// instruction, 10.85 -> 9.68 ms per sweep (bit-identical: only the order of independent instructions changes);
// (2) the warps stay within a few hundred instructions of each other, so a line of the ~110 KB stage body is fetched from
// beyond the SM once for all of them.  A compiler-only fence (asm volatile("" ::: "memory"), -DB3_USE_FENCE) at the same
// points does NOT have effect (1) — ptxas still interleaves the sections: 412 bytes of spills, 3 % slower than no fence.
// -DB3_NO_LOCKSTEP removes the marks (A/B).
#if defined(B3_NO_LOCKSTEP) && defined(B3_USE_FENCE)
#define B3_FENCE() asm volatile("" ::: "memory")
#define B3_FENCE_ALL() asm volatile("" ::: "memory")
#elif defined(B3_NO_LOCKSTEP)
#define B3_FENCE()
#define B3_FENCE_ALL()
#else
#define B3_FENCE()                                                                           \
    do {                                                                                     \
        if constexpr (b3_marks(KIND)) {                                                      \
            if (warp == 0) asm volatile("bar.sync 1, %0;" ::"r"(lock_role) : "memory");     \
            else asm volatile("bar.sync 2, %0;" ::"r"(lock_role) : "memory");               \
        }                                                                                    \
    } while (0)
#define B3_FENCE_ALL() do { if constexpr (b3_marks(KIND)) asm volatile("bar.sync 3, %0;" ::"r"(lock_all) : "memory"); } while (0)
#endif
// The quadrotor's two-warp sweep (four inputs: a smaller factor, smaller gains) spills 116 bytes per thread without the
// marks and runs 4 % FASTER without them (131072 x 150: 240.1 against 250.7 ms for the 27 sweeps of a solve); the
// six-input families spill 580-850 bytes without them.
__host__ __device__ constexpr bool b3_marks(int kind) { return kind != TRAJOPT_DRONE; }
// which warp of a group forms and stores the feed-forward gain k (A/B: -DB3_KFF_WARP=1)
#ifndef B3_KFF_WARP
#define B3_KFF_WARP 0
#endif
constexpr int kB3KffWarp = B3_KFF_WARP;
constexpr int kB3Warps = 2;
constexpr int kB3Threads = kB3Warps * 32;

template <int KIND> struct B3Smem {
    using D = Dims<KIND>;
    using LR = LinRec<KIND>;
    static constexpr int NT = D::NX * (D::NX + 1) / 2;
    static constexpr int NYC = 9;
    static constexpr int V_OFF = 0;
    static constexpr int VX_OFF = V_OFF + NT;
    static constexpr int Y_OFF = VX_OFF + D::NX;
    static constexpr int REC_OFF = Y_OFF + D::NU * NYC;
    static constexpr int DOUBLES = REC_OFF + LR::STAGE_LEN;
    static constexpr int REC_BYTES = LR::STAGE_LEN * 32 * 8;
    static constexpr size_t BYTES = (size_t)DOUBLES * 32 * 8 + 16 + 32 * sizeof(int);   // + mbarrier + per-lane flags
    static constexpr size_t GROUP_BYTES = (BYTES + 127) / 128 * 128;                     // stride between the groups of one CTA
};

__host__ __device__ constexpr int b3_ycol(int c) { return c < 6 ? c : c - 3; }   // smem slot of Y column c (c in 0..5, 9..11)

// (mbarrier / TMA bulk copy helpers: backward.cuh)

// ---- building blocks ----------------------------------------------------------------------------
// Plain fully-unrolled loops: after unrolling every index, block kind and record offset is a
// constant, zero blocks disappear and all arrays stay in registers.  Shared-memory columns are
// lane-private (element e of a lane lives at e * 32 doubles from the lane's base).

// 3x3 block (KB, CB) of A = f_x from the staged record
template <int KIND>
TO_DEV void b3_load_blk(int KB, int CB, const double* __restrict__ rec, double (&m)[9]) {
    const int kind = blk_kind<KIND>(KB, CB), off = blk_off<KIND>(KB, CB);
    if (kind == BK_DENSE) {
#pragma unroll
        for (int t = 0; t < 9; ++t) m[t] = rec[(off + t) * kRecStride];
    } else {
        const double v0 = rec[off * kRecStride], v1 = rec[(off + 1) * kRecStride], v2 = rec[(off + 2) * kRecStride];
        if (kind == BK_SKEW) {
            m[0] = 0.0; m[1] = -v2; m[2] = v1;
            m[3] = v2;  m[4] = 0.0; m[5] = -v0;
            m[6] = -v1; m[7] = v0;  m[8] = 0.0;
        } else {
            m[0] = 1.0; m[1] = v2;  m[2] = -v1;
            m[3] = -v2; m[4] = 1.0; m[5] = v0;
            m[6] = v1;  m[7] = -v0; m[8] = 1.0;
        }
    }
}

// Rows [R0, R1) of X = V A[:, CB]  (NX x 3)
template <int KIND, int CB, int R0, int R1>
TO_DEV void b3_compute_X(const double* __restrict__ Vs, const double* __restrict__ rec, double (&X)[Dims<KIND>::NX][3]) {
    constexpr int NX = Dims<KIND>::NX, NB = NX / 3;
#pragma unroll
    for (int r = R0; r < R1; ++r)
#pragma unroll
        for (int j = 0; j < 3; ++j) X[r][j] = 0.0;
#pragma unroll
    for (int KB = 0; KB < NB; ++KB) {
        const int kind = blk_kind<KIND>(KB, CB);
        if (kind == BK_ZERO) continue;
        double m[9];
        b3_load_blk<KIND>(KB, CB, rec, m);
#pragma unroll
        for (int r = R0; r < R1; ++r)
#pragma unroll
            for (int ii = 0; ii < 3; ++ii) {
                const double t = Vs[sym_idx(NX, r, KB * 3 + ii) * 32];
#pragma unroll
                for (int j = 0; j < 3; ++j)
                    if (blk_nz(kind, ii, j)) X[r][j] = fma(t, m[3 * ii + j], X[r][j]);
            }
    }
}

// qx[j] += sum_k A[k, CB j] V_x[k]   (the V_x part of Q_x = l_x + A^T (V_x + V_xx d))
template <int KIND, int CB>
TO_DEV void b3_qx_A(const double* __restrict__ Vxs, const double* __restrict__ rec, double (&qx)[3]) {
    constexpr int NB = Dims<KIND>::NX / 3;
#pragma unroll
    for (int KB = 0; KB < NB; ++KB) {
        const int kind = blk_kind<KIND>(KB, CB);
        if (kind == BK_ZERO) continue;
        double m[9];
        b3_load_blk<KIND>(KB, CB, rec, m);
#pragma unroll
        for (int ii = 0; ii < 3; ++ii) {
            const double vx = Vxs[(KB * 3 + ii) * 32];
#pragma unroll
            for (int j = 0; j < 3; ++j)
                if (blk_nz(kind, ii, j)) qx[j] = fma(m[3 * ii + j], vx, qx[j]);
        }
    }
}

// qx[j] += sum_{r in [R0, R1)} X[r][j] d[r]   (multiple shooting: (V A)^T d = A^T V d, taken row range by row range)
template <int KIND, int R0, int R1>
TO_DEV void b3_qx_Xd(const double* __restrict__ rec, const double (&X)[Dims<KIND>::NX][3], double (&qx)[3]) {
    using LR = LinRec<KIND>;
#pragma unroll
    for (int r = R0; r < R1; ++r) {
        const double dr = rec[(LR::D_OFF + r) * kRecStride];
#pragma unroll
        for (int j = 0; j < 3; ++j) qx[j] = fma(X[r][j], dr, qx[j]);
    }
}

// qx[j] += sum_k A[k, CB j] (V_xx d)[k], formed from V directly (for the warp that only has the velocity rows of X)
template <int KIND, int CB>
TO_DEV void b3_qx_AVd(const double* __restrict__ Vs, const double* __restrict__ rec, double (&qx)[3]) {
    using LR = LinRec<KIND>;
    constexpr int NX = Dims<KIND>::NX, NB = NX / 3;
#pragma unroll
    for (int KB = 0; KB < NB; ++KB) {
        const int kind = blk_kind<KIND>(KB, CB);
        if (kind == BK_ZERO) continue;
        double m[9];
        b3_load_blk<KIND>(KB, CB, rec, m);
#pragma unroll
        for (int ii = 0; ii < 3; ++ii) {
            double vd = 0.0;
#pragma unroll
            for (int c = 0; c < NX; ++c) vd = fma(Vs[sym_idx(NX, KB * 3 + ii, c) * 32], rec[(LR::D_OFF + c) * kRecStride], vd);
#pragma unroll
            for (int j = 0; j < 3; ++j)
                if (blk_nz(kind, ii, j)) qx[j] = fma(m[3 * ii + j], vd, qx[j]);
        }
    }
}

// Column block CB of Q_ux (from the velocity rows of X) -> Y = L^-1 Q_ux, K = -L^-T Y (stored);
// yq[j] = Y[:, j]^T y is what V_x(i)[CB j] = Q_x - Y^T y subtracts.
template <int KIND, int CB, bool Y_TO_SMEM, bool Y_TO_REGS>
TO_DEV void b3_gains(const Params& prm, const double* __restrict__ rec, const double (&X)[Dims<KIND>::NX][3],
                     const double (&Lc)[Dims<KIND>::NU * Dims<KIND>::NU], const double (&y)[Dims<KIND>::NU], double mu_used,
                     double (&yq)[3], double* __restrict__ Ys, double (&Yk)[Dims<KIND>::NU][3], double* __restrict__ Kout,
                     bool act) {
    using D = Dims<KIND>;
    constexpr int NX = D::NX, NP = D::NP, NU = D::NU, NV = NX - NP, NB = NX / 3, NYC = B3Smem<KIND>::NYC;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        double q[NU];
#pragma unroll
        for (int a = 0; a < NU; ++a) {
            double s = 0.0;
#pragma unroll
            for (int r = 0; r < NV; ++r)
                if (bv_nz<KIND>(r, a)) s = fma(prm.Bv[r * NU + a], X[NP + r][j], s);
            q[a] = s;
        }
        if (mu_used != 0.0) {   // + mu Bv^T A_vel[:, CB j]  (:2311-2312); mu is 0 after the first stages of a solve
#pragma unroll
            for (int KB = NP / 3; KB < NB; ++KB) {
                const int kind = blk_kind<KIND>(KB, CB);
                if (kind == BK_ZERO) continue;
                double m[9];
                b3_load_blk<KIND>(KB, CB, rec, m);
#pragma unroll
                for (int a = 0; a < NU; ++a)
#pragma unroll
                    for (int ii = 0; ii < 3; ++ii)
                        if (bv_nz<KIND>(KB * 3 + ii - NP, a) && blk_nz(kind, ii, j))
                            q[a] = fma(mu_used * prm.Bv[(KB * 3 + ii - NP) * NU + a], m[3 * ii + j], q[a]);
            }
        }
        // Y[:, j] = L^-1 q
#pragma unroll
        for (int r = 0; r < NU; ++r) {
            double s = q[r];
#pragma unroll
            for (int k = 0; k < r; ++k) s = fma(-Lc[r * NU + k], q[k], s);
            q[r] = s * Lc[r * NU + r];
        }
        // K[:, j] = -L^-T Y[:, j]
        double kk[NU];
#pragma unroll
        for (int r = NU - 1; r >= 0; --r) {
            double s = q[r];
#pragma unroll
            for (int k = r + 1; k < NU; ++k) s = fma(-Lc[k * NU + r], kk[k], s);
            kk[r] = s * Lc[r * NU + r];
        }
        if (act) {
#pragma unroll
            for (int a = 0; a < NU; ++a) Kout[(a * NX + CB * 3 + j) * kRecStride] = -kk[a];
        }
        double s = 0.0;
#pragma unroll
        for (int a = 0; a < NU; ++a) s = fma(q[a], y[a], s);
        yq[j] = s;
#pragma unroll
        for (int a = 0; a < NU; ++a) {
            if (Y_TO_SMEM) Ys[(a * NYC + b3_ycol(CB * 3 + j)) * 32] = q[a];
            if (Y_TO_REGS) Yk[a][j] = q[a];
        }
    }
}

// l_xx entry (r, c) of the stage cost (Gauss-Newton pose block from the record, constant velocity block)
template <int KIND>
TO_DEV double b3_lxx(const Params& prm, const double* __restrict__ grec, int r, int c) {
    using D = Dims<KIND>;
    using LR = LinRec<KIND>;
    constexpr int NP = D::NP, NV = D::NX - D::NP;
    if (c < NP) return grec[(size_t)(LR::LXX_OFF + tri_idx(NP, r, c)) * kRecStride];
    if (r >= NP) return 2.0 * prm.W2[(r - NP) * NV + (c - NP)];
    return 0.0;
}

// Y[a][column block CBY, t]: from shared memory, or (REGS) from the caller's own register copy
template <int KIND, bool REGS>
TO_DEV double b3_y(const double* __restrict__ Ys, const double (&Yk)[Dims<KIND>::NU][3], int cby, int a, int t) {
    if (REGS) return Yk[a][t];
    return Ys[(a * B3Smem<KIND>::NYC + b3_ycol(cby * 3 + t)) * 32];
}

// V(RB, CB) for the owner of X = V A[:, CB]:  acc = l_xx + A[:, RB]^T X - Y_RB^T Y_CB   (RB <= CB)
template <int KIND, int RB, int CB, bool YR_REGS, bool YC_REGS>
TO_DEV void b3_block_cb(const Params& prm, const double* __restrict__ rec, const double* __restrict__ grec,
                        const double (&X)[Dims<KIND>::NX][3], const double* __restrict__ Ys,
                        const double (&Yk)[Dims<KIND>::NU][3], double (&acc)[3][3]) {
    constexpr int NX = Dims<KIND>::NX, NU = Dims<KIND>::NU, NB = NX / 3;
#pragma unroll
    for (int ii = 0; ii < 3; ++ii)
#pragma unroll
        for (int j = 0; j < 3; ++j)
            acc[ii][j] = (RB < CB || ii <= j) ? b3_lxx<KIND>(prm, grec, RB * 3 + ii, CB * 3 + j) : 0.0;
#pragma unroll
    for (int KB = 0; KB < NB; ++KB) {
        const int kind = blk_kind<KIND>(KB, RB);
        if (kind == BK_ZERO) continue;
        double m[9];
        b3_load_blk<KIND>(KB, RB, rec, m);
#pragma unroll
        for (int k = 0; k < 3; ++k)
#pragma unroll
            for (int ii = 0; ii < 3; ++ii)
                if (blk_nz(kind, k, ii)) {
#pragma unroll
                    for (int j = 0; j < 3; ++j)
                        if (RB < CB || ii <= j) acc[ii][j] = fma(m[3 * k + ii], X[KB * 3 + k][j], acc[ii][j]);
                }
    }
#pragma unroll
    for (int a = 0; a < NU; ++a) {
        double r3[3], c3[3];
#pragma unroll
        for (int t = 0; t < 3; ++t) {
            r3[t] = b3_y<KIND, YR_REGS>(Ys, Yk, RB, a, t);
            c3[t] = b3_y<KIND, YC_REGS>(Ys, Yk, CB, a, t);
        }
#pragma unroll
        for (int ii = 0; ii < 3; ++ii)
#pragma unroll
            for (int j = 0; j < 3; ++j)
                if (RB < CB || ii <= j) acc[ii][j] = fma(-r3[ii], c3[j], acc[ii][j]);
    }
}

// V(RB, CB), RB < CB, for the owner of X = V A[:, RB]:  acc = l_xx + X^T A[:, CB] - Y_RB^T Y_CB
template <int KIND, int RB, int CB, bool YR_REGS, bool YC_REGS>
TO_DEV void b3_block_rb(const Params& prm, const double* __restrict__ rec, const double* __restrict__ grec,
                        const double (&X)[Dims<KIND>::NX][3], const double* __restrict__ Ys,
                        const double (&Yk)[Dims<KIND>::NU][3], double (&acc)[3][3]) {
    constexpr int NX = Dims<KIND>::NX, NU = Dims<KIND>::NU, NB = NX / 3;
#pragma unroll
    for (int ii = 0; ii < 3; ++ii)
#pragma unroll
        for (int j = 0; j < 3; ++j) acc[ii][j] = b3_lxx<KIND>(prm, grec, RB * 3 + ii, CB * 3 + j);
#pragma unroll
    for (int KB = 0; KB < NB; ++KB) {
        const int kind = blk_kind<KIND>(KB, CB);
        if (kind == BK_ZERO) continue;
        double m[9];
        b3_load_blk<KIND>(KB, CB, rec, m);
#pragma unroll
        for (int k = 0; k < 3; ++k)
#pragma unroll
            for (int j = 0; j < 3; ++j)
                if (blk_nz(kind, k, j)) {
#pragma unroll
                    for (int ii = 0; ii < 3; ++ii) acc[ii][j] = fma(X[KB * 3 + k][ii], m[3 * k + j], acc[ii][j]);
                }
    }
#pragma unroll
    for (int a = 0; a < NU; ++a) {
        double r3[3], c3[3];
#pragma unroll
        for (int t = 0; t < 3; ++t) {
            r3[t] = b3_y<KIND, YR_REGS>(Ys, Yk, RB, a, t);
            c3[t] = b3_y<KIND, YC_REGS>(Ys, Yk, CB, a, t);
        }
#pragma unroll
        for (int ii = 0; ii < 3; ++ii)
#pragma unroll
            for (int j = 0; j < 3; ++j) acc[ii][j] = fma(-r3[ii], c3[j], acc[ii][j]);
    }
}

template <int KIND, int RB, int CB>
TO_DEV void b3_store_block(double* __restrict__ Vs, const double (&acc)[3][3], bool live) {
    constexpr int NX = Dims<KIND>::NX;
    if (!live) return;   // a problem whose horizon has not started yet keeps its terminal V
#pragma unroll
    for (int ii = 0; ii < 3; ++ii)
#pragma unroll
        for (int j = 0; j < 3; ++j)
            if (RB < CB || ii <= j) Vs[tri_idx(NX, RB * 3 + ii, CB * 3 + j) * 32] = acc[ii][j];
}

// VH: per-problem horizons (trajopt_set_horizons).  Without it every `vlive` below folds to true and the stores are
// unpredicated, which is worth 9% of the sweep at the headline size.
// G: groups of 32 problems per CTA (2 warps each).  With G > 1 the groups share the CTA-wide barriers and so march through
// the stage body in step: a line of its ~110 KB of straight-line code is fetched once for all of them.
template <int KIND, bool MS, bool VH, int G>
__global__ void __launch_bounds__(kB3Threads * G, 4 / G) k_backward3(const Params prm, Work w, int it_arg) {
    static_assert(!on_so3(KIND), "the 3-warp sweep is for the 12-dimensional families");
    using D = Dims<KIND>;
    using LR = LinRec<KIND>;
    using SM = B3Smem<KIND>;
    constexpr int NX = D::NX, NP = D::NP, NU = D::NU, NV = NX - NP, F = LR::LEN;
    extern __shared__ __align__(128) double sm_cta[];
    const int lane = threadIdx.x & 31;
    // G = 4 (-DB3_WITH_G4, A/B): the roles alternate so that every scheduler hosts one warp of each role (warp w runs on
    // scheduler w % 4): two warps of the same role in step on one scheduler want the same pipe at the same time
    const int wq = threadIdx.x >> 5;
    const int warp = (G == 4) ? ((wq & 1) ^ ((wq >> 2) & 1)) : (wq & 1);
    const int tid = warp * 32 + lane;                 // thread within its group
    const int grp = wq >> 1;
    double* sm = sm_cta + (size_t)grp * (SM::GROUP_BYTES / 8);
    const int b = (blockIdx.x * G + grp) * 32 + lane;
    const int N = prm.N, Np1 = N + 1;     // record layout / loop extent; the problem's own horizon is Nb <= N
    const int Nb = VH ? w.Nb[b < prm.Bp ? b : 0] : N;
    const size_t Bp = (size_t)prm.Bp;

    double* Vs = sm + SM::V_OFF * 32 + lane;
    double* Vxs = sm + SM::VX_OFF * 32 + lane;
    double* Ys = sm + SM::Y_OFF * 32 + lane;
    const double* rec = sm + SM::REC_OFF * 32 + lane;
    const uint32_t rec_addr = b3_smem_addr(sm + SM::REC_OFF * 32);
    const uint32_t bar = b3_smem_addr(sm + SM::DOUBLES * 32);
    int* flags = reinterpret_cast<int*>(sm + SM::DOUBLES * 32 + 2);

    bool act = (b < prm.B) && (w.status[b] == TRAJOPT_RUNNING);
    const int it = it_arg >= 0 ? it_arg : (act ? w.iters[b] : 0);   // < 0: per-slot iteration counts (trajopt_solve_stream)
    bool alive = __ballot_sync(0xffffffffu, act) != 0u;   // same lanes in both warps of a group: they agree
    if constexpr (G == 1) {
        if (!alive) return;                                // the whole CTA leaves together
    } else {
        if (!__syncthreads_or(alive)) return;              // a finished group stays for the barriers of the others
    }

    const double* __restrict__ lin = w.lin;
    // record of stage s of this lane (global); fields are kRecStride apart
    auto grec_of = [&](int s) { return lin + lsoa(s, 0, F, Np1, b); };
    const double* group_base = lin + lsoa(0, 0, F, Np1, b - lane);          // stage 0 of the group
    constexpr size_t kStageDoubles = (size_t)F * 32;

    // ---- cost / defect of the current trajectory (warp 0), terminal condition (warps 1, 2) --------
    if (!alive) {
    } else if (warp == 0) {
        int ok = act ? 1 : 0;
        if (act) {
            double Jcur, dn = 0.0;
            if constexpr (MS) {
                double s = 0.0;   // J_new of the previous iteration: left to right, + terminal (:2742-2754)
                for (int i = 0; i < Nb; ++i) s += w.Lc[(size_t)i * Bp + b];
                Jcur = s + w.Lc[(size_t)Nb * Bp + b];
                double q = 0.0;
                for (int i = 0; i < Nb; ++i) q += w.Dsq[(size_t)i * Bp + b];
                dn = sqrt(q);
                w.dnorm[b] = dn;
                if (it > 0) w.Jhist[(size_t)(it - 1) * Bp + b] = Jcur;
                w.defhist[(size_t)it * Bp + b] = dn;
            } else {
                Jcur = pairwise_sum(w.Lc + b, Bp, Nb + 1);   // J_opt = L.sum() (:1935)
            }
            w.J[b] = Jcur;
            if (!isfinite(Jcur)) {
                w.status[b] = TRAJOPT_NO_DESCENT | TRAJOPT_FLAG_NONFINITE;
                ok = 0;
            } else if (it >= prm.max_iters) {   // MS only: closing pass after the last rollout
                w.status[b] = TRAJOPT_MAX_ITER | (w.status[b] & ~15);
                ok = 0;
            }
        }
        flags[lane] = ok;
    } else {
        const double* g = grec_of(Nb);
#pragma unroll
        for (int r = 0; r < NX; ++r)
#pragma unroll
            for (int c = r; c < NX; ++c) {
                double v;
                if (c < NP) v = g[(size_t)(LR::LXX_OFF + tri_idx(NP, r, c)) * kRecStride];
                else if (r >= NP) v = 2.0 * prm.P2[(r - NP) * NV + (c - NP)] + ((r == c && prm.has_state_bounds) ? w.lxxv[soa(Nb, r - NP, NV, (int)Bp, b)] : 0.0);
                else v = 0.0;
                Vs[tri_idx(NX, r, c) * 32] = v;
            }
#pragma unroll
        for (int j = 0; j < NX; ++j) Vxs[j * 32] = g[(size_t)(LR::LX_OFF + j) * kRecStride];
    }
    if (alive && tid == 0) b3_mbar_init(bar, 1);
    __syncthreads();
    if (alive) {
        act = flags[lane] != 0;
        alive = __ballot_sync(0xffffffffu, act) != 0u;
    }
    [[maybe_unused]] int lock_role = 32, lock_all = 64;     // threads behind a role-wide / CTA-wide section mark (B3_FENCE)
    if constexpr (G == 1) {
        if (!alive) return;
    } else {
        const int n_alive = __syncthreads_count(alive && tid == 0);   // groups of this CTA that take part in the sweep
        if (n_alive == 0) return;
        lock_role = 32 * n_alive;
        lock_all = 64 * n_alive;
    }
    if (alive && tid == 0) b3_tma_load(rec_addr, group_base + (size_t)(N - 1) * kStageDoubles, SM::REC_BYTES, bar);

    double pad[NX];   // SS, warp 1: adjoint variable p (:2339)
    if constexpr (!MS) {
        if (warp == 1 && alive) {
#pragma unroll
            for (int j = 0; j < NX; ++j) pad[j] = Vxs[j * 32];
        }
    }

    double mu = 0.0, delta = 0.0;
    if (alive) {
        mu = w.mu[b];
        delta = w.delta[b];
    }
    double gsum = 0.0;
    int flag_bits = 0;
    uint32_t parity = 0;

    for (int i = N - 1; i >= 0; --i) {
        if constexpr (G > 1) {
            if (!alive) {   // the three barriers of a stage
                __syncthreads();
                __syncthreads();
                __syncthreads();
                continue;
            }
        }
        const double* __restrict__ grec = grec_of(i);
        bool live = act && (!VH || i < Nb);   // stages beyond a problem's horizon only keep the barriers company
#define vlive (VH ? live : true)
        // pull the record two stages ahead towards L2 (the next one is already on its way through TMA)
        if (i >= 2) {
            const char* nxt = (const char*)(group_base + (size_t)(i - 2) * kStageDoubles);
#pragma unroll
            for (int t = 0; t < (F * 2 + kB3Threads - 1) / kB3Threads; ++t) {
                const int line = t * kB3Threads + tid;
                if (line < F * 2) asm volatile("prefetch.global.L2 [%0];" ::"l"(nxt + (size_t)line * 128));
            }
        }

#ifdef B3_L1_PREFETCH
        {   // A/B: the stage's l_x / l_xx rows (read straight from global memory further down) towards L1
            const char* lxrow = (const char*)(group_base + (size_t)i * kStageDoubles + (size_t)LR::LX_OFF * 32);
            constexpr int kLines = (NX + LR::LXX_LEN) * 2;
#pragma unroll
            for (int t = 0; t < (kLines + kB3Threads - 1) / kB3Threads; ++t) {
                const int line = t * kB3Threads + tid;
                if (line < kLines) asm volatile("prefetch.global.L1 [%0];" ::"l"(lxrow + (size_t)line * 128));
            }
        }
#endif
        // ---- (S0a) Q_uu0 = l_uu + Bv^T V_vv Bv and its regularised Cholesky factor (every warp) -------
        double Lc[NU * NU];   // lower factor, Lc[r][c] for c < r; the diagonal stores 1 / L_rr
        double mu_used;
        {
            double Quu0[NU * NU];
#pragma unroll
            for (int a = 0; a < NU; ++a)
#pragma unroll
                for (int c = a; c < NU; ++c) {
                    double s = 2.0 * prm.R[a * NU + c];
                    if (a == c && prm.has_constraints) s += grec[(size_t)(LR::LUU_OFF + a) * kRecStride];
                    Quu0[a * NU + c] = s;
                }
#pragma unroll
            for (int r = 0; r < NV; ++r) {   // row r of T = V_vv Bv, folded into Bv^T T at once
                double vr[NV], T[NU];
#pragma unroll
                for (int k = 0; k < NV; ++k) vr[k] = Vs[sym_idx(NX, NP + r, NP + k) * 32];
#pragma unroll
                for (int a = 0; a < NU; ++a) {
                    double s = 0.0;
#pragma unroll
                    for (int k = 0; k < NV; ++k)
                        if (bv_nz<KIND>(k, a)) s = fma(vr[k], prm.Bv[k * NU + a], s);
                    T[a] = s;
                }
#pragma unroll
                for (int a = 0; a < NU; ++a)
                    if (bv_nz<KIND>(r, a)) {
#pragma unroll
                        for (int c = a; c < NU; ++c) Quu0[a * NU + c] = fma(prm.Bv[r * NU + a], T[c], Quu0[a * NU + c]);
                    }
            }
            B3_FENCE_ALL();
            // regularisation loop (:2221-2246 / :2964-2991): Cholesky of Q_uu0 + mu B^T B
            while (true) {
                mu_used = mu;
                bool pd = true;
                // (compile-time indices throughout: a rolled loop here would push Lc into local memory)
                sfor<0, NU>([&](auto cc) {
                    constexpr int c = decltype(cc)::value;
                    double dg = fma(mu_used, prm.BtB[c * NU + c], Quu0[c * NU + c]);
                    sfor<0, c>([&](auto kc) {
                        constexpr int k = decltype(kc)::value;
                        dg = fma(-Lc[c * NU + k], Lc[c * NU + k], dg);
                    });
                    if (!(dg > 0.0)) pd = false;
                    const double inv = rsqrt(dg);
                    Lc[c * NU + c] = inv;
                    sfor<c + 1, NU>([&](auto rc) {
                        constexpr int r = decltype(rc)::value;
                        double sacc = fma(mu_used, prm.BtB[c * NU + r], Quu0[c * NU + r]);
                        sfor<0, c>([&](auto kc) {
                            constexpr int k = decltype(kc)::value;
                            sacc = fma(-Lc[r * NU + k], Lc[c * NU + k], sacc);
                        });
                        Lc[r * NU + c] = sacc * inv;
                    });
                });
                if (!live) break;   // finished / padded / not-yet-started lanes only keep the barriers company
                if (!pd) {
                    delta = fmax(1.0, delta) * prm.delta0;
                    mu = fmax(prm.mu_min, mu * delta);
                    if (prm.mu_max > 0.0 && mu >= prm.mu_max) {
                        // The reference warns and carries on with a non-PD Q_uu (:2238-2240); such a problem has
                        // already diverged.  It is stopped here and flagged.
                        flag_bits |= TRAJOPT_FLAG_REG_EXCEEDED;
                        act = false;
                        live = false;
                        break;
                    }
                } else {
                    delta = fmin(1.0, delta) / prm.delta0;
                    mu *= delta;
                    if (mu <= prm.mu_min) mu = 0.0;
                    break;
                }
            }
        }

        B3_FENCE_ALL();
        b3_mbar_wait(bar, parity);   // the stage's record prefix is in shared memory
        parity ^= 1u;

        // ---- (S0b) Q_u = l_u + Bv^T (V_x + V_xx d)_vel;  y = L^-1 Q_u;  k = -L^-T y (every warp) ------
        double y[NU];
        {
            double vv[NV];
#pragma unroll
            for (int r = 0; r < NV; ++r) vv[r] = Vxs[(NP + r) * 32];
            if constexpr (MS) {
#pragma unroll
                for (int c = 0; c < NX; ++c) {
                    const double dc = rec[(LR::D_OFF + c) * kRecStride];
#pragma unroll
                    for (int r = 0; r < NV; ++r) vv[r] = fma(Vs[sym_idx(NX, NP + r, c) * 32], dc, vv[r]);
                }
            }
            double Qu[NU];
#pragma unroll
            for (int a = 0; a < NU; ++a) {
                double s = rec[(LR::LU_OFF + a) * kRecStride];
#pragma unroll
                for (int r = 0; r < NV; ++r)
                    if (bv_nz<KIND>(r, a)) s = fma(prm.Bv[r * NU + a], vv[r], s);
                Qu[a] = s;
            }
            if constexpr (MS) {   // g_t = L_u + F_u^T (V_x + V_xx^T d) = Q_u (:3090)
                double q = 0.0;
#pragma unroll
                for (int a = 0; a < NU; ++a) q += Qu[a] * Qu[a];
                if (vlive) gsum += sqrt(q);
            }
#pragma unroll
            for (int r = 0; r < NU; ++r) {
                double s = Qu[r];
#pragma unroll
                for (int k = 0; k < r; ++k) s = fma(-Lc[r * NU + k], y[k], s);
                y[r] = s * Lc[r * NU + r];
            }
            if (warp == kB3KffWarp) {
                double kk[NU];
#pragma unroll
                for (int r = NU - 1; r >= 0; --r) {
                    double s = y[r];
#pragma unroll
                    for (int k = r + 1; k < NU; ++k) s = fma(-Lc[k * NU + r], kk[k], s);
                    kk[r] = s * Lc[r * NU + r];
                }
                if (live) {
#pragma unroll
                    for (int a = 0; a < NU; ++a) w.gains[lsoa(i, GainRec<KIND>::KFF_OFF + a, GainRec<KIND>::LEN, N, b)] = -kk[a];
                }
            }
        }

        B3_FENCE_ALL();
        double* Kout = w.gains + lsoa(i, 0, GainRec<KIND>::LEN, N, b);
        double Yk[NU][3];   // Y_2 in registers (warp 0)

        if (warp == 0) {
            double h00[3][3], h22[3][3], vx0[3], vx2[3], yq[3];
            {
                double X[NX][3];
#pragma unroll
                for (int j = 0; j < 3; ++j) vx0[j] = grec[(size_t)(LR::LX_OFF + j) * kRecStride];
                b3_compute_X<KIND, 0, NP, NX>(Vs, rec, X);
                B3_FENCE();
                b3_gains<KIND, 0, true, false>(prm, rec, X, Lc, y, mu_used, yq, Ys, Yk, Kout, live);
                B3_FENCE();
                b3_qx_A<KIND, 0>(Vxs, rec, vx0);
                if constexpr (MS) b3_qx_Xd<KIND, NP, NX>(rec, X, vx0);
                B3_FENCE();
                b3_compute_X<KIND, 0, 0, NP>(Vs, rec, X);
                if constexpr (MS) b3_qx_Xd<KIND, 0, NP>(rec, X, vx0);
                B3_FENCE();
#pragma unroll
                for (int j = 0; j < 3; ++j) vx0[j] -= yq[j];
                b3_block_cb<KIND, 0, 0, false, false>(prm, rec, grec, X, Ys, Yk, h00);
            }
            __syncthreads();   // (1) Y_0, Y_1 visible
            double X[NX][3];
#pragma unroll
            for (int j = 0; j < 3; ++j) vx2[j] = grec[(size_t)(LR::LX_OFF + 6 + j) * kRecStride];
            b3_compute_X<KIND, 2, NP, NX>(Vs, rec, X);
            B3_FENCE();
            b3_gains<KIND, 2, false, true>(prm, rec, X, Lc, y, mu_used, yq, Ys, Yk, Kout, live);
            B3_FENCE();
            b3_qx_A<KIND, 2>(Vxs, rec, vx2);
            if constexpr (MS) b3_qx_Xd<KIND, NP, NX>(rec, X, vx2);
            B3_FENCE();
            b3_compute_X<KIND, 2, 0, NP>(Vs, rec, X);
            if constexpr (MS) b3_qx_Xd<KIND, 0, NP>(rec, X, vx2);
            B3_FENCE();
#pragma unroll
            for (int j = 0; j < 3; ++j) vx2[j] -= yq[j];
            b3_block_cb<KIND, 2, 2, true, true>(prm, rec, grec, X, Ys, Yk, h22);
            if (prm.has_state_bounds) {
#pragma unroll
                for (int j = 0; j < 3; ++j) h22[j][j] += w.lxxv[soa(i, j, NV, (int)Bp, b)];
            }
            __syncthreads();   // (2) every X is formed: V and V_x may be overwritten; Y_3 visible
            b3_store_block<KIND, 0, 0>(Vs, h00, vlive);
            b3_store_block<KIND, 2, 2>(Vs, h22, vlive);
            if (vlive) {
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    Vxs[j * 32] = vx0[j];
                    Vxs[(6 + j) * 32] = vx2[j];
                }
            }
            {
                double acc[3][3];
                b3_block_cb<KIND, 0, 2, false, true>(prm, rec, grec, X, Ys, Yk, acc);
                b3_store_block<KIND, 0, 2>(Vs, acc, vlive);
                B3_FENCE();
                b3_block_cb<KIND, 1, 2, false, true>(prm, rec, grec, X, Ys, Yk, acc);
                b3_store_block<KIND, 1, 2>(Vs, acc, vlive);
                B3_FENCE();
                b3_block_rb<KIND, 2, 3, true, false>(prm, rec, grec, X, Ys, Yk, acc);
                b3_store_block<KIND, 2, 3>(Vs, acc, vlive);
            }
        } else {
            double h11[3][3], h01[3][3], h33[3][3], vx1[3], vx3[3], yq[3];
            {
                double X[NX][3];
#pragma unroll
                for (int j = 0; j < 3; ++j) vx1[j] = grec[(size_t)(LR::LX_OFF + 3 + j) * kRecStride];
                b3_compute_X<KIND, 1, NP, NX>(Vs, rec, X);
                B3_FENCE();
                b3_gains<KIND, 1, true, false>(prm, rec, X, Lc, y, mu_used, yq, Ys, Yk, Kout, live);
                B3_FENCE();
                b3_qx_A<KIND, 1>(Vxs, rec, vx1);
                if constexpr (MS) b3_qx_Xd<KIND, NP, NX>(rec, X, vx1);
                B3_FENCE();
                b3_compute_X<KIND, 1, 0, NP>(Vs, rec, X);
                if constexpr (MS) b3_qx_Xd<KIND, 0, NP>(rec, X, vx1);
                B3_FENCE();
#pragma unroll
                for (int j = 0; j < 3; ++j) vx1[j] -= yq[j];
                b3_block_cb<KIND, 1, 1, false, false>(prm, rec, grec, X, Ys, Yk, h11);
                __syncthreads();   // (1)
                b3_block_cb<KIND, 0, 1, false, false>(prm, rec, grec, X, Ys, Yk, h01);
            }
            double X[NX][3];
#pragma unroll
            for (int j = 0; j < 3; ++j) vx3[j] = grec[(size_t)(LR::LX_OFF + 9 + j) * kRecStride];
            b3_compute_X<KIND, 3, NP, NX>(Vs, rec, X);
            B3_FENCE();
            b3_gains<KIND, 3, true, false>(prm, rec, X, Lc, y, mu_used, yq, Ys, Yk, Kout, live);
            B3_FENCE();
            b3_qx_A<KIND, 3>(Vxs, rec, vx3);
            if constexpr (MS) b3_qx_Xd<KIND, NP, NX>(rec, X, vx3);
            B3_FENCE();
            b3_compute_X<KIND, 3, 0, NP>(Vs, rec, X);
            if constexpr (MS) b3_qx_Xd<KIND, 0, NP>(rec, X, vx3);
            B3_FENCE();
#pragma unroll
            for (int j = 0; j < 3; ++j) vx3[j] -= yq[j];
            b3_block_cb<KIND, 3, 3, false, false>(prm, rec, grec, X, Ys, Yk, h33);
            if (prm.has_state_bounds) {
#pragma unroll
                for (int j = 0; j < 3; ++j) h33[j][j] += w.lxxv[soa(i, 3 + j, NV, (int)Bp, b)];
            }
            __syncthreads();   // (2)
            b3_store_block<KIND, 1, 1>(Vs, h11, vlive);
            b3_store_block<KIND, 0, 1>(Vs, h01, vlive);
            b3_store_block<KIND, 3, 3>(Vs, h33, vlive);
            if (vlive) {
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    Vxs[(3 + j) * 32] = vx1[j];
                    Vxs[(9 + j) * 32] = vx3[j];
                }
            }
            {
                double acc[3][3];
                b3_block_cb<KIND, 0, 3, false, false>(prm, rec, grec, X, Ys, Yk, acc);
                b3_store_block<KIND, 0, 3>(Vs, acc, vlive);
                B3_FENCE();
                b3_block_cb<KIND, 1, 3, false, false>(prm, rec, grec, X, Ys, Yk, acc);
                b3_store_block<KIND, 1, 3>(Vs, acc, vlive);
            }
            if constexpr (!MS) if (vlive) {
                // adjoint recursion of the single-shooting gradient (:2343-2346): g = l_u + B^T p;  p <- l_x + A^T p
                double g2 = 0.0;
#pragma unroll
                for (int a = 0; a < NU; ++a) {
                    double s = rec[(LR::LU_OFF + a) * kRecStride];
#pragma unroll
                    for (int r = 0; r < NV; ++r)
                        if (bv_nz<KIND>(r, a)) s = fma(prm.Bv[r * NU + a], pad[NP + r], s);
                    g2 += s * s;
                }
                gsum += sqrt(g2);
                double pn[NX];
#pragma unroll
                for (int CB = 0; CB < NX / 3; ++CB) {
                    double q[3];
#pragma unroll
                    for (int j = 0; j < 3; ++j) q[j] = grec[(size_t)(LR::LX_OFF + CB * 3 + j) * kRecStride];
#pragma unroll
                    for (int RB = 0; RB < NX / 3; ++RB) {
                        const int kind = blk_kind<KIND>(RB, CB);
                        if (kind == BK_ZERO) continue;
                        double m[9];
                        b3_load_blk<KIND>(RB, CB, rec, m);
#pragma unroll
                        for (int ii = 0; ii < 3; ++ii)
#pragma unroll
                            for (int j = 0; j < 3; ++j)
                                if (blk_nz(kind, ii, j)) q[j] = fma(m[3 * ii + j], pad[RB * 3 + ii], q[j]);
                    }
#pragma unroll
                    for (int j = 0; j < 3; ++j) pn[CB * 3 + j] = q[j];
                }
#pragma unroll
                for (int c = 0; c < NX; ++c) pad[c] = pn[c];
            }
        }
        __syncthreads();   // V(i), V_x(i) complete; nobody reads this stage's record any more
        if (tid == 0 && i > 0)
            b3_tma_load(rec_addr, group_base + (size_t)(i - 1) * kStageDoubles, SM::REC_BYTES, bar);
    }

    const bool owner = MS ? (warp == 0) : (warp == 1);   // who accumulated the gradient norm
    const bool was_running = alive && flags[lane] != 0;
    if (warp == 0 && was_running) {
        w.mu[b] = mu;
        w.delta[b] = delta;
    }
    if (owner && was_running) {
        const double g = gsum / (double)Nb;
        w.grad[b] = g;
        w.gradhist[(size_t)it * Bp + b] = g;
        int st = TRAJOPT_RUNNING;
        if (flag_bits & TRAJOPT_FLAG_REG_EXCEEDED) st = TRAJOPT_NO_DESCENT;
        else if (MS ? (g < prm.tol_grad && w.dnorm[b] < prm.tol_defect) : (g < prm.tol_grad)) st = TRAJOPT_CONVERGED;
        w.status[b] = st | flag_bits;
    }
#undef vlive
}

}  // namespace trajopt
