// Backward Riccati sweep for the 12-dimensional families when at most ONE group of 32 problems lands on an SM (a single
// solve, a strong-scaling shard of <= 148 x 32 problems, the last tail iterations of a batch): one CTA of SIX warps per
// group, lane = problem.  Same recursion and — operation for operation — the same arithmetic as k_backward3 / k_backward4
// (traopt_controller.py:2178-2321 / 2912-3068, gradient norms :2323-2349 / :3070-3093): every output element comes from
// the same building block (b3_compute_X, b3_block_cb/rb, the two halves of b3_gains ...) with the same inputs, so the three sweeps are
// bit-identical and the host picks one launch by launch (run_backward, host_impl.cuh).
//
// What a launch costs in this regime is the LENGTH of a stage's longest dependent instruction stream, not throughput.
// In the two- and four-warp sweeps every warp first repeats the serial part of the stage (Q_uu, its regularised 6x6
// Cholesky factor with six dependent rsqrt, y = L^-1 Q_u): ~2.5 k of the ~5.8 k statically scheduled cycles of a
// four-warp stage.  Here that part gets its own two warps and runs BESIDE the column work that does not need it:
//
//            until barrier (F)                           | until barrier (A)                          | until barrier (B)
//   warp c   X_c = V A[:,c], Q_x part of V_x, Q_ux[:,c],  | y, Y_c = L^-1 Q_ux, K_c, V_x,               | one off-diagonal block of V
//            l_xx + A[:,c]^T X_c of V(c,c)                | V(c,c) -= Y_c^T Y_c                          |
//   warp 4   Q_uu0, Cholesky (mu, delta state)            | SS: adjoint p                                | V(0,2)
//   warp 5   v = V_x + V d, Q_u, |Q_u|                    | y, k -> gains                                | V(1,2)
//
// (what a column warp finishes before (F) is pinned there with empty asm statements: left alone, ptxas sinks those
// register-only computations behind the barrier, i.e. back onto the critical path.)
// (F): the factor L, Q_u and mu go through shared memory.  (A): every X is formed and every Y is visible, V and V_x are
// overwritten in place.  X_2, X_3 and Y_2 also go through shared memory so that the six off-diagonal blocks are one per
// warp: V(0,3) warp 0, V(0,1) warp 1, V(2,3) warp 2 (= X_2^T A[:,3]), V(1,3) warp 3.  One CTA per SM leaves room to stage
// the WHOLE record of a stage (prefix + l_uu + l_x + l_xx, 29 KB) by TMA, double-buffered and two stages ahead: nothing in
// the stage loop reads global memory.
#pragma once
#include "backward4.cuh"

namespace trajopt {

constexpr int kB6Warps = 6;
constexpr int kB6Threads = kB6Warps * 32;

template <int KIND> struct B6Smem {
    using D = Dims<KIND>;
    using LR = LinRec<KIND>;
    static constexpr int NT = D::NX * (D::NX + 1) / 2;
    static constexpr int NYC = B3Smem<KIND>::NYC;         // the b3_* blocks address Y with this pitch (columns 0..5, 9..11)
    static constexpr int V_OFF = 0;
    static constexpr int VX_OFF = V_OFF + NT;
    static constexpr int Y_OFF = VX_OFF + D::NX;
    static constexpr int Y2_OFF = Y_OFF + D::NU * NYC;    // Y columns 6..8 (NU x 3), for the warps that form V(0,2), V(1,2)
    static constexpr int LC_OFF = Y2_OFF + D::NU * 3;     // Cholesky factor (NU x NU slots, lower triangle + inverse diagonal used)
    static constexpr int QU_OFF = LC_OFF + D::NU * D::NU;
    static constexpr int MU_OFF = QU_OFF + D::NU;         // mu of this stage's factorisation
    static constexpr int X2_OFF = MU_OFF + 1;             // X_2 = V A[:,2]  (NX x 3)
    static constexpr int X3_OFF = X2_OFF + D::NX * 3;
    static constexpr int REC_OFF = X3_OFF + D::NX * 3;    // two whole stage records
    static constexpr int DOUBLES = REC_OFF + 2 * LR::LEN;
    static constexpr int REC_BYTES = LR::LEN * 32 * 8;
    static constexpr size_t BYTES = (size_t)DOUBLES * 32 * 8 + 16 + 3 * 32 * sizeof(int);   // + 2 mbarriers + flags, live, bits
};

template <int KIND>
TO_DEV void b6_x_to_smem(double* __restrict__ Xs, const double (&X)[Dims<KIND>::NX][3]) {
#pragma unroll
    for (int r = 0; r < Dims<KIND>::NX; ++r)
#pragma unroll
        for (int j = 0; j < 3; ++j) Xs[(r * 3 + j) * 32] = X[r][j];
}
template <int KIND>
TO_DEV void b6_x_from_smem(const double* __restrict__ Xs, double (&X)[Dims<KIND>::NX][3]) {
#pragma unroll
    for (int r = 0; r < Dims<KIND>::NX; ++r)
#pragma unroll
        for (int j = 0; j < 3; ++j) X[r][j] = Xs[(r * 3 + j) * 32];
}

// b3_gains in two halves.  First half (needs no factor): the velocity rows of X folded into Q_ux[:, CB j] = Bv^T X_vel.
template <int KIND>
TO_DEV void b6_qux(const Params& prm, const double (&X)[Dims<KIND>::NX][3], double (&q)[3][Dims<KIND>::NU]) {
    using D = Dims<KIND>;
    constexpr int NX = D::NX, NP = D::NP, NU = D::NU, NV = NX - NP;
#pragma unroll
    for (int j = 0; j < 3; ++j)
#pragma unroll
        for (int a = 0; a < NU; ++a) {
            double s = 0.0;
#pragma unroll
            for (int r = 0; r < NV; ++r)
                if (bv_nz<KIND>(r, a)) s = fma(prm.Bv[r * NU + a], X[NP + r][j], s);
            q[j][a] = s;
        }
}
// Second half: + mu Bv^T A_vel, Y = L^-1 Q_ux, K = -L^-T Y (stored), yq = Y^T y — statement for statement b3_gains.
template <int KIND, int CB, bool Y_TO_SMEM, bool Y_TO_REGS>
TO_DEV void b6_gains(const Params& prm, const double* __restrict__ rec, double (&qq)[3][Dims<KIND>::NU],
                     const double (&Lc)[Dims<KIND>::NU * Dims<KIND>::NU], const double (&y)[Dims<KIND>::NU], double mu_used,
                     double (&yq)[3], double* __restrict__ Ys, double (&Yk)[Dims<KIND>::NU][3], double* __restrict__ Kout,
                     bool act) {
    using D = Dims<KIND>;
    constexpr int NX = D::NX, NP = D::NP, NU = D::NU, NB = NX / 3, NYC = B3Smem<KIND>::NYC;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        double (&q)[NU] = qq[j];
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
#pragma unroll
        for (int r = 0; r < NU; ++r) {   // Y[:, j] = L^-1 q
            double s = q[r];
#pragma unroll
            for (int k = 0; k < r; ++k) s = fma(-Lc[r * NU + k], q[k], s);
            q[r] = s * Lc[r * NU + r];
        }
        double kk[NU];                    // K[:, j] = -L^-T Y[:, j]
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

// b3_block_cb for a diagonal block in two halves: acc = l_xx + A[:, CB]^T X (needs no factor), then acc -= Y_CB^T Y_CB.
template <int KIND, int CB>
TO_DEV void b6_diag_ax(const Params& prm, const double* __restrict__ rec, const double (&X)[Dims<KIND>::NX][3], double (&acc)[3][3]) {
    constexpr int NX = Dims<KIND>::NX, NB = NX / 3;
#pragma unroll
    for (int ii = 0; ii < 3; ++ii)
#pragma unroll
        for (int j = 0; j < 3; ++j) acc[ii][j] = (ii <= j) ? b3_lxx<KIND>(prm, rec, CB * 3 + ii, CB * 3 + j) : 0.0;
#pragma unroll
    for (int KB = 0; KB < NB; ++KB) {
        const int kind = blk_kind<KIND>(KB, CB);
        if (kind == BK_ZERO) continue;
        double m[9];
        b3_load_blk<KIND>(KB, CB, rec, m);
#pragma unroll
        for (int k = 0; k < 3; ++k)
#pragma unroll
            for (int ii = 0; ii < 3; ++ii)
                if (blk_nz(kind, k, ii)) {
#pragma unroll
                    for (int j = 0; j < 3; ++j)
                        if (ii <= j) acc[ii][j] = fma(m[3 * k + ii], X[KB * 3 + k][j], acc[ii][j]);
                }
    }
}
template <int KIND, int CB, bool Y_REGS>
TO_DEV void b6_diag_yy(const double* __restrict__ Ys, const double (&Yk)[Dims<KIND>::NU][3], double (&acc)[3][3]) {
#pragma unroll
    for (int a = 0; a < Dims<KIND>::NU; ++a) {
        double c3[3];
#pragma unroll
        for (int t = 0; t < 3; ++t) c3[t] = b3_y<KIND, Y_REGS>(Ys, Yk, CB, a, t);
#pragma unroll
        for (int ii = 0; ii < 3; ++ii)
#pragma unroll
            for (int j = 0; j < 3; ++j)
                if (ii <= j) acc[ii][j] = fma(-c3[ii], c3[j], acc[ii][j]);
    }
}

// Column block CB up to barrier (A): returns the diagonal block and V_x[CB] in registers
template <int KIND, bool MS, int CB>
TO_DEV void b6_column(const Params& prm, const Work& w, const double* __restrict__ Vs, const double* __restrict__ Vxs,
                      double* __restrict__ Ys, const double* __restrict__ rec, const double* __restrict__ LCs,
                      const double* __restrict__ QUs, const double* __restrict__ MUs, const int* __restrict__ lives,
                      double* __restrict__ Kout, double (&X)[Dims<KIND>::NX][3], double (&Yk)[Dims<KIND>::NU][3],
                      double (&vx)[3], double (&hd)[3][3], bool& live, size_t Bp, int i, int b, int lane) {
    using D = Dims<KIND>;
    using LR = LinRec<KIND>;
    constexpr int NX = D::NX, NP = D::NP, NU = D::NU, NV = NX - NP;
    constexpr bool Y_REGS = (CB == 2);
    // ---- nothing here needs the factor: runs beside the Cholesky of warp 4 ----
#pragma unroll
    for (int j = 0; j < 3; ++j) vx[j] = rec[(LR::LX_OFF + CB * 3 + j) * kRecStride];
    b3_compute_X<KIND, CB, NP, NX>(Vs, rec, X);
    b3_qx_A<KIND, CB>(Vxs, rec, vx);
    if constexpr (MS) b3_qx_Xd<KIND, NP, NX>(rec, X, vx);
    b3_compute_X<KIND, CB, 0, NP>(Vs, rec, X);
    if constexpr (MS) b3_qx_Xd<KIND, 0, NP>(rec, X, vx);
    double qq[3][NU];
    b6_qux<KIND>(prm, X, qq);
    b6_diag_ax<KIND, CB>(prm, rec, X, hd);
    // pin what was just computed in front of the barrier: left alone, ptxas sinks these register-only computations behind
    // it (shorter live ranges), i.e. back onto the critical path of the stage
#pragma unroll
    for (int j = 0; j < 3; ++j) {
#pragma unroll
        for (int a = 0; a < NU; ++a) asm volatile("" : "+d"(qq[j][a]));
#pragma unroll
        for (int ii = 0; ii <= j; ++ii) asm volatile("" : "+d"(hd[ii][j]));
        asm volatile("" : "+d"(vx[j]));
    }
    __syncthreads();   // (F) L, Q_u, mu of this stage are in shared memory
    double Lc[NU * NU], y[NU], yq[3];
#pragma unroll
    for (int r = 0; r < NU; ++r)
#pragma unroll
        for (int c = 0; c <= r; ++c) Lc[r * NU + c] = LCs[(r * NU + c) * 32];
    const double mu_used = MUs[0];
    live = lives[lane] != 0;
#pragma unroll
    for (int r = 0; r < NU; ++r) {   // y = L^-1 Q_u
        double s = QUs[r * 32];
#pragma unroll
        for (int k = 0; k < r; ++k) s = fma(-Lc[r * NU + k], y[k], s);
        y[r] = s * Lc[r * NU + r];
    }
    b6_gains<KIND, CB, !Y_REGS, Y_REGS>(prm, rec, qq, Lc, y, mu_used, yq, Ys, Yk, Kout, live);
#pragma unroll
    for (int j = 0; j < 3; ++j) vx[j] -= yq[j];
    b6_diag_yy<KIND, CB, Y_REGS>(Ys, Yk, hd);
    if constexpr (CB >= 2) {
        if (prm.has_state_bounds) {
#pragma unroll
            for (int j = 0; j < 3; ++j) hd[j][j] += w.lxxv[soa(i, (CB - 2) * 3 + j, NV, (int)Bp, b)];
        }
    }
}

template <int KIND, bool MS, bool VH>
__global__ void __launch_bounds__(kB6Threads, 1) k_backward6(const Params prm, Work w, int it_arg) {
    static_assert(!on_so3(KIND), "the 6-warp sweep is for the 12-dimensional families");
    using D = Dims<KIND>;
    using LR = LinRec<KIND>;
    using SM = B6Smem<KIND>;
    constexpr int NX = D::NX, NP = D::NP, NU = D::NU, NV = NX - NP, F = LR::LEN;
    extern __shared__ __align__(128) double sm[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int tid = threadIdx.x;
    const int b = blockIdx.x * 32 + lane;
    const int N = prm.N, Np1 = N + 1;     // record layout / loop extent; the problem's own horizon is Nb <= N
    const int Nb = VH ? w.Nb[b < prm.Bp ? b : 0] : N;
    const size_t Bp = (size_t)prm.Bp;

    double* Vs = sm + SM::V_OFF * 32 + lane;
    double* Vxs = sm + SM::VX_OFF * 32 + lane;
    double* Ys = sm + SM::Y_OFF * 32 + lane;
    double* Y2s = sm + SM::Y2_OFF * 32 + lane;
    double* LCs = sm + SM::LC_OFF * 32 + lane;
    double* QUs = sm + SM::QU_OFF * 32 + lane;
    double* MUs = sm + SM::MU_OFF * 32 + lane;
    double* X2s = sm + SM::X2_OFF * 32 + lane;
    double* X3s = sm + SM::X3_OFF * 32 + lane;
    const uint32_t rec_addr0 = b3_smem_addr(sm + SM::REC_OFF * 32);
    const uint32_t bar0 = b3_smem_addr(sm + SM::DOUBLES * 32);
    int* flags = reinterpret_cast<int*>(sm + SM::DOUBLES * 32 + 2);
    int* lives = flags + 32;
    int* bits = flags + 64;

    bool act = (b < prm.B) && (w.status[b] == TRAJOPT_RUNNING);
    const int it = it_arg >= 0 ? it_arg : (act ? w.iters[b] : 0);   // < 0: per-slot iteration counts (trajopt_solve_stream)
    if (__ballot_sync(0xffffffffu, act) == 0u) return;              // same lanes in every warp: the whole CTA leaves

    const double* __restrict__ lin = w.lin;
    auto grec_of = [&](int s) { return lin + lsoa(s, 0, F, Np1, b); };
    const double* group_base = lin + lsoa(0, 0, F, Np1, b - lane);          // stage 0 of the group
    constexpr size_t kStageDoubles = (size_t)F * 32;

    // ---- cost / defect of the current trajectory (warp 0), terminal condition (warp 1): as in k_backward3 ----
    if (warp == 0) {
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
    } else if (warp == 1) {
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
    if (tid == 0) {
        b3_mbar_init(bar0, 1);
        b3_mbar_init(bar0 + 8, 1);
    }
    __syncthreads();
    act = flags[lane] != 0;
    if (__ballot_sync(0xffffffffu, act) == 0u) return;
    if (tid == 0) {   // records of the first two stages of the recursion
        b3_tma_load(rec_addr0, group_base + (size_t)(N - 1) * kStageDoubles, SM::REC_BYTES, bar0);
        if (N >= 2) b3_tma_load(rec_addr0 + SM::REC_BYTES, group_base + (size_t)(N - 2) * kStageDoubles, SM::REC_BYTES, bar0 + 8);
    }

    double pad[NX];   // SS, warp 4: adjoint variable p (:2339)
    if constexpr (!MS) {
        if (warp == 4) {
#pragma unroll
            for (int j = 0; j < NX; ++j) pad[j] = Vxs[j * 32];
        }
    }

    double mu = 0.0, delta = 0.0;   // Levenberg-Marquardt state of the problem: carried by warp 4
    if (warp == 4) {
        mu = w.mu[b < prm.Bp ? b : 0];
        delta = w.delta[b < prm.Bp ? b : 0];
    }
    double gsum = 0.0;      // MS: warp 5, SS: warp 4
    int flag_bits = 0;      // warp 4

    for (int i = N - 1; i >= 0; --i) {
        const int use = N - 1 - i, p = use & 1;
        const double* rec = sm + (SM::REC_OFF + p * F) * 32 + lane;
        b3_mbar_wait(bar0 + 8 * p, (uint32_t)((use >> 1) & 1));   // this stage's record is in shared memory
#define vlive (VH ? live : true)
        bool live = act && (!VH || i < Nb);   // warps 0..3, 5 replace it after (F) by what warp 4 decided
        double* Kout = w.gains + lsoa(i, 0, GainRec<KIND>::LEN, N, b);

        if (warp < 4) {
            double X[NX][3], Yk[NU][3], vx[3], hd[3][3], acc[3][3];
            if (warp == 0) {
                b6_column<KIND, MS, 0>(prm, w, Vs, Vxs, Ys, rec, LCs, QUs, MUs, lives, Kout, X, Yk, vx, hd, live, Bp, i, b, lane);
                __syncthreads();   // (A)
                b3_store_block<KIND, 0, 0>(Vs, hd, vlive);
                if (vlive) {
#pragma unroll
                    for (int j = 0; j < 3; ++j) Vxs[j * 32] = vx[j];
                }
                b6_x_from_smem<KIND>(X3s, X);
                b3_block_cb<KIND, 0, 3, false, false>(prm, rec, rec, X, Ys, Yk, acc);
                b3_store_block<KIND, 0, 3>(Vs, acc, vlive);
            } else if (warp == 1) {
                b6_column<KIND, MS, 1>(prm, w, Vs, Vxs, Ys, rec, LCs, QUs, MUs, lives, Kout, X, Yk, vx, hd, live, Bp, i, b, lane);
                __syncthreads();   // (A)
                b3_store_block<KIND, 1, 1>(Vs, hd, vlive);
                if (vlive) {
#pragma unroll
                    for (int j = 0; j < 3; ++j) Vxs[(3 + j) * 32] = vx[j];
                }
                b3_block_cb<KIND, 0, 1, false, false>(prm, rec, rec, X, Ys, Yk, acc);
                b3_store_block<KIND, 0, 1>(Vs, acc, vlive);
            } else if (warp == 2) {
                b6_column<KIND, MS, 2>(prm, w, Vs, Vxs, Ys, rec, LCs, QUs, MUs, lives, Kout, X, Yk, vx, hd, live, Bp, i, b, lane);
                b6_x_to_smem<KIND>(X2s, X);
#pragma unroll
                for (int a = 0; a < NU; ++a)
#pragma unroll
                    for (int t = 0; t < 3; ++t) Y2s[(a * 3 + t) * 32] = Yk[a][t];
                __syncthreads();   // (A)
                b3_store_block<KIND, 2, 2>(Vs, hd, vlive);
                if (vlive) {
#pragma unroll
                    for (int j = 0; j < 3; ++j) Vxs[(6 + j) * 32] = vx[j];
                }
                b3_block_rb<KIND, 2, 3, true, false>(prm, rec, rec, X, Ys, Yk, acc);
                b3_store_block<KIND, 2, 3>(Vs, acc, vlive);
            } else {
                b6_column<KIND, MS, 3>(prm, w, Vs, Vxs, Ys, rec, LCs, QUs, MUs, lives, Kout, X, Yk, vx, hd, live, Bp, i, b, lane);
                b6_x_to_smem<KIND>(X3s, X);
                __syncthreads();   // (A)
                b3_store_block<KIND, 3, 3>(Vs, hd, vlive);
                if (vlive) {
#pragma unroll
                    for (int j = 0; j < 3; ++j) Vxs[(9 + j) * 32] = vx[j];
                }
                b3_block_cb<KIND, 1, 3, false, false>(prm, rec, rec, X, Ys, Yk, acc);
                b3_store_block<KIND, 1, 3>(Vs, acc, vlive);
            }
        } else if (warp == 4) {
            // ---- Q_uu0 = l_uu + Bv^T V_vv Bv and its regularised Cholesky factor (identical to k_backward3) ----
            double Lc[NU * NU];   // lower factor, Lc[r][c] for c < r; the diagonal stores 1 / L_rr
            double mu_used;
            {
                double Quu0[NU * NU];
#pragma unroll
                for (int a = 0; a < NU; ++a)
#pragma unroll
                    for (int c = a; c < NU; ++c) {
                        double s = 2.0 * prm.R[a * NU + c];
                        if (a == c && prm.has_constraints) s += rec[(LR::LUU_OFF + a) * kRecStride];
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
                while (true) {   // regularisation loop (:2221-2246 / :2964-2991): Cholesky of Q_uu0 + mu B^T B
                    mu_used = mu;
                    bool pd = true;
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
                        if (prm.mu_max > 0.0 && mu >= prm.mu_max) {   // give up: the problem is stopped and flagged (see k_backward3)
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
            sfor<0, NU>([&](auto rc) {
                constexpr int r = decltype(rc)::value;
                sfor<0, r + 1>([&](auto cc) {
                    constexpr int c = decltype(cc)::value;
                    LCs[(r * NU + c) * 32] = Lc[r * NU + c];
                });
            });
            MUs[0] = mu_used;
            lives[lane] = live ? 1 : 0;
            __syncthreads();   // (F)
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
                    for (int j = 0; j < 3; ++j) q[j] = rec[(LR::LX_OFF + CB * 3 + j) * kRecStride];
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
            __syncthreads();   // (A)
            {
                double X[NX][3], Yk[NU][3], acc[3][3];
                b6_x_from_smem<KIND>(X2s, X);
#pragma unroll
                for (int a = 0; a < NU; ++a)
#pragma unroll
                    for (int t = 0; t < 3; ++t) Yk[a][t] = Y2s[(a * 3 + t) * 32];
                b3_block_cb<KIND, 0, 2, false, true>(prm, rec, rec, X, Ys, Yk, acc);
                b3_store_block<KIND, 0, 2>(Vs, acc, vlive);
            }
        } else {
            // ---- Q_u = l_u + Bv^T (V_x + V_xx d)_vel;  after (F): y = L^-1 Q_u;  k = -L^-T y -> gains ----
            double Qu[NU], qn = 0.0;
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
                    qn = sqrt(q);
                }
#pragma unroll
                for (int a = 0; a < NU; ++a) QUs[a * 32] = Qu[a];
            }
            __syncthreads();   // (F)
            live = lives[lane] != 0;
            if constexpr (MS) {
                if (vlive) gsum += qn;
            }
            {
                double Lc[NU * NU], y[NU], kk[NU];
#pragma unroll
                for (int r = 0; r < NU; ++r)
#pragma unroll
                    for (int c = 0; c <= r; ++c) Lc[r * NU + c] = LCs[(r * NU + c) * 32];
#pragma unroll
                for (int r = 0; r < NU; ++r) {
                    double s = Qu[r];
#pragma unroll
                    for (int k = 0; k < r; ++k) s = fma(-Lc[r * NU + k], y[k], s);
                    y[r] = s * Lc[r * NU + r];
                }
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
            __syncthreads();   // (A)
            {
                double X[NX][3], Yk[NU][3], acc[3][3];
                b6_x_from_smem<KIND>(X2s, X);
#pragma unroll
                for (int a = 0; a < NU; ++a)
#pragma unroll
                    for (int t = 0; t < 3; ++t) Yk[a][t] = Y2s[(a * 3 + t) * 32];
                b3_block_cb<KIND, 1, 2, false, true>(prm, rec, rec, X, Ys, Yk, acc);
                b3_store_block<KIND, 1, 2>(Vs, acc, vlive);
            }
        }
        __syncthreads();   // (B) V(i), V_x(i) complete; nobody reads this stage's record, X or Y any more
        if (tid == 0 && i >= 2)
            b3_tma_load(rec_addr0 + p * SM::REC_BYTES, group_base + (size_t)(i - 2) * kStageDoubles, SM::REC_BYTES, bar0 + 8 * p);
#undef vlive
    }

    const bool was_running = flags[lane] != 0;
    if (warp == 4) {
        if (was_running) {
            w.mu[b] = mu;
            w.delta[b] = delta;
        }
        bits[lane] = flag_bits;
    }
    __syncthreads();
    if (warp == (MS ? 5 : 4) && was_running) {   // who accumulated the gradient norm
        const int fb = bits[lane];
        const double g = gsum / (double)Nb;
        w.grad[b] = g;
        w.gradhist[(size_t)it * Bp + b] = g;
        int st = TRAJOPT_RUNNING;
        if (fb & TRAJOPT_FLAG_REG_EXCEEDED) st = TRAJOPT_NO_DESCENT;
        else if (MS ? (g < prm.tol_grad && w.dnorm[b] < prm.tol_defect) : (g < prm.tol_grad)) st = TRAJOPT_CONVERGED;
        w.status[b] = st | fb;
    }
}

}  // namespace trajopt
