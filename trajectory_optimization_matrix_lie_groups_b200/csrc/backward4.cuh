// Backward Riccati sweep for the 12-dimensional families when the GPU is NOT full (a shard of a strong-scaling run, the
// tail iterations of a batch, a single solve): one CTA of FOUR warps per group of 32 problems, lane = problem,
// warp c = 3-column block c of V_xx(i) = Q_xx - Y^T Y.  Same recursion and — operation for operation — the same arithmetic
// as k_backward3 (traopt_controller.py:2178-2321 / 2912-3068, gradient norms :2323-2349 / :3070-3093): every output
// element is produced by the same building block (b3_compute_X, b3_gains, b3_block_cb/rb ...) with the same inputs, so the
// two sweeps are bit-identical and the host may pick either one launch by launch.
//
// A sweep is a chain of N dependent stages and a lone warp issues one instruction every ~4.4 cycles (FP64 dependency
// latency + instruction fetch of a non-repeating stream), so what a launch costs when few problems are running is the
// length of the longest per-warp instruction stream of a stage.  Four warps instead of two cut it from ~3.5 k to ~2.1 k
// instructions (the replicated serial part — Q_uu, its Cholesky factor, y — is ~0.6 k of that), and the stage needs two
// CTA barriers instead of three:
//
//            before barrier (A)                               | after (A)                                 | (B)
//   warp 0   X0 = V A[:,0] -> Y0, K0, V(0,0), V_x[0:3];  k    | SS: adjoint p                              |
//   warp 1   X1            -> Y1, K1, V(1,1), V_x[3:6]        | V(0,1)                                     |
//   warp 2   X2            -> Y2, K2, V(2,2), V_x[6:9]        | V(0,2), V(1,2), V(2,3) (= X2^T A[:,3])     |
//   warp 3   X3            -> Y3, K3, V(3,3), V_x[9:12]       | V(0,3), V(1,3)                             |
//
// Every X is formed before (A), so V and V_x are overwritten in place right after it.  255 registers x 128 threads: two
// CTAs per SM, i.e. room for 296 groups = 9472 problems in one wave; the host uses this sweep below that (see
// run_backward in host_impl.cuh) and k_backward3 — all 512 groups of a 16 k batch resident at two warps each — above.
#pragma once
#include "backward3.cuh"

namespace trajopt {

// Section marks of a stage, as in k_backward3: named barriers (one per warp inside its column block, CTA-wide in the
// replicated serial part) that ptxas does not schedule across.  -DB4_NO_FENCE removes them (A/B).
#ifdef B4_NO_FENCE
#define B4_FENCE_COL(CB)
#define B4_FENCE_ALL()
#else
#define B4_FENCE_COL(CB) asm volatile("bar.sync %0, 32;" ::"n"(1 + (CB)) : "memory")
#define B4_FENCE_ALL() asm volatile("bar.sync 5, 128;" ::: "memory")
#endif
constexpr int kB4Warps = 4;
constexpr int kB4Threads = kB4Warps * 32;

// one column block: everything warp CB does between the replicated serial part and the end-of-stage barrier
template <int KIND, bool MS, bool VH, int CB>
TO_DEV void b4_column(const Params& prm, const Work& w, double* __restrict__ Vs, double* __restrict__ Vxs, double* __restrict__ Ys,
                      const double* __restrict__ rec, const double* __restrict__ grec,
                      const double (&Lc)[Dims<KIND>::NU * Dims<KIND>::NU], const double (&y)[Dims<KIND>::NU], double mu_used,
                      double* __restrict__ Kout, size_t Bp, bool live, int i, int b) {
    using D = Dims<KIND>;
    using LR = LinRec<KIND>;
    constexpr int NX = D::NX, NP = D::NP, NU = D::NU, NV = NX - NP;
    constexpr bool Y_REGS = (CB == 2);          // Y_2 is only ever needed by its own warp
#define vlive4 (VH ? live : true)
    double X[NX][3], Yk[NU][3], vx[3], yq[3], hd[3][3];
#pragma unroll
    for (int j = 0; j < 3; ++j) vx[j] = grec[(size_t)(LR::LX_OFF + CB * 3 + j) * kRecStride];
    b3_compute_X<KIND, CB, NP, NX>(Vs, rec, X);
    B4_FENCE_COL(CB);
    b3_gains<KIND, CB, !Y_REGS, Y_REGS>(prm, rec, X, Lc, y, mu_used, yq, Ys, Yk, Kout, live);
    B4_FENCE_COL(CB);
    b3_qx_A<KIND, CB>(Vxs, rec, vx);
    if constexpr (MS) b3_qx_Xd<KIND, NP, NX>(rec, X, vx);
    B4_FENCE_COL(CB);
    b3_compute_X<KIND, CB, 0, NP>(Vs, rec, X);
    if constexpr (MS) b3_qx_Xd<KIND, 0, NP>(rec, X, vx);
#pragma unroll
    for (int j = 0; j < 3; ++j) vx[j] -= yq[j];
    B4_FENCE_COL(CB);
    b3_block_cb<KIND, CB, CB, Y_REGS, Y_REGS>(prm, rec, grec, X, Ys, Yk, hd);
    if constexpr (CB >= 2) {
        if (prm.has_state_bounds) {
#pragma unroll
            for (int j = 0; j < 3; ++j) hd[j][j] += w.lxxv[soa(i, (CB - 2) * 3 + j, NV, (int)Bp, b)];
        }
    }
    __syncthreads();   // (A) every X is formed and Y_0, Y_1, Y_3 are visible: V and V_x may be overwritten
    b3_store_block<KIND, CB, CB>(Vs, hd, vlive4);
    if (vlive4) {
#pragma unroll
        for (int j = 0; j < 3; ++j) Vxs[(CB * 3 + j) * 32] = vx[j];
    }
    double acc[3][3];
    if constexpr (CB == 1) {
        b3_block_cb<KIND, 0, 1, false, false>(prm, rec, grec, X, Ys, Yk, acc);
        b3_store_block<KIND, 0, 1>(Vs, acc, vlive4);
    } else if constexpr (CB == 2) {
        b3_block_cb<KIND, 0, 2, false, true>(prm, rec, grec, X, Ys, Yk, acc);
        b3_store_block<KIND, 0, 2>(Vs, acc, vlive4);
        B4_FENCE_COL(CB);
        b3_block_cb<KIND, 1, 2, false, true>(prm, rec, grec, X, Ys, Yk, acc);
        b3_store_block<KIND, 1, 2>(Vs, acc, vlive4);
        B4_FENCE_COL(CB);
        b3_block_rb<KIND, 2, 3, true, false>(prm, rec, grec, X, Ys, Yk, acc);
        b3_store_block<KIND, 2, 3>(Vs, acc, vlive4);
    } else if constexpr (CB == 3) {
        b3_block_cb<KIND, 0, 3, false, false>(prm, rec, grec, X, Ys, Yk, acc);
        b3_store_block<KIND, 0, 3>(Vs, acc, vlive4);
        B4_FENCE_COL(CB);
        b3_block_cb<KIND, 1, 3, false, false>(prm, rec, grec, X, Ys, Yk, acc);
        b3_store_block<KIND, 1, 3>(Vs, acc, vlive4);
    }
#undef vlive4
}

template <int KIND, bool MS, bool VH>
__global__ void __launch_bounds__(kB4Threads, 2) k_backward4(const Params prm, Work w, int it_arg) {
    static_assert(!on_so3(KIND), "the 4-warp sweep is for the 12-dimensional families");
    using D = Dims<KIND>;
    using LR = LinRec<KIND>;
    using SM = B3Smem<KIND>;
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
    const double* rec = sm + SM::REC_OFF * 32 + lane;
    const uint32_t rec_addr = b3_smem_addr(sm + SM::REC_OFF * 32);
    const uint32_t bar = b3_smem_addr(sm + SM::DOUBLES * 32);
    int* flags = reinterpret_cast<int*>(sm + SM::DOUBLES * 32 + 2);

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
    if (tid == 0) b3_mbar_init(bar, 1);
    __syncthreads();
    act = flags[lane] != 0;
    if (__ballot_sync(0xffffffffu, act) == 0u) return;
    if (tid == 0) b3_tma_load(rec_addr, group_base + (size_t)(N - 1) * kStageDoubles, SM::REC_BYTES, bar);

    double pad[NX];   // SS, warp 0: adjoint variable p (:2339)
    if constexpr (!MS) {
        if (warp == 0) {
#pragma unroll
            for (int j = 0; j < NX; ++j) pad[j] = Vxs[j * 32];
        }
    }

    double mu = w.mu[b < prm.Bp ? b : 0], delta = w.delta[b < prm.Bp ? b : 0];
    double gsum = 0.0;
    int flag_bits = 0;
    uint32_t parity = 0;

    for (int i = N - 1; i >= 0; --i) {
        const double* __restrict__ grec = grec_of(i);
        bool live = act && (!VH || i < Nb);   // stages beyond a problem's horizon only keep the barriers company
#define vlive (VH ? live : true)
        if (i >= 2) {   // pull the record two stages ahead towards L2 (the next one is already on its way through TMA)
            const char* nxt = (const char*)(group_base + (size_t)(i - 2) * kStageDoubles);
#pragma unroll
            for (int t = 0; t < (F * 2 + kB4Threads - 1) / kB4Threads; ++t) {
                const int line = t * kB4Threads + tid;
                if (line < F * 2) asm volatile("prefetch.global.L2 [%0];" ::"l"(nxt + (size_t)line * 128));
            }
        }

        // ---- Q_uu0 = l_uu + Bv^T V_vv Bv and its regularised Cholesky factor (every warp; identical to k_backward3) ----
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
            B4_FENCE_ALL();
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

        B4_FENCE_ALL();
        b3_mbar_wait(bar, parity);   // the stage's record prefix is in shared memory
        parity ^= 1u;

        // ---- Q_u = l_u + Bv^T (V_x + V_xx d)_vel;  y = L^-1 Q_u;  k = -L^-T y (every warp; warp 0 stores k) ----
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
            if (warp == 0) {
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

        B4_FENCE_ALL();
        double* Kout = w.gains + lsoa(i, 0, GainRec<KIND>::LEN, N, b);
        if (warp == 0) {
            b4_column<KIND, MS, VH, 0>(prm, w, Vs, Vxs, Ys, rec, grec, Lc, y, mu_used, Kout, Bp, live, i, b);
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
        } else if (warp == 1) {
            b4_column<KIND, MS, VH, 1>(prm, w, Vs, Vxs, Ys, rec, grec, Lc, y, mu_used, Kout, Bp, live, i, b);
        } else if (warp == 2) {
            b4_column<KIND, MS, VH, 2>(prm, w, Vs, Vxs, Ys, rec, grec, Lc, y, mu_used, Kout, Bp, live, i, b);
        } else {
            b4_column<KIND, MS, VH, 3>(prm, w, Vs, Vxs, Ys, rec, grec, Lc, y, mu_used, Kout, Bp, live, i, b);
        }
        __syncthreads();   // (B) V(i), V_x(i) complete; nobody reads this stage's record or Y any more
        if (tid == 0 && i > 0)
            b3_tma_load(rec_addr, group_base + (size_t)(i - 1) * kStageDoubles, SM::REC_BYTES, bar);
#undef vlive
    }

    const bool was_running = flags[lane] != 0;
    if (warp == 0 && was_running) {
        w.mu[b] = mu;
        w.delta[b] = delta;
        const double g = gsum / (double)Nb;
        w.grad[b] = g;
        w.gradhist[(size_t)it * Bp + b] = g;
        int st = TRAJOPT_RUNNING;
        if (flag_bits & TRAJOPT_FLAG_REG_EXCEEDED) st = TRAJOPT_NO_DESCENT;
        else if (MS ? (g < prm.tol_grad && w.dnorm[b] < prm.tol_defect) : (g < prm.tol_grad)) st = TRAJOPT_CONVERGED;
        w.status[b] = st | flag_bits;
    }
}

}  // namespace trajopt
