// Backward Riccati sweep with in-loop regularisation: one problem per thread, one warp per CTA,
// the horizon runs sequentially inside the thread (traopt_controller.py:2178-2321 single shooting,
// :2912-3068 multiple shooting; gradient norms :2323-2349 / :3070-3093; J_opt / defect norm of the
// current trajectory :1935, :2504-2507).
//
// Everything is laid out so that no array is ever indexed with a run-time value: the 12x12 (6x6)
// algebra is organised in 3x3 blocks whose coordinates are template constants, so the working set
// stays in registers and the three per-problem matrices that do not fit live in shared memory,
// one column per thread with a stride of 32 doubles (conflict-free 64-bit accesses):
//   Vs   packed upper triangle of V_xx(i+1)                     NX(NX+1)/2
//   Vn   V_xx(i) under construction                             NX(NX+1)/2
//   Ys   Y = L^-1 Q_ux, columns of the earlier column blocks    NU (NX-3)
//   Vxs  V_x(i+1), then Q_x, then V_x(i)                        NX
// The dynamics Jacobian A = f_x is never assembled: its nonzero 3x3 blocks are read from the
// linearisation record when a column block needs them,
//   SE3/drone   [ a   0   c    0     ]        SO3   [ a  c ]      pendulum  [ a  c ]
//               [ b   a   e    c     ]              [ 0  h ]                [ l  h ]
//               [ 0   0   h11  h12   ]
//               [(s^) 0   vdt^ I-vdt^]
// Algebra per stage (equal to :3052-3060 and :2993-3004 up to rounding):
//   v = V_x + V_xx d;  Q_x = l_x + A^T v;  Q_u = l_u + B^T v;  X = V_xx A;
//   Q_xx = l_xx + A^T X;  Q_ux = B^T (X + mu A);  Q_uu = l_uu + B^T (V_xx + mu I) B = L L^T
//   Y = L^-1 Q_ux, y = L^-1 Q_u;  K = -L^-T Y, k = -L^-T y;
//   V_x(i)  = Q_x - Y^T y   (= Q_x + K^T Q_uu k + K^T Q_u + Q_ux^T k)
//   V_xx(i) = Q_xx - Y^T Y  (= sym(Q_xx + K^T Q_uu K + K^T Q_ux + Q_ux^T K)), symmetric by construction.
#pragma once
#include <type_traits>

#include "kernels.cuh"

namespace trajopt {

template <int Beg, int End, class F>
__device__ __forceinline__ void sfor(F&& f) {
    if constexpr (Beg < End) {
        f(std::integral_constant<int, Beg>{});
        sfor<Beg + 1, End>(f);
    }
}

enum { BK_ZERO = 0, BK_DENSE = 1, BK_SKEW = 2, BK_IMSKEW = 3 };

// kind and record offset of the 3x3 block (rb, cb) of A
template <int KIND> __host__ __device__ constexpr int blk_kind(int rb, int cb) {
    if (on_so3(KIND)) return (rb == 1 && cb == 0) ? (KIND == TRAJOPT_PEND ? BK_DENSE : BK_ZERO) : BK_DENSE;
    if (cb == 0) return (rb == 0 || rb == 1) ? BK_DENSE : ((rb == 3 && has_gravity(KIND)) ? BK_SKEW : BK_ZERO);
    if (cb == 1) return rb == 1 ? BK_DENSE : BK_ZERO;
    if (cb == 2) return rb == 3 ? BK_SKEW : BK_DENSE;
    return rb == 0 ? BK_ZERO : (rb == 3 ? BK_IMSKEW : BK_DENSE);
}
template <int KIND> __host__ __device__ constexpr int blk_off(int rb, int cb) {
    if (on_so3(KIND)) return cb == 0 ? (rb == 0 ? 0 : 27) : (rb == 0 ? 9 : 18);
    if (cb == 0) return rb == 0 ? 0 : (rb == 1 ? 9 : 57);
    if (cb == 1) return 0;
    if (cb == 2) return rb == 0 ? 18 : (rb == 1 ? 27 : (rb == 2 ? 36 : 54));
    return rb == 1 ? 18 : (rb == 2 ? 45 : 54);
}
__host__ __device__ constexpr bool blk_nz(int kind, int i, int j) {
    return kind == BK_DENSE || kind == BK_IMSKEW || (kind == BK_SKEW && i != j);
}

// rec points at field 0 of this thread's record of one stage; fields are Bp doubles apart
template <int KIND, int RB, int CB>
TO_DEV void load_blk(const double* __restrict__ rec, size_t Bp, double (&m)[9]) {
    constexpr int kind = blk_kind<KIND>(RB, CB), off = blk_off<KIND>(RB, CB);
    if constexpr (kind == BK_DENSE) {
#pragma unroll
        for (int t = 0; t < 9; ++t) m[t] = rec[(size_t)(off + t) * Bp];
    } else {
        const double v0 = rec[(size_t)off * Bp], v1 = rec[(size_t)(off + 1) * Bp], v2 = rec[(size_t)(off + 2) * Bp];
        if constexpr (kind == BK_SKEW) {
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

template <int KIND> __host__ __device__ constexpr int bwd_smem_doubles() {
    using D = Dims<KIND>;
    return D::NX * (D::NX + 1) + D::NU * (D::NX - 3) + D::NX;
}
// + the whole linearisation record of one stage of the warp's 32-problem group, double-buffered (TMA), + two mbarriers
template <int KIND, int LPW> constexpr size_t bwd_smem_bytes() {
    return ((size_t)bwd_smem_doubles<KIND>() * LPW + 2 * (size_t)LinRec<KIND>::LEN * 32) * sizeof(double) + 16;
}

// mbarrier / TMA bulk copy (cp.async.bulk) helpers shared by the sweeps and the rollouts
__device__ __forceinline__ uint32_t b3_smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void b3_mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void b3_tma_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void b3_mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
    }
}

// LPW = problems (active lanes) per warp: 32, or 16 to double the number of independent recursions in flight
// when the batch alone cannot fill the schedulers (each CTA is ONE warp of LPW threads)
template <int KIND, bool MS, int LPW>
__global__ void __launch_bounds__(LPW) k_backward(const Params prm, Work w, int it_arg) {
    using D = Dims<KIND>;
    using LR = LinRec<KIND>;
    constexpr int NX = D::NX, NP = D::NP, NU = D::NU, NV = NX - NP, F = LR::LEN;
    constexpr int NT = NX * (NX + 1) / 2;
    constexpr int NB = NX / 3;       // 3x3 blocks per side
    constexpr int NYC = NX - 3;      // columns of Y kept in shared memory
    extern __shared__ __align__(128) double sm[];
    const int lane = threadIdx.x;
    const int b = blockIdx.x * LPW + lane;
    // every lane stays for the warp-wide record staging; `act` = this lane's problem takes part in the sweep
    bool act = (b < prm.B) && (w.status[b] == TRAJOPT_RUNNING);
    const int it = it_arg >= 0 ? it_arg : (act ? w.iters[b] : 0);   // < 0: every slot counts its own iterations (trajopt_solve_stream)
    const int N = prm.N;                 // record layout stride; the problem's own horizon is Nb
    const int Nb = w.Nb[b < prm.Bp ? b : 0];
    const size_t Bp = (size_t)prm.Bp;
    double* Vs = sm + lane;
    double* Vn = sm + NT * LPW + lane;
    double* Ys = sm + 2 * NT * LPW + lane;
    double* Vxs = sm + (2 * NT + NU * NYC) * LPW + lane;
    double* recbuf = sm + (size_t)bwd_smem_doubles<KIND>() * LPW;      // 2 x [LEN][32]
    const uint32_t bar0 = b3_smem_addr(recbuf + 2 * (size_t)F * 32);
    const double* __restrict__ lin = w.lin;

    // ---- cost / defect of the current trajectory ----------------------------------------
    double Jcur = 0.0, dn = 0.0;
    if (act) {
        if constexpr (MS) {
            // J_new of the previous iteration: Python sum, left to right, + terminal (:2742-2754)
            double s = 0.0;
            for (int i = 0; i < Nb; ++i) s += w.Lc[(size_t)i * Bp + b];
            Jcur = s + w.Lc[(size_t)Nb * Bp + b];
            double q = 0.0;
            for (int i = 0; i < Nb; ++i) q += w.Dsq[(size_t)i * Bp + b];
            dn = sqrt(q);
            w.dnorm[b] = dn;
            if (it > 0) w.Jhist[(size_t)(it - 1) * Bp + b] = Jcur;
            w.defhist[(size_t)it * Bp + b] = dn;
        } else {
            Jcur = pairwise_sum(w.Lc + b, Bp, Nb + 1);     // J_opt = L.sum() (:1935)
        }
        w.J[b] = Jcur;
        if (!isfinite(Jcur)) {
            w.status[b] = TRAJOPT_NO_DESCENT | TRAJOPT_FLAG_NONFINITE;
            act = false;
        } else if (it >= prm.max_iters) {          // MS only: closing pass after the last rollout
            w.status[b] = TRAJOPT_MAX_ITER | (w.status[b] & ~15);
            act = false;
        }
    }
    const bool was_running = act;
    // stages the warp stays together for (per-problem horizons); nobody left: the warp leaves
    int Nmax = act ? Nb : 0;
#pragma unroll
    for (int o = LPW / 2; o > 0; o >>= 1) Nmax = max(Nmax, __shfl_xor_sync(0xffffffffu >> (32 - LPW), Nmax, o));
    if (Nmax == 0) return;

    // The record of a stage of the warp's group (one contiguous chunk, see LinRec) arrives by ONE TMA bulk copy, two stages
    // ahead, double-buffered: the recursion never waits on global memory for it.  (Until round 2 this kernel read the
    // record straight from L2 inside the stage: 2.9 us per stage of a 6-dimensional problem, most of it load latency.)
    const double* group_base = lin + lsoa(0, 0, F, N + 1, b - (b & 31));
    constexpr uint32_t kRecBytes = (uint32_t)F * 32 * 8;
    if (lane == 0) {
        b3_mbar_init(bar0, 1);
        b3_mbar_init(bar0 + 8, 1);
    }
    __syncwarp(0xffffffffu >> (32 - LPW));
    auto issue = [&](int stage, int buf) {
        if (lane == 0)
            b3_tma_load(b3_smem_addr(recbuf + (size_t)buf * F * 32), group_base + (size_t)stage * F * 32, kRecBytes, bar0 + 8 * buf);
    };
    issue(Nmax - 1, 0);
    if (Nmax > 1) issue(Nmax - 2, 1);

    // ---- terminal condition: V_x = l_x(N), V_xx = l_xx(N) --------------------------------
    if (act) {
        const double* rec = lin + lsoa(Nb, 0, F, N + 1, b);
#pragma unroll
        for (int j = 0; j < NX; ++j) Vxs[j * LPW] = rec[(size_t)(LR::LX_OFF + j) * kRecStride];
#pragma unroll
        for (int r = 0; r < NX; ++r)
#pragma unroll
            for (int c = r; c < NX; ++c) {
                double v;
                if (c < NP) v = rec[(size_t)(LR::LXX_OFF + tri_idx(NP, r, c)) * kRecStride];
                else if (r >= NP) v = 2.0 * prm.P2[(r - NP) * NV + (c - NP)] + ((r == c && prm.has_state_bounds) ? w.lxxv[soa(Nb, r - NP, NV, (int)Bp, b)] : 0.0);
                else v = 0.0;
                Vs[tri_idx(NX, r, c) * LPW] = v;
            }
    }
    double pad[NX];                      // SS: adjoint variable p (:2339)
    if constexpr (!MS) {
#pragma unroll
        for (int j = 0; j < NX; ++j) pad[j] = act ? Vxs[j * LPW] : 0.0;
    }

    double mu = 0.0, delta = 0.0;
    if (act) {
        mu = w.mu[b];
        delta = w.delta[b];
    }
    double gsum = 0.0;
    int flags = 0;

    for (int i = Nmax - 1; i >= 0; --i) {
        const int use = Nmax - 1 - i, buf = use & 1;
        b3_mbar_wait(bar0 + 8 * buf, (uint32_t)((use >> 1) & 1));      // this stage's record is in shared memory
        if (act && i < Nb) do {          // stages beyond this problem's horizon only keep the warp company
        const double* __restrict__ rec = recbuf + (size_t)buf * F * 32 + (b & 31);
        const BvStage<KIND> Bv(prm, rec, kRecStride);   // velocity rows of f_u: constant, or per stage (pendulum)

        // (1) v = V_x + V_xx d
        double v[NX];
#pragma unroll
        for (int r = 0; r < NX; ++r) v[r] = Vxs[r * LPW];
        if constexpr (MS) {
            double d[NX];
#pragma unroll
            for (int j = 0; j < NX; ++j) d[j] = rec[(size_t)(LR::D_OFF + j) * kRecStride];
#pragma unroll
            for (int r = 0; r < NX; ++r)
#pragma unroll
                for (int c = r; c < NX; ++c) {
                    const double t = Vs[tri_idx(NX, r, c) * LPW];
                    v[r] = fma(t, d[c], v[r]);
                    if (c != r) v[c] = fma(t, d[r], v[c]);
                }
        }

        // (2) Q_u = l_u + Bv^T v_vel ; gradient-norm term
        double Qu[NU], lu[NU];
#pragma unroll
        for (int a = 0; a < NU; ++a) lu[a] = rec[(size_t)(LR::LU_OFF + a) * kRecStride];
#pragma unroll
        for (int a = 0; a < NU; ++a) {
            double s = lu[a];
#pragma unroll
            for (int r = 0; r < NV; ++r)
                if (bv_nz<KIND>(r, a)) s = fma(Bv.get(r, a), v[NP + r], s);
            Qu[a] = s;
        }
        if constexpr (MS) {
            // g_t = L_u + F_u^T (V_x + V_xx^T d) = Q_u (:3090)
            double q = 0.0;
#pragma unroll
            for (int a = 0; a < NU; ++a) q += Qu[a] * Qu[a];
            gsum += sqrt(q);
        } else {
            // adjoint recursion of the SS gradient (:2343-2346): g = l_u + B^T p
            double g2 = 0.0;
#pragma unroll
            for (int a = 0; a < NU; ++a) {
                double s = lu[a];
#pragma unroll
                for (int r = 0; r < NV; ++r)
                    if (bv_nz<KIND>(r, a)) s = fma(Bv.get(r, a), pad[NP + r], s);
                g2 += s * s;
            }
            gsum += sqrt(g2);
        }

        // (3) Q_x = l_x + A^T v -> Vxs ;  SS: p = l_x + A^T p
        {
            double pn[NX];
            sfor<0, NB>([&](auto cbc) {
                constexpr int CB = decltype(cbc)::value;
                double q[3], qp[3];
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    q[j] = rec[(size_t)(LR::LX_OFF + CB * 3 + j) * kRecStride];
                    qp[j] = q[j];
                }
                sfor<0, NB>([&](auto rbc) {
                    constexpr int RB = decltype(rbc)::value;
                    constexpr int kind = blk_kind<KIND>(RB, CB);
                    if constexpr (kind != BK_ZERO) {
                        double m[9];
                        load_blk<KIND, RB, CB>(rec, kRecStride, m);
#pragma unroll
                        for (int ii = 0; ii < 3; ++ii)
#pragma unroll
                            for (int j = 0; j < 3; ++j)
                                if (blk_nz(kind, ii, j)) {
                                    q[j] = fma(m[3 * ii + j], v[RB * 3 + ii], q[j]);
                                    if constexpr (!MS) qp[j] = fma(m[3 * ii + j], pad[RB * 3 + ii], qp[j]);
                                }
                    }
                });
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    Vxs[(CB * 3 + j) * LPW] = q[j];
                    pn[CB * 3 + j] = qp[j];
                }
            });
            if constexpr (!MS) {
#pragma unroll
                for (int c = 0; c < NX; ++c) pad[c] = pn[c];
            }
        }

        // (4) Q_uu0 = l_uu + Bv^T V_vv Bv  (upper triangle)
        double Quu0[NU * NU];
        {
            double T[NV * NU];           // V_vv Bv
#pragma unroll
            for (int r = 0; r < NV; ++r) {
                double vr[NV];
#pragma unroll
                for (int k = 0; k < NV; ++k) vr[k] = Vs[sym_idx(NX, NP + r, NP + k) * LPW];
#pragma unroll
                for (int a = 0; a < NU; ++a) {
                    double s = 0.0;
#pragma unroll
                    for (int k = 0; k < NV; ++k)
                        if (bv_nz<KIND>(k, a)) s = fma(vr[k], Bv.get(k, a), s);
                    T[r * NU + a] = s;
                }
            }
#pragma unroll
            for (int a = 0; a < NU; ++a)
#pragma unroll
                for (int c = a; c < NU; ++c) {
                    double s = 2.0 * prm.R[a * NU + c];
                    if (a == c && prm.has_constraints) s += rec[(size_t)(LR::LUU_OFF + a) * kRecStride];
#pragma unroll
                    for (int r = 0; r < NV; ++r)
                        if (bv_nz<KIND>(r, a)) s = fma(Bv.get(r, a), T[r * NU + c], s);
                    Quu0[a * NU + c] = s;
                }
        }

        // (5) regularisation loop (:2221-2246 / :2964-2991): Cholesky of Q_uu0 + mu B^T B
        double Lc[NU * NU];              // lower factor, Lc[r][c] for c < r; diagonal stores 1/L_rr
        double mu_used;                  // the mu this stage's Q_ux / Q_uu are formed with (:2311-2313)
        bool gave_up = false;
        while (true) {
            mu_used = mu;
            bool pd = true;
            // (compile-time indices throughout: a rolled loop here would push Lc into local memory)
            sfor<0, NU>([&](auto cc) {
                constexpr int c = decltype(cc)::value;
                double dg = fma(mu_used, Bv.btb(prm, c, c), Quu0[c * NU + c]);
                sfor<0, c>([&](auto kc) {
                    constexpr int k = decltype(kc)::value;
                    dg = fma(-Lc[c * NU + k], Lc[c * NU + k], dg);
                });
                if (!(dg > 0.0)) pd = false;
                const double inv = rsqrt(dg);
                Lc[c * NU + c] = inv;
                sfor<c + 1, NU>([&](auto rc) {
                    constexpr int r = decltype(rc)::value;
                    double sacc = fma(mu_used, Bv.btb(prm, c, r), Quu0[c * NU + r]);
                    sfor<0, c>([&](auto kc) {
                        constexpr int k = decltype(kc)::value;
                        sacc = fma(-Lc[r * NU + k], Lc[c * NU + k], sacc);
                    });
                    Lc[r * NU + c] = sacc * inv;
                });
            });
            if (!pd) {
                delta = fmax(1.0, delta) * prm.delta0;
                mu = fmax(prm.mu_min, mu * delta);
                if (prm.mu_max > 0.0 && mu >= prm.mu_max) { gave_up = true; break; }
            } else {
                delta = fmin(1.0, delta) / prm.delta0;
                mu *= delta;
                if (mu <= prm.mu_min) mu = 0.0;
                break;
            }
        }
        if (gave_up) {
            // The reference warns and carries on with a non-PD Q_uu (:2238-2240); such a problem
            // has already diverged.  It is stopped here and flagged.
            flags |= TRAJOPT_FLAG_REG_EXCEEDED;
            act = false;
            break;
        }

        // (6) y = L^-1 Q_u ;  k = -L^-T y
        double y[NU];
#pragma unroll
        for (int r = 0; r < NU; ++r) {
            double s = Qu[r];
#pragma unroll
            for (int k = 0; k < r; ++k) s = fma(-Lc[r * NU + k], y[k], s);
            y[r] = s * Lc[r * NU + r];
        }
        {
            double kk[NU];
#pragma unroll
            for (int r = NU - 1; r >= 0; --r) {
                double s = y[r];
#pragma unroll
                for (int k = r + 1; k < NU; ++k) s = fma(-Lc[k * NU + r], kk[k], s);
                kk[r] = s * Lc[r * NU + r];
            }
#pragma unroll
            for (int a = 0; a < NU; ++a) w.gains[lsoa(i, GainRec<KIND>::KFF_OFF + a, GainRec<KIND>::LEN, N, b)] = -kk[a];
        }

        // (7) column blocks of 3
        sfor<0, NB>([&](auto cbc) {
            constexpr int CB = decltype(cbc)::value;
            // X = V_xx A[:, CB]
            double X[NX][3];
#pragma unroll
            for (int r = 0; r < NX; ++r)
#pragma unroll
                for (int j = 0; j < 3; ++j) X[r][j] = 0.0;
            double Qux[NU][3];
#pragma unroll
            for (int a = 0; a < NU; ++a)
#pragma unroll
                for (int j = 0; j < 3; ++j) Qux[a][j] = 0.0;
            sfor<0, NB>([&](auto kbc) {
                constexpr int KB = decltype(kbc)::value;
                constexpr int kind = blk_kind<KIND>(KB, CB);
                if constexpr (kind != BK_ZERO) {
                    double m[9];
                    load_blk<KIND, KB, CB>(rec, kRecStride, m);
#pragma unroll
                    for (int r = 0; r < NX; ++r)
#pragma unroll
                        for (int ii = 0; ii < 3; ++ii) {
                            const double t = Vs[sym_idx(NX, r, KB * 3 + ii) * LPW];
#pragma unroll
                            for (int j = 0; j < 3; ++j)
                                if (blk_nz(kind, ii, j)) X[r][j] = fma(t, m[3 * ii + j], X[r][j]);
                        }
                    // regularisation term of Q_ux: mu Bv^T A_vel (:2311-2312); mu is 0 after the first stages
                    if constexpr (KB * 3 >= NP) {
                        if (mu_used != 0.0) {
#pragma unroll
                            for (int a = 0; a < NU; ++a)
#pragma unroll
                                for (int ii = 0; ii < 3; ++ii)
                                    if (bv_nz<KIND>(KB * 3 + ii - NP, a)) {
                                        const double bm = mu_used * Bv.get(KB * 3 + ii - NP, a);
#pragma unroll
                                        for (int j = 0; j < 3; ++j)
                                            if (blk_nz(kind, ii, j)) Qux[a][j] = fma(bm, m[3 * ii + j], Qux[a][j]);
                                    }
                        }
                    }
                }
            });
            // Q_ux[:, CB] += Bv^T X_vel
#pragma unroll
            for (int a = 0; a < NU; ++a)
#pragma unroll
                for (int r = 0; r < NV; ++r)
                    if (bv_nz<KIND>(r, a)) {
#pragma unroll
                        for (int j = 0; j < 3; ++j) Qux[a][j] = fma(Bv.get(r, a), X[NP + r][j], Qux[a][j]);
                    }
            // Y = L^-1 Q_ux (in place), K = -L^-T Y
#pragma unroll
            for (int j = 0; j < 3; ++j) {
#pragma unroll
                for (int r = 0; r < NU; ++r) {
                    double s = Qux[r][j];
#pragma unroll
                    for (int k = 0; k < r; ++k) s = fma(-Lc[r * NU + k], Qux[k][j], s);
                    Qux[r][j] = s * Lc[r * NU + r];
                }
                double kk[NU];
#pragma unroll
                for (int r = NU - 1; r >= 0; --r) {
                    double s = Qux[r][j];
#pragma unroll
                    for (int k = r + 1; k < NU; ++k) s = fma(-Lc[k * NU + r], kk[k], s);
                    kk[r] = s * Lc[r * NU + r];
                }
#pragma unroll
                for (int a = 0; a < NU; ++a) w.gains[lsoa(i, a * NX + CB * 3 + j, GainRec<KIND>::LEN, N, b)] = -kk[a];
                // V_x(i)[c] = Q_x[c] - Y[:,c]^T y
                double s = Vxs[(CB * 3 + j) * LPW];
#pragma unroll
                for (int a = 0; a < NU; ++a) s = fma(-Qux[a][j], y[a], s);
                Vxs[(CB * 3 + j) * LPW] = s;
            }
            // V_xx(i)[RB block, CB block] = l_xx + A[:, RB]^T X - Y[:, RB]^T Y[:, CB],  RB <= CB
            sfor<0, CB + 1>([&](auto rbc) {
                constexpr int RB = decltype(rbc)::value;
                double acc[3][3];
#pragma unroll
                for (int ii = 0; ii < 3; ++ii)
#pragma unroll
                    for (int j = 0; j < 3; ++j) {
                        const int r = RB * 3 + ii, c = CB * 3 + j;
                        double s = 0.0;
                        if (r <= c) {
                            if (c < NP) s = rec[(size_t)(LR::LXX_OFF + tri_idx(NP, r, c)) * kRecStride];
                            else if (r >= NP) s = 2.0 * prm.W2[(r - NP) * NV + (c - NP)] + ((r == c && prm.has_state_bounds) ? w.lxxv[soa(i, r - NP, NV, (int)Bp, b)] : 0.0);
                        }
                        acc[ii][j] = s;
                    }
                sfor<0, NB>([&](auto kbc) {
                    constexpr int KB = decltype(kbc)::value;
                    constexpr int kind = blk_kind<KIND>(KB, RB);
                    if constexpr (kind != BK_ZERO) {
                        double m[9];
                        load_blk<KIND, KB, RB>(rec, kRecStride, m);
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
                });
#pragma unroll
                for (int a = 0; a < NU; ++a) {
                    double yr[3];
#pragma unroll
                    for (int ii = 0; ii < 3; ++ii) {
                        if constexpr (RB == CB) yr[ii] = Qux[a][ii];
                        else yr[ii] = Ys[(a * NYC + RB * 3 + ii) * LPW];
                    }
#pragma unroll
                    for (int ii = 0; ii < 3; ++ii)
#pragma unroll
                        for (int j = 0; j < 3; ++j)
                            if (RB < CB || ii <= j) acc[ii][j] = fma(-yr[ii], Qux[a][j], acc[ii][j]);
                }
#pragma unroll
                for (int ii = 0; ii < 3; ++ii)
#pragma unroll
                    for (int j = 0; j < 3; ++j)
                        if (RB < CB || ii <= j) Vn[tri_idx(NX, RB * 3 + ii, CB * 3 + j) * LPW] = acc[ii][j];
            });
            if constexpr (CB * 3 < NYC) {
#pragma unroll
                for (int a = 0; a < NU; ++a)
#pragma unroll
                    for (int j = 0; j < 3; ++j) Ys[(a * NYC + CB * 3 + j) * LPW] = Qux[a][j];
            }
        });
        { double* t = Vs; Vs = Vn; Vn = t; }
        } while (0);
        __syncwarp(0xffffffffu >> (32 - LPW));          // every lane is done with this buffer
        if (i >= 2) issue(i - 2, buf);
    }

    if (!was_running) return;
    w.mu[b] = mu;
    w.delta[b] = delta;
    const double g = gsum / (double)Nb;
    w.grad[b] = g;
    w.gradhist[(size_t)it * Bp + b] = g;
    int st = TRAJOPT_RUNNING;
    if (flags & TRAJOPT_FLAG_REG_EXCEEDED) st = TRAJOPT_NO_DESCENT;
    else if (MS ? (g < prm.tol_grad && dn < prm.tol_defect) : (g < prm.tol_grad)) st = TRAJOPT_CONVERGED;
    w.status[b] = st | flags;
}

}  // namespace trajopt
