// DDP/iLQR kernels.  One problem per thread; the time recursion runs sequentially inside the
// thread; the batch is the parallel dimension (SoA, problem index fastest => every global
// access of a warp is one coalesced 256-byte row).
//
//   k_linearize   stage-parallel: thread = (stage, problem).  traopt_controller.py:2098-2176 / 2823-2910
//   k_backward    problem-parallel Riccati sweep with in-loop regularisation.   :2178-2321 / 2912-3068,
//                 gradient norms :2323-2349 / 3070-3093, cost / defect reductions :1935, 2504-2507
//   k_forward     problem-parallel rollout (one step size per thread).          :2030-2082 / 2641-2740
//   k_init_*      initial rollout :2015-2028 / initial guess :3123-3136
//   k_ls_*        line-search bookkeeping (accept first J_new < J_opt)          :1972-1990
//   k_al_update   multiplier / penalty update                                   :3270-3290
#pragma once
#include "model.cuh"

namespace trajopt {

struct Work {
    double* X[2];        // [N+1][NS][Bp] x 2 (current / candidate, selected per problem by sel)
    double* U[2];        // [N][NU][Bp]  x 2
    int* sel;            // [Bp] which buffer holds the current trajectory
    const double* ref;   // [N+1][RefRow]
    double* lin;         // [N+1][LinRec::LEN][Bp]
    double* Lc;          // [N+1][Bp] stage costs of the current trajectory
    double* Dsq;         // [N][Bp]   squared defect norm per stage
    double* kff;         // [N][NU][Bp]
    double* Kfb;         // [N][NU*NX][Bp]
    double* J;           // [Bp] cost of the current trajectory (J_opt)
    double* Jcand;       // [n_alphas][Bp] candidate costs of the line search
    double* grad;        // [Bp]
    double* dnorm;       // [Bp]
    double* mu;          // [Bp] Levenberg-Marquardt state, persists across stages and iterations
    double* delta;       // [Bp]
    int* iters;          // [Bp] completed iterations (= len(J_hist))
    int* status;         // [Bp]
    int* ls_state;       // [Bp] line search: -2 idle, -1 pending (needs more step sizes), >=0 accepted index
    double* Jhist;       // [max_iters][Bp]
    double* gradhist;    // [max_iters+1][Bp]
    double* defhist;     // [max_iters+1][Bp]
    int* alphahist;      // [max_iters][Bp]
    double* x0;          // [NS][Bp]
    const double* us_init;   // device, layout per us_mode (may be null)
    int us_mode;
    // augmented Lagrangian
    double* lam;         // [N+1][2NU][Bp]
    double* imu;         // [N+1][2NU][Bp]
    double* al_mu;       // [Bp]
    int* al_outer;       // [Bp]
    double* al_viol;     // [Bp]
    int* al_done;        // [Bp]
    int* counters;       // [0] running problems, [1] pending line searches, [2] AL problems not converged
};

constexpr int kBlock = 32;

// ------------------------------------------------------------------------------------------
// Structure of the dynamics Jacobian A = f_x as 3x3 blocks (zero blocks are never touched):
//   SE3/drone   [ a   0   c    0  ]        SO3   [ a  c ]
//               [ b   a   e    c  ]              [ 0  h ]
//               [ 0   0   h11  h12]
//               [(s)  0   vdt^ I-vdt^]
// ------------------------------------------------------------------------------------------
template <int KIND> struct AMat;

__host__ __device__ constexpr double skew_sign(int i, int j) {
    // coefficient of v[k] in skew(v)(i,j), k = 3 - i - j
    return (i == j) ? 0.0 : (((j - i + 3) % 3 == 1) ? -1.0 : 1.0);
}

template <> struct AMat<TRAJOPT_SO3> {
    double a[9], c[9], h[9];
    static __host__ __device__ constexpr bool nz(int r, int cc) { return !(r >= 3 && cc < 3); }
    TO_DEV double get(int r, int cc) const {
        const int i = r % 3, j = cc % 3;
        if (r < 3 && cc < 3) return a[3 * i + j];
        if (r < 3) return c[3 * i + j];
        return h[3 * i + j];
    }
    TO_DEV void load(const double* __restrict__ lin, int stage, int Bp, int b) {
        constexpr int F = LinRec<TRAJOPT_SO3>::LEN;
#pragma unroll
        for (int i = 0; i < 9; ++i) {
            a[i] = lin[soa(stage, i, F, Bp, b)];
            c[i] = lin[soa(stage, 9 + i, F, Bp, b)];
            h[i] = lin[soa(stage, 18 + i, F, Bp, b)];
        }
    }
};

template <int KIND> struct AMat {   // SE3 and DRONE
    double a[9], b[9], c[9], e[9], h11[9], h12[9], vdt[3], s[3];
    static __host__ __device__ constexpr bool nz(int r, int cc) {
        const int br = r / 3, bc = cc / 3, i = r % 3, j = cc % 3;
        if (br == 0) return bc == 0 || bc == 2;
        if (br == 1) return true;
        if (br == 2) return bc >= 2;
        // br == 3
        if (bc == 0) return KIND == TRAJOPT_DRONE && i != j;
        if (bc == 1) return false;
        if (bc == 2) return i != j;
        return true;
    }
    TO_DEV double get(int r, int cc) const {
        const int br = r / 3, bc = cc / 3, i = r % 3, j = cc % 3;
        if (br == 0) return (bc == 0) ? a[3 * i + j] : c[3 * i + j];
        if (br == 1) return (bc == 0) ? b[3 * i + j] : (bc == 1) ? a[3 * i + j] : (bc == 2) ? e[3 * i + j] : c[3 * i + j];
        if (br == 2) return (bc == 2) ? h11[3 * i + j] : h12[3 * i + j];
        const int k = (3 - i - j) % 3;
        if (bc == 0) return skew_sign(i, j) * s[k];
        if (bc == 2) return skew_sign(i, j) * vdt[k];
        return ((i == j) ? 1.0 : 0.0) - skew_sign(i, j) * vdt[k];
    }
    TO_DEV void load(const double* __restrict__ lin, int stage, int Bp, int bb) {
        constexpr int F = LinRec<KIND>::LEN;
#pragma unroll
        for (int i = 0; i < 9; ++i) {
            a[i] = lin[soa(stage, i, F, Bp, bb)];
            b[i] = lin[soa(stage, 9 + i, F, Bp, bb)];
            c[i] = lin[soa(stage, 18 + i, F, Bp, bb)];
            e[i] = lin[soa(stage, 27 + i, F, Bp, bb)];
            h11[i] = lin[soa(stage, 36 + i, F, Bp, bb)];
            h12[i] = lin[soa(stage, 45 + i, F, Bp, bb)];
        }
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            vdt[i] = lin[soa(stage, 54 + i, F, Bp, bb)];
            s[i] = (KIND == TRAJOPT_DRONE) ? lin[soa(stage, 57 + i, F, Bp, bb)] : 0.0;
        }
    }
};

// Velocity rows of f_u: Bv = Jinv Pu dt  (NV x NU), compile-time sparsity
template <int KIND> __host__ __device__ constexpr bool bv_nz(int r, int c) {
    if (KIND == TRAJOPT_SO3) return true;
    if (KIND == TRAJOPT_SE3) return (r < 3 && c < 3) || (r >= 3 && r == c);
    return (r < 3 && c < 3) || (r == 5 && c == 3);   // drone: torques + body-z thrust
}

__host__ __device__ constexpr int tri_idx(int n, int r, int c) {   // packed upper triangle, r <= c
    return r * n - (r * (r - 1)) / 2 + (c - r);
}
__host__ __device__ constexpr int sym_idx(int n, int r, int c) { return r <= c ? tri_idx(n, r, c) : tri_idx(n, c, r); }

// ------------------------------------------------------------------------------------------
// Initialisation
// ------------------------------------------------------------------------------------------
template <int KIND>
__device__ __forceinline__ double us_init_value(const Work& w, const Params& prm, int stage, int j, int b) {
    constexpr int NU = Dims<KIND>::NU;
    if (w.us_mode == 0 || w.us_init == nullptr) return 0.0;
    if (w.us_mode == 1) return w.us_init[(size_t)stage * NU + j];
    return w.us_init[((size_t)b * prm.N + stage) * NU + j];
}

// x0 arrives problem-major [B][NS]; keep a normalised SoA copy
template <int KIND>
__global__ void k_load_x0(const Params prm, Work w, const double* __restrict__ x0_aos) {
    constexpr int NS = Dims<KIND>::NS;
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= prm.Bp) return;
    double v[NS];
    if (b < prm.B) {
#pragma unroll
        for (int j = 0; j < NS; ++j) v[j] = x0_aos[(size_t)b * NS + j];
        quat_normalize(v);
    } else {   // padding lanes: identity pose, zero velocity (never run, but keep memory finite)
#pragma unroll
        for (int j = 0; j < NS; ++j) v[j] = 0.0;
        v[3] = 1.0;
    }
#pragma unroll
    for (int j = 0; j < NS; ++j) w.x0[(size_t)j * prm.Bp + b] = v[j];
}

// reset per-fit state (regulariser :1899-1900, histories)
__global__ void k_reset(const Params prm, Work w) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= prm.Bp) return;
    w.sel[b] = 0;
    w.mu[b] = 1.0;
    w.delta[b] = prm.delta0;
    w.iters[b] = 0;
    w.status[b] = (b < prm.B) ? TRAJOPT_RUNNING : TRAJOPT_MAX_ITER;
    w.ls_state[b] = -2;
    w.J[b] = 0.0;
    w.grad[b] = 0.0;
    w.dnorm[b] = 0.0;
}

// only the problems whose AL outer loop is still running are restarted
__global__ void k_reset_al_inner(const Params prm, Work w) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= prm.Bp) return;
    if (b < prm.B && !w.al_done[b]) {
        w.sel[b] = 0;
        w.mu[b] = 1.0;
        w.delta[b] = prm.delta0;
        w.iters[b] = 0;
        w.status[b] = TRAJOPT_RUNNING;
        w.ls_state[b] = -2;
    }
}

// Multiple shooting initial guess: shooting nodes = reference (:3123-3136); controls = us_init
template <int KIND>
__global__ void k_init_ms(const Params prm, Work w, bool only_running) {
    constexpr int NS = Dims<KIND>::NS, NU = Dims<KIND>::NU;
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    const int stage = blockIdx.y;
    if (b >= prm.Bp) return;
    if (only_running && w.status[b] != TRAJOPT_RUNNING) return;
    State<KIND> s;
    if (stage == 0) {
#pragma unroll
        for (int j = 0; j < 4; ++j) s.q[j] = w.x0[(size_t)j * prm.Bp + b];
        if constexpr (KIND != TRAJOPT_SO3) {
#pragma unroll
            for (int j = 0; j < 3; ++j) s.p[j] = w.x0[(size_t)(4 + j) * prm.Bp + b];
#pragma unroll
            for (int j = 0; j < 6; ++j) s.xi[j] = w.x0[(size_t)(7 + j) * prm.Bp + b];
        } else {
#pragma unroll
            for (int j = 0; j < 3; ++j) s.xi[j] = w.x0[(size_t)(4 + j) * prm.Bp + b];
        }
    } else {
        load_ref_state<KIND>(w.ref, stage, s);
    }
    store_state<KIND>(w.X[0], stage, prm.Bp, b, s);
    if (stage < prm.N) {
#pragma unroll
        for (int j = 0; j < NU; ++j) w.U[0][soa(stage, j, NU, prm.Bp, b)] = us_init_value<KIND>(w, prm, stage, j, b);
    }
    (void)NS;
}

// Single shooting initial rollout (:2015-2028)
template <int KIND>
__global__ void __launch_bounds__(kBlock) k_init_ss(const Params prm, Work w) {
    constexpr int NU = Dims<KIND>::NU;
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= prm.Bp) return;
    State<KIND> x, xn;
#pragma unroll
    for (int j = 0; j < 4; ++j) x.q[j] = w.x0[(size_t)j * prm.Bp + b];
    if constexpr (KIND != TRAJOPT_SO3) {
#pragma unroll
        for (int j = 0; j < 3; ++j) x.p[j] = w.x0[(size_t)(4 + j) * prm.Bp + b];
#pragma unroll
        for (int j = 0; j < 6; ++j) x.xi[j] = w.x0[(size_t)(7 + j) * prm.Bp + b];
    } else {
#pragma unroll
        for (int j = 0; j < 3; ++j) x.xi[j] = w.x0[(size_t)(4 + j) * prm.Bp + b];
    }
    store_state<KIND>(w.X[0], 0, prm.Bp, b, x);
    for (int i = 0; i < prm.N; ++i) {
        double u[NU];
#pragma unroll
        for (int j = 0; j < NU; ++j) {
            u[j] = us_init_value<KIND>(w, prm, i, j, b);
            w.U[0][soa(i, j, NU, prm.Bp, b)] = u[j];
        }
        dyn_step<KIND>(prm, x, u, xn);
        store_state<KIND>(w.X[0], i + 1, prm.Bp, b, xn);
        x = xn;
    }
}

// ------------------------------------------------------------------------------------------
// Stage-parallel linearisation.  grid = (ceil(Bp/128), N+1), thread = (problem, stage)
// ------------------------------------------------------------------------------------------
template <int KIND, bool MS>
__global__ void __launch_bounds__(128) k_linearize(const Params prm, Work w) {
    using D = Dims<KIND>;
    using LR = LinRec<KIND>;
    constexpr int NX = D::NX, NU = D::NU, F = LR::LEN;
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    const int stage = blockIdx.y;
    if (b >= prm.B || w.status[b] != TRAJOPT_RUNNING) return;
    const int Bp = prm.Bp;
    const double* X = w.X[w.sel[b]];
    const double* U = w.U[w.sel[b]];
    const double* refrow = w.ref + (size_t)stage * RefRow<KIND>::N;
    double* out = w.lin;

    State<KIND> x;
    load_state<KIND>(X, stage, Bp, b, x);
    const bool terminal = (stage == prm.N);

    double lx[NX], lxx[LR::LXX_LEN];
    double val = cost_expand<KIND>(prm, x, refrow, terminal, lx, lxx);

    if (!terminal) {
        double u[NU];
#pragma unroll
        for (int j = 0; j < NU; ++j) u[j] = U[soa(stage, j, NU, Bp, b)];
        // control cost u^T R u, l_u = 2 R u
        double lu[NU], cu = 0.0;
#pragma unroll
        for (int i = 0; i < NU; ++i) {
            double s = 0.0;
#pragma unroll
            for (int j = 0; j < NU; ++j) s += prm.R[i * NU + j] * u[j];
            lu[i] = 2.0 * s;
            cu += u[i] * s;
        }
        val += cu;
        double luu_add[NU];
#pragma unroll
        for (int j = 0; j < NU; ++j) luu_add[j] = 0.0;
        if (prm.has_constraints) {
            double lam[2 * NU], imu[2 * NU], lu_add[NU];
#pragma unroll
            for (int j = 0; j < 2 * NU; ++j) {
                lam[j] = w.lam[soa(stage, j, 2 * NU, Bp, b)];
                imu[j] = w.imu[soa(stage, j, 2 * NU, Bp, b)];
            }
            val += al_terms<NU>(prm, u, lam, imu, lu_add, luu_add);
#pragma unroll
            for (int j = 0; j < NU; ++j) lu[j] += lu_add[j];
        }
#pragma unroll
        for (int j = 0; j < NU; ++j) {
            out[soa(stage, LR::LU_OFF + j, F, Bp, b)] = lu[j];
            out[soa(stage, LR::LUU_OFF + j, F, Bp, b)] = luu_add[j];
        }
        // dynamics Jacobian
        double rec[LR::A_LEN];
        dyn_jacobian<KIND>(prm, x, rec);
#pragma unroll
        for (int j = 0; j < LR::A_LEN; ++j) out[soa(stage, LR::A_OFF + j, F, Bp, b)] = rec[j];
        // defect against the next shooting node
        if constexpr (MS) {
            State<KIND> fx, xnext;
            dyn_step<KIND>(prm, x, u, fx);
            load_state<KIND>(X, stage + 1, Bp, b, xnext);
            double d[NX], dsq = 0.0;
            defect<KIND>(fx, xnext, d);
#pragma unroll
            for (int j = 0; j < NX; ++j) {
                out[soa(stage, LR::D_OFF + j, F, Bp, b)] = d[j];
                dsq += d[j] * d[j];
            }
            w.Dsq[(size_t)stage * Bp + b] = dsq;
        }
    }
#pragma unroll
    for (int j = 0; j < NX; ++j) out[soa(stage, LR::LX_OFF + j, F, Bp, b)] = lx[j];
#pragma unroll
    for (int j = 0; j < LR::LXX_LEN; ++j) out[soa(stage, LR::LXX_OFF + j, F, Bp, b)] = lxx[j];
    w.Lc[(size_t)stage * Bp + b] = val;
}

// ------------------------------------------------------------------------------------------
// NumPy pairwise summation (numpy/_core/src/umath/loops_utils.h.src, DOUBLE_pairwise_sum) over a
// strided column: the reference's J_opt = L.sum() (traopt_controller.py:1935).
// ------------------------------------------------------------------------------------------
__device__ inline double pairwise_leaf(const double* a, size_t stride, int n) {   // n <= 128
    if (n < 8) {
        double res = 0.0;
        for (int i = 0; i < n; ++i) res += a[i * stride];
        return res;
    }
    double r[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = a[j * stride];
    int i;
    for (i = 8; i < n - (n % 8); i += 8) {
#pragma unroll
        for (int j = 0; j < 8; ++j) r[j] += a[(i + j) * stride];
    }
    double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
    for (; i < n; ++i) res += a[i * stride];
    return res;
}
// the recursion `sum(a[:n2]) + sum(a[n2:])`, n2 = n/2 rounded down to a multiple of 8, unrolled
// onto an explicit stack (device recursion would leave the kernel's stack size undetermined)
__device__ inline double pairwise_sum(const double* a, size_t stride, int n) {
    if (n <= 128) return pairwise_leaf(a, stride, n);
    int start[12], len[12], state[12];
    double left[12];
    int sp = 1;
    start[0] = 0; len[0] = n; state[0] = 0;
    double ret = 0.0;
    while (sp > 0) {
        const int t = sp - 1;
        if (len[t] <= 128) {
            ret = pairwise_leaf(a + (size_t)start[t] * stride, stride, len[t]);
            --sp;
            continue;
        }
        int n2 = len[t] / 2;
        n2 -= n2 % 8;
        if (state[t] == 0) {
            state[t] = 1;
            start[sp] = start[t]; len[sp] = n2; state[sp] = 0;
            ++sp;
        } else if (state[t] == 1) {
            left[t] = ret;
            state[t] = 2;
            start[sp] = start[t] + n2; len[sp] = len[t] - n2; state[sp] = 0;
            ++sp;
        } else {
            ret = left[t] + ret;
            --sp;
        }
    }
    return ret;
}

// ------------------------------------------------------------------------------------------
// Backward Riccati sweep.  grid = Bp/32 blocks of one warp; thread = problem.
// Shared memory (per thread column, stride 32 doubles => conflict-free 64-bit accesses):
//   Vs  packed upper triangle of V_xx(i+1)            NX(NX+1)/2
//   Vn  V_xx(i) under construction                    NX(NX+1)/2
//   Ys  Y = L^-1 Q_ux, first NX-3 columns             NU (NX-3)
// Algebra per stage (equivalent to :3052-3060, :2993-3004 up to rounding):
//   v = V_x + V_xx d;  Q_x = l_x + A^T v;  Q_u = l_u + B^T v;  X = V_xx A;
//   Q_xx = l_xx + A^T X;  Q_ux = B^T (X + mu A);  Q_uu = l_uu + B^T (V_xx + mu I) B = L L^T
//   Y = L^-1 Q_ux, y = L^-1 Q_u;  K = -L^-T Y, k = -L^-T y;
//   V_x(i) = Q_x - Y^T y  (= Q_x + K^T Q_uu k + K^T Q_u + Q_ux^T k);
//   V_xx(i) = Q_xx - Y^T Y (= sym(Q_xx + K^T Q_uu K + K^T Q_ux + Q_ux^T K)), symmetric by construction.
// ------------------------------------------------------------------------------------------
template <int KIND> constexpr int bwd_smem_doubles() {
    using D = Dims<KIND>;
    return D::NX * (D::NX + 1) + D::NU * (D::NX - 3);
}

template <int KIND, bool MS>
__global__ void __launch_bounds__(kBlock) k_backward(const Params prm, Work w, int it) {
    using D = Dims<KIND>;
    using LR = LinRec<KIND>;
    constexpr int NX = D::NX, NP = D::NP, NU = D::NU, NV = NX - NP, F = LR::LEN;
    constexpr int NT = NX * (NX + 1) / 2;
    constexpr int NYC = NX - 3;      // columns of Y kept in shared memory
    extern __shared__ double sm[];
    const int lane = threadIdx.x;
    const int b = blockIdx.x * kBlock + lane;
    if (b >= prm.B || w.status[b] != TRAJOPT_RUNNING) return;
    const int Bp = prm.Bp, N = prm.N;
    double* Vs = sm + lane;
    double* Vn = sm + NT * kBlock + lane;
    double* Ys = sm + 2 * NT * kBlock + lane;
    const double* lin = w.lin;

    // ---- cost / defect of the current trajectory ----------------------------------------
    double Jcur, dn = 0.0;
    if constexpr (MS) {
        // J_new of the previous iteration: Python sum, left to right, + terminal (:2742-2754)
        double s = 0.0;
        for (int i = 0; i < N; ++i) s += w.Lc[(size_t)i * Bp + b];
        Jcur = s + w.Lc[(size_t)N * Bp + b];
        double q = 0.0;
        for (int i = 0; i < N; ++i) q += w.Dsq[(size_t)i * Bp + b];
        dn = sqrt(q);
        w.dnorm[b] = dn;
        if (it > 0) w.Jhist[(size_t)(it - 1) * Bp + b] = Jcur;
        w.defhist[(size_t)it * Bp + b] = dn;
    } else {
        Jcur = pairwise_sum(w.Lc + b, (size_t)Bp, N + 1);      // J_opt = L.sum() (:1935)
    }
    w.J[b] = Jcur;
    if (!isfinite(Jcur)) {
        w.status[b] = TRAJOPT_NO_DESCENT | TRAJOPT_FLAG_NONFINITE;
        return;
    }
    if (it >= prm.max_iters) {          // MS only: closing pass after the last rollout
        w.status[b] = TRAJOPT_MAX_ITER | (w.status[b] & ~15);
        return;
    }

    // ---- terminal condition: V_x = l_x(N), V_xx = l_xx(N) --------------------------------
    double Vx[NX];
#pragma unroll
    for (int j = 0; j < NX; ++j) Vx[j] = lin[soa(N, LR::LX_OFF + j, F, Bp, b)];
#pragma unroll
    for (int r = 0; r < NX; ++r)
#pragma unroll
        for (int c = r; c < NX; ++c) {
            double v;
            if (c < NP) v = lin[soa(N, LR::LXX_OFF + tri_idx(NP, r, c), F, Bp, b)];
            else if (r >= NP) v = 2.0 * prm.P2[(r - NP) * NV + (c - NP)];
            else v = 0.0;
            Vs[tri_idx(NX, r, c) * kBlock] = v;
        }
    double pad[NX];                      // SS: adjoint variable p (:2339)
#pragma unroll
    for (int j = 0; j < NX; ++j) pad[j] = Vx[j];

    double mu = w.mu[b], delta = w.delta[b];
    double gsum = 0.0;
    int flags = 0;

    // B^T B (constant)
    double BtB[NU * NU];
#pragma unroll
    for (int a = 0; a < NU; ++a)
#pragma unroll
        for (int c = a; c < NU; ++c) {
            double s = 0.0;
#pragma unroll
            for (int r = 0; r < NV; ++r)
                if (bv_nz<KIND>(r, a) && bv_nz<KIND>(r, c)) s += prm.Bv[r * NU + a] * prm.Bv[r * NU + c];
            BtB[a * NU + c] = s;
        }

    for (int i = N - 1; i >= 0; --i) {
        AMat<KIND> A;
        A.load(lin, i, Bp, b);

        // (1) v = V_x + V_xx d
        double v[NX];
        if constexpr (MS) {
            double d[NX];
#pragma unroll
            for (int j = 0; j < NX; ++j) d[j] = lin[soa(i, LR::D_OFF + j, F, Bp, b)];
#pragma unroll
            for (int r = 0; r < NX; ++r) {
                double s = Vx[r];
#pragma unroll
                for (int c = 0; c < NX; ++c) s += Vs[sym_idx(NX, r, c) * kBlock] * d[c];
                v[r] = s;
            }
        } else {
#pragma unroll
            for (int r = 0; r < NX; ++r) v[r] = Vx[r];
        }

        // (2) Q_x = l_x + A^T v ;  Q_u = l_u + Bv^T v_vel
        double Qx[NX], Qu[NU], lu[NU];
#pragma unroll
        for (int c = 0; c < NX; ++c) {
            double s = lin[soa(i, LR::LX_OFF + c, F, Bp, b)];
#pragma unroll
            for (int r = 0; r < NX; ++r)
                if (AMat<KIND>::nz(r, c)) s += A.get(r, c) * v[r];
            Qx[c] = s;
        }
#pragma unroll
        for (int a = 0; a < NU; ++a) lu[a] = lin[soa(i, LR::LU_OFF + a, F, Bp, b)];
#pragma unroll
        for (int a = 0; a < NU; ++a) {
            double s = lu[a];
#pragma unroll
            for (int r = 0; r < NV; ++r)
                if (bv_nz<KIND>(r, a)) s += prm.Bv[r * NU + a] * v[NP + r];
            Qu[a] = s;
        }
        if constexpr (MS) {
            // g_t = L_u + F_u^T (V_x + V_xx^T d) = Q_u (:3090)
            double q = 0.0;
#pragma unroll
            for (int a = 0; a < NU; ++a) q += Qu[a] * Qu[a];
            gsum += sqrt(q);
        } else {
            // adjoint recursion of the SS gradient (:2343-2346): g = l_u + B^T p ; p = l_x + A^T p
            double g2 = 0.0;
#pragma unroll
            for (int a = 0; a < NU; ++a) {
                double s = lu[a];
#pragma unroll
                for (int r = 0; r < NV; ++r)
                    if (bv_nz<KIND>(r, a)) s += prm.Bv[r * NU + a] * pad[NP + r];
                g2 += s * s;
            }
            gsum += sqrt(g2);
            double pn[NX];
#pragma unroll
            for (int c = 0; c < NX; ++c) {
                double s = lin[soa(i, LR::LX_OFF + c, F, Bp, b)];
#pragma unroll
                for (int r = 0; r < NX; ++r)
                    if (AMat<KIND>::nz(r, c)) s += A.get(r, c) * pad[r];
                pn[c] = s;
            }
#pragma unroll
            for (int c = 0; c < NX; ++c) pad[c] = pn[c];
        }

        // (3) Q_uu0 = l_uu + Bv^T V_vv Bv  (upper triangle)
        double Quu0[NU * NU];
        {
            double T[NV * NU];           // V_vv Bv
#pragma unroll
            for (int r = 0; r < NV; ++r)
#pragma unroll
                for (int a = 0; a < NU; ++a) {
                    double s = 0.0;
#pragma unroll
                    for (int k = 0; k < NV; ++k)
                        if (bv_nz<KIND>(k, a)) s += Vs[sym_idx(NX, NP + r, NP + k) * kBlock] * prm.Bv[k * NU + a];
                    T[r * NU + a] = s;
                }
#pragma unroll
            for (int a = 0; a < NU; ++a)
#pragma unroll
                for (int c = a; c < NU; ++c) {
                    double s = 2.0 * prm.R[a * NU + c];
                    if (a == c) s += lin[soa(i, LR::LUU_OFF + a, F, Bp, b)];
#pragma unroll
                    for (int r = 0; r < NV; ++r)
                        if (bv_nz<KIND>(r, a)) s += prm.Bv[r * NU + a] * T[r * NU + c];
                    Quu0[a * NU + c] = s;
                }
        }

        // (4) regularisation loop (:2221-2246 / :2964-2991): Cholesky of Q_uu0 + mu B^T B
        double Lc[NU * NU];              // lower factor, Lc[r][c] for c < r; diagonal stores 1/L_rr
        double mu_used;                  // the mu this stage's Q_ux / Q_uu are formed with (:2311-2313)
        bool gave_up = false;
        while (true) {
            mu_used = mu;
            bool pd = true;
#pragma unroll
            for (int c = 0; c < NU; ++c) {
                double dg = Quu0[c * NU + c] + mu_used * BtB[c * NU + c];
#pragma unroll
                for (int k = 0; k < c; ++k) dg -= Lc[c * NU + k] * Lc[c * NU + k];
                if (!(dg > 0.0)) pd = false;
                const double inv = rsqrt(dg);
                Lc[c * NU + c] = inv;
#pragma unroll
                for (int r = c + 1; r < NU; ++r) {
                    double s = Quu0[c * NU + r] + mu_used * BtB[c * NU + r];
#pragma unroll
                    for (int k = 0; k < c; ++k) s -= Lc[r * NU + k] * Lc[c * NU + k];
                    Lc[r * NU + c] = s * inv;
                }
            }
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
            break;
        }

        // y = L^-1 Q_u
        double y[NU];
#pragma unroll
        for (int r = 0; r < NU; ++r) {
            double s = Qu[r];
#pragma unroll
            for (int k = 0; k < r; ++k) s -= Lc[r * NU + k] * y[k];
            y[r] = s * Lc[r * NU + r];
        }
        // k = -L^-T y
        {
            double kk[NU];
#pragma unroll
            for (int r = NU - 1; r >= 0; --r) {
                double s = y[r];
#pragma unroll
                for (int k = r + 1; k < NU; ++k) s -= Lc[k * NU + r] * kk[k];
                kk[r] = s * Lc[r * NU + r];
            }
#pragma unroll
            for (int a = 0; a < NU; ++a) w.kff[soa(i, a, NU, Bp, b)] = -kk[a];
        }

        // (5) column blocks of 3
        double Vxn[NX];
#pragma unroll
        for (int cb = 0; cb < NX / 3; ++cb) {
            double X[NX][3];
#pragma unroll
            for (int r = 0; r < NX; ++r)
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    double s = 0.0;
#pragma unroll
                    for (int k = 0; k < NX; ++k)
                        if (AMat<KIND>::nz(k, cb * 3 + j)) s += Vs[sym_idx(NX, r, k) * kBlock] * A.get(k, cb * 3 + j);
                    X[r][j] = s;
                }
            double Y[NU][3];
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                const int c = cb * 3 + j;
                // Q_ux[:, c] = Bv^T (X_vel[:, c] + mu A_vel[:, c])
                double q[NU];
#pragma unroll
                for (int a = 0; a < NU; ++a) {
                    double s = 0.0;
#pragma unroll
                    for (int r = 0; r < NV; ++r)
                        if (bv_nz<KIND>(r, a)) {
                            double t = X[NP + r][j];
                            if (AMat<KIND>::nz(NP + r, c)) t += mu_used * A.get(NP + r, c);
                            s += prm.Bv[r * NU + a] * t;
                        }
                    q[a] = s;
                }
#pragma unroll
                for (int r = 0; r < NU; ++r) {
                    double s = q[r];
#pragma unroll
                    for (int k = 0; k < r; ++k) s -= Lc[r * NU + k] * Y[k][j];
                    Y[r][j] = s * Lc[r * NU + r];
                }
                double kk[NU];
#pragma unroll
                for (int r = NU - 1; r >= 0; --r) {
                    double s = Y[r][j];
#pragma unroll
                    for (int k = r + 1; k < NU; ++k) s -= Lc[k * NU + r] * kk[k];
                    kk[r] = s * Lc[r * NU + r];
                }
#pragma unroll
                for (int a = 0; a < NU; ++a) w.Kfb[soa(i, a * NX + c, NU * NX, Bp, b)] = -kk[a];
                if (c < NYC) {
#pragma unroll
                    for (int a = 0; a < NU; ++a) Ys[(a * NYC + c) * kBlock] = Y[a][j];
                }
                // V_x(i)[c] = Q_x[c] - Y[:,c]^T y
                double s = Qx[c];
#pragma unroll
                for (int a = 0; a < NU; ++a) s -= Y[a][j] * y[a];
                Vxn[c] = s;
            }
            // V_xx(i)[r][c], r <= c
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                const int c = cb * 3 + j;
#pragma unroll
                for (int r = 0; r <= c; ++r) {
                    double s;
                    if (c < NP) s = lin[soa(i, LR::LXX_OFF + tri_idx(NP, r, c), F, Bp, b)];
                    else if (r >= NP) s = 2.0 * prm.W2[(r - NP) * NV + (c - NP)];
                    else s = 0.0;
#pragma unroll
                    for (int k = 0; k < NX; ++k)
                        if (AMat<KIND>::nz(k, r)) s += A.get(k, r) * X[k][j];
#pragma unroll
                    for (int a = 0; a < NU; ++a) {
                        const double yr = (r >= cb * 3) ? Y[a][r - cb * 3] : Ys[(a * NYC + r) * kBlock];
                        s -= yr * Y[a][j];
                    }
                    Vn[tri_idx(NX, r, c) * kBlock] = s;
                }
            }
        }
#pragma unroll
        for (int c = 0; c < NX; ++c) Vx[c] = Vxn[c];
        { double* t = Vs; Vs = Vn; Vn = t; }
    }

    w.mu[b] = mu;
    w.delta[b] = delta;
    const double g = gsum / (double)N;
    w.grad[b] = g;
    w.gradhist[(size_t)it * Bp + b] = g;
    int st = TRAJOPT_RUNNING;
    if (flags & TRAJOPT_FLAG_REG_EXCEEDED) st = TRAJOPT_NO_DESCENT;
    else if (MS ? (g < prm.tol_grad && dn < prm.tol_defect) : (g < prm.tol_grad)) st = TRAJOPT_CONVERGED;
    w.status[b] = st | flags;
}

}  // namespace trajopt
