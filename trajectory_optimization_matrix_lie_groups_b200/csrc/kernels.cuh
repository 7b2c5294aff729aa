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
//   k_al_*        multiplier / penalty update (stage-parallel)                  :3270-3290
#pragma once
#include "model.cuh"

namespace trajopt {

struct Work {
    double* X[2];        // [N+1][NS][Bp] x 2 (current / candidate, selected per problem by sel)
    double* U[2];        // [N][NU][Bp]  x 2
    int* sel;            // [Bp] which buffer holds the current trajectory
    const double* ref;   // [N+1][RefRow]  reference shared by the batch
    const double* ref_batch;   // [N+1][RefRow][Bp] per-problem references (NULL: shared), see trajopt_set_reference_batch
    double* lin;         // [Bp/32][N+1][LinRec::LEN][32] (group-major, see LinRec)
    double* Lc;          // [N+1][Bp] stage costs of the current trajectory
    double* Dsq;         // [N][Bp]   squared defect norm per stage
    double* Gpre;        // [N][GPre::LEN][Bp]  MS: x(i+1) Exp(d_i) f(x_i,u_i)^-1 (pose) and f(x_i,u_i).xi of the current trajectory
    double* gains;       // [Bp/32][N][GainRec::LEN][32]  feedback gains K and feedforward terms k (group-major, see GainRec)
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
    double* lam_s;       // [N+1][2NV][Bp]  multipliers / penalties of the velocity bounds (has_state_bounds)
    double* imu_s;       // [N+1][2NV][Bp]
    double* lxxv;        // [N+1][NV][Bp]   what the velocity bounds add to the diagonal of l_xx's velocity block
    double* al_mu;       // [Bp]
    int* al_outer;       // [Bp]
    double* al_viol;     // [Bp]
    int* al_done;        // [Bp]
    int* counters;       // [0] running problems, [1] pending line searches, [2] AL problems not converged
    // trajopt_solve_stream (stream.cuh): problem id held by a slot (-1: empty), refill flags, slot lists, counters
    int* slot_id;        // [Bp]
    int* fresh;          // [Bp]
    int* free_list;      // [Bp]
    int* done_list;      // [Bp]
    int* scnt;           // [4] free slots, finished problems to export, running slots
    int* Nb;             // [Bp] horizon of each problem, 1 <= Nb <= N (trajopt_set_horizons; default N for all)
    int* orig;           // [Bp] slot -> problem index of the caller (identity until a compaction moves problems)
    // small batches (ensure_cand): one trajectory buffer per line-search step size, so that ONE launch rolls out and keeps
    // every candidate ([n_alphas][N+1][NS][Bp], [n_alphas][N][NU][Bp]); NULL otherwise
    double* Xc;
    double* Uc;
};

constexpr int kBlock = 32;

// What the full-step (alpha = 1) multiple-shooting rollout needs from the CURRENT trajectory at stage i
// besides x(i), u(i): the pose G_i = q(i+1) Exp(d_q) f(x_i,u_i).q^-1 and f(x_i,u_i).xi (traopt_controller.py:2697-2718).
// Both are independent of the rollout's recursion, so the stage-parallel linearisation forms them.
template <int KIND> struct GPre {
    static constexpr int NPOSE = (on_so3(KIND)) ? 4 : 7;
    static constexpr int LEN = NPOSE + (Dims<KIND>::NX - Dims<KIND>::NP);
};

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

// attitude families: [ a  c ; l  h ], l = 0 except for the pendulum
template <int KIND> struct AMatSO3 {
    double a[9], c[9], h[9], l[9];
    static __host__ __device__ constexpr bool nz(int r, int cc) { return KIND == TRAJOPT_PEND || !(r >= 3 && cc < 3); }
    TO_DEV double get(int r, int cc) const {
        const int i = r % 3, j = cc % 3;
        if (r < 3 && cc < 3) return a[3 * i + j];
        if (r < 3) return c[3 * i + j];
        if (cc < 3) return l[3 * i + j];
        return h[3 * i + j];
    }
    TO_DEV void load(const double* __restrict__ lin, int stage, int Np1, int b) {
        constexpr int F = LinRec<KIND>::LEN;
#pragma unroll
        for (int i = 0; i < 9; ++i) {
            a[i] = lin[lsoa(stage, i, F, Np1, b)];
            c[i] = lin[lsoa(stage, 9 + i, F, Np1, b)];
            h[i] = lin[lsoa(stage, 18 + i, F, Np1, b)];
            l[i] = (KIND == TRAJOPT_PEND) ? lin[lsoa(stage, 27 + i, F, Np1, b)] : 0.0;
        }
    }
};
template <> struct AMat<TRAJOPT_SO3> : AMatSO3<TRAJOPT_SO3> {};
template <> struct AMat<TRAJOPT_PEND> : AMatSO3<TRAJOPT_PEND> {};

template <int KIND> struct AMat {   // SE3, drone, rigid body
    double a[9], b[9], c[9], e[9], h11[9], h12[9], vdt[3], s[3];
    static __host__ __device__ constexpr bool nz(int r, int cc) {
        const int br = r / 3, bc = cc / 3, i = r % 3, j = cc % 3;
        if (br == 0) return bc == 0 || bc == 2;
        if (br == 1) return true;
        if (br == 2) return bc >= 2;
        // br == 3
        if (bc == 0) return has_gravity(KIND) && i != j;
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
    TO_DEV void load(const double* __restrict__ lin, int stage, int Np1, int bb) {
        constexpr int F = LinRec<KIND>::LEN;
#pragma unroll
        for (int i = 0; i < 9; ++i) {
            a[i] = lin[lsoa(stage, i, F, Np1, bb)];
            b[i] = lin[lsoa(stage, 9 + i, F, Np1, bb)];
            c[i] = lin[lsoa(stage, 18 + i, F, Np1, bb)];
            e[i] = lin[lsoa(stage, 27 + i, F, Np1, bb)];
            h11[i] = lin[lsoa(stage, 36 + i, F, Np1, bb)];
            h12[i] = lin[lsoa(stage, 45 + i, F, Np1, bb)];
        }
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            vdt[i] = lin[lsoa(stage, 54 + i, F, Np1, bb)];
            s[i] = has_gravity(KIND) ? lin[lsoa(stage, 57 + i, F, Np1, bb)] : 0.0;
        }
    }
};

// Velocity rows of f_u: Bv = Jinv Pu dt  (NV x NU), compile-time sparsity
template <int KIND> __host__ __device__ constexpr bool bv_nz(int r, int c) {
    if (on_so3(KIND)) return true;
    if (KIND == TRAJOPT_SE3 || KIND == TRAJOPT_RIGID) return (r < 3 && c < 3) || (r >= 3 && r == c);
    return (r < 3 && c < 3) || (r == 5 && c == 3);   // drone: torques + body-z thrust
}

// Velocity rows of f_u for one stage.  Constant for the rigid-body families (Params::Bv); the pendulum's
// depend on the attitude (J^-1 skew(m rho) R^T dt, traopt_dynamics.py:590-603) and come from the record.
template <int KIND> struct BvStage {
    const Params& prm;
    TO_DEV BvStage(const Params& p, const double*, int) : prm(p) {}
    TO_DEV double get(int r, int a) const { return prm.Bv[r * Dims<KIND>::NU + a]; }
    TO_DEV double btb(const Params&, int a, int c) const { return prm.BtB[a * Dims<KIND>::NU + c]; }   // (Bv^T Bv)[a][c]
};
template <> struct BvStage<TRAJOPT_PEND> {
    double v[9];
    TO_DEV BvStage(const Params&, const double* __restrict__ rec, int stride) {
#pragma unroll
        for (int t = 0; t < 9; ++t) v[t] = rec[(size_t)(LinRec<TRAJOPT_PEND>::BV_OFF + t) * stride];
    }
    TO_DEV double get(int r, int a) const { return v[r * 3 + a]; }
    TO_DEV double btb(const Params&, int a, int c) const { return fma(v[a], v[c], fma(v[3 + a], v[3 + c], v[6 + a] * v[6 + c])); }
};

__host__ __device__ constexpr int tri_idx(int n, int r, int c) {   // packed upper triangle, r <= c
    return r * n - (r * (r - 1)) / 2 + (c - r);
}
__host__ __device__ constexpr int sym_idx(int n, int r, int c) { return r <= c ? tri_idx(n, r, c) : tri_idx(n, c, r); }

// ------------------------------------------------------------------------------------------
// Reference row of (stage, problem): the batch-shared row (uniform loads) or the problem's own (coalesced SoA)
// ------------------------------------------------------------------------------------------
template <int KIND>
TO_DEV void fetch_ref_row(const Work& w, int Bp, int stage, int b, double (&rr)[RefRow<KIND>::N]) {
    constexpr int RR = RefRow<KIND>::N;
    if (w.ref_batch) {
        const double* p = w.ref_batch + (size_t)stage * RR * Bp + b;
#pragma unroll
        for (int j = 0; j < RR; ++j) rr[j] = p[(size_t)j * Bp];
    } else {
        const double* p = w.ref + (size_t)stage * RR;
#pragma unroll
        for (int j = 0; j < RR; ++j) rr[j] = p[j];
    }
}

// Per-problem references arrive problem-major ([B][N+1][7|4] poses as quaternion (+ position), [B][N+1][6|3] twists);
// this packs them into the rows the cost reads (same arithmetic, operation by operation, as the host packing of the
// shared reference in trajopt_set_reference: no FMA contraction, IEEE sqrt and division).
template <int KIND>
__global__ void k_pack_ref_batch(int B, int Bp, int Np1, const double* __restrict__ q_in, const double* __restrict__ xi_in,
                                 double* __restrict__ out) {
    constexpr int RR = RefRow<KIND>::N, NPOSE = on_so3(KIND) ? 4 : 7, NV = Dims<KIND>::NX - Dims<KIND>::NP;
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    const int stage = blockIdx.y;
    if (b >= Bp) return;
    double r[RR];
#pragma unroll
    for (int j = 0; j < RR; ++j) r[j] = 0.0;
    r[3] = 1.0;
    if (b < B) {
        const double* q = q_in + ((size_t)b * Np1 + stage) * NPOSE;
        const double* xi = xi_in + ((size_t)b * Np1 + stage) * NV;
        const double n2 = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(q[0], q[0]), __dmul_rn(q[1], q[1])), __dmul_rn(q[2], q[2])),
                                    __dmul_rn(q[3], q[3]));
        const double nq = __dsqrt_rn(n2);
#pragma unroll
        for (int j = 0; j < 4; ++j) r[j] = __ddiv_rn(q[j], nq);
        double* R;
        if constexpr (on_so3(KIND)) {
#pragma unroll
            for (int j = 0; j < 3; ++j) r[4 + j] = xi[j];
            R = r + 7;
        } else {
#pragma unroll
            for (int j = 0; j < 3; ++j) r[4 + j] = q[4 + j];
#pragma unroll
            for (int j = 0; j < 6; ++j) r[7 + j] = xi[j];
            R = r + 13;
        }
        {   // Eigen::Quaternion::toRotationMatrix, operation by operation
            const double x = r[0], y = r[1], z = r[2], w_ = r[3];
            const double tx = __dmul_rn(2.0, x), ty = __dmul_rn(2.0, y), tz = __dmul_rn(2.0, z);
            const double twx = __dmul_rn(tx, w_), twy = __dmul_rn(ty, w_), twz = __dmul_rn(tz, w_);
            const double txx = __dmul_rn(tx, x), txy = __dmul_rn(ty, x), txz = __dmul_rn(tz, x);
            const double tyy = __dmul_rn(ty, y), tyz = __dmul_rn(tz, y), tzz = __dmul_rn(tz, z);
            R[0] = __dsub_rn(1.0, __dadd_rn(tyy, tzz)); R[1] = __dsub_rn(txy, twz); R[2] = __dadd_rn(txz, twy);
            R[3] = __dadd_rn(txy, twz); R[4] = __dsub_rn(1.0, __dadd_rn(txx, tzz)); R[5] = __dsub_rn(tyz, twx);
            R[6] = __dsub_rn(txz, twy); R[7] = __dadd_rn(tyz, twx); R[8] = __dsub_rn(1.0, __dadd_rn(txx, tyy));
        }
        if constexpr (!on_so3(KIND)) {
            const double* pp = r + 4;   // [p]x R
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                r[22 + j] = __dsub_rn(__dmul_rn(pp[1], R[6 + j]), __dmul_rn(pp[2], R[3 + j]));
                r[25 + j] = __dsub_rn(__dmul_rn(pp[2], R[j]), __dmul_rn(pp[0], R[6 + j]));
                r[28 + j] = __dsub_rn(__dmul_rn(pp[0], R[3 + j]), __dmul_rn(pp[1], R[j]));
            }
        }
    }
#pragma unroll
    for (int j = 0; j < RR; ++j) out[((size_t)stage * RR + j) * Bp + b] = r[j];
}

// ------------------------------------------------------------------------------------------
// Initialisation
// ------------------------------------------------------------------------------------------
template <int KIND>
__device__ __forceinline__ double us_init_value(const Work& w, const Params& prm, int stage, int j, int b) {
    constexpr int NU = Dims<KIND>::NU;
    if (w.us_mode == 0 || w.us_init == nullptr) return 0.0;
    if (w.us_mode == 1) return w.us_init[(size_t)stage * NU + j];
    return w.us_init[((size_t)w.orig[b] * prm.N + stage) * NU + j];
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
static __global__ void k_reset(const Params prm, Work w) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= prm.Bp) return;
    w.sel[b] = 0;
    w.mu[b] = 1.0;
    w.delta[b] = prm.delta0;
    w.iters[b] = 0;
    // n_iterations = 0: single shooting returns the initial rollout untouched; multiple shooting
    // still runs its closing pass (cost / defect of the initial guess)
    const bool run = (b < prm.B) && (prm.max_iters > 0 || prm.method != TRAJOPT_SS);
    w.status[b] = run ? TRAJOPT_RUNNING : TRAJOPT_MAX_ITER;
    w.ls_state[b] = -2;
    w.J[b] = 0.0;
    w.grad[b] = 0.0;
    w.dnorm[b] = 0.0;
}

// only the problems whose AL outer loop is still running are restarted
static __global__ void k_reset_al_inner(const Params prm, Work w) {
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
__global__ void k_init_ms(const Params prm, Work w, bool only_running, const int* __restrict__ mask) {
    constexpr int NS = Dims<KIND>::NS, NU = Dims<KIND>::NU;
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    const int stage = blockIdx.y;
    if (b >= prm.Bp) return;
    if (only_running && w.status[b] != TRAJOPT_RUNNING) return;
    if (mask && !mask[b]) return;            // trajopt_solve_stream: only the slots that were just refilled
    State<KIND> s;
    if (stage == 0) {
#pragma unroll
        for (int j = 0; j < 4; ++j) s.q[j] = w.x0[(size_t)j * prm.Bp + b];
        if constexpr (!on_so3(KIND)) {
#pragma unroll
            for (int j = 0; j < 3; ++j) s.p[j] = w.x0[(size_t)(4 + j) * prm.Bp + b];
#pragma unroll
            for (int j = 0; j < 6; ++j) s.xi[j] = w.x0[(size_t)(7 + j) * prm.Bp + b];
        } else {
#pragma unroll
            for (int j = 0; j < 3; ++j) s.xi[j] = w.x0[(size_t)(4 + j) * prm.Bp + b];
        }
    } else {
        double rr[RefRow<KIND>::N];
        fetch_ref_row<KIND>(w, prm.Bp, stage, b, rr);
        load_ref_state<KIND>(rr, 0, s);
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
__global__ void __launch_bounds__(kBlock) k_init_ss(const Params prm, Work w, const int* __restrict__ mask) {
    constexpr int NU = Dims<KIND>::NU;
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= prm.Bp) return;
    if (mask && !mask[b]) return;
    State<KIND> x, xn;
#pragma unroll
    for (int j = 0; j < 4; ++j) x.q[j] = w.x0[(size_t)j * prm.Bp + b];
    if constexpr (!on_so3(KIND)) {
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
// REFB: per-problem references (a separate instantiation keeps the shared-reference kernel's register budget intact)
// 3 blocks per SM (168 registers, 52 bytes of spills): 3.60 -> 3.29 ms at 16384 x 955 once the state part of the record is
// stored before the dynamics part is formed; 4 blocks (128 registers, 450 bytes of spills) is slower again (3.67 ms).
template <int KIND, bool MS, bool REFB>
__global__ void __launch_bounds__(128, 3) k_linearize(const Params prm, Work w, int stage0, int flip, const int* __restrict__ mask) {
    using D = Dims<KIND>;
    using LR = LinRec<KIND>;
    constexpr int NX = D::NX, NU = D::NU, F = LR::LEN;
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    const int stage = stage0 + blockIdx.y;   // stage0 > 0: one chunk of the horizon (see run_forward_overlapped)
    // A warp is one record group (32 consecutive problems).  It leaves only when NONE of its problems needs this
    // stage: as long as one does, the finished ones ride along (their records are never read), because a row written
    // by part of a warp is a partial-sector write — with 10 % of the problems running, spread over every warp, the
    // kernel took 7.4 ms instead of 3.9 (scripts/iter_times.py).
    const bool valid = b < prm.B;
    const bool wanted = valid && w.status[b] == TRAJOPT_RUNNING && !(mask && !mask[b]) && stage <= w.Nb[b];
    if (__ballot_sync(0xffffffffu, wanted) == 0u || !valid) return;
    const int Nb = max(w.Nb[b], stage);      // a problem riding along beyond its own horizon: treated as its terminal stage
    const int Bp = prm.Bp;
    const int cur = w.sel[b] ^ flip;         // flip = 1: the trajectory a rollout is writing, before it is accepted
    const double* X = w.X[cur];
    const double* U = w.U[cur];
    double rr[REFB ? RefRow<KIND>::N : 1];
    const double* refrow;
    if constexpr (REFB) {
        constexpr int RR = RefRow<KIND>::N;
        const double* p = w.ref_batch + (size_t)stage * RR * Bp + b;
#pragma unroll
        for (int j = 0; j < RR; ++j) rr[j] = p[(size_t)j * Bp];
        refrow = rr;
    } else {
        (void)rr;
        refrow = w.ref + (size_t)stage * RefRow<KIND>::N;
    }
    double* out = w.lin;

    State<KIND> x;
    load_state<KIND>(X, stage, Bp, b, x);
    const bool terminal = (stage == Nb);

    double lx[NX], lxx[LR::LXX_LEN];
    double val = cost_expand<KIND>(prm, x, refrow, terminal, lx, lxx);
    if (prm.has_state_bounds) {   // AL terms of the velocity bounds (every stage, the terminal one included)
        constexpr int NV = NX - D::NP;
        double lam[2 * NV], imu[2 * NV], lxa[NV], lxxa[NV];
#pragma unroll
        for (int j = 0; j < 2 * NV; ++j) {
            lam[j] = w.lam_s[soa(stage, j, 2 * NV, Bp, b)];
            imu[j] = w.imu_s[soa(stage, j, 2 * NV, Bp, b)];
        }
        val += al_box_terms<NV>(prm.xlb, prm.xub, x.xi, lam, imu, lxa, lxxa);
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            lx[D::NP + j] += lxa[j];
            w.lxxv[soa(stage, j, NV, Bp, b)] = lxxa[j];
        }
    }

    // the state part of the record is complete: store it now, its 33 doubles need not stay live through the dynamics
#pragma unroll
    for (int j = 0; j < NX; ++j) out[lsoa(stage, LR::LX_OFF + j, F, prm.N + 1, b)] = lx[j];
#pragma unroll
    for (int j = 0; j < LR::LXX_LEN; ++j) out[lsoa(stage, LR::LXX_OFF + j, F, prm.N + 1, b)] = lxx[j];

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
        for (int j = 0; j < NU; ++j) out[lsoa(stage, LR::LU_OFF + j, F, prm.N + 1, b)] = lu[j];
        if (prm.has_constraints) {   // only the AL cost adds to l_uu, and only then does the sweep read these rows
#pragma unroll
            for (int j = 0; j < NU; ++j) out[lsoa(stage, LR::LUU_OFF + j, F, prm.N + 1, b)] = luu_add[j];
        }
        // dynamics Jacobian
        double rec[LR::A_LEN];
        dyn_jacobian<KIND>(prm, x, u, rec);
#pragma unroll
        for (int j = 0; j < LR::A_LEN; ++j) out[lsoa(stage, LR::A_OFF + j, F, prm.N + 1, b)] = rec[j];
        // defect against the next shooting node
        if constexpr (MS) {
            State<KIND> fx, xnext;
            dyn_step<KIND>(prm, x, u, fx);
            load_state<KIND>(X, stage + 1, Bp, b, xnext);
            double d[NX], dsq = 0.0;
            defect<KIND>(fx, xnext, d);
#pragma unroll
            for (int j = 0; j < NX; ++j) {
                out[lsoa(stage, LR::D_OFF + j, F, prm.N + 1, b)] = d[j];
                dsq += d[j] * d[j];
            }
            w.Dsq[(size_t)stage * Bp + b] = dsq;
            // G_i for the alpha = 1 rollout, composed left to right exactly as k_forward does for any alpha
            {
                constexpr int GL = GPre<KIND>::LEN, GP = GPre<KIND>::NPOSE;
                double* gp = w.Gpre + soa(stage, 0, GL, Bp, b);
                if constexpr (on_so3(KIND)) {
                    double qe[4], q1[4], q2[4];
                    so3_exp(d, qe);
                    quat_compose(xnext.q, qe, q1);
                    quat_compose_inv_r(q1, fx.q, q2);
#pragma unroll
                    for (int j = 0; j < 4; ++j) gp[(size_t)j * Bp] = q2[j];
                } else {
                    double qe[4], pe[3], q1[4], p1[3], q2[4], p2[3];
                    se3_exp(d, qe, pe);
                    se3_compose(xnext.q, xnext.p, qe, pe, q1, p1);
                    se3_compose_inv_r(q1, p1, fx.q, fx.p, q2, p2);
#pragma unroll
                    for (int j = 0; j < 4; ++j) gp[(size_t)j * Bp] = q2[j];
#pragma unroll
                    for (int j = 0; j < 3; ++j) gp[(size_t)(4 + j) * Bp] = p2[j];
                }
#pragma unroll
                for (int j = 0; j < NX - D::NP; ++j) gp[(size_t)(GP + j) * Bp] = fx.xi[j];
            }
        }
    }
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

}  // namespace trajopt
