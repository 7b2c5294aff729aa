// Shared definitions: problem kinds, parameter block, SoA layouts.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/trajopt_b200.h"

namespace trajopt {

// ------------------------------------------------------------------------------------------
// Compile-time description of the three problem families.
//   NX   tangent dimension of the state (pose + velocity)
//   NP   pose tangent dimension (3 for SO3, 6 for SE3)
//   NU   control dimension
//   NS   stored doubles per state: unit quaternion (4) [+ position (3)] + velocity
// ------------------------------------------------------------------------------------------
template <int KIND> struct Dims;
template <> struct Dims<TRAJOPT_SO3>   { static constexpr int NX = 6,  NP = 3, NU = 3, NS = 7;  };
template <> struct Dims<TRAJOPT_SE3>   { static constexpr int NX = 12, NP = 6, NU = 6, NS = 13; };
template <> struct Dims<TRAJOPT_DRONE> { static constexpr int NX = 12, NP = 6, NU = 4, NS = 13; };
template <> struct Dims<TRAJOPT_RIGID> { static constexpr int NX = 12, NP = 6, NU = 6, NS = 13; };
template <> struct Dims<TRAJOPT_PEND>  { static constexpr int NX = 6,  NP = 3, NU = 3, NS = 7;  };

// attitude-only families: pose = unit quaternion, 6-dimensional tangent state
__host__ __device__ constexpr bool on_so3(int kind) { return kind == TRAJOPT_SO3 || kind == TRAJOPT_PEND; }

// gravity acts on the body (quadrotor, rigid body): f gets m g R^T(-e3), f_x a lower-left skew block
__host__ __device__ constexpr bool has_gravity(int kind) { return kind == TRAJOPT_DRONE || kind == TRAJOPT_RIGID; }

// Doubles per reference-trajectory row (shared by the whole batch, read with uniform loads):
//   SE3/drone: q_ref quat(4) p_ref(3) xi_ref(6) R_ref(9) [p_ref]x R_ref (9)            = 31
//   SO3      : q_ref quat(4) w_ref(3) R_ref(9)                                          = 16
template <int KIND> struct RefRow { static constexpr int N = on_so3(KIND) ? 16 : 31; };

// Per-stage linearisation record written by the stage-parallel kernel and read by the
// backward sweep.  Layout in HBM is group-major: [group of 32 problems][stage][field][32], so the
// record of one stage of one warp's 32 problems is ONE contiguous chunk of LEN * 256 bytes whose
// first STAGE_LEN fields (what every warp of the backward CTA needs repeatedly) are brought into
// shared memory by a single TMA bulk copy per stage.  Field offsets:
//   SE3/drone:  a(9) b(9) c(9) e(9) h11(9) h12(9) vdt(3) [s(3) drone] | d(12) lu(NU) | luu_add(NU, AL) lx(12) lxx(21)
//   SO3      :  a(9) c(9) h(9) [pendulum: l(9) = lower-left block, bv(9)]  | d(6)  lu(3)  | luu_add(3)      lx(6)  lxx(6)
template <int KIND> struct LinRec {
    using D = Dims<KIND>;
    static constexpr int A_OFF = 0;
    static constexpr int A_LEN = (KIND == TRAJOPT_SO3) ? 27 : (KIND == TRAJOPT_PEND) ? 45 : (has_gravity(KIND) ? 60 : 57);
    static constexpr int BV_OFF = 36;                   // pendulum only: per-stage velocity rows of f_u (3 x 3) inside the A part
    static constexpr int D_OFF = A_OFF + A_LEN;
    static constexpr int LU_OFF = D_OFF + D::NX;
    static constexpr int STAGE_LEN = LU_OFF + D::NU;    // prefix staged in shared memory by the backward sweep
    static constexpr int LUU_OFF = STAGE_LEN;           // AL only: diagonal addition to l_uu
    static constexpr int LX_OFF = LUU_OFF + D::NU;
    static constexpr int LXX_OFF = LX_OFF + D::NX;
    static constexpr int LXX_LEN = D::NP * (D::NP + 1) / 2;
    static constexpr int LEN = LXX_OFF + LXX_LEN;
};

// Gains of the backward sweep (what the rollouts read): group-major like the linearisation records,
//   [group of 32 problems][stage][field][32],   fields: K (NU x NX, row-major) then k (NU),
// so the gains of one stage of one warp's 32 problems are ONE contiguous chunk of LEN * 256 bytes: the rollout stages them
// with a single TMA bulk copy per stage.  (Round 1 kept them as [stage][field][Bp] and issued one 256-byte bulk copy per
// field: `cp.async.bulk` takes uniform registers, so 84 copies with lane-varying addresses compile into a loop that
// elects one lane at a time — ncu attributed 48 % of the rollout's stall samples to that loop.)
template <int KIND> struct GainRec {
    using D = Dims<KIND>;
    static constexpr int K_OFF = 0;
    static constexpr int KFF_OFF = D::NU * D::NX;
    static constexpr int LEN = KFF_OFF + D::NU;
};

// ------------------------------------------------------------------------------------------
// Parameter block, passed to every kernel by value (lives in the constant bank).
// ------------------------------------------------------------------------------------------
struct Params {
    int kind, N, B, Bp;          // Bp = batch padded to a multiple of 32 (SoA pitch)
    int method;                  // TRAJOPT_SS / MS / AL_MS
    int rollout_linear;          // 0 = 'nonlinear', 1 = 'linear'
    int line_search;             // MS merit line search on/off
    int n_alphas;
    int max_iters;
    int has_constraints;         // AL: box bounds on u
    double dt;
    double Ib[9], Ibinv[9], mass, grav, length;
    double W1[36], W2[36];       // stage weights: pose block / velocity block (NP x NP, row-major, leading dim NP)
    double P1[36], P2[36];       // terminal weights
    double R[36];                // NU x NU
    double Bv[36];               // (NP x NU) velocity rows of f_u = Jinv * Pu * dt  (row-major, leading dim NU)
    double BtB[36];              // Bv^T Bv (NU x NU): what mu * I adds to Q_uu (traopt_controller.py:2311-2313)
    double lb[6], ub[6];
    double xlb[6], xub[6];       // AL: box bounds on the velocity part of the state
    int has_state_bounds;
    double tol_grad, tol_defect;
    double mu_min, mu_max, delta0;           // Levenberg-Marquardt schedule
    double defect_mu0, defect_rho, defect_gamma, defect_kappa;   // MS merit line search constants
    int so3_terminal_quirk;      // SO3: terminal value/gradient use Q, Hessian uses P
};

// SoA addressing: element (stage, field, problem) of an array with F fields per stage
__device__ __forceinline__ size_t soa(int stage, int field, int F, int Bp, int b) {
    return ((size_t)stage * F + field) * (size_t)Bp + b;
}

// linearisation records: group-major (see LinRec); Np1 = N + 1 stages per problem
__device__ __forceinline__ size_t lsoa(int stage, int field, int F, int Np1, int b) {
    return ((((size_t)(b >> 5) * Np1 + stage) * F + field) << 5) + (size_t)(b & 31);
}
constexpr int kRecStride = 32;   // doubles between consecutive fields of one problem's record

#define CUDA_OK(call)                                                      \
    do {                                                                   \
        cudaError_t e_ = (call);                                           \
        if (e_ != cudaSuccess) return trajopt_set_error_(e_, __FILE__, __LINE__); \
    } while (0)

}  // namespace trajopt

extern "C" int trajopt_set_error_(cudaError_t e, const char* file, int line);
