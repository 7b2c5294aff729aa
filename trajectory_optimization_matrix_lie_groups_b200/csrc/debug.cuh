// Lie-group primitives exposed one by one for the parity tests (trajopt_debug_lie).
// Rows are problem-major: input row r at d_in + r * in_width(op), output at d_out + r * out_width(op).
#pragma once
#include "backward.cuh"

namespace trajopt {

enum LieOp {
    LIE_SO3_EXP = 0,      // w(3)            -> quat(4)
    LIE_SO3_LOG = 1,      // quat(4)         -> w(3)
    LIE_SO3_JR = 2,       // w(3)            -> 3x3
    LIE_SO3_JR_INV = 3,   // w(3)            -> 3x3
    LIE_SO3_JL = 4,       // w(3)            -> 3x3
    LIE_SO3_JL_INV = 5,   // w(3)            -> 3x3
    LIE_SE3_EXP = 6,      // tau(6)          -> quat(4) p(3)
    LIE_SE3_LOG = 7,      // quat(4) p(3)    -> tau(6)
    LIE_SE3_Q = 8,        // tau(6)          -> 3x3
    LIE_SE3_JR = 9,       // tau(6)          -> 6x6
    LIE_SE3_JR_INV = 10,  // tau(6)          -> 6x6
    LIE_SE3_ADJ = 11,     // quat(4) p(3)    -> 6x6
    LIE_SE3_COMPOSE = 12, // a(7) b(7)       -> a b (7)
    LIE_SE3_RMINUS = 13,  // a(7) b(7)       -> Log(b^-1 a) (6)
    LIE_SE3_LMINUS = 14,  // a(7) b(7)       -> Log(a b^-1) (6)
    LIE_OP_COUNT = 15
};

__host__ __device__ inline int lie_in_width(int op) {
    const int w[LIE_OP_COUNT] = {3, 4, 3, 3, 3, 3, 6, 7, 6, 6, 6, 7, 14, 14, 14};
    return w[op];
}
__host__ __device__ inline int lie_out_width(int op) {
    const int w[LIE_OP_COUNT] = {4, 3, 9, 9, 9, 9, 7, 6, 9, 36, 36, 36, 7, 6, 6};
    return w[op];
}

__device__ inline void put_block(double* M6, int br, int bc, const double* B3, double sign) {
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) M6[(3 * br + i) * 6 + 3 * bc + j] = sign * B3[3 * i + j];
}

static __global__ void k_debug_lie(int op, int n, const double* __restrict__ in, double* __restrict__ out) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const double* x = in + (size_t)r * lie_in_width(op);
    double* y = out + (size_t)r * lie_out_width(op);
    double a[14], o[36];
    for (int j = 0; j < lie_in_width(op); ++j) a[j] = x[j];
    for (int j = 0; j < 36; ++j) o[j] = 0.0;
    switch (op) {
        case LIE_SO3_EXP: so3_exp(a, o); break;
        case LIE_SO3_LOG: so3_log(a, o); break;
        case LIE_SO3_JR: so3_jr(a, o); break;
        case LIE_SO3_JR_INV: so3_jr_inv(a, o); break;
        case LIE_SO3_JL: so3_jl(a, o); break;
        case LIE_SO3_JL_INV: so3_jl_inv(a, o); break;
        case LIE_SE3_EXP: se3_exp(a, o, o + 4); break;
        case LIE_SE3_LOG: se3_log(a, a + 4, o); break;
        case LIE_SE3_Q: se3_Q(a, a + 3, o); break;
        case LIE_SE3_JR: {
            // Jr(tau) = Jl(-tau) = [[Jr(w), 0], [Q(-w, -v), Jr(w)]]
            double J[9], Q[9], na[6];
            for (int j = 0; j < 6; ++j) na[j] = -a[j];
            so3_jr(a, J);
            se3_Q(na, na + 3, Q);
            put_block(o, 0, 0, J, 1.0);
            put_block(o, 1, 1, J, 1.0);
            put_block(o, 1, 0, Q, 1.0);
            break;
        }
        case LIE_SE3_JR_INV: {
            double Ji[9], Q[9], T[9], Z[9], na[6];
            for (int j = 0; j < 6; ++j) na[j] = -a[j];
            so3_jr_inv(a, Ji);
            se3_Q(na, na + 3, Q);
            mm3(Ji, Q, T);
            mm3(T, Ji, Z);
            put_block(o, 0, 0, Ji, 1.0);
            put_block(o, 1, 1, Ji, 1.0);
            put_block(o, 1, 0, Z, -1.0);
            break;
        }
        case LIE_SE3_ADJ: {
            double R[9], PR[9];
            quat_to_rot(a, R);
            skew_mul(a + 4, R, PR);
            put_block(o, 0, 0, R, 1.0);
            put_block(o, 1, 1, R, 1.0);
            put_block(o, 1, 0, PR, 1.0);
            break;
        }
        case LIE_SE3_COMPOSE: se3_compose(a, a + 4, a + 7, a + 11, o, o + 4); break;
        case LIE_SE3_RMINUS: {
            double q[4], p[3];
            se3_compose_inv_l(a + 7, a + 11, a, a + 4, q, p);
            se3_log(q, p, o);
            break;
        }
        case LIE_SE3_LMINUS: {
            double q[4], p[3];
            se3_compose_inv_r(a, a + 4, a + 7, a + 11, q, p);
            se3_log(q, p, o);
            break;
        }
        default: break;
    }
    for (int j = 0; j < lie_out_width(op); ++j) y[j] = o[j];
}

// Per-stage callbacks of the reference's Dynamics / Cost classes on n independent (x, u) rows against
// reference row i: f (traopt_dynamics.py:763-787, 369-380, 1373-1401), f_x / f_u (:802-850, 385-403,
// 1416-1482), l / l_x / l_xx / l_u (traopt_cost.py:675-867, 381-564) and the tracking error _err.
template <int KIND>
__global__ void k_debug_stage(const Params prm, const double* __restrict__ ref, int i, int terminal, int n,
                              const double* __restrict__ xin, const double* __restrict__ uin, double* f_out,
                              double* Fx, double* Fu, double* l_out, double* lx_out, double* lxx_out, double* lu_out,
                              double* err_out) {
    using D = Dims<KIND>;
    using LR = LinRec<KIND>;
    constexpr int NX = D::NX, NP = D::NP, NU = D::NU, NS = D::NS, NV = NX - NP, NB = NX / 3;
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const double* refrow = ref + (size_t)i * RefRow<KIND>::N;
    State<KIND> x;
    {
        double row[NS];
        for (int j = 0; j < NS; ++j) row[j] = xin[(size_t)r * NS + j];
        quat_normalize(row);
        for (int j = 0; j < 4; ++j) x.q[j] = row[j];
        if constexpr (!on_so3(KIND)) {
            for (int j = 0; j < 3; ++j) x.p[j] = row[4 + j];
            for (int j = 0; j < 6; ++j) x.xi[j] = row[7 + j];
        } else {
            for (int j = 0; j < 3; ++j) x.xi[j] = row[4 + j];
        }
    }
    double u[NU];
    for (int j = 0; j < NU; ++j) u[j] = (uin && !terminal) ? uin[(size_t)r * NU + j] : 0.0;
    if (f_out) {
        State<KIND> xn;
        dyn_step<KIND>(prm, x, u, xn);
        double* o = f_out + (size_t)r * NS;
        for (int j = 0; j < 4; ++j) o[j] = xn.q[j];
        if constexpr (!on_so3(KIND)) {
            for (int j = 0; j < 3; ++j) o[4 + j] = xn.p[j];
            for (int j = 0; j < 6; ++j) o[7 + j] = xn.xi[j];
        } else {
            for (int j = 0; j < 3; ++j) o[4 + j] = xn.xi[j];
        }
    }
    if (Fx) {
        double rec[LR::A_LEN];
        dyn_jacobian<KIND>(prm, x, u, rec);
        double* o = Fx + (size_t)r * NX * NX;
        for (int j = 0; j < NX * NX; ++j) o[j] = 0.0;
        sfor<0, NB>([&](auto rbc) {
            sfor<0, NB>([&](auto cbc) {
                constexpr int RB = decltype(rbc)::value, CB = decltype(cbc)::value;
                if constexpr (blk_kind<KIND>(RB, CB) != BK_ZERO) {
                    double m[9];
                    load_blk<KIND, RB, CB>(rec, 1, m);
                    for (int a = 0; a < 3; ++a)
                        for (int c = 0; c < 3; ++c) o[(RB * 3 + a) * NX + CB * 3 + c] = m[3 * a + c];
                }
            });
        });
    }
    if (Fu) {
        double rec[LR::A_LEN];
        if constexpr (KIND == TRAJOPT_PEND) dyn_jacobian<KIND>(prm, x, u, rec);    // f_u depends on the attitude
        const BvStage<KIND> Bv(prm, rec, 1);
        double* o = Fu + (size_t)r * NX * NU;
        for (int a = 0; a < NX; ++a)
            for (int c = 0; c < NU; ++c) o[a * NU + c] = (a >= NP && bv_nz<KIND>(a - NP, c)) ? Bv.get(a - NP, c) : 0.0;
    }
    if (l_out || lx_out || lxx_out) {
        double lx[NX], lxx[LR::LXX_LEN];
        double val = cost_expand<KIND>(prm, x, refrow, terminal != 0, lx, lxx);
        if (!terminal) {
            for (int a = 0; a < NU; ++a) {
                double sacc = 0.0;
                for (int c = 0; c < NU; ++c) sacc += prm.R[a * NU + c] * u[c];
                val += u[a] * sacc;
            }
        }
        if (l_out) l_out[r] = val;
        if (lx_out) for (int j = 0; j < NX; ++j) lx_out[(size_t)r * NX + j] = lx[j];
        if (lxx_out) {
            const double* W2 = terminal ? prm.P2 : prm.W2;
            for (int a = 0; a < NX; ++a)
                for (int c = 0; c < NX; ++c) {
                    double v = 0.0;
                    if (a < NP && c < NP) v = lxx[sym_idx(NP, a, c)];
                    else if (a >= NP && c >= NP) v = 2.0 * W2[(a - NP) * NV + (c - NP)];
                    lxx_out[((size_t)r * NX + a) * NX + c] = v;
                }
        }
    }
    if (lu_out) {
        for (int a = 0; a < NU; ++a) {
            double sacc = 0.0;
            for (int c = 0; c < NU; ++c) sacc += prm.R[a * NU + c] * u[c];
            lu_out[(size_t)r * NU + a] = terminal ? 0.0 : 2.0 * sacc;
        }
    }
    if (err_out) {
        double e[NP], dxi[NV];
        tracking_error<KIND>(x, refrow, e, dxi);
        for (int j = 0; j < NP; ++j) err_out[(size_t)r * NX + j] = e[j];
        for (int j = 0; j < NV; ++j) err_out[(size_t)r * NX + NP + j] = dxi[j];
    }
}

// FP64 FMA peak: 8 independent dependent-chains of DFMA per thread, enough warps to fill every SM.
// 2 flop per DFMA; the result is stored so the chains cannot be removed.
static __global__ void __launch_bounds__(256) k_fp64_peak(int iters, double seed, double* __restrict__ sink) {
    double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double m = 0.999999, c = 1e-9;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
            a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
        }
    }
    sink[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
}

}  // namespace trajopt
