// Lie-group primitives exposed one by one for the parity tests (trajopt_debug_lie).
// Rows are problem-major: input row r at d_in + r * in_width(op), output at d_out + r * out_width(op).
#pragma once
#include "lie.cuh"

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

__global__ void k_debug_lie(int op, int n, const double* __restrict__ in, double* __restrict__ out) {
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

// FP64 FMA peak: 8 independent dependent-chains of DFMA per thread, enough warps to fill every SM.
// 2 flop per DFMA; the result is stored so the chains cannot be removed.
__global__ void __launch_bounds__(256) k_fp64_peak(int iters, double seed, double* __restrict__ sink) {
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
