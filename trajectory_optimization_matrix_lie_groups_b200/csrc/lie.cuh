// SO(3)/SE(3) closed forms in registers, FP64, one problem per thread.
//
// Replaces, on the device, what the reference does through manifpy objects and
// scipy Rotation at every stage (traoptlibrary/traopt_utilis.py:331-399 and the manif
// rplus / rminus / lminus / exp / log calls cited per function below).
//
// Conventions:
//   * tangent = [omega, v] (angular first);
//   * right perturbations: X (+) tau = X Exp(tau), A (-) B = Log(B^-1 A);
//   * pose = unit quaternion [x, y, z, w] + translation; every compose renormalises the
//     quaternion, which is the reference's implicit matrix -> quaternion re-projection.
//   * 3x3 matrices are row-major double[9].
//
// Small angles: below kSmall on theta^2 the Jacobian coefficients are evaluated by Taylor series
// (more accurate than the cancelling closed forms manif uses down to theta^2 = 1e-10; the two
// agree to <= 1e-12 relative, far inside the 1e-9 parity bar).
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace trajopt {

#define TO_DEV __device__ __forceinline__

constexpr double kSmall = 1e-2;     // theta^2 threshold for series evaluation

// ------------------------------------------------------------------------------------------
// 3-vectors / 3x3 blocks
// ------------------------------------------------------------------------------------------
TO_DEV void cross3(const double* a, const double* b, double* c) {
    c[0] = a[1] * b[2] - a[2] * b[1];
    c[1] = a[2] * b[0] - a[0] * b[2];
    c[2] = a[0] * b[1] - a[1] * b[0];
}
TO_DEV double dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

// C = A * B
TO_DEV void mm3(const double* A, const double* B, double* C) {
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j)
            C[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
}
// C = A^T * B
TO_DEV void mtm3(const double* A, const double* B, double* C) {
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j)
            C[3 * i + j] = A[i] * B[j] + A[3 + i] * B[3 + j] + A[6 + i] * B[6 + j];
}
// y = A x ; y = A^T x
TO_DEV void mv3(const double* A, const double* x, double* y) {
#pragma unroll
    for (int i = 0; i < 3; ++i) y[i] = A[3 * i] * x[0] + A[3 * i + 1] * x[1] + A[3 * i + 2] * x[2];
}
TO_DEV void mtv3(const double* A, const double* x, double* y) {
#pragma unroll
    for (int i = 0; i < 3; ++i) y[i] = A[i] * x[0] + A[3 + i] * x[1] + A[6 + i] * x[2];
}
// skew(w) * M  (row-major 3x3) : rows are cross products
TO_DEV void skew_mul(const double* w, const double* M, double* C) {
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        C[j] = w[1] * M[6 + j] - w[2] * M[3 + j];
        C[3 + j] = w[2] * M[j] - w[0] * M[6 + j];
        C[6 + j] = w[0] * M[3 + j] - w[1] * M[j];
    }
}
// M * skew(w)
TO_DEV void mul_skew(const double* M, const double* w, double* C) {
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        C[3 * i] = M[3 * i + 1] * w[2] - M[3 * i + 2] * w[1];
        C[3 * i + 1] = M[3 * i + 2] * w[0] - M[3 * i] * w[2];
        C[3 * i + 2] = M[3 * i] * w[1] - M[3 * i + 1] * w[0];
    }
}

// ------------------------------------------------------------------------------------------
// quaternions [x, y, z, w]
// ------------------------------------------------------------------------------------------
TO_DEV void quat_mul(const double* a, const double* b, double* c) {
    const double ax = a[0], ay = a[1], az = a[2], aw = a[3];
    const double bx = b[0], by = b[1], bz = b[2], bw = b[3];
    c[0] = aw * bx + ax * bw + ay * bz - az * by;
    c[1] = aw * by + ay * bw + az * bx - ax * bz;
    c[2] = aw * bz + az * bw + ax * by - ay * bx;
    c[3] = aw * bw - ax * bx - ay * by - az * bz;
}
TO_DEV void quat_normalize(double* q) {
    const double s = rsqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    q[0] *= s; q[1] *= s; q[2] *= s; q[3] *= s;
}
// a * b, renormalised (manif SO3::compose)
TO_DEV void quat_compose(const double* a, const double* b, double* c) {
    quat_mul(a, b, c);
    quat_normalize(c);
}
// conj(a) * b, renormalised
TO_DEV void quat_compose_inv_l(const double* a, const double* b, double* c) {
    const double ai[4] = {-a[0], -a[1], -a[2], a[3]};
    quat_mul(ai, b, c);
    quat_normalize(c);
}
// a * conj(b), renormalised
TO_DEV void quat_compose_inv_r(const double* a, const double* b, double* c) {
    const double bi[4] = {-b[0], -b[1], -b[2], b[3]};
    quat_mul(a, bi, c);
    quat_normalize(c);
}
// Eigen::Quaternion::toRotationMatrix
TO_DEV void quat_to_rot(const double* q, double* R) {
    const double x = q[0], y = q[1], z = q[2], w = q[3];
    const double tx = 2.0 * x, ty = 2.0 * y, tz = 2.0 * z;
    const double twx = tx * w, twy = ty * w, twz = tz * w;
    const double txx = tx * x, txy = ty * x, txz = tz * x;
    const double tyy = ty * y, tyz = tz * y, tzz = tz * z;
    R[0] = 1.0 - (tyy + tzz); R[1] = txy - twz;         R[2] = txz + twy;
    R[3] = txy + twz;         R[4] = 1.0 - (txx + tzz); R[5] = tyz - twx;
    R[6] = txz - twy;         R[7] = tyz + twx;         R[8] = 1.0 - (txx + tyy);
}
// v' = R(q) v  and  v' = R(q)^T v without forming R
TO_DEV void quat_rotate(const double* q, const double* v, double* r) {
    double t[3], u[3];
    cross3(q, v, t);
    t[0] *= 2.0; t[1] *= 2.0; t[2] *= 2.0;
    cross3(q, t, u);
    r[0] = v[0] + q[3] * t[0] + u[0];
    r[1] = v[1] + q[3] * t[1] + u[1];
    r[2] = v[2] + q[3] * t[2] + u[2];
}
TO_DEV void quat_rotate_inv(const double* q, const double* v, double* r) {
    double t[3], u[3];
    cross3(q, v, t);
    t[0] *= 2.0; t[1] *= 2.0; t[2] *= 2.0;
    cross3(q, t, u);
    r[0] = v[0] - q[3] * t[0] + u[0];
    r[1] = v[1] - q[3] * t[1] + u[1];
    r[2] = v[2] - q[3] * t[2] + u[2];
}

// ------------------------------------------------------------------------------------------
// SO(3)
// ------------------------------------------------------------------------------------------
// manif SO3Tangent::exp -> unit quaternion
TO_DEV void so3_exp(const double* w, double* q) {
    const double th2 = dot3(w, w);
    double s, c;   // s = sin(th/2)/th, c = cos(th/2)
    if (th2 < kSmall) {
        const double h2 = 0.25 * th2;   // (th/2)^2
        s = 0.5 * (1.0 + h2 * (-1.0 / 6 + h2 * (1.0 / 120 + h2 * (-1.0 / 5040 + h2 * (1.0 / 362880 + h2 * (-1.0 / 39916800))))));
        c = 1.0 + h2 * (-0.5 + h2 * (1.0 / 24 + h2 * (-1.0 / 720 + h2 * (1.0 / 40320 + h2 * (-1.0 / 3628800 + h2 * (1.0 / 479001600))))));
    } else {
        const double th = sqrt(th2);
        double sh;
        sincos(0.5 * th, &sh, &c);
        s = sh / th;
    }
    q[0] = s * w[0]; q[1] = s * w[1]; q[2] = s * w[2]; q[3] = c;
}

// manif SO3::log (atan2 form on the unit quaternion), angle in (-pi, pi]
TO_DEV void so3_log(const double* q, double* w) {
    const double s2 = dot3(q, q);
    double k;
    if (s2 > 1e-10) {
        const double s = sqrt(s2);
        const double c = q[3];
        const double two_angle = 2.0 * ((c < 0.0) ? atan2(-s, -c) : atan2(s, c));
        k = two_angle / s;
    } else {
        // theta/s with s = sin(theta/2):  2 (1 + s^2/6 + 3 s^4/40), sign from w
        k = ((q[3] < 0.0) ? -2.0 : 2.0) * (1.0 + s2 * (1.0 / 6 + s2 * (3.0 / 40)));
    }
    w[0] = k * q[0]; w[1] = k * q[1]; w[2] = k * q[2];
}

// a = (1-cos th)/th^2,  b = (th - sin th)/th^3
TO_DEV void so3_coef_ab(double th2, double& a, double& b) {
    if (th2 < kSmall) {
        a = 0.5 + th2 * (-1.0 / 24 + th2 * (1.0 / 720 + th2 * (-1.0 / 40320 + th2 * (1.0 / 3628800 + th2 * (-1.0 / 479001600)))));
        b = 1.0 / 6 + th2 * (-1.0 / 120 + th2 * (1.0 / 5040 + th2 * (-1.0 / 362880 + th2 * (1.0 / 39916800 + th2 * (-1.0 / 6227020800.0)))));
    } else {
        const double th = sqrt(th2);
        double s, c;
        sincos(th, &s, &c);
        a = (1.0 - c) / th2;
        b = (th - s) / (th2 * th);
    }
}
// g = 1/th^2 - (1+cos th)/(2 th sin th)   (coefficient of W^2 in Jr^-1 / Jl^-1)
TO_DEV double so3_coef_inv(double th2) {
    if (th2 < kSmall) {
        return 1.0 / 12 + th2 * (1.0 / 720 + th2 * (1.0 / 30240 + th2 * (1.0 / 1209600 + th2 * (1.0 / 47900160 + th2 * (691.0 / 1307674368000.0)))));
    }
    const double th = sqrt(th2);
    double s, c;
    sincos(th, &s, &c);
    return 1.0 / th2 - (1.0 + c) / (2.0 * th * s);
}

// M = I + a*sgn*W + b*W^2 with W = skew(w):  W^2 = w w^T - th2 I
TO_DEV void so3_jac_build(const double* w, double ca, double cb, double* M) {
    const double th2 = dot3(w, w);
    const double d = 1.0 - cb * th2;
    M[0] = d + cb * w[0] * w[0];
    M[4] = d + cb * w[1] * w[1];
    M[8] = d + cb * w[2] * w[2];
    const double xy = cb * w[0] * w[1], xz = cb * w[0] * w[2], yz = cb * w[1] * w[2];
    M[1] = xy - ca * w[2]; M[3] = xy + ca * w[2];
    M[2] = xz + ca * w[1]; M[6] = xz - ca * w[1];
    M[5] = yz - ca * w[0]; M[7] = yz + ca * w[0];
}
// Jl(w) (manif ljac), Jr(w) = Jl(-w) (rjac), Jr^-1 (rjacinv), Jl^-1 (ljacinv)
TO_DEV void so3_jl(const double* w, double* M) { double a, b; so3_coef_ab(dot3(w, w), a, b); so3_jac_build(w, a, b, M); }
TO_DEV void so3_jr(const double* w, double* M) { double a, b; so3_coef_ab(dot3(w, w), a, b); so3_jac_build(w, -a, b, M); }
TO_DEV void so3_jr_inv(const double* w, double* M) { so3_jac_build(w, 0.5, so3_coef_inv(dot3(w, w)), M); }
TO_DEV void so3_jl_inv(const double* w, double* M) { so3_jac_build(w, -0.5, so3_coef_inv(dot3(w, w)), M); }

// y = Jl(w) v and y = Jl(w)^-1 v without forming the matrix:  (I + a W + b W^2) v
TO_DEV void so3_jac_apply(const double* w, double ca, double cb, const double* v, double* y) {
    double wv[3], wwv[3];
    cross3(w, v, wv);
    cross3(w, wv, wwv);
#pragma unroll
    for (int i = 0; i < 3; ++i) y[i] = v[i] + ca * wv[i] + cb * wwv[i];
}

// ------------------------------------------------------------------------------------------
// SE(3)
// ------------------------------------------------------------------------------------------
// manif SE3Tangent::exp: (Exp(w), Jl(w) v)
TO_DEV void se3_exp(const double* tau, double* q, double* p) {
    so3_exp(tau, q);
    double a, b;
    so3_coef_ab(dot3(tau, tau), a, b);
    so3_jac_apply(tau, a, b, tau + 3, p);
}
// manif SE3::log: [Log(R), Jl(w)^-1 p]
TO_DEV void se3_log(const double* q, const double* p, double* tau) {
    so3_log(q, tau);
    so3_jac_apply(tau, -0.5, so3_coef_inv(dot3(tau, tau)), p, tau + 3);
}
// (qa, pa) * (qb, pb)
TO_DEV void se3_compose(const double* qa, const double* pa, const double* qb, const double* pb, double* q, double* p) {
    double r[3];
    quat_rotate(qa, pb, r);
    quat_compose(qa, qb, q);
    p[0] = pa[0] + r[0]; p[1] = pa[1] + r[1]; p[2] = pa[2] + r[2];
}
// (qa, pa)^-1 * (qb, pb)      -- rminus argument:  Log(a^-1 b)
TO_DEV void se3_compose_inv_l(const double* qa, const double* pa, const double* qb, const double* pb, double* q, double* p) {
    const double dp[3] = {pb[0] - pa[0], pb[1] - pa[1], pb[2] - pa[2]};
    quat_rotate_inv(qa, dp, p);
    quat_compose_inv_l(qa, qb, q);
}
// (qa, pa) * (qb, pb)^-1      -- lminus argument:  Log(a b^-1)
TO_DEV void se3_compose_inv_r(const double* qa, const double* pa, const double* qb, const double* pb, double* q, double* p) {
    quat_compose_inv_r(qa, qb, q);
    double r[3];
    quat_rotate(q, pb, r);
    p[0] = pa[0] - r[0]; p[1] = pa[1] - r[1]; p[2] = pa[2] - r[2];
}

// Barfoot's Q(w, v): lower-left block of the SE(3) left Jacobian (manif SE3Tangent::fillQ)
TO_DEV void se3_Q(const double* w, const double* v, double* Q) {
    const double th2 = dot3(w, w);
    double B, C, D;
    if (th2 < kSmall) {
        B = 1.0 / 6 + th2 * (-1.0 / 120 + th2 * (1.0 / 5040 + th2 * (-1.0 / 362880 + th2 * (1.0 / 39916800 + th2 * (-1.0 / 6227020800.0)))));
        C = -1.0 / 24 + th2 * (1.0 / 720 + th2 * (-1.0 / 40320 + th2 * (1.0 / 3628800 + th2 * (-1.0 / 479001600 + th2 * (1.0 / 87178291200.0)))));
        // D = C - 3 (th - sin th - th^3/6)/th^5
        D = -1.0 / 60 + th2 * (1.0 / 1260 + th2 * (-1.0 / 60480 + th2 * (1.0 / 4989600 + th2 * (-1.0 / 622702080.0 + th2 * (1.0 / 108972864000.0)))));
    } else {
        const double th = sqrt(th2);
        double s, c;
        sincos(th, &s, &c);
        B = (th - s) / (th2 * th);
        C = (1.0 - 0.5 * th2 - c) / (th2 * th2);
        D = C - 3.0 * (th - s - th2 * th / 6.0) / (th2 * th2 * th);
    }
    // W = skew(w), V = skew(v)
    double Wm[9] = {0, -w[2], w[1], w[2], 0, -w[0], -w[1], w[0], 0};
    double Vm[9] = {0, -v[2], v[1], v[2], 0, -v[0], -v[1], v[0], 0};
    double WV[9], VW[9], WVW[9], WW[9], T1[9], T2[9], T3[9], T4[9];
    mm3(Wm, Vm, WV);
    mm3(Vm, Wm, VW);
    mm3(WV, Wm, WVW);
    mm3(Wm, Wm, WW);
    mm3(WW, Vm, T1);     // W^2 V
    mm3(Vm, WW, T2);     // V W^2
    mm3(WVW, Wm, T3);    // W V W^2
    mm3(Wm, WVW, T4);    // W^2 V W
#pragma unroll
    for (int i = 0; i < 9; ++i)
        Q[i] = 0.5 * Vm[i] + B * (WV[i] + VW[i] + WVW[i]) - C * (T1[i] + T2[i] - 3.0 * WVW[i]) - 0.5 * D * (T3[i] + T4[i]);
}

}  // namespace trajopt
