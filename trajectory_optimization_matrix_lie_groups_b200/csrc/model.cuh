// Per-stage model evaluation in registers: rigid-body dynamics on SO(3)/SE(3) (+ quadrotor),
// their analytic state Jacobians, the log-map tracking cost with Gauss-Newton Hessian, the
// multiple-shooting defect and the augmented-Lagrangian terms.
//
// Reference functions replaced (file:line in /root/reference/traoptlibrary):
//   SE3Dynamics.fd_euler  traopt_dynamics.py:763-787     f_x :802-837    f_u :839-850
//   DroneDynamics.fd_euler               :1373-1401      f_x :1416-1469  f_u :1471-1482
//   SO3Dynamics.fd_euler                 :369-380        f_x :385-400    f_u :402-403
//   SE3TrackingQuadraticGaussNewtonCost  traopt_cost.py:675-867
//   SO3TrackingQuadraticGaussNewtonCost  traopt_cost.py:381-564
//   ALConstrainedCost                    traopt_cost.py:1236-1320  + InputConstraint traopt_constraints.py:116-169
//   defect d_i                           traopt_controller.py:2882-2888 (SE3), :1464-1466 (SO3)
//
// The three reference quirks that the shipped results depend on are reproduced (SURVEY.md
// finding 4): (1) f_x uses ad of the swapped twist [v, omega]; (2) the drone gravity Jacobian
// has no m*g factor; (3) the SO3 terminal value and gradient use Q while the Hessian uses P.
#pragma once
#include "common.cuh"
#include "lie.cuh"

namespace trajopt {

template <int KIND> struct State {
    double q[4];
    double p[3];     // unused for SO3
    double xi[Dims<KIND>::NX - Dims<KIND>::NP];   // [omega, v] (SE3) or omega (SO3)
};

template <int KIND>
TO_DEV void load_state(const double* __restrict__ X, int stage, int Bp, int b, State<KIND>& s) {
    constexpr int NS = Dims<KIND>::NS;
    const double* base = X + soa(stage, 0, NS, Bp, b);
#pragma unroll
    for (int j = 0; j < 4; ++j) s.q[j] = base[(size_t)j * Bp];
    if constexpr (!on_so3(KIND)) {
#pragma unroll
        for (int j = 0; j < 3; ++j) s.p[j] = base[(size_t)(4 + j) * Bp];
#pragma unroll
        for (int j = 0; j < 6; ++j) s.xi[j] = base[(size_t)(7 + j) * Bp];
    } else {
#pragma unroll
        for (int j = 0; j < 3; ++j) s.xi[j] = base[(size_t)(4 + j) * Bp];
    }
}
template <int KIND>
TO_DEV void store_state(double* __restrict__ X, int stage, int Bp, int b, const State<KIND>& s) {
    constexpr int NS = Dims<KIND>::NS;
    double* base = X + soa(stage, 0, NS, Bp, b);
#pragma unroll
    for (int j = 0; j < 4; ++j) base[(size_t)j * Bp] = s.q[j];
    if constexpr (!on_so3(KIND)) {
#pragma unroll
        for (int j = 0; j < 3; ++j) base[(size_t)(4 + j) * Bp] = s.p[j];
#pragma unroll
        for (int j = 0; j < 6; ++j) base[(size_t)(7 + j) * Bp] = s.xi[j];
    } else {
#pragma unroll
        for (int j = 0; j < 3; ++j) base[(size_t)(4 + j) * Bp] = s.xi[j];
    }
}
template <int KIND>
TO_DEV void load_ref_state(const double* __restrict__ ref, int stage, State<KIND>& s) {
    const double* r = ref + (size_t)stage * RefRow<KIND>::N;
#pragma unroll
    for (int j = 0; j < 4; ++j) s.q[j] = r[j];
    if constexpr (!on_so3(KIND)) {
#pragma unroll
        for (int j = 0; j < 3; ++j) s.p[j] = r[4 + j];
#pragma unroll
        for (int j = 0; j < 6; ++j) s.xi[j] = r[7 + j];
    } else {
#pragma unroll
        for (int j = 0; j < 3; ++j) s.xi[j] = r[4 + j];
    }
}

// ------------------------------------------------------------------------------------------
// Dynamics: x+ = f(x, u)
// ------------------------------------------------------------------------------------------
template <int KIND>
TO_DEV void dyn_step(const Params& prm, const State<KIND>& x, const double* u, State<KIND>& xn) {
    const double dt = prm.dt;
    if constexpr (on_so3(KIND)) {
        // q+ = q Exp(w dt);  w+ = w + J^-1 (w^T^ J w + u) dt      (traopt_dynamics.py:375-379)
        const double th[3] = {x.xi[0] * dt, x.xi[1] * dt, x.xi[2] * dt};
        double qe[4];
        so3_exp(th, qe);
        quat_compose(x.q, qe, xn.q);
        double Jw[3], c[3], r[3], a[3];
        mv3(prm.Ib, x.xi, Jw);
        cross3(x.xi, Jw, c);                 // w^ J w ;  (w^)^T J w = -w x Jw
        if constexpr (KIND == TRAJOPT_PEND) {
            // torque = (m g rho) x R^T(-e3) + (m rho) x R^T u,  rho = (0, 0, -l/2)      (traopt_dynamics.py:520-541)
            const double down[3] = {0.0, 0.0, -1.0};
            const double mrho[3] = {0.0, 0.0, -0.5 * prm.length * prm.mass};
            const double mgrho[3] = {0.0, 0.0, mrho[2] * prm.grav};
            double gb[3], ub[3], t1[3], t2[3];
            quat_rotate_inv(x.q, down, gb);
            quat_rotate_inv(x.q, u, ub);
            cross3(mgrho, gb, t1);
            cross3(mrho, ub, t2);
#pragma unroll
            for (int i = 0; i < 3; ++i) r[i] = (t1[i] + t2[i]) - c[i];
        } else {
#pragma unroll
            for (int i = 0; i < 3; ++i) r[i] = u[i] - c[i];
        }
        mv3(prm.Ibinv, r, a);
#pragma unroll
        for (int i = 0; i < 3; ++i) xn.xi[i] = x.xi[i] + a[i] * dt;
    } else {
        // q+ = q Exp(xi dt)                                        (traopt_dynamics.py:783, 1397)
        const double* w = x.xi;
        const double* v = x.xi + 3;
        const double tau[6] = {w[0] * dt, w[1] * dt, w[2] * dt, v[0] * dt, v[1] * dt, v[2] * dt};
        double qe[4], pe[3];
        se3_exp(tau, qe, pe);
        se3_compose(x.q, x.p, qe, pe, xn.q, xn.p);
        // xi+ = xi + J^-1 (ad_xi^T J xi + F) dt                    (:785, :1399)
        //   ad_xi^T J xi = [-w x (Ib w) ; -m w x v]
        double Jw[3], c1[3], c2[3], fw[3], fv[3];
        mv3(prm.Ib, w, Jw);
        cross3(w, Jw, c1);
        cross3(w, v, c2);
        if constexpr (has_gravity(KIND)) {
            // drone: F = [tau_u ; (0,0,f_z) + m g R^T (-e3)]        (:1393-1399, Pu :1250-1254)
            // rigid: F = [u_w   ; u_v       + m g R^T (-e3)]        (:1066-1072)
            const double down[3] = {0.0, 0.0, -1.0};
            double gb[3];
            quat_rotate_inv(x.q, down, gb);
            const double mg = prm.mass * prm.grav;
            fw[0] = u[0]; fw[1] = u[1]; fw[2] = u[2];
            if constexpr (KIND == TRAJOPT_DRONE) {
                fv[0] = mg * gb[0]; fv[1] = mg * gb[1]; fv[2] = mg * gb[2] + u[3];
            } else {
                fv[0] = mg * gb[0] + u[3]; fv[1] = mg * gb[1] + u[4]; fv[2] = mg * gb[2] + u[5];
            }
        } else {
            fw[0] = u[0]; fw[1] = u[1]; fw[2] = u[2];
            fv[0] = u[3]; fv[1] = u[4]; fv[2] = u[5];
        }
        double rw[3], aw[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) rw[i] = fw[i] - c1[i];
        mv3(prm.Ibinv, rw, aw);
        const double im = 1.0 / prm.mass;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            xn.xi[i] = w[i] + aw[i] * dt;
            xn.xi[3 + i] = v[i] + (fv[i] - prm.mass * c2[i]) * im * dt;
        }
    }
}

// ------------------------------------------------------------------------------------------
// Tracking error e = Log(q q_ref^-1) (global / left error, manif lminus) and velocity error
// ------------------------------------------------------------------------------------------
template <int KIND>
TO_DEV void tracking_error(const State<KIND>& x, const double* __restrict__ refrow, double* e, double* dxi) {
    if constexpr (on_so3(KIND)) {
        double qe[4];
        quat_compose_inv_r(x.q, refrow, qe);
        so3_log(qe, e);
#pragma unroll
        for (int i = 0; i < 3; ++i) dxi[i] = x.xi[i] - refrow[4 + i];
    } else {
        double qe[4], pe[3];
        se3_compose_inv_r(x.q, x.p, refrow, refrow + 4, qe, pe);
        se3_log(qe, pe, e);
#pragma unroll
        for (int i = 0; i < 6; ++i) dxi[i] = x.xi[i] - refrow[7 + i];
    }
}

// AL terms of InputConstraint g = [lb - u; u - ub] with multipliers lam[2NU], penalties imu[2NU]
//   value: lam^T g + 1/2 g^T Imu g ;  l_u += g_u^T (lam + Imu g) ;  l_uu += g_u^T Imu g_u (diagonal)
template <int M>
TO_DEV double al_box_terms(const double* lb, const double* ub, const double* v, const double* lam, const double* imu,
                           double* l_add, double* ll_add) {
    double val = 0.0;
#pragma unroll
    for (int j = 0; j < M; ++j) {
        const double g0 = lb[j] - v[j];
        const double g1 = v[j] - ub[j];
        val += lam[j] * g0 + lam[M + j] * g1;
        const double t0 = lam[j] + imu[j] * g0;
        const double t1 = lam[M + j] + imu[M + j] * g1;
        l_add[j] = t1 - t0;
        ll_add[j] = imu[j] + imu[M + j];
    }
    double quad = 0.0;
#pragma unroll
    for (int j = 0; j < M; ++j) {
        const double g0 = lb[j] - v[j];
        const double g1 = v[j] - ub[j];
        quad += g0 * imu[j] * g0 + g1 * imu[M + j] * g1;
    }
    return val + 0.5 * quad;
}
template <int NU>
TO_DEV double al_terms(const Params& prm, const double* u, const double* lam, const double* imu, double* lu_add, double* luu_add) {
    return al_box_terms<NU>(prm.lb, prm.ub, u, lam, imu, lu_add, luu_add);
}

// Stage cost value only (forward passes):  e^T W1 e + dxi^T W2 dxi (+ u^T R u)
template <int KIND>
TO_DEV double stage_cost(const Params& prm, const State<KIND>& x, const double* u, const double* __restrict__ refrow, bool terminal) {
    constexpr int NP = Dims<KIND>::NP, NU = Dims<KIND>::NU, NV = Dims<KIND>::NX - Dims<KIND>::NP;
    double e[NP], dxi[NV];
    tracking_error<KIND>(x, refrow, e, dxi);
    // SO3 quirk 3: the terminal value uses Q
    const bool useP = terminal && !(on_so3(KIND) && prm.so3_terminal_quirk);
    const double* W1 = useP ? prm.P1 : prm.W1;
    const double* W2 = useP ? prm.P2 : prm.W2;
    double c = 0.0;
#pragma unroll
    for (int i = 0; i < NP; ++i) {
        double g = 0.0;
#pragma unroll
        for (int j = 0; j < NP; ++j) g += W1[i * NP + j] * e[j];
        c += e[i] * g;
    }
    double cv = 0.0;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        double g = 0.0;
#pragma unroll
        for (int j = 0; j < NV; ++j) g += W2[i * NV + j] * dxi[j];
        cv += dxi[i] * g;
    }
    c += cv;
    if (!terminal) {
        double cu = 0.0;
#pragma unroll
        for (int i = 0; i < NU; ++i) {
            double g = 0.0;
#pragma unroll
            for (int j = 0; j < NU; ++j) g += prm.R[i * NU + j] * u[j];
            cu += u[i] * g;
        }
        c += cu;
    }
    return c;
}

// ------------------------------------------------------------------------------------------
// Cost expansion: value, l_x (NX), pose block of l_xx (packed upper triangle, NP(NP+1)/2)
//   J_e = Jr^-1(e) Ad(q_ref);   l_x = [2 J_e^T W1 e ; 2 W2 dxi];   l_xx = blkdiag(2 J_e^T W1 J_e, 2 W2)
// ------------------------------------------------------------------------------------------
template <int KIND>
TO_DEV double cost_expand(const Params& prm, const State<KIND>& x, const double* __restrict__ refrow, bool terminal,
                          double* lx, double* lxx) {
    constexpr int NP = Dims<KIND>::NP, NV = Dims<KIND>::NX - Dims<KIND>::NP;
    double e[NP], dxi[NV];
    tracking_error<KIND>(x, refrow, e, dxi);
    const bool quirk = (on_so3(KIND)) && prm.so3_terminal_quirk;
    const double* W1v = (terminal && !quirk) ? prm.P1 : prm.W1;    // value + gradient
    const double* W2v = (terminal && !quirk) ? prm.P2 : prm.W2;
    const double* W1h = terminal ? prm.P1 : prm.W1;                // Hessian
    double Je[NP * NP];
    if constexpr (on_so3(KIND)) {
        double Ji[9];
        so3_jr_inv(e, Ji);
        mm3(Ji, refrow + 7, Je);                                   // Jr^-1(e) R_ref
    } else {
        // Jr^-1(e) = [[Ji, 0], [-Ji Q(-w,-v) Ji, Ji]] ;  Ad(ref) = [[Rr, 0], [pr^ Rr, Rr]]
        double Ji[9], Qm[9], T[9], Z[9], X[9], Y[9], Y2[9];
        so3_jr_inv(e, Ji);
        const double ne[6] = {-e[0], -e[1], -e[2], -e[3], -e[4], -e[5]};
        se3_Q(ne, ne + 3, Qm);
        mm3(Ji, Qm, T);
        mm3(T, Ji, Z);                                             // Ji Q Ji  (sign applied below)
        mm3(Ji, refrow + 13, X);                                   // Ji Rr
        mm3(Z, refrow + 13, Y);                                    // (Ji Q Ji) Rr
        mm3(Ji, refrow + 22, Y2);                                  // Ji (pr^ Rr)
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                Je[i * 6 + j] = X[3 * i + j];
                Je[i * 6 + 3 + j] = 0.0;
                Je[(3 + i) * 6 + j] = Y2[3 * i + j] - Y[3 * i + j];
                Je[(3 + i) * 6 + 3 + j] = X[3 * i + j];
            }
    }
    // g = W1 e ; value
    double g[NP], val = 0.0;
#pragma unroll
    for (int i = 0; i < NP; ++i) {
        double s = 0.0;
#pragma unroll
        for (int j = 0; j < NP; ++j) s += W1v[i * NP + j] * e[j];
        g[i] = s;
        val += e[i] * s;
    }
    double valv = 0.0;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        double s = 0.0;
#pragma unroll
        for (int j = 0; j < NV; ++j) s += W2v[i * NV + j] * dxi[j];
        lx[NP + i] = 2.0 * s;
        valv += dxi[i] * s;
    }
    val += valv;
    // l_x pose = 2 Je^T g
#pragma unroll
    for (int c = 0; c < NP; ++c) {
        double s = 0.0;
#pragma unroll
        for (int r = 0; r < NP; ++r) {
            if (!on_so3(KIND) && r < 3 && c >= 3) continue;  // zero block of Je
            s += Je[r * NP + c] * g[r];
        }
        lx[c] = 2.0 * s;
    }
    // M = W1h Je ;  l_xx = 2 Je^T M  (upper triangle, row-major packed)
    double M[NP * NP];
#pragma unroll
    for (int r = 0; r < NP; ++r)
#pragma unroll
        for (int c = 0; c < NP; ++c) {
            double s = 0.0;
#pragma unroll
            for (int k = 0; k < NP; ++k) {
                if (!on_so3(KIND) && k < 3 && c >= 3) continue;
                s += W1h[r * NP + k] * Je[k * NP + c];
            }
            M[r * NP + c] = s;
        }
    int idx = 0;
#pragma unroll
    for (int r = 0; r < NP; ++r)
#pragma unroll
        for (int c = r; c < NP; ++c) {
            double s = 0.0;
#pragma unroll
            for (int k = 0; k < NP; ++k) {
                if (!on_so3(KIND) && k < 3 && r >= 3) continue;
                s += Je[k * NP + r] * M[k * NP + c];
            }
            lxx[idx++] = 2.0 * s;
        }
    return val;
}

// ------------------------------------------------------------------------------------------
// Dynamics Jacobian blocks (see LinRec in common.cuh for the record layout)
// ------------------------------------------------------------------------------------------
template <int KIND>
TO_DEV void dyn_jacobian(const Params& prm, const State<KIND>& x, const double* u, double* rec) {
    const double dt = prm.dt;
    if constexpr (on_so3(KIND)) {
        // [[Exp(w dt)^T, Jr(w dt) dt], [0, I + J^-1 (w^T^ J + s(Jw)) dt]]   (:385-400)
        const double* w = x.xi;
        const double nth[3] = {-w[0] * dt, -w[1] * dt, -w[2] * dt};
        const double th[3] = {w[0] * dt, w[1] * dt, w[2] * dt};
        double qe[4];
        so3_exp(nth, qe);
        quat_to_rot(qe, rec + 0);                                  // a
        double Jr[9];
        so3_jr(th, Jr);
#pragma unroll
        for (int i = 0; i < 9; ++i) rec[9 + i] = Jr[i] * dt;       // c
        double Jw[3], M[9], T[9], H[9];
        mv3(prm.Ib, w, Jw);
        skew_mul(w, prm.Ib, T);                                    // w^ Ib
        const double S[9] = {0, -Jw[2], Jw[1], Jw[2], 0, -Jw[0], -Jw[1], Jw[0], 0};
#pragma unroll
        for (int i = 0; i < 9; ++i) M[i] = S[i] - T[i];            // (w^)^T Ib + s(Ib w)
        mm3(prm.Ibinv, M, H);
#pragma unroll
        for (int i = 0; i < 9; ++i) rec[18 + i] = ((i % 4 == 0) ? 1.0 : 0.0) + dt * H[i];   // h
        if constexpr (KIND == TRAJOPT_PEND) {
            // l = J^-1 (s(m g rho) s(R^T(-e3)) + s(m rho) s(R^T u)) dt          (traopt_dynamics.py:563-577)
            // bv = J^-1 s(m rho) R^T dt                                          (:590-603)
            const double down[3] = {0.0, 0.0, -1.0};
            const double mrho[3] = {0.0, 0.0, -0.5 * prm.length * prm.mass};
            const double mgrho[3] = {0.0, 0.0, mrho[2] * prm.grav};
            double gb[3], ub[3], S1[9], S2[9], L1[9], L2[9], Lm[9], Rm[9], Rt[9], B0[9], Bm[9];
            quat_rotate_inv(x.q, down, gb);
            quat_rotate_inv(x.q, u, ub);
            const double Sg[9] = {0, -gb[2], gb[1], gb[2], 0, -gb[0], -gb[1], gb[0], 0};
            const double Su[9] = {0, -ub[2], ub[1], ub[2], 0, -ub[0], -ub[1], ub[0], 0};
            skew_mul(mgrho, Sg, S1);
            skew_mul(mrho, Su, S2);
#pragma unroll
            for (int i = 0; i < 9; ++i) L1[i] = S1[i] + S2[i];
            mm3(prm.Ibinv, L1, Lm);
#pragma unroll
            for (int i = 0; i < 9; ++i) rec[27 + i] = Lm[i] * dt;                               // l
            (void)L2;
            quat_to_rot(x.q, Rm);
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int j = 0; j < 3; ++j) Rt[3 * i + j] = Rm[3 * j + i];
            skew_mul(mrho, Rt, B0);
            mm3(prm.Ibinv, B0, Bm);
#pragma unroll
            for (int i = 0; i < 9; ++i) rec[36 + i] = Bm[i] * dt;                               // bv
        }
    } else {
        const double* w = x.xi;
        const double* v = x.xi + 3;
        const double ntau[6] = {-w[0] * dt, -w[1] * dt, -w[2] * dt, -v[0] * dt, -v[1] * dt, -v[2] * dt};
        const double th[3] = {w[0] * dt, w[1] * dt, w[2] * dt};
        // A11 = Ad(Exp(-tau)) = [[R', 0], [p'^ R', R']]             (:821-825)
        double qe[4], pe[3];
        se3_exp(ntau, qe, pe);
        quat_to_rot(qe, rec + 0);                                  // a
        skew_mul(pe, rec + 0, rec + 9);                            // b
        // A12 = Jr(tau) dt = Jl(-tau) dt = [[Jr(w dt), 0], [Q(-w dt, -v dt), Jr(w dt)]] dt   (:826)
        double Jr[9], Qm[9];
        so3_jr(th, Jr);
        se3_Q(ntau, ntau + 3, Qm);
#pragma unroll
        for (int i = 0; i < 9; ++i) {
            rec[18 + i] = Jr[i] * dt;                              // c
            rec[27 + i] = Qm[i] * dt;                              // e
        }
        // A22 = I + H dt,  H = J^-1 (ad_{[v,w]}^T J + G)           (QUIRK 1, :828-832)
        //     = [[Ib^-1 (s(Ib w) - v^ Ib),  m Ib^-1 (v^ - w^)], [v^, -v^]]
        double Jw[3], T[9], M[9], H11[9], H12[9];
        mv3(prm.Ib, w, Jw);
        skew_mul(v, prm.Ib, T);
        const double S[9] = {0, -Jw[2], Jw[1], Jw[2], 0, -Jw[0], -Jw[1], Jw[0], 0};
#pragma unroll
        for (int i = 0; i < 9; ++i) M[i] = S[i] - T[i];
        mm3(prm.Ibinv, M, H11);
        const double dvw[3] = {v[0] - w[0], v[1] - w[1], v[2] - w[2]};
        mul_skew(prm.Ibinv, dvw, H12);
        const double md = prm.mass * dt;
#pragma unroll
        for (int i = 0; i < 9; ++i) {
            rec[36 + i] = ((i % 4 == 0) ? 1.0 : 0.0) + dt * H11[i];    // h11
            rec[45 + i] = md * H12[i];                                 // h12
        }
        rec[54] = v[0] * dt; rec[55] = v[1] * dt; rec[56] = v[2] * dt; // vdt: h21 = vdt^, h22 = I - vdt^
        if constexpr (has_gravity(KIND)) {
            // A21 = J^-1 [[0,0],[s(R^T(-e3)),0]] dt, WITHOUT m g      (QUIRK 2, :1445-1458; rigid body :1121-1134)
            const double down[3] = {0.0, 0.0, -1.0};
            double gb[3];
            quat_rotate_inv(x.q, down, gb);
            const double sdt = dt / prm.mass;
            rec[57] = gb[0] * sdt; rec[58] = gb[1] * sdt; rec[59] = gb[2] * sdt;
        }
    }
}

// Multiple-shooting defect d = [Log(x_next.q^-1 f.q) ; f.xi - x_next.xi]   (:2882-2888)
template <int KIND>
TO_DEV void defect(const State<KIND>& fx, const State<KIND>& xnext, double* d) {
    if constexpr (on_so3(KIND)) {
        double qd[4];
        quat_compose_inv_l(xnext.q, fx.q, qd);
        so3_log(qd, d);
#pragma unroll
        for (int i = 0; i < 3; ++i) d[3 + i] = fx.xi[i] - xnext.xi[i];
    } else {
        double qd[4], pd[3];
        se3_compose_inv_l(xnext.q, xnext.p, fx.q, fx.p, qd, pd);
        se3_log(qd, pd, d);
#pragma unroll
        for (int i = 0; i < 6; ++i) d[6 + i] = fx.xi[i] - xnext.xi[i];
    }
}

// State difference  a (-) b = [Log(b.q^-1 a.q) ; a.xi - b.xi]  (manif rminus, :2056-2062)
template <int KIND>
TO_DEV void state_minus(const State<KIND>& a, const State<KIND>& b, double* dx) { defect<KIND>(a, b, dx); }

}  // namespace trajopt
