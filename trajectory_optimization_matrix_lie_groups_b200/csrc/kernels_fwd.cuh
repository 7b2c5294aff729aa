// Forward passes, line-search bookkeeping, augmented-Lagrangian update and layout exports.
#pragma once
#include "kernels.cuh"
#include "backward.cuh"
#include "backward3.cuh"
#include "backward4.cuh"
#include "backward6.cuh"

namespace trajopt {

// ------------------------------------------------------------------------------------------
// Forward rollout of one candidate step size per thread.
//   thread t -> problem b = t % Bp, candidate a = a_lo + t / Bp
//   MS  (:2641-2740):  dx = x_new(i) (-) x(i);  du = alpha k + K dx;  u_new = u + du
//        nonlinear:  q_new(i+1) = q(i+1) Exp(alpha d_q) f(x,u).q^-1 f(x_new,u_new).q
//                    xi_new(i+1) = xi(i+1) + f_new.xi - f.xi + alpha d_xi
//        linear   :  x_new(i+1) = x(i+1) (+) (A dx + B du + alpha d)
//   SS  (:2030-2082):  nonlinear: x_new(i+1) = f(x_new(i), u_new(i));  linear: x(i+1) (+) (A dx + B du)
//   WRITE  : store the candidate trajectory into the problem's other buffer
//   COST   : accumulate J_new = sum_i l(x_new_i, u_new_i) + l_N left to right (:2084-2096);
//            MS additionally accumulates the candidate's defect norm (:2565-2566)
// Which problems run: `need` < 0 -> every running problem; otherwise those with ls_state == need.
// The three passes of a line search (step size 0; the remaining ones; the accepted one again) all run the WRITE && COST
// instantiation, so that a candidate's cost and its kept trajectory come from the same instruction sequence whichever
// pass produced them (separate instantiations may contract a * b + c * d differently); the pass that only needs costs
// (need == -1) skips the stores at run time.
// need == -4 (small batches, WRITE && COST): every candidate keeps its trajectory in ITS OWN buffer (Work::Xc / Uc) so that
// one launch serves the whole line search; k_ls_copy_cand then moves the accepted one into the problem's other buffer.
// ------------------------------------------------------------------------------------------
// value of the AL terms of the velocity bounds at (stage, problem) for velocity xi
template <int KIND>
TO_DEV double al_state_value(const Params& prm, const Work& w, int stage, int Bp, int b, const double* xi) {
    constexpr int NV = Dims<KIND>::NX - Dims<KIND>::NP;
    double lam[2 * NV], imu[2 * NV], t1[NV], t2[NV];
#pragma unroll
    for (int j = 0; j < 2 * NV; ++j) {
        lam[j] = w.lam_s[soa(stage, j, 2 * NV, Bp, b)];
        imu[j] = w.imu_s[soa(stage, j, 2 * NV, Bp, b)];
    }
    return al_box_terms<NV>(prm.xlb, prm.xub, xi, lam, imu, t1, t2);
}

// shared memory of the rollouts: the gains of one stage of 32 problems (GainRec), double-buffered
template <int KIND> struct FwdSmem {
    using GR = GainRec<KIND>;
    static constexpr int BUF_DOUBLES = GR::LEN * 32;                     // one stage of gains of 32 problems
    static constexpr uint32_t BUF_BYTES = (uint32_t)BUF_DOUBLES * 8;
    static constexpr size_t BYTES = (size_t)2 * BUF_BYTES + 16;          // double buffer + two mbarriers
};

template <int KIND, bool MS, bool LINEAR, bool WRITE, bool COST>
__global__ void __launch_bounds__(kBlock) k_forward(const Params prm, Work w, int a_lo, int a_cnt, int need,
                                                    int use_ls_state_as_alpha) {
    using D = Dims<KIND>;
    using LR = LinRec<KIND>;
    using GR = GainRec<KIND>;
    using FS = FwdSmem<KIND>;
    constexpr int NX = D::NX, NP = D::NP, NU = D::NU, NV = NX - NP, F = LR::LEN;
    extern __shared__ __align__(128) double fsm[];
    const int lane = threadIdx.x;
    const int t = blockIdx.x * kBlock + lane;
    const int Bp = prm.Bp, N = prm.N;
    // a block is 32 consecutive threads and Bp is a multiple of 32: its lanes are 32 consecutive problems (one gain group)
    // with the same candidate index t / Bp
    const int b = t % Bp;
    int ai = a_lo + t / Bp;
    bool run = (t / Bp < a_cnt) && (b < prm.B) && (w.status[b] == TRAJOPT_RUNNING);
    if (run && need >= -1 && w.ls_state[b] != need && !use_ls_state_as_alpha) run = false;
    if (run && use_ls_state_as_alpha) {
        ai = w.ls_state[b];
        if (ai < a_lo) run = false;  // a_lo = 1: step size 0 was already written by the first pass
    }
    const unsigned runmask = __ballot_sync(0xffffffffu, run);
    if (runmask == 0u) return;
    // alpha = 1.1 ** (-ai**2)  (:1908, :2472); index 0 is exactly 1.0
    const double alpha = (ai == 0) ? 1.0 : pow(1.1, -(double)(ai * ai));

    const int cur = w.sel[b];
    const double* X = w.X[cur];
    const double* U = w.U[cur];
    double* Xn = w.X[1 - cur];
    double* Un = w.U[1 - cur];
    if (WRITE && COST && need == -4) {
        Xn = w.Xc + (size_t)ai * (size_t)(N + 1) * D::NS * Bp;
        Un = w.Uc + (size_t)ai * (size_t)N * NU * Bp;
    }
    const double* lin = w.lin;
    const bool store = WRITE && !(COST && need == -1);

    // gains of the warp's 32 problems, one stage at a time, double-buffered by one TMA bulk copy per stage (as in
    // k_forward_ms_full); the lanes that do not run keep the warp's staging company
    const uint32_t bar0 = b3_smem_addr(fsm + 2 * FS::BUF_DOUBLES);
    const uint32_t bar1 = bar0 + 8;
    if (lane == 0) {
        b3_mbar_init(bar0, 1);
        b3_mbar_init(bar1, 1);
    }
    __syncwarp();
    const double* gains0 = w.gains + lsoa(0, 0, GR::LEN, N, b - lane);
    auto issue = [&](int stage, int buf) {
        if (lane == 0)
            b3_tma_load(b3_smem_addr(fsm + buf * FS::BUF_DOUBLES), gains0 + (size_t)stage * FS::BUF_DOUBLES, FS::BUF_BYTES, buf ? bar1 : bar0);
    };
    uint32_t ph0 = 0, ph1 = 0;

    State<KIND> xnew, x, xnext;
    load_state<KIND>(X, 0, Bp, b, x);
    xnew = x;
    if (store && run) store_state<KIND>(Xn, 0, Bp, b, xnew);
    double J = 0.0, dsq = 0.0;
    const int Nb = w.Nb[b];              // this problem's horizon (<= N)
    int Nmax = run ? Nb : 0;             // stages the warp stays together for
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) Nmax = max(Nmax, __shfl_xor_sync(0xffffffffu, Nmax, o));
    issue(0, 0);
    if (Nmax > 1) issue(1, 1);

    for (int i = 0; i < Nmax; ++i) {
        const int buf = i & 1;
        if (buf) { b3_mbar_wait(bar1, ph1); ph1 ^= 1u; } else { b3_mbar_wait(bar0, ph0); ph0 ^= 1u; }
        const double* sb = fsm + buf * FS::BUF_DOUBLES + lane;
        if (run && i < Nb) {
        double refrow[RefRow<KIND>::N];
        if (COST) fetch_ref_row<KIND>(w, Bp, i, b, refrow);
        double dx[NX];
        state_minus<KIND>(xnew, x, dx);
        double u[NU], unew[NU], du[NU];
#pragma unroll
        for (int a = 0; a < NU; ++a) {
            u[a] = U[soa(i, a, NU, Bp, b)];
            double s = alpha * sb[(GR::KFF_OFF + a) * 32];
#pragma unroll
            for (int c = 0; c < NX; ++c) s += sb[(GR::K_OFF + a * NX + c) * 32] * dx[c];
            du[a] = s;
            unew[a] = u[a] + s;
        }
        if (store) {
#pragma unroll
            for (int a = 0; a < NU; ++a) Un[soa(i, a, NU, Bp, b)] = unew[a];
        }
        if (COST) {
            double c = stage_cost<KIND>(prm, xnew, unew, refrow, false);
            if (prm.has_constraints) {
                double lam[2 * NU], imu[2 * NU], t1[NU], t2[NU];
#pragma unroll
                for (int j = 0; j < 2 * NU; ++j) {
                    lam[j] = w.lam[soa(i, j, 2 * NU, Bp, b)];
                    imu[j] = w.imu[soa(i, j, 2 * NU, Bp, b)];
                }
                c += al_terms<NU>(prm, unew, lam, imu, t1, t2);
            }
            if (prm.has_state_bounds) c += al_state_value<KIND>(prm, w, i, Bp, b, xnew.xi);
            J = J + c;
        }
        load_state<KIND>(X, i + 1, Bp, b, xnext);

        State<KIND> xn1;
        if constexpr (LINEAR) {
            // step = A dx + B du (+ alpha d), pose part retracts x(i+1)
            AMat<KIND> A;
            A.load(lin, i, N + 1, b);
            const BvStage<KIND> Bv(prm, lin + lsoa(i, 0, F, N + 1, b), kRecStride);
            double st[NX];
#pragma unroll
            for (int r = 0; r < NX; ++r) {
                double s = 0.0;
#pragma unroll
                for (int c = 0; c < NX; ++c)
                    if (AMat<KIND>::nz(r, c)) s += A.get(r, c) * dx[c];
                if (r >= NP) {
#pragma unroll
                    for (int a = 0; a < NU; ++a)
                        if (bv_nz<KIND>(r - NP, a)) s += Bv.get(r - NP, a) * du[a];
                }
                if constexpr (MS) s += alpha * lin[lsoa(i, LR::D_OFF + r, F, N + 1, b)];
                st[r] = s;
            }
            if constexpr (on_so3(KIND)) {
                double qe[4];
                so3_exp(st, qe);
                quat_compose(xnext.q, qe, xn1.q);
            } else {
                double qe[4], pe[3];
                se3_exp(st, qe, pe);
                se3_compose(xnext.q, xnext.p, qe, pe, xn1.q, xn1.p);
            }
#pragma unroll
            for (int j = 0; j < NV; ++j) xn1.xi[j] = xnext.xi[j] + st[NP + j];
        } else {
            State<KIND> fnew;
            dyn_step<KIND>(prm, xnew, unew, fnew);
            if constexpr (MS) {
                State<KIND> fold;
                dyn_step<KIND>(prm, x, u, fold);
                double d[NX];
#pragma unroll
                for (int j = 0; j < NX; ++j) d[j] = alpha * lin[lsoa(i, LR::D_OFF + j, F, N + 1, b)];
                if constexpr (on_so3(KIND)) {
                    double qe[4], q1[4], q2[4];
                    so3_exp(d, qe);
                    quat_compose(xnext.q, qe, q1);
                    quat_compose_inv_r(q1, fold.q, q2);
                    quat_compose(q2, fnew.q, xn1.q);
                } else {
                    double qe[4], pe[3], q1[4], p1[3], q2[4], p2[3];
                    se3_exp(d, qe, pe);
                    se3_compose(xnext.q, xnext.p, qe, pe, q1, p1);
                    se3_compose_inv_r(q1, p1, fold.q, fold.p, q2, p2);
                    se3_compose(q2, p2, fnew.q, fnew.p, xn1.q, xn1.p);
                }
#pragma unroll
                for (int j = 0; j < NV; ++j) xn1.xi[j] = xnext.xi[j] + fnew.xi[j] - fold.xi[j] + d[NP + j];
            } else {
                xn1 = fnew;
            }
        }
        if (MS && COST) {
            // defect of the candidate: f(x_new_i, u_new_i) (-) x_new(i+1)   (:2790-2810)
            State<KIND> fnew;
            dyn_step<KIND>(prm, xnew, unew, fnew);
            double dd[NX];
            defect<KIND>(fnew, xn1, dd);
#pragma unroll
            for (int j = 0; j < NX; ++j) dsq += dd[j] * dd[j];
        }
        if (store) store_state<KIND>(Xn, i + 1, Bp, b, xn1);
        xnew = xn1;
        x = xnext;
        }   // run && i < Nb
        __syncwarp();                       // every lane is done with this buffer
        if (i + 2 < Nmax) issue(i + 2, buf);
    }
    if (COST && run) {
        double refrow[RefRow<KIND>::N];
        fetch_ref_row<KIND>(w, Bp, Nb, b, refrow);
        J = J + stage_cost<KIND>(prm, xnew, nullptr, refrow, true);
        if (prm.has_state_bounds) J = J + al_state_value<KIND>(prm, w, Nb, Bp, b, xnew.xi);
        w.Jcand[(size_t)ai * Bp + b] = J;
        if (MS) w.Jcand[(size_t)(prm.n_alphas + ai) * Bp + b] = sqrt(dsq);
    }
}

// ------------------------------------------------------------------------------------------
// The hot rollout: multiple shooting, rollout='nonlinear', full step (alpha = 1), written to the other
// buffer, no cost (the next linearisation evaluates it).  One warp per 32 problems, the horizon runs
// sequentially inside the thread.  Per stage the gains K_i (NU x NX) and k_i of the warp's 32 problems
// — one contiguous chunk of the group-major gain array (GainRec) — are staged in shared memory,
// double-buffered, by ONE TMA bulk copy issued two stages ahead (mbarrier completion), so the recursion
// never waits on HBM for them; u_i, G_i and f(x_i,u_i).xi (precomputed by the linearisation, GPre) are
// loaded at the top of the stage, before the chain needs them.
//   dx = x_new(i) (-) x(i);  u_new = u + k + K dx;  f_new = f(x_new, u_new)
//   q_new(i+1) = G_i f_new.q;  xi_new(i+1) = xi(i+1) + f_new.xi - f.xi + d_xi          (:2697-2718)
// ------------------------------------------------------------------------------------------
template <int KIND>
__global__ void __launch_bounds__(kBlock) k_forward_ms_full(const Params prm, Work w, int i0, int i1) {
    using D = Dims<KIND>;
    using LR = LinRec<KIND>;
    using GR = GainRec<KIND>;
    using FS = FwdSmem<KIND>;
    constexpr int NX = D::NX, NP = D::NP, NU = D::NU, NV = NX - NP, F = LR::LEN;
    constexpr int GL = GPre<KIND>::LEN, GP = GPre<KIND>::NPOSE;
    extern __shared__ __align__(128) double fsm[];
    const int lane = threadIdx.x;
    const int b = blockIdx.x * kBlock + lane;
    const int Bp = prm.Bp, N = prm.N;
    const bool act = (b < prm.B) && (w.status[b] == TRAJOPT_RUNNING);
    if (__ballot_sync(0xffffffffu, act) == 0u) return;
    const int cur = w.sel[b];
    const double* X = w.X[cur];
    const double* U = w.U[cur];
    double* Xn = w.X[1 - cur];
    double* Un = w.U[1 - cur];
    const uint32_t bar0 = b3_smem_addr(fsm + 2 * FS::BUF_DOUBLES);
    const uint32_t bar1 = bar0 + 8;
    if (lane == 0) {
        b3_mbar_init(bar0, 1);
        b3_mbar_init(bar1, 1);
    }
    __syncwarp();
    // the gains of stage s of this warp's 32 problems: one contiguous chunk (GainRec), one bulk copy by one lane
    const double* gains0 = w.gains + lsoa(0, 0, GR::LEN, N, b - lane);
    auto issue = [&](int stage, int buf) {
        if (lane == 0)
            b3_tma_load(b3_smem_addr(fsm + buf * FS::BUF_DOUBLES), gains0 + (size_t)stage * FS::BUF_DOUBLES, FS::BUF_BYTES, buf ? bar1 : bar0);
    };
    // stages [i0, i1) of the horizon: the whole of it, or one chunk (run_forward_overlapped), resuming from x_new(i0)
    issue(i0, i0 & 1);
    if (i0 + 1 < i1) issue(i0 + 1, (i0 + 1) & 1);

    State<KIND> xnew, x, xnext;
    load_state<KIND>(X, i0, Bp, b, x);
    if (i0 == 0) {
        xnew = x;
        if (act) store_state<KIND>(Xn, 0, Bp, b, xnew);
    } else {
        load_state<KIND>(Xn, i0, Bp, b, xnew);
    }
    uint32_t ph0 = 0, ph1 = 0;
    const int Nb = w.Nb[b];              // this problem's horizon; the staging below runs over the group's stages
    // What a stage needs from global memory besides the gains — x(i+1), G_i, f(x_i,u_i).xi, d_xi, u_i — is independent of
    // the recursion: it is loaded ONE STAGE AHEAD into a second set of registers, so that the chain of a stage never waits
    // for it (loaded at the top of its own stage it cost 1.8 of the 5.0 stall cycles per issued instruction: the chain
    // reaches u_i after a few hundred cycles, a miss takes longer).
    struct StageIn {
        State<KIND> xnext;
        double G[GP], fxi[NV], dxi[NV], u[NU];
    };
    auto fetch = [&](int i, StageIn& in) {
        load_state<KIND>(X, i + 1, Bp, b, in.xnext);
        const double* gp = w.Gpre + soa(i, 0, GL, Bp, b);
#pragma unroll
        for (int j = 0; j < GP; ++j) in.G[j] = gp[(size_t)j * Bp];
#pragma unroll
        for (int j = 0; j < NV; ++j) in.fxi[j] = gp[(size_t)(GP + j) * Bp];
#pragma unroll
        for (int j = 0; j < NV; ++j) in.dxi[j] = w.lin[lsoa(i, LR::D_OFF + NP + j, F, N + 1, b)];
#pragma unroll
        for (int a = 0; a < NU; ++a) in.u[a] = U[soa(i, a, NU, Bp, b)];
    };
    StageIn in, nxt;
    fetch(i0, in);
    for (int i = i0; i < i1; ++i) {
        const int buf = i & 1;
        const bool live = act && (i < Nb);
        if (i + 1 < N) fetch(i + 1, nxt);    // (the last stage of a chunk fetches the first of the next: valid rows, unused)
        double dx[NX];
        state_minus<KIND>(xnew, x, dx);
        if (buf) { b3_mbar_wait(bar1, ph1); ph1 ^= 1u; } else { b3_mbar_wait(bar0, ph0); ph0 ^= 1u; }
        const double* sb = fsm + buf * FS::BUF_DOUBLES + lane;
        double unew[NU];
#pragma unroll
        for (int a = 0; a < NU; ++a) {
            // same order as k_forward: s = alpha k (alpha = 1), then += K[a][c] dx[c] for c ascending
            double s = sb[(GR::KFF_OFF + a) * 32];
#pragma unroll
            for (int c = 0; c < NX; ++c) s += sb[(GR::K_OFF + a * NX + c) * 32] * dx[c];
            unew[a] = in.u[a] + s;
        }
        __syncwarp();                       // every lane is done with this buffer
        if (i + 2 < i1) issue(i + 2, buf);
        if (live) {
#pragma unroll
            for (int a = 0; a < NU; ++a) Un[soa(i, a, NU, Bp, b)] = unew[a];
        }
        State<KIND> fnew, xn1;
        dyn_step<KIND>(prm, xnew, unew, fnew);
        if constexpr (on_so3(KIND)) {
            quat_compose(in.G, fnew.q, xn1.q);
        } else {
            se3_compose(in.G, in.G + 4, fnew.q, fnew.p, xn1.q, xn1.p);
        }
#pragma unroll
        for (int j = 0; j < NV; ++j) xn1.xi[j] = in.xnext.xi[j] + fnew.xi[j] - in.fxi[j] + in.dxi[j];
        if (live) store_state<KIND>(Xn, i + 1, Bp, b, xn1);
        xnew = xn1;
        x = in.xnext;
        in = nxt;
    }
}

// MS without line search: alpha = 1 is always accepted (:2592-2600)
static __global__ void k_accept_all(const Params prm, Work w, int it_arg) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= prm.B || w.status[b] != TRAJOPT_RUNNING) return;
    const int it = it_arg >= 0 ? it_arg : w.iters[b];
    w.sel[b] = 1 - w.sel[b];
    w.alphahist[(size_t)it * prm.Bp + b] = 0;
    w.iters[b] = it + 1;
}

// Four counters -> pinned host memory, written by the kernel over PCIe.  A cudaMemcpyAsync would queue behind whatever the
// device->host copy engine is doing — a 1.6 GB trajectory copy of this or another solver held every iteration's 16-byte
// read-back (and with it the iteration loop) for 30 ms.
static __global__ void k_publish4(const int* __restrict__ src, volatile int* dst_host) {
    if (threadIdx.x < 4) dst_host[threadIdx.x] = src[threadIdx.x];
    __threadfence_system();
}

static __global__ void k_ints_to_host(const int* __restrict__ src, volatile int* dst_host, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst_host[i] = src[i];
}

// count problems that are still running into counters[0]
static __global__ void k_count_running(const Params prm, Work w) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    const bool run = (b < prm.B) && (w.status[b] == TRAJOPT_RUNNING);
    const unsigned m = __ballot_sync(0xffffffffu, run);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(&w.counters[0], __popc(m));
}

// SS line search (:1972-1990): examine candidates [a_lo, a_lo+a_cnt) in index order, accept the
// first J_new < J_opt.  first pass (a_lo == 0): the accepted trajectory is already in the other
// buffer; later passes only pick the index and a final k_forward<WRITE> materialises it.
static __global__ void k_ls_select_ss(const Params prm, Work w, int it_arg, int a_lo, int a_cnt, int last) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    const bool in = (b < prm.B) && (w.status[b] == TRAJOPT_RUNNING);
    const int it = it_arg >= 0 ? it_arg : (in ? w.iters[b] : 0);
    bool pending = false;
    if (in && (a_lo == 0 || w.ls_state[b] == -1)) {
        const double Jopt = w.J[b];
        int found = -1;
        for (int a = a_lo; a < a_lo + a_cnt; ++a) {
            if (w.Jcand[(size_t)a * prm.Bp + b] < Jopt) { found = a; break; }
        }
        if (found >= 0) {
            w.ls_state[b] = found;
        } else if (!last) {
            w.ls_state[b] = -1;
            pending = true;
        } else {
            // "Couldn't find descent direction" (:2005-2007): J_hist gets J_opt, solve stops
            w.ls_state[b] = -2;
            w.Jhist[(size_t)it * prm.Bp + b] = Jopt;
            w.alphahist[(size_t)it * prm.Bp + b] = -1;
            w.iters[b] = it + 1;
            w.status[b] = TRAJOPT_NO_DESCENT;
        }
    }
    const unsigned m = __ballot_sync(0xffffffffu, pending);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(&w.counters[1], __popc(m));
}

// small batches: the accepted candidate's trajectory (kept by k_forward, need == -4) -> the problem's other buffer.
// thread = (problem, stage); rows beyond the problem's horizon were not written and are not copied.
template <int KIND>
__global__ void k_ls_copy_cand(const Params prm, Work w) {
    using D = Dims<KIND>;
    const int b = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
    if (b >= prm.B || w.status[b] != TRAJOPT_RUNNING) return;
    const int a = w.ls_state[b];
    if (a < 0 || i > w.Nb[b]) return;
    const int Bp = prm.Bp, N = prm.N, other = 1 - w.sel[b];
    const double* xs = w.Xc + (size_t)a * (size_t)(N + 1) * D::NS * Bp;
    double* xd = w.X[other];
#pragma unroll
    for (int f = 0; f < D::NS; ++f) xd[soa(i, f, D::NS, Bp, b)] = xs[soa(i, f, D::NS, Bp, b)];
    if (i < w.Nb[b]) {
        const double* us = w.Uc + (size_t)a * (size_t)N * D::NU * Bp;
        double* ud = w.U[other];
#pragma unroll
        for (int f = 0; f < D::NU; ++f) ud[soa(i, f, D::NU, Bp, b)] = us[soa(i, f, D::NU, Bp, b)];
    }
}

// commit an accepted SS candidate whose trajectory is in the other buffer
static __global__ void k_ls_commit_ss(const Params prm, Work w, int it_arg) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= prm.B || w.status[b] != TRAJOPT_RUNNING) return;
    const int it = it_arg >= 0 ? it_arg : w.iters[b];
    const int a = w.ls_state[b];
    if (a < 0) return;
    const double Jn = w.Jcand[(size_t)a * prm.Bp + b];
    w.sel[b] = 1 - w.sel[b];
    w.J[b] = Jn;
    w.Jhist[(size_t)it * prm.Bp + b] = Jn;
    w.alphahist[(size_t)it * prm.Bp + b] = a;
    w.iters[b] = it + 1;
    w.ls_state[b] = -2;
    if (it + 1 >= prm.max_iters) w.status[b] = TRAJOPT_MAX_ITER;
}

// ------------------------------------------------------------------------------------------
// MS merit line search (:2549-2590).  k_ms_expected: linear rollout with alpha = 1 for the
// expected cost change (c1, c2) (:2756-2769) and the defect weight (:2774-2788); stored in
// Jcand rows [2 n_alphas .. 2 n_alphas + 3] = c1, c2, d_weight, merit.
// ------------------------------------------------------------------------------------------
template <int KIND>
__global__ void __launch_bounds__(kBlock) k_ms_expected(const Params prm, Work w, double* dweight_prev) {
    using D = Dims<KIND>;
    using LR = LinRec<KIND>;
    constexpr int NX = D::NX, NP = D::NP, NU = D::NU, NV = NX - NP, F = LR::LEN;
    const int b = blockIdx.x * kBlock + threadIdx.x;
    const int Bp = prm.Bp, N = prm.N;
    if (b >= prm.B || w.status[b] != TRAJOPT_RUNNING) return;
    const int Nb = w.Nb[b];
    const double* lin = w.lin;
    const int cur = w.sel[b];
    const double* U = w.U[cur];
    (void)U;
    // In the linear rollout the state error obeys dx(i+1) = A dx + B du + d exactly in the
    // tangent of x(i+1):  x_new(i+1) = x(i+1) (+) step  =>  x_new(i+1) (-) x(i+1) = Log(Exp(step)).
    // The reference evaluates Log(Exp(.)) numerically through manif; so do we.
    double dx[NX];
#pragma unroll
    for (int j = 0; j < NX; ++j) dx[j] = 0.0;
    double c1 = 0.0, c2 = 0.0;
    for (int i = 0; i < Nb; ++i) {
        double du[NU];
#pragma unroll
        for (int a = 0; a < NU; ++a) {
            double s = w.gains[lsoa(i, GainRec<KIND>::KFF_OFF + a, GainRec<KIND>::LEN, N, b)];
#pragma unroll
            for (int c = 0; c < NX; ++c) s += w.gains[lsoa(i, a * NX + c, GainRec<KIND>::LEN, N, b)] * dx[c];
            du[a] = s;
        }
        // first / second order terms
        double f1 = 0.0, f2 = 0.0;
#pragma unroll
        for (int c = 0; c < NX; ++c) f1 += lin[lsoa(i, LR::LX_OFF + c, F, N + 1, b)] * dx[c];
#pragma unroll
        for (int a = 0; a < NU; ++a) f1 += lin[lsoa(i, LR::LU_OFF + a, F, N + 1, b)] * du[a];
        // dx^T l_xx dx
#pragma unroll
        for (int r = 0; r < NP; ++r)
#pragma unroll
            for (int c = 0; c < NP; ++c)
                f2 += dx[r] * lin[lsoa(i, LR::LXX_OFF + sym_idx(NP, r, c), F, N + 1, b)] * dx[c];
#pragma unroll
        for (int r = 0; r < NV; ++r)
#pragma unroll
            for (int c = 0; c < NV; ++c) f2 += dx[NP + r] * 2.0 * prm.W2[r * NV + c] * dx[NP + c];
        // du^T l_uu du  (l_ux = 0)
#pragma unroll
        for (int r = 0; r < NU; ++r)
#pragma unroll
            for (int c = 0; c < NU; ++c) {
                double l = 2.0 * prm.R[r * NU + c];
                if (r == c && prm.has_constraints) l += lin[lsoa(i, LR::LUU_OFF + r, F, N + 1, b)];
                f2 += du[r] * l * du[c];
            }
        c1 += f1;
        c2 += f2;
        // propagate
        AMat<KIND> A;
        A.load(lin, i, N + 1, b);
        const BvStage<KIND> Bv(prm, lin + lsoa(i, 0, F, N + 1, b), kRecStride);
        double st[NX];
#pragma unroll
        for (int r = 0; r < NX; ++r) {
            double s = lin[lsoa(i, LR::D_OFF + r, F, N + 1, b)];
#pragma unroll
            for (int c = 0; c < NX; ++c)
                if (AMat<KIND>::nz(r, c)) s += A.get(r, c) * dx[c];
            if (r >= NP) {
#pragma unroll
                for (int a = 0; a < NU; ++a)
                    if (bv_nz<KIND>(r - NP, a)) s += Bv.get(r - NP, a) * du[a];
            }
            st[r] = s;
        }
        // pose part goes through Exp then Log (rminus of the retracted pose against the node)
        if constexpr (on_so3(KIND)) {
            double qe[4];
            so3_exp(st, qe);
            so3_log(qe, dx);
        } else {
            double qe[4], pe[3];
            se3_exp(st, qe, pe);
            se3_log(qe, pe, dx);
        }
#pragma unroll
        for (int j = NP; j < NX; ++j) dx[j] = st[j];
    }
    {
        double f1 = 0.0, f2 = 0.0;
#pragma unroll
        for (int c = 0; c < NX; ++c) f1 += lin[lsoa(Nb, LR::LX_OFF + c, F, N + 1, b)] * dx[c];
#pragma unroll
        for (int r = 0; r < NP; ++r)
#pragma unroll
            for (int c = 0; c < NP; ++c)
                f2 += dx[r] * lin[lsoa(Nb, LR::LXX_OFF + sym_idx(NP, r, c), F, N + 1, b)] * dx[c];
#pragma unroll
        for (int r = 0; r < NV; ++r)
#pragma unroll
            for (int c = 0; c < NV; ++c) f2 += dx[NP + r] * 2.0 * prm.P2[r * NV + c] * dx[NP + c];
        c1 += f1;
        c2 += f2;
    }
    const double dn = w.dnorm[b];
    double dw;
    if (dn < prm.defect_kappa) dw = dweight_prev[b];
    else dw = fmax(prm.defect_mu0, prm.defect_mu0 + fabs(c1 + 0.5 * c2) / ((1.0 - prm.defect_rho) * dn));
    dweight_prev[b] = dw;
    const size_t base = (size_t)(2 * prm.n_alphas) * Bp + b;
    w.Jcand[base] = c1;
    w.Jcand[base + Bp] = c2;
    w.Jcand[base + 2 * (size_t)Bp] = dw;
    // merit uses J_opt = L.sum() (pairwise, :2507) of the current trajectory
    w.Jcand[base + 3 * (size_t)Bp] = pairwise_sum(w.Lc + b, (size_t)Bp, Nb + 1) + dw * dn;
}

// accept the first alpha with merit_new - merit < gamma (dJ_exp(alpha) - alpha w ||d||) (:2576)
static __global__ void k_ls_select_ms(const Params prm, Work w, int it_arg, int a_lo, int a_cnt, int last) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    const bool in = (b < prm.B) && (w.status[b] == TRAJOPT_RUNNING);
    const int it = it_arg >= 0 ? it_arg : (in ? w.iters[b] : 0);
    bool pending = false;
    const int Bp = prm.Bp;
    if (in && (a_lo == 0 || w.ls_state[b] == -1)) {
        const size_t base = (size_t)(2 * prm.n_alphas) * Bp + b;
        const double c1 = w.Jcand[base], c2 = w.Jcand[base + Bp], dw = w.Jcand[base + 2 * (size_t)Bp];
        const double merit = w.Jcand[base + 3 * (size_t)Bp];
        const double dn = w.dnorm[b];
        int found = -1;
        for (int a = a_lo; a < a_lo + a_cnt; ++a) {
            const double alpha = (a == 0) ? 1.0 : pow(1.1, -(double)(a * a));
            const double Jn = w.Jcand[(size_t)a * Bp + b];
            const double dnn = w.Jcand[(size_t)(prm.n_alphas + a) * Bp + b];
            const double merit_new = Jn + dw * dnn;
            const double Jexp = alpha * c1 + 0.5 * (alpha * alpha) * c2;
            if (merit_new - merit < prm.defect_gamma * (Jexp - alpha * dw * dn)) { found = a; break; }
        }
        if (found >= 0) {
            w.ls_state[b] = found;
        } else if (!last) {
            w.ls_state[b] = -1;
            pending = true;
        } else {
            w.ls_state[b] = -2;
            w.alphahist[(size_t)it * Bp + b] = -1;
            w.iters[b] = it + 1;
            w.status[b] = TRAJOPT_NO_DESCENT;
            // the callback still records J_opt (pairwise sum of the unchanged trajectory) (:2621-2626)
            w.Jhist[(size_t)it * Bp + b] = merit - dw * dn;
        }
    }
    const unsigned m = __ballot_sync(0xffffffffu, pending);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(&w.counters[1], __popc(m));
}

static __global__ void k_ls_commit_ms(const Params prm, Work w, int it_arg) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= prm.B || w.status[b] != TRAJOPT_RUNNING) return;
    const int it = it_arg >= 0 ? it_arg : w.iters[b];
    const int a = w.ls_state[b];
    if (a < 0) return;
    w.sel[b] = 1 - w.sel[b];
    w.alphahist[(size_t)it * prm.Bp + b] = a;
    w.iters[b] = it + 1;
    w.ls_state[b] = -2;
}

// ------------------------------------------------------------------------------------------
// Augmented Lagrangian outer update (:3242-3290) for InputConstraint g = [lb - u; u - ub]
// ------------------------------------------------------------------------------------------
// Stage-parallel (thread = (problem, chunk of stages)): with one thread per problem walking its N stages twice the update
// of a 2048-problem shard at N = 1400 took ~10 ms per outer iteration, an eighth of the configuration's solve.
//   k_al_viol_zero -> k_al_viol (max g over the stages, exact whatever the order: atomicMax on the bit pattern of a
//   non-negative double) -> k_al_decide (converged? else mu <- min(mu scale, mu_max); counters[2]) -> k_al_apply
constexpr int kAlChunk = 32;     // stages per thread

static __global__ void k_al_viol_zero(const Params prm, Work w) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < prm.B && !w.al_done[b]) w.al_viol[b] = 0.0;
}

template <int KIND>
__global__ void k_al_viol(const Params prm, Work w) {
    constexpr int NU = Dims<KIND>::NU, NS = Dims<KIND>::NS, NV = Dims<KIND>::NX - Dims<KIND>::NP, XI0 = NS - NV;
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    const int Bp = prm.Bp;
    if (b >= prm.B || w.al_done[b]) return;
    const int N = w.Nb[b];          // this problem's horizon
    const int i0 = blockIdx.y * kAlChunk, i1 = min(i0 + kAlChunk, N + 1);
    if (i0 > N) return;
    const double* U = w.U[w.sel[b]];
    const double* X = w.X[w.sel[b]];
    // max over stages of g (the input rows of the terminal stage are zeros, :3245-3247)
    double gmax = 0.0;
    for (int i = i0; i < i1; ++i) {
        if (i < N) {
#pragma unroll
            for (int j = 0; j < NU; ++j) {
                const double u = U[soa(i, j, NU, Bp, b)];
                gmax = fmax(gmax, fmax(prm.lb[j] - u, u - prm.ub[j]));
            }
        }
        if (prm.has_state_bounds) {
#pragma unroll
            for (int j = 0; j < NV; ++j) {
                const double v = X[soa(i, XI0 + j, NS, Bp, b)];
                gmax = fmax(gmax, fmax(prm.xlb[j] - v, v - prm.xub[j]));
            }
        }
    }
    if (gmax > 0.0) atomicMax(reinterpret_cast<unsigned long long*>(w.al_viol + b), (unsigned long long)__double_as_longlong(gmax));
}

static __global__ void k_al_decide(const Params prm, Work w, double tol_constr, double mu_scale, double mu_max, int outer_it) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    bool notdone = false;
    if (b < prm.B && !w.al_done[b]) {
        w.al_outer[b] = outer_it + 1;
        if (w.al_viol[b] < tol_constr) {
            w.al_done[b] = 1;
        } else {
            notdone = true;
            w.al_mu[b] = fmin(w.al_mu[b] * mu_scale, mu_max);
        }
    }
    const unsigned m = __ballot_sync(0xffffffffu, notdone);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(&w.counters[2], __popc(m));
}

template <int KIND>
__global__ void k_al_apply(const Params prm, Work w) {
    constexpr int NU = Dims<KIND>::NU, NS = Dims<KIND>::NS, NV = Dims<KIND>::NX - Dims<KIND>::NP, XI0 = NS - NV;
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    const int Bp = prm.Bp;
    if (b >= prm.B || w.al_done[b]) return;     // converged problems (this round's included) keep their multipliers
    const int N = w.Nb[b];
    const int i0 = blockIdx.y * kAlChunk, i1 = min(i0 + kAlChunk, N + 1);
    if (i0 > N) return;
    const double* U = w.U[w.sel[b]];
    const double* X = w.X[w.sel[b]];
    const double mu_new = w.al_mu[b];
    for (int i = i0; i < i1; ++i) {
        if (i < N) {
#pragma unroll
            for (int j = 0; j < NU; ++j) {
                const double u = U[soa(i, j, NU, Bp, b)];
                const double g[2] = {prm.lb[j] - u, u - prm.ub[j]};
#pragma unroll
                for (int s = 0; s < 2; ++s) {
                    const size_t idx = soa(i, s * NU + j, 2 * NU, Bp, b);
                    const double ln = fmax(0.0, w.lam[idx] + w.imu[idx] * g[s]);
                    w.lam[idx] = ln;
                    w.imu[idx] = (g[s] < 0.0 && ln == 0.0) ? 0.0 : mu_new;
                }
            }
        } else {
            // terminal row: g = 0 -> lambda stays 0, Imu = mu_new (no effect on the cost)
#pragma unroll
            for (int j = 0; j < 2 * NU; ++j) w.imu[soa(N, j, 2 * NU, Bp, b)] = mu_new;
        }
        if (prm.has_state_bounds) {
#pragma unroll
            for (int j = 0; j < NV; ++j) {
                const double v = X[soa(i, XI0 + j, NS, Bp, b)];
                const double g[2] = {prm.xlb[j] - v, v - prm.xub[j]};
#pragma unroll
                for (int s = 0; s < 2; ++s) {
                    const size_t idx = soa(i, s * NV + j, 2 * NV, Bp, b);
                    const double ln = fmax(0.0, w.lam_s[idx] + w.imu_s[idx] * g[s]);
                    w.lam_s[idx] = ln;
                    w.imu_s[idx] = (g[s] < 0.0 && ln == 0.0) ? 0.0 : mu_new;
                }
            }
        }
    }
}

template <int KIND>
__global__ void k_al_init(const Params prm, Work w, double mu0) {
    constexpr int NU = Dims<KIND>::NU;
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    const int stage = blockIdx.y;
    if (b >= prm.Bp) return;
#pragma unroll
    for (int j = 0; j < 2 * NU; ++j) {
        w.lam[soa(stage, j, 2 * NU, prm.Bp, b)] = 0.0;
        w.imu[soa(stage, j, 2 * NU, prm.Bp, b)] = mu0;
    }
    if (prm.has_state_bounds) {
        constexpr int NV = Dims<KIND>::NX - Dims<KIND>::NP;
#pragma unroll
        for (int j = 0; j < 2 * NV; ++j) {
            w.lam_s[soa(stage, j, 2 * NV, prm.Bp, b)] = 0.0;
            w.imu_s[soa(stage, j, 2 * NV, prm.Bp, b)] = mu0;
        }
    }
    if (stage == 0) {
        w.al_mu[b] = mu0;
        w.al_outer[b] = 0;
        w.al_viol[b] = 0.0;
        w.al_done[b] = (b < prm.B) ? 0 : 1;
    }
}

// ------------------------------------------------------------------------------------------
// Layout exports: SoA [stage][field][Bp] -> problem-major
// ------------------------------------------------------------------------------------------
// out[b][stage][f] = src_sel[b][stage][f][b]   (grid.y = stage)
// One warp per problem and chunk of its row: element e = stage * F + f of problem (slot) b sits at src[e * Bp + b], so
// the warp's writes are 32 consecutive doubles and its reads 32 sectors that the block's four warps (four consecutive
// slots = one 32-byte sector) share through L1.  (One thread per (slot, stage) writing F doubles took 4-9 ms for the
// 2.4 GB of a headline batch.)  grid = (ceil(Bp / 4), chunks of kExportChunk elements), block = 128.
constexpr int kExportChunk = 2048;
// Nb (NULL: every problem uses all nstage rows): rows of stages > Nb[b] - nb_shift were never written by a rollout; they
// are taken from buffer 0, which holds the initial guess there (nb_shift = 0 for states, 1 for controls).
static __global__ void k_export_traj(int B, int Bp, int F, const double* s0, const double* s1, const int* sel,
                                     const int* __restrict__ orig, double* out, int nstage, const int* __restrict__ Nb, int nb_shift) {
    const int b = blockIdx.x * 4 + (threadIdx.x >> 5);      // slot
    if (b >= Bp) return;
    const int o = orig[b];                                  // the caller's problem index
    if (o >= B) return;
    const double* src = ((sel && sel[b]) ? s1 : s0) + b;
    const int len = nstage * F, lane = threadIdx.x & 31;
    const int e0 = blockIdx.y * kExportChunk, e1 = min(e0 + kExportChunk, len);
    const int e_own = Nb ? min(len, (Nb[b] - nb_shift + 1) * F) : len;      // elements [0, e_own) belong to the solved horizon
    double* dst = out + (size_t)o * len;
    for (int e = e0 + lane; e < e1; e += 32) dst[e] = (e < e_own ? src : s0 + b)[(size_t)e * Bp];
}
// the same for the problems that were still running when `snap` (status by caller index) was taken
static __global__ void k_export_traj_late(int B, int Bp, int F, const double* s0, const double* s1, const int* sel,
                                          const int* __restrict__ orig, const int* __restrict__ snap, double* out, int nstage,
                                          const int* __restrict__ Nb, int nb_shift) {
    const int b = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (b >= Bp) return;
    const int o = orig[b];
    if (o >= B || (snap[o] & 15) != TRAJOPT_RUNNING) return;
    const double* src = ((sel && sel[b]) ? s1 : s0) + b;
    const int len = nstage * F, lane = threadIdx.x & 31;
    const int e0 = blockIdx.y * kExportChunk, e1 = min(e0 + kExportChunk, len);
    const int e_own = Nb ? min(len, (Nb[b] - nb_shift + 1) * F) : len;
    double* dst = out + (size_t)o * len;
    for (int e = e0 + lane; e < e1; e += 32) dst[e] = (e < e_own ? src : s0 + b)[(size_t)e * Bp];
}
// gains (group-major, GainRec): fields [f0, f0 + nf) of every stage -> out[orig[b]][stage][f - f0]   (grid.y = stage)
static __global__ void k_export_gains(int B, int Bp, int N, int LEN, int f0, int nf, const double* __restrict__ gains,
                                      const int* __restrict__ orig, double* __restrict__ out) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    const int stage = blockIdx.y;
    if (b >= Bp) return;
    const int o = orig[b];
    if (o >= B) return;
    for (int f = 0; f < nf; ++f) out[((size_t)o * N + stage) * nf + f] = gains[lsoa(stage, f0 + f, LEN, N, b)];
}
// out[orig[b]][row] = src[row][b]
template <typename T>
__global__ void k_export_rows(int B, int Bp, int rows, const T* src, const int* __restrict__ orig, T* out) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y;
    if (b >= Bp) return;
    const int o = orig[b];
    if (o >= B) return;
    out[(size_t)o * rows + r] = src[(size_t)r * Bp + b];
}

// rows `rows[0..n)` of a [B][row_doubles] device array -> the same rows of a HOST array (pinned, written over PCIe by
// the kernel itself: thousands of small copy-engine transfers cost ~18 us each, these rows go at link speed)
static __global__ void k_rows_to_host(const int* __restrict__ rows, size_t row_doubles, const double* __restrict__ src,
                                      double* __restrict__ dst_host) {
    const size_t base = (size_t)rows[blockIdx.x] * row_doubles;
    for (size_t j = threadIdx.x; j < row_doubles; j += blockDim.x) dst_host[base + j] = src[base + j];
}

// Compaction (see maybe_compact in host_impl.cuh): gather the leading `front` slots of a [rows][Bp] array into their
// new order, in two steps through a scratch buffer:  scratch[r][n] = data[r][src_of[n]];  data[r][n] = scratch[r][n].
template <typename T>
__global__ void k_permute_gather(int rows, int Bp, int front, const T* __restrict__ data, const int* __restrict__ src_of,
                                 T* __restrict__ scratch) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= front) return;
    const int src = src_of[n];
    for (int r = blockIdx.y; r < rows; r += gridDim.y) scratch[(size_t)r * front + n] = data[(size_t)r * Bp + src];
}
template <typename T>
__global__ void k_permute_scatter(int rows, int Bp, int front, T* __restrict__ data, const T* __restrict__ scratch) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= front) return;
    for (int r = blockIdx.y; r < rows; r += gridDim.y) data[(size_t)r * Bp + n] = scratch[(size_t)r * front + n];
}
// the inverse of every compaction so far: scratch[r][orig[n]] = data[r][n]  (orig is a permutation of [0, Bp))
template <typename T>
__global__ void k_unpermute_scatter(int rows, int Bp, const T* __restrict__ data, const int* __restrict__ orig, T* __restrict__ scratch) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= Bp) return;
    const int o = orig[n];
    for (int r = blockIdx.y; r < rows; r += gridDim.y) scratch[(size_t)r * Bp + o] = data[(size_t)r * Bp + n];
}
// per-problem horizons: out[b] = clamp(in[b], 1, N) for b < B (in == NULL: N), N for the padding slots
static __global__ void k_set_horizons(int B, int Bp, int N, const int* __restrict__ in, int* __restrict__ out) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= Bp) return;
    int v = N;
    if (in && b < B) v = min(max(in[b], 1), N);
    out[b] = v;
}
static __global__ void k_fill_double(int n, double v, double* out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = v;
}
static __global__ void k_identity(int n, int* out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = i;
}

// dense linearisation for parity tests
template <int KIND>
__global__ void k_export_lin(const Params prm, Work w, double* Fx, double* Fu, double* dd, double* L, double* Lx, double* Lxx, double* Lu) {
    using D = Dims<KIND>;
    using LR = LinRec<KIND>;
    constexpr int NX = D::NX, NP = D::NP, NU = D::NU, NV = NX - NP, F = LR::LEN;
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y;
    const int Bp = prm.Bp, N = prm.N;
    if (b >= Bp) return;
    const int o = w.orig[b];           // output row = the caller's problem index
    if (o >= prm.B) return;
    const double* lin = w.lin;
    if (L) L[(size_t)o * (N + 1) + i] = w.Lc[(size_t)i * Bp + b];
    if (Lx)
        for (int c = 0; c < NX; ++c) Lx[((size_t)o * (N + 1) + i) * NX + c] = lin[lsoa(i, LR::LX_OFF + c, F, N + 1, b)];
    if (Lxx) {
        const double* W2 = (i == N) ? prm.P2 : prm.W2;
        for (int r = 0; r < NX; ++r)
            for (int c = 0; c < NX; ++c) {
                double v = 0.0;
                if (r < NP && c < NP) v = lin[lsoa(i, LR::LXX_OFF + sym_idx(NP, r, c), F, N + 1, b)];
                else if (r >= NP && c >= NP) v = 2.0 * W2[(r - NP) * NV + (c - NP)];
                Lxx[(((size_t)o * (N + 1) + i) * NX + r) * NX + c] = v;
            }
    }
    if (i == N) return;
    AMat<KIND> A;
    A.load(lin, i, N + 1, b);
    if (Fx) {
#pragma unroll
        for (int r = 0; r < NX; ++r)
#pragma unroll
            for (int c = 0; c < NX; ++c)
                Fx[(((size_t)o * N + i) * NX + r) * NX + c] = AMat<KIND>::nz(r, c) ? A.get(r, c) : 0.0;
    }
    if (Fu) {
        const BvStage<KIND> Bv(prm, lin + lsoa(i, 0, F, N + 1, b), kRecStride);
        for (int r = 0; r < NX; ++r)
            for (int a = 0; a < NU; ++a)
                Fu[(((size_t)o * N + i) * NX + r) * NU + a] = (r >= NP && bv_nz<KIND>(r - NP, a)) ? Bv.get(r - NP, a) : 0.0;
    }
    if (dd)
        for (int c = 0; c < NX; ++c) dd[((size_t)o * N + i) * NX + c] = lin[lsoa(i, LR::D_OFF + c, F, N + 1, b)];
    if (Lu)
        for (int a = 0; a < NU; ++a) Lu[((size_t)o * N + i) * NU + a] = lin[lsoa(i, LR::LU_OFF + a, F, N + 1, b)];
}

}  // namespace trajopt
