// Continuous batching of DDP solves (trajopt_solve_stream): M problems through the handle's B slots.
//
// A batch solved as one unit runs as long as its slowest problem: on the headline workload 90 % of the problems
// stop after 20 iterations and the rest need up to 27, so a quarter of the launches work on a tenth of the slots
// (the sweeps are latency bound: a launch costs the same for 1600 running problems as for 16384).  Here a slot whose
// problem has finished is handed the next problem of the queue before the following iteration, so every launch
// works on (nearly) B running problems until the queue is empty.  A problem's arithmetic does not depend on its
// slot or on its neighbours (tests/test_gpu_fullsize.py, test_gpu_compaction.py), so each result is bit-identical
// to the one trajopt_solve gives for the same x0; only the order of completion changes.
//
// Per DDP iteration, on the caller's stream:
//   k_stream_collect   one CTA, deterministic: list of slots that are not running (free) and of those among them
//                      that still hold an un-exported problem (done); counters
//   k_stream_export_*  rows of the finished problems -> the caller's arrays at [problem id]
//   k_stream_admit     free slots <- next problems of the queue: x0, regulariser / counters reset, `fresh` flag
//   k_init_ms/_ss      initial guess of the fresh slots (same kernels as trajopt_begin, masked)
//   k_linearize        of the fresh slots when the records of the others are already current (overlapped rollout)
//   (host variant)     rows of the completed prefix of problem ids are copied to the host on a side stream meanwhile
//   inner_iteration    unchanged kernels; the iteration index of a slot is its own `iters` (it < 0 convention)
// The host reads the three counters once per iteration (the same single synchronisation trajopt_iterate has).
#pragma once
// (included by host_impl.cuh after inner_iteration is defined)

namespace trajopt {

constexpr int kCollectThreads = 1024;

static __global__ void k_stream_clear(const Params prm, Work w) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= prm.Bp) return;
    w.status[b] = TRAJOPT_MAX_ITER;   // anything but RUNNING: the slot is free
    w.slot_id[b] = -1;
    w.fresh[b] = 0;
    w.iters[b] = 0;
    w.sel[b] = 0;
}

static __global__ void __launch_bounds__(kCollectThreads) k_stream_collect(const Params prm, Work w) {
    __shared__ int s_free[kCollectThreads], s_done[kCollectThreads];
    __shared__ int s_run, s_min;
    const int t = threadIdx.x;
    const int per = (prm.B + kCollectThreads - 1) / kCollectThreads;
    const int lo = t * per, hi = min(lo + per, prm.B);
    if (t == 0) { s_run = 0; s_min = 0x7fffffff; }
    int nf = 0, nd = 0, nr = 0, lowest = 0x7fffffff;
    for (int b = lo; b < hi; ++b) {
        if (w.status[b] == TRAJOPT_RUNNING) { ++nr; lowest = min(lowest, w.slot_id[b]); continue; }
        ++nf;
        if (w.slot_id[b] >= 0) ++nd;
    }
    s_free[t] = nf;
    s_done[t] = nd;
    __syncthreads();
    if (nr) { atomicAdd(&s_run, nr); atomicMin(&s_min, lowest); }
    for (int off = 1; off < kCollectThreads; off <<= 1) {   // inclusive scans
        const int a = (t >= off) ? s_free[t - off] : 0, c = (t >= off) ? s_done[t - off] : 0;
        __syncthreads();
        s_free[t] += a;
        s_done[t] += c;
        __syncthreads();
    }
    int pf = s_free[t] - nf, pd = s_done[t] - nd;
    for (int b = lo; b < hi; ++b) {
        if (w.status[b] == TRAJOPT_RUNNING) continue;
        w.free_list[pf++] = b;
        if (w.slot_id[b] >= 0) w.done_list[pd++] = b;
    }
    if (t == kCollectThreads - 1) {
        w.scnt[0] = s_free[t];
        w.scnt[1] = s_done[t];
    }
    __syncthreads();
    if (t == 0) { w.scnt[2] = s_run; w.scnt[3] = s_min; }   // [3]: lowest problem id still running
}

// rows [stage][field] of the finished problems -> out[id][stage][field]
static __global__ void k_stream_export_traj(const Params prm, Work w, int F, const double* s0, const double* s1, double* out,
                                            int nstage) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= w.scnt[1]) return;
    const int slot = w.done_list[idx];
    const long long id = w.slot_id[slot];
    const int stage = blockIdx.y;
    const double* src = w.sel[slot] ? s1 : s0;
    for (int f = 0; f < F; ++f) out[((size_t)id * nstage + stage) * F + f] = src[((size_t)stage * F + f) * prm.Bp + slot];
}

static __global__ void k_stream_export_summary(const Params prm, Work w, double* J, int* iters, int* status, double* grad,
                                               double* defect) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= w.scnt[1]) return;
    const int slot = w.done_list[idx];
    const int id = w.slot_id[slot];
    if (J) J[id] = w.J[slot];
    if (iters) iters[id] = w.iters[slot];
    if (status) status[id] = w.status[slot];
    if (grad) grad[id] = w.grad[slot];
    if (defect) defect[id] = w.dnorm[slot];
}

// free slots <- problems [q_head, q_head + n) of the queue, n = min(free slots, problems left)
template <int KIND>
__global__ void k_stream_admit(const Params prm, Work w, const double* __restrict__ x0_aos, int q_head, int q_total,
                               double* dweight) {
    constexpr int NS = Dims<KIND>::NS;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= w.scnt[0]) return;
    const int slot = w.free_list[idx];
    if (idx >= q_total - q_head) {     // nothing left to admit: the slot stays empty
        w.slot_id[slot] = -1;
        return;
    }
    const int id = q_head + idx;
    double v[NS];
#pragma unroll
    for (int j = 0; j < NS; ++j) v[j] = x0_aos[(size_t)id * NS + j];
    quat_normalize(v);
#pragma unroll
    for (int j = 0; j < NS; ++j) w.x0[(size_t)j * prm.Bp + slot] = v[j];
    // what k_reset does for a whole batch
    w.sel[slot] = 0;
    w.mu[slot] = 1.0;
    w.delta[slot] = prm.delta0;
    w.iters[slot] = 0;
    w.status[slot] = (prm.max_iters > 0 || prm.method != TRAJOPT_SS) ? TRAJOPT_RUNNING : TRAJOPT_MAX_ITER;
    w.ls_state[slot] = -2;
    w.J[slot] = 0.0;
    w.grad[slot] = 0.0;
    w.dnorm[slot] = 0.0;
    if (dweight) dweight[slot] = prm.defect_mu0;
    w.slot_id[slot] = id;
    w.fresh[slot] = 1;
}

}  // namespace trajopt

namespace trajopt_host {

template <int KIND>
int solve_stream_impl(trajopt_handle* h, const double* d_x0, int M, double* d_xs, double* d_us, double* d_J, int* d_iters,
                      int* d_status, double* d_grad, double* d_defect, cudaStream_t s) {
    Work& w = h->w;
    int rc0_ = 0;
    const int bg = blocks_for(h->Bp, 128);
    if ((rc0_ = restore_caller_order(h, s))) return rc0_;
    LAUNCH(k_identity, bg, 128, 0, s, h->Bp, w.orig);
    LAUNCH(k_stream_clear, bg, 128, 0, s, h->prm, w);
    h->it = 0;
    h->inner_done = false;
    h->lin_ready = false;
    h->front = h->Bp;
    h->streaming = true;
    struct Off { trajopt_handle* h; ~Off() { h->streaming = false; h->lin_ready = false; h->inner_done = true; } } off{h};
    int q_head = 0, rc;
    long long guard = (long long)(M / std::max(h->B, 1) + 2) * (h->prm.max_iters + 2) + 8;   // iterations this can take at most
    while (guard-- > 0) {
        {
            PhaseTimer t(h, s, PH_OTHER);
            LAUNCH(k_stream_collect, 1, kCollectThreads, 0, s, h->prm, w);
            LAUNCH(k_publish4, 1, 32, 0, s, (const int*)w.scnt, (volatile int*)h->h_counters_dev);
            if (d_xs) LAUNCH(k_stream_export_traj, dim3(bg, h->N + 1), 128, 0, s, h->prm, w, h->NS, w.X[0], w.X[1], d_xs, h->N + 1);
            if (d_us) LAUNCH(k_stream_export_traj, dim3(bg, h->N), 128, 0, s, h->prm, w, h->NU, w.U[0], w.U[1], d_us, h->N);
            LAUNCH(k_stream_export_summary, bg, 128, 0, s, h->prm, w, d_J, d_iters, d_status, d_grad, d_defect);
            if (q_head < M) {
                LAUNCH((k_stream_admit<KIND>), bg, 128, 0, s, h->prm, w, d_x0, q_head, M, h->prm.line_search ? h->d_dweight : nullptr);
                if (h->method == TRAJOPT_SS) {
                    LAUNCH((k_init_ss<KIND>), h->Bp / kBlock, kBlock, 0, s, h->prm, w, (const int*)w.fresh);
                } else {
                    LAUNCH((k_init_ms<KIND>), dim3(bg, h->N + 1), 128, 0, s, h->prm, w, false, (const int*)w.fresh);
                    if (h->lin_ready) {   // the other slots' records are current: linearise the newcomers only
                        dim3 grid(bg, h->N + 1);
                        LAUNCH((k_linearize<KIND, true, false>), grid, 128, 0, s, h->prm, w, 0, 0, (const int*)w.fresh);
                    }
                }
                CUDA_OK(cudaMemsetAsync(w.fresh, 0, (size_t)h->Bp * sizeof(int), s));
            } else {
                LAUNCH((k_stream_admit<KIND>), bg, 128, 0, s, h->prm, w, d_x0, M, M, (double*)nullptr);   // mark the exported slots empty
            }
        }
        // the iteration is queued behind; the host learns the counters while it runs
        if ((rc = inner_iteration<KIND>(h, s))) return rc;
        CUDA_OK(cudaStreamSynchronize(s));
        const int n_free = h->h_counters[0], n_run = h->h_counters[2];
        const int admitted = std::min(n_free, M - q_head);
        if (h->stream_progress) {   // every problem below this id has been exported (trajopt_solve_stream_host copies them out)
            if ((rc = h->stream_progress(std::min(h->h_counters[3], q_head)))) return rc;
        }
        q_head += admitted;
        if (q_head >= M && n_run + admitted == 0) return 0;   // queue empty, nothing was running: everything is exported
    }
    return fail(TRAJOPT_E_STATE, "trajopt_solve_stream: iteration guard exceeded");
}

}  // namespace trajopt_host
