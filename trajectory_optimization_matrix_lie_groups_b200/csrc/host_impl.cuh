// Host-side orchestration shared by api.cu (the exported C ABI) and the per-family translation units
// kind_*.cu (explicit instantiations of the kernel sequences for ONE problem family each, so that the
// families compile in parallel).  Everything here is templated on the family or inline.
#pragma once
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>
#include <functional>

#include "kernels_fwd.cuh"
#include "debug.cuh"

extern std::atomic<long long> g_trajopt_launches;
int trajopt_fail_(int code, const char* msg);

#define LAUNCH(kern, grid, block, smem, stream, ...)                 \
    do {                                                             \
        kern<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);    \
        g_trajopt_launches.fetch_add(1, std::memory_order_relaxed);  \
        CUDA_OK(cudaPeekAtLastError());                              \
    } while (0)

enum { PH_LIN = 0, PH_BWD = 1, PH_FWD = 2, PH_OTHER = 3, PH_COUNT = 4 };

using namespace trajopt;

constexpr int kMaxChunks = 32;
constexpr int kHostTickets = 4;     // host-buffer solves of one handle whose copies may be outstanding at once

struct trajopt_handle {
    int kind = 0, method = 0, N = 0, B = 0, Bp = 0, device = 0;
    int NX = 0, NP = 0, NU = 0, NS = 0, LEN = 0, REFROW = 0;
    Params prm{};
    Work w{};
    trajopt_params user{};
    bool have_params = false, have_ref = false, begun = false;
    bool var_horizons = false;   // trajopt_set_horizons was given a non-NULL array
    // overlapped rollout / linearisation (run_forward_overlapped): second stream, one event per chunk, join event;
    // lin_ready: the records already belong to the current trajectory, the next iteration skips its linearisation
    cudaStream_t s2 = nullptr;
    cudaEvent_t ev_chunk[kMaxChunks] = {};
    cudaEvent_t ev_join = nullptr;
    cudaEvent_t ev_block = nullptr;   // wait_stream: blocking (sleeping) wait when many lanes share the host's cores
    bool lin_ready = false;
    bool streaming = false;      // inside trajopt_solve_stream
    std::function<int(int)> stream_progress;   // called once per iteration with the completed prefix of problem ids
    cudaStream_t s_copy = nullptr;              // trajopt_solve_host / _stream_host: device -> host copies beside the solve
    cudaEvent_t ev_host[2] = {nullptr, nullptr};
    cudaEvent_t ev_done[kHostTickets] = {};     // trajopt_solve_host_begin / _wait: all copies of a ticket have landed
    int next_ticket = 0;
    long long tickets_issued = 0;
    int32_t* s_snap = nullptr;                  // status snapshot of the early export (by caller index)
    int* d_late = nullptr;                      // rows copied again at the end
    void* so_buf[7] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};   // its device staging (xs, us, J, iters, status, grad, defect)
    size_t so_bytes[7] = {0, 0, 0, 0, 0, 0, 0};
    double* so_x0 = nullptr;
    size_t so_x0_bytes = 0;
    int it = 0;              // next inner iteration
    bool inner_done = true;  // the inner loop of the current fit() / AL outer iteration has ended
    int al_outer = 0;        // AL outer iterations completed
    bool al_finished = false;
    bool al_inner_open = false;   // AL: an inner solve has been started and its multiplier update has not been applied yet
    int hist_cap = -1, cand_rows = -1;
    // small batches with a line search: one trajectory buffer per step size (Work::Xc / Uc), see ensure_cand
    void* cand_allocs[2] = {nullptr, nullptr};
    int cand_alphas = 0, cand_max_batch = 256;
    // compaction: leading slots that may hold running problems; thresholds (see maybe_compact)
    int front = 0, compact_min_batch = 1024, compact_ratio = 4;
    // Nb / ref_batch are settings that outlive a solve ("set before trajopt_begin"); a compaction moves them with their
    // problems, so the next begin puts them back into the caller's order first (restore_caller_order)
    bool nb_permuted = false, ref_permuted = false;
    // backward sweep of the 12-dimensional families: 0 = auto (six-warp CTAs while the slots in use fit one wave of them,
    // four-warp CTAs up to one wave of those, two-warp CTAs above), 2 / 4 / 6 = always that one; lanes = solver handles
    // sharing this GPU (trajopt_set_sweep)
    int sweep_variant = 0, sweep_lanes = 1, sms = 148;
    int* d_perm = nullptr;
    void* d_scratch = nullptr;
    size_t scratch_bytes = 0;
    std::vector<void*> allocs;
    std::vector<const void*> smem_ready;   // kernels whose dynamic shared-memory limit has been raised on this handle's device
    std::vector<int> compact_src;          // maybe_compact: the new order, kept to avoid an allocation per compaction
    void* hist_allocs[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    double* d_ref = nullptr;
    double* d_ref_long = nullptr;    // trajopt_set_reference_long: a reference longer than the horizon; w.ref points at a window of it
    int64_t ref_long_rows = 0;
    double* d_ref_batch = nullptr;   // per-problem references, allocated on first use
    double* d_dweight = nullptr;
    int* h_counters = nullptr;   // pinned, written by k_publish4 through its device alias
    int* h_counters_dev = nullptr;
    int* h_ints = nullptr;       // [Bp] pinned + mapped: status snapshots (compaction, early host copies), same reason
    int* h_ints_dev = nullptr;
    // host staging for trajopt_solve_host
    double *s_x0 = nullptr, *s_us0 = nullptr, *s_xs = nullptr, *s_us = nullptr, *s_J = nullptr, *s_grad = nullptr, *s_def = nullptr;
    int32_t *s_iters = nullptr, *s_status = nullptr;
    size_t s_us0_bytes = 0;
    // phase profiling
    bool profiling = false;
    cudaEvent_t ev[2] = {nullptr, nullptr};
    double phase_ms[PH_COUNT] = {0, 0, 0, 0};
    long long phase_cnt[PH_COUNT] = {0, 0, 0, 0};
};

namespace trajopt_host {

inline int fail(int code, const char* msg) { return trajopt_fail_(code, msg); }


struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) ok = false;
        if (ok && prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

template <typename T>
int dalloc(trajopt_handle* h, T** p, size_t count, bool zero = true) {
    void* q = nullptr;
    CUDA_OK(cudaMalloc(&q, count * sizeof(T)));
    if (zero) CUDA_OK(cudaMemset(q, 0, count * sizeof(T)));
    h->allocs.push_back(q);
    *p = (T*)q;
    return 0;
}

struct PhaseTimer {
    trajopt_handle* h;
    cudaStream_t s;
    int ph;
    PhaseTimer(trajopt_handle* h_, cudaStream_t s_, int ph_) : h(h_), s(s_), ph(ph_) {
        if (h->profiling) cudaEventRecord(h->ev[0], s);
    }
    ~PhaseTimer() {
        if (h->profiling) {
            cudaEventRecord(h->ev[1], s);
            cudaEventSynchronize(h->ev[1]);
            float ms = 0.f;
            cudaEventElapsedTime(&ms, h->ev[0], h->ev[1]);
            h->phase_ms[ph] += ms;
            h->phase_cnt[ph] += 1;
        }
    }
};

inline int blocks_for(int n, int bs) { return (n + bs - 1) / bs; }

// cudaFuncAttributeMaxDynamicSharedMemorySize once per kernel and handle (it is a driver call; at B <= 1024 the path is
// launch bound and it was made before every launch)
template <typename K>
int ensure_smem(trajopt_handle* h, K kern, size_t bytes) {
    const void* key = (const void*)kern;
    for (const void* p : h->smem_ready)
        if (p == key) return 0;
    CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    h->smem_ready.push_back(key);
    return 0;
}

// host copies of the small algebra needed to pack parameters
inline void inv3(const double* A, double* Ai) {
    const double a = A[0], b = A[1], c = A[2], d = A[3], e = A[4], f = A[5], g = A[6], hh = A[7], i = A[8];
    const double det = a * (e * i - f * hh) - b * (d * i - f * g) + c * (d * hh - e * g);
    const double id = 1.0 / det;
    Ai[0] = (e * i - f * hh) * id; Ai[1] = (c * hh - b * i) * id; Ai[2] = (b * f - c * e) * id;
    Ai[3] = (f * g - d * i) * id;  Ai[4] = (a * i - c * g) * id;  Ai[5] = (c * d - a * f) * id;
    Ai[6] = (d * hh - e * g) * id; Ai[7] = (b * g - a * hh) * id; Ai[8] = (a * e - b * d) * id;
}
inline void host_quat_to_rot(const double* q, double* R) {
    const double x = q[0], y = q[1], z = q[2], w = q[3];
    const double tx = 2.0 * x, ty = 2.0 * y, tz = 2.0 * z;
    const double twx = tx * w, twy = ty * w, twz = tz * w;
    const double txx = tx * x, txy = ty * x, txz = tz * x;
    const double tyy = ty * y, tyz = tz * y, tzz = tz * z;
    R[0] = 1.0 - (tyy + tzz); R[1] = txy - twz;         R[2] = txz + twy;
    R[3] = txy + twz;         R[4] = 1.0 - (txx + tzz); R[5] = tyz - twx;
    R[6] = txz - twy;         R[7] = tyz + twx;         R[8] = 1.0 - (txx + tyy);
}

// ------------------------------------------------------------------------------------------
// kernel sequences, templated on the problem family
// ------------------------------------------------------------------------------------------
template <int KIND, bool MS>
int run_linearize(trajopt_handle* h, cudaStream_t s) {
    PhaseTimer t(h, s, PH_LIN);
    dim3 grid(blocks_for(h->Bp, 128), h->N + 1);
    if (h->w.ref_batch) LAUNCH((k_linearize<KIND, MS, true>), grid, 128, 0, s, h->prm, h->w, 0, 0, (const int*)nullptr);
    else LAUNCH((k_linearize<KIND, MS, false>), grid, 128, 0, s, h->prm, h->w, 0, 0, (const int*)nullptr);
    return 0;
}

// Small batches are launch bound (1024 SO3 problems: 17 launches per iteration instead of 4 cost 3.8x in throughput with
// eight batches in flight); the overlap is for batches that fill the device.
constexpr int kOverlapMinBatch = 4096;
// TRAJOPT_OVERLAP_MIN_BATCH overrides it (tests run the overlapped path at small sizes through it)
inline int overlap_min_batch() {
    static const int v = [] { const char* e = getenv("TRAJOPT_OVERLAP_MIN_BATCH"); return e ? atoi(e) : kOverlapMinBatch; }();
    return v;
}

// TRAJOPT_OVERLAP: chunks of the horizon for the overlapped rollout / linearisation (default 8; 0 or 1 = off)
inline int overlap_chunks() {
    static const int v = [] { const char* e = getenv("TRAJOPT_OVERLAP"); return e ? atoi(e) : 8; }();
    return v < kMaxChunks ? v : kMaxChunks;
}

// Where the overlapped rollout is used: plain multiple shooting with one shared reference and one horizon (what the
// full-size parity tests and the bench's serial-vs-pipelined assertion cover), batches that fill the device.
inline bool overlap_applies(const trajopt_handle* h) {
    return !h->profiling && h->s2 && h->method == TRAJOPT_MS && !h->var_horizons && !h->w.ref_batch && overlap_chunks() > 1 &&
           h->N >= 8 * overlap_chunks() && h->Bp >= overlap_min_batch();
}

// The multiple-shooting full-step rollout and the linearisation of the trajectory it writes, overlapped.
// The rollout is one sequential chain per problem (16384 problems are 512 warps, under one warp per scheduler); the
// linearisation is stage-parallel.  The horizon is cut into chunks: rollout chunk c on a second stream, then the linearisation of the
// same stages on the caller's stream, which runs beside rollout chunk c + 1.  Stage i of the linearisation reads x_new(i),
// x_new(i+1), u_new(i) (all written by rollout chunks <= c) and overwrites record i and G_i, which the rollout has
// finished reading.  The arithmetic of both kernels is untouched.  Afterwards the records are those of the NEW
// trajectory, and the next iteration starts at its Riccati sweep (h->lin_ready).
// Measured (B200, 16384 x 955, per iteration): rollout alone 4.2 ms, linearisation alone 3.3 ms, overlapped 7.5 -> the
// section barely shrinks: side by side the rollout chunks take 0.70 ms instead of 0.52 and the linearisation chunks
// 1.1 ms instead of 0.5 — both draw on the same FP64 pipe, the single rollout warp per scheduler keeps it busier than its
// occupancy suggests.  What is gained (520 -> 498 ms per solve, +4% throughput) is tails and launch gaps.
template <int KIND>
int run_forward_overlapped(trajopt_handle* h, cudaStream_t s, int chunks) {
    constexpr size_t fsmem = FwdSmem<KIND>::BYTES;
    { int rc_ = ensure_smem(h, k_forward_ms_full<KIND>, fsmem); if (rc_) return rc_; }
    const int N = h->N;
    // The rollout goes to the handle's own HIGH-PRIORITY stream: its 512 one-warp CTAs are placed first, and the
    // linearisation (caller's stream) fills the registers they leave.  64-thread blocks for the latter: at 254
    // registers a 128-thread block is half an SM, and two of them leave no room for a rollout CTA.
    static const int lin_block = [] { const char* e = getenv("TRAJOPT_LIN_BLOCK"); const int v = e ? atoi(e) : 64; return (v == 32 || v == 64 || v == 96 || v == 128) ? v : 64; }();
    const int bx = blocks_for(h->Bp, lin_block);
    CUDA_OK(cudaEventRecord(h->ev_join, s));
    CUDA_OK(cudaStreamWaitEvent(h->s2, h->ev_join, 0));
    for (int c = 0; c < chunks; ++c) {
        const int i0 = (int)((long long)N * c / chunks), i1 = (int)((long long)N * (c + 1) / chunks);
        LAUNCH((k_forward_ms_full<KIND>), h->Bp / kBlock, kBlock, fsmem, h->s2, h->prm, h->w, i0, i1);
        CUDA_OK(cudaEventRecord(h->ev_chunk[c], h->s2));
        CUDA_OK(cudaStreamWaitEvent(s, h->ev_chunk[c], 0));
        const int s1 = (c == chunks - 1) ? N + 1 : i1;      // the last chunk takes the terminal stage along
        dim3 grid(bx, s1 - i0);
        if (h->w.ref_batch) LAUNCH((k_linearize<KIND, true, true>), grid, lin_block, 0, s, h->prm, h->w, i0, 1, (const int*)nullptr);
        else LAUNCH((k_linearize<KIND, true, false>), grid, lin_block, 0, s, h->prm, h->w, i0, 1, (const int*)nullptr);
    }
    h->lin_ready = true;
    return 0;
}

// TRAJOPT_BACKWARD=1 selects the one-warp-per-group sweep for the 12-dimensional families too (A/B measurements)
// TRAJOPT_BACKWARD: 0 (default) two-warp CTA per 32 problems; 1 one warp per 32 problems (A/B measurements).
// (16 problems per warp, i.e. twice the warps, was measured too: 16.2 ms per sweep against 12.2 — more warps per
// scheduler streaming this straight-line code only fight over instruction fetch.)
inline int backward_variant() {
    static const int v = [] { const char* e = getenv("TRAJOPT_BACKWARD"); return e ? atoi(e) : 0; }();
    return v;
}
inline bool use_one_warp_sweep() { return backward_variant() != 0; }
// TRAJOPT_B3_GROUPS: groups of 32 problems per CTA of the two-warp sweep: 2 (default) or 1.  Two groups behind the same
// barriers fetch each line of the stage body once: 10.78 -> 10.35 ms per sweep (four groups: 10.37, not kept).
inline int backward_groups() {
    static const int v = [] { const char* e = getenv("TRAJOPT_B3_GROUPS"); return e ? atoi(e) : 2; }();
    return v;
}

// Four-warp CTAs (backward4.cuh) are resident two per SM: while the slots that may hold running problems fit ONE wave of
// them — counting the other solver handles that share the GPU — a launch is bound by the latency of a stage, which four
// warps cut by ~1.6x; beyond that the two-warp sweep keeps every group of a 16 k batch resident.  Bit-identical results.
// TRAJOPT_SWEEP=2|4 forces one variant (A/B measurements).
// Six-warp CTAs (backward6.cuh: the serial part of a stage on its own two warps beside the column work) are resident one
// per SM: they are the sweep while the slots in use fit ONE such wave (148 x 32 problems: a single solve, a strong-scaling
// shard, the last tail iterations).  TRAJOPT_SWEEP=2|4|6 forces one variant (A/B measurements).
inline int sweep_shape(const trajopt_handle* h) {
    static const int forced = [] { const char* e = getenv("TRAJOPT_SWEEP"); return e ? atoi(e) : 0; }();
    const int variant = forced ? forced : h->sweep_variant;
    if (variant == 2 || variant == 4 || variant == 6) return variant;
    const long long slots = (long long)h->front * h->sweep_lanes;
    if (slots <= (long long)h->sms * 32) return 6;
    if (slots <= (long long)h->sms * 2 * 32) return 4;
    return 2;
}

template <int KIND, bool MS>
int run_backward(trajopt_handle* h, cudaStream_t s, int it) {
    PhaseTimer t(h, s, PH_BWD);
    if constexpr (!on_so3(KIND)) {
        const int shape = use_one_warp_sweep() ? 1 : sweep_shape(h);
        if (shape == 6) {
            constexpr size_t smem6 = B6Smem<KIND>::BYTES;
            auto launch6 = [&](auto kern) -> int {
                { int rc_ = ensure_smem(h, kern, smem6); if (rc_) return rc_; }
                LAUNCH(kern, h->Bp / 32, kB6Threads, smem6, s, h->prm, h->w, it);
                return 0;
            };
            return h->var_horizons ? launch6(k_backward6<KIND, MS, true>) : launch6(k_backward6<KIND, MS, false>);
        }
        if (shape == 4) {
            constexpr size_t smem4 = B3Smem<KIND>::BYTES;
            auto launch4 = [&](auto kern) -> int {
                { int rc_ = ensure_smem(h, kern, smem4); if (rc_) return rc_; }
                LAUNCH(kern, h->Bp / 32, kB4Threads, smem4, s, h->prm, h->w, it);
                return 0;
            };
            return h->var_horizons ? launch4(k_backward4<KIND, MS, true>) : launch4(k_backward4<KIND, MS, false>);
        }
        if (shape == 2) {
            constexpr size_t smem3 = B3Smem<KIND>::BYTES;
            auto launch = [&](auto kern, int groups) -> int {
                const size_t bytes = groups == 1 ? smem3 : (size_t)groups * B3Smem<KIND>::GROUP_BYTES;
                { int rc_ = ensure_smem(h, kern, bytes); if (rc_) return rc_; }
                LAUNCH(kern, blocks_for(h->Bp / 32, groups), kB3Threads * groups, bytes, s, h->prm, h->w, it);
                return 0;
            };
            if (h->var_horizons) return launch(k_backward3<KIND, MS, true, 1>, 1);
            const int groups = backward_groups();
            if (groups == 2) return launch(k_backward3<KIND, MS, false, 2>, 2);
#ifdef B3_WITH_G4
            if (groups == 4) return launch(k_backward3<KIND, MS, false, 4>, 4);
#endif
            return launch(k_backward3<KIND, MS, false, 1>, 1);
        }
    }
    constexpr size_t smem = bwd_smem_bytes<KIND, kBlock>();
    { int rc_ = ensure_smem(h, k_backward<KIND, MS, kBlock>, smem); if (rc_) return rc_; }
    LAUNCH((k_backward<KIND, MS, kBlock>), h->Bp / kBlock, kBlock, smem, s, h->prm, h->w, it);
    return 0;
}

template <int KIND, bool MS, bool WRITE, bool COST>
int run_forward(trajopt_handle* h, cudaStream_t s, int a_lo, int a_cnt, int need, int from_state) {
    PhaseTimer t(h, s, PH_FWD);
    const int grid = blocks_for(h->Bp * a_cnt, kBlock);
    constexpr size_t fsmem = FwdSmem<KIND>::BYTES;       // < 48 KB: no attribute needed
    if (h->prm.rollout_linear)
        LAUNCH((k_forward<KIND, MS, true, WRITE, COST>), grid, kBlock, fsmem, s, h->prm, h->w, a_lo, a_cnt, need, from_state);
    else
        LAUNCH((k_forward<KIND, MS, false, WRITE, COST>), grid, kBlock, fsmem, s, h->prm, h->w, a_lo, a_cnt, need, from_state);
    return 0;
}

// The once-per-iteration wait of the host loop.  cudaStreamSynchronize spins on a core, which is the lowest latency and
// fine for a few lanes; with many solver handles per GPU and one process per GPU (8 lanes x 8 ranks on the 32 cores of the
// measurement box) the spinning threads starve each other, so from six lanes up the thread sleeps on a blocking event.
inline cudaError_t wait_stream(trajopt_handle* h, cudaStream_t s) {
    if (h->sweep_lanes < 6) return cudaStreamSynchronize(s);
    if (!h->ev_block) {
        cudaError_t e = cudaEventCreateWithFlags(&h->ev_block, cudaEventDisableTiming | cudaEventBlockingSync);
        if (e != cudaSuccess) return e;
    }
    cudaError_t e = cudaEventRecord(h->ev_block, s);
    return e != cudaSuccess ? e : cudaEventSynchronize(h->ev_block);
}

// number of problems still running -> host (synchronises the stream)
inline int count_running(trajopt_handle* h, cudaStream_t s, int* out) {
    PhaseTimer t(h, s, PH_OTHER);
    CUDA_OK(cudaMemsetAsync(h->w.counters, 0, 4 * sizeof(int), s));
    LAUNCH(k_count_running, blocks_for(h->Bp, 128), 128, 0, s, h->prm, h->w);
    LAUNCH(k_publish4, 1, 32, 0, s, (const int*)h->w.counters, (volatile int*)h->h_counters_dev);
    CUDA_OK(wait_stream(h, s));
    *out = h->h_counters[0];
    return 0;
}

// one pass of the `for iteration in range(n_iterations)` body for every running problem
template <int KIND>
int inner_iteration(trajopt_handle* h, cudaStream_t s) {
    const int it = h->streaming ? -1 : h->it;    // streaming: every slot counts its own iterations (stream.cuh)
    const bool roll = h->streaming || h->it < h->prm.max_iters;
    const int na = h->prm.n_alphas;
    const int bgrid = blocks_for(h->Bp, 128);
    int rc;
    if (h->method == TRAJOPT_SS) {
        if ((rc = run_linearize<KIND, false>(h, s))) return rc;
        if ((rc = run_backward<KIND, false>(h, s, it))) return rc;
        if (h->w.Xc) {
            // small batch: the GPU is empty, so ONE launch rolls out every step size and keeps every candidate; what a solve
            // costs here is the number of dependent rollouts (each a chain of N Exp / Log / f evaluations), and the tail of
            // a single-shooting solve otherwise needs three per iteration
            if ((rc = run_forward<KIND, false, true, true>(h, s, 0, na, -4, 0))) return rc;
            LAUNCH(k_ls_select_ss, bgrid, 128, 0, s, h->prm, h->w, it, 0, na, 1);
            LAUNCH((k_ls_copy_cand<KIND>), dim3(bgrid, h->N + 1), 128, 0, s, h->prm, h->w);
            LAUNCH(k_ls_commit_ss, bgrid, 128, 0, s, h->prm, h->w, it);
            h->it = h->it + 1;
            return 0;
        }
        // line search (:1972-1990): step size 0 first (accepted by nearly every problem), then
        // all remaining step sizes at once for the problems that rejected it
        if ((rc = run_forward<KIND, false, true, true>(h, s, 0, 1, -2, 0))) return rc;
        LAUNCH(k_ls_select_ss, bgrid, 128, 0, s, h->prm, h->w, it, 0, 1, na == 1 ? 1 : 0);
        if (na > 1) {
            if ((rc = run_forward<KIND, false, true, true>(h, s, 1, na - 1, -1, 0))) return rc;      // costs only
            LAUNCH(k_ls_select_ss, bgrid, 128, 0, s, h->prm, h->w, it, 1, na - 1, 1);
            if ((rc = run_forward<KIND, false, true, true>(h, s, 1, 1, -3, 1))) return rc;            // the accepted one, kept
        }
        LAUNCH(k_ls_commit_ss, bgrid, 128, 0, s, h->prm, h->w, it);
    } else {
        if (!h->lin_ready && (rc = run_linearize<KIND, true>(h, s))) return rc;
        h->lin_ready = false;
        if ((rc = run_backward<KIND, true>(h, s, it))) return rc;
        if (roll) {
            if (h->prm.line_search) {
                {
                    PhaseTimer t(h, s, PH_FWD);
                    LAUNCH((k_ms_expected<KIND>), h->Bp / kBlock, kBlock, 0, s, h->prm, h->w, h->d_dweight);
                }
                if (h->w.Xc) {   // small batch: every step size in one launch (see the single-shooting branch)
                    if ((rc = run_forward<KIND, true, true, true>(h, s, 0, na, -4, 0))) return rc;
                    LAUNCH(k_ls_select_ms, bgrid, 128, 0, s, h->prm, h->w, it, 0, na, 1);
                    LAUNCH((k_ls_copy_cand<KIND>), dim3(bgrid, h->N + 1), 128, 0, s, h->prm, h->w);
                    LAUNCH(k_ls_commit_ms, bgrid, 128, 0, s, h->prm, h->w, it);
                    h->it = h->it + 1;
                    return 0;
                }
                if ((rc = run_forward<KIND, true, true, true>(h, s, 0, 1, -2, 0))) return rc;
                LAUNCH(k_ls_select_ms, bgrid, 128, 0, s, h->prm, h->w, it, 0, 1, na == 1 ? 1 : 0);
                if (na > 1) {
                    if ((rc = run_forward<KIND, true, true, true>(h, s, 1, na - 1, -1, 0))) return rc;
                    LAUNCH(k_ls_select_ms, bgrid, 128, 0, s, h->prm, h->w, it, 1, na - 1, 1);
                    if ((rc = run_forward<KIND, true, true, true>(h, s, 1, 1, -3, 1))) return rc;
                }
                LAUNCH(k_ls_commit_ms, bgrid, 128, 0, s, h->prm, h->w, it);
            } else {
                if (h->prm.rollout_linear) {
                    if ((rc = run_forward<KIND, true, true, false>(h, s, 0, 1, -2, 0))) return rc;
                } else if (overlap_applies(h)) {
                    if ((rc = run_forward_overlapped<KIND>(h, s, overlap_chunks()))) return rc;
                } else {
                    PhaseTimer t(h, s, PH_FWD);
                    constexpr size_t fsmem = FwdSmem<KIND>::BYTES;
                    { int rc_ = ensure_smem(h, k_forward_ms_full<KIND>, fsmem); if (rc_) return rc_; }
                    LAUNCH((k_forward_ms_full<KIND>), h->Bp / kBlock, kBlock, fsmem, s, h->prm, h->w, 0, h->N);
                }
                LAUNCH(k_accept_all, bgrid, 128, 0, s, h->prm, h->w, it);
            }
        }
    }
    h->it = h->it + 1;
    return 0;
}

// ------------------------------------------------------------------------------------------
// Compaction.  Every launch of the sweep kernels costs the same whether 16384 or 300 problems are still
// running (one sequential recursion per problem), but a CTA is only released when all of its 32 problems
// have finished — and the slow problems of a batch are usually spread over every warp.  Once the running
// problems are few, they are moved (stable partition) to the leading slots: the other CTAs exit at once and
// leave their SMs to whatever else is in flight on the device (the next batch of a PipelinedSolver).
// A problem's result does not depend on its slot (tests/test_gpu_fullsize.py), so this changes no number.
// Everything that is per problem and survives an iteration boundary moves; linearisation records, gains and
// G_i are rebuilt by the next iteration anyway.  `orig` remembers the caller's index for the exports.
// ------------------------------------------------------------------------------------------
template <typename T>
int permute_array(trajopt_handle* h, cudaStream_t s, T* data, size_t rows, int front) {
    if (!data || rows == 0) return 0;
    const size_t need = rows * (size_t)front * sizeof(T);
    if (need > h->scratch_bytes) {
        if (h->d_scratch) cudaFree(h->d_scratch);
        h->d_scratch = nullptr;
        h->scratch_bytes = 0;
        CUDA_OK(cudaMalloc(&h->d_scratch, need));
        h->scratch_bytes = need;
    }
    dim3 grid(blocks_for(front, 128), (unsigned)std::min<size_t>(rows, 4096));
    LAUNCH((k_permute_gather<T>), grid, 128, 0, s, (int)rows, h->Bp, front, (const T*)data, (const int*)h->d_perm, (T*)h->d_scratch);
    LAUNCH((k_permute_scatter<T>), grid, 128, 0, s, (int)rows, h->Bp, front, data, (const T*)h->d_scratch);
    return 0;
}

// data[r][orig[n]] <- data[r][n] for every slot n: undoes every compaction since `orig` was the identity.  Row chunks
// through the scratch buffer (at least 64 MiB of it), so a 4 GB per-problem reference does not need a 4 GB copy.
template <typename T>
int unpermute_array(trajopt_handle* h, cudaStream_t s, T* data, size_t rows) {
    if (!data || rows == 0) return 0;
    const size_t row_bytes = (size_t)h->Bp * sizeof(T);
    const size_t want = std::max<size_t>(row_bytes, std::min<size_t>(rows * row_bytes, (size_t)64 << 20));
    if (want > h->scratch_bytes) {
        if (h->d_scratch) cudaFree(h->d_scratch);
        h->d_scratch = nullptr;
        h->scratch_bytes = 0;
        CUDA_OK(cudaMalloc(&h->d_scratch, want));
        h->scratch_bytes = want;
    }
    const size_t chunk = std::max<size_t>(1, h->scratch_bytes / row_bytes);
    for (size_t r0 = 0; r0 < rows; r0 += chunk) {
        const size_t n = std::min(chunk, rows - r0);
        dim3 grid(blocks_for(h->Bp, 128), (unsigned)std::min<size_t>(n, 4096));
        LAUNCH((k_unpermute_scatter<T>), grid, 128, 0, s, (int)n, h->Bp, (const T*)(data + r0 * (size_t)h->Bp), (const int*)h->w.orig,
               (T*)h->d_scratch);
        CUDA_OK(cudaMemcpyAsync(data + r0 * (size_t)h->Bp, h->d_scratch, n * row_bytes, cudaMemcpyDeviceToDevice, s));
    }
    return 0;
}

// Called before `orig` is reset to the identity (trajopt_begin, trajopt_solve_stream): the per-problem settings a
// compaction of the previous solve moved go back to the caller's order.
inline int restore_caller_order(trajopt_handle* h, cudaStream_t s) {
    int rc = 0;
    if (h->nb_permuted) rc = unpermute_array(h, s, h->w.Nb, 1);
    if (!rc && h->ref_permuted && h->d_ref_batch) rc = unpermute_array(h, s, h->d_ref_batch, ((size_t)h->N + 1) * h->REFROW);
    h->nb_permuted = h->ref_permuted = false;
    return rc;
}

inline int maybe_compact(trajopt_handle* h, cudaStream_t s, int act) {
    if (h->compact_min_batch < 0 || h->Bp < h->compact_min_batch || act <= 0) return 0;
    const int front = h->front;
    if ((long long)act * h->compact_ratio > front || act >= front) return 0;
    PhaseTimer t(h, s, PH_OTHER);
    std::vector<int>& src_of = h->compact_src;
    src_of.resize((size_t)front);
    LAUNCH(k_ints_to_host, blocks_for(front, 256), 256, 0, s, (const int*)h->w.status, (volatile int*)h->h_ints_dev, front);
    CUDA_OK(cudaStreamSynchronize(s));
    const int* st = h->h_ints;
    int n = 0, last_running = -1;
    for (int b = 0; b < front; ++b)
        if ((st[b] & 15) == TRAJOPT_RUNNING) { src_of[n++] = b; last_running = b; }
    const int new_front = (n + 31) / 32 * 32;
    if (last_running < new_front) {      // already packed: just shrink the front
        h->front = std::max(new_front, 32);
        return 0;
    }
    for (int b = 0; b < front; ++b)
        if ((st[b] & 15) != TRAJOPT_RUNNING) src_of[n++] = b;
    if (!h->d_perm) CUDA_OK(cudaMalloc((void**)&h->d_perm, (size_t)h->Bp * sizeof(int)));
    CUDA_OK(cudaMemcpyAsync(h->d_perm, src_of.data(), (size_t)front * sizeof(int), cudaMemcpyHostToDevice, s));
    Work& w = h->w;
    const size_t Np1 = (size_t)h->N + 1, N = (size_t)h->N, mi = (size_t)std::max(h->prm.max_iters, 1);
    int rc = 0;
#define P_(ptr, rows) if (!rc) rc = permute_array(h, s, (ptr), (size_t)(rows), front)
    P_(w.X[0], Np1 * h->NS); P_(w.X[1], Np1 * h->NS); P_(w.U[0], N * h->NU); P_(w.U[1], N * h->NU);
    P_(w.x0, h->NS); P_(w.sel, 1); P_(w.orig, 1); P_(w.Nb, 1);
    P_(w.J, 1); P_(w.grad, 1); P_(w.dnorm, 1); P_(w.mu, 1); P_(w.delta, 1);
    P_(w.iters, 1); P_(w.status, 1); P_(w.ls_state, 1);
    P_(w.Jhist, mi); P_(w.gradhist, mi + 1); P_(w.defhist, mi + 1); P_(w.alphahist, mi);
    P_(h->d_dweight, 1);
    if (h->w.ref_batch) P_(h->d_ref_batch, Np1 * h->REFROW);
    h->nb_permuted = true;
    if (h->w.ref_batch) h->ref_permuted = true;
    if (h->method == TRAJOPT_AL_MS) {
        P_(w.lam, Np1 * 2 * h->NU); P_(w.imu, Np1 * 2 * h->NU);
        if (h->prm.has_state_bounds) { P_(w.lam_s, Np1 * 2 * (h->NX - h->NP)); P_(w.imu_s, Np1 * 2 * (h->NX - h->NP)); }
        P_(w.al_mu, 1); P_(w.al_outer, 1); P_(w.al_viol, 1); P_(w.al_done, 1);
    }
#undef P_
    if (rc) return rc;
    CUDA_OK(cudaStreamSynchronize(s));    // src_of (host) was read by an async copy
    h->lin_ready = false;                 // the records stayed behind in their old slots
    h->front = std::max(new_front, 32);
    return 0;
}

// run the inner loop for up to n_iters iterations; *active = problems still running
template <int KIND>
int run_inner(trajopt_handle* h, cudaStream_t s, int n_iters, int* active) {
    // MS needs one closing pass after the last rollout (cost / defect of the final trajectory)
    const int last = (h->method == TRAJOPT_SS) ? h->prm.max_iters : h->prm.max_iters + 1;
    int rc, act = -1;
    for (int j = 0; j < n_iters && h->it < last; ++j) {
        if ((rc = inner_iteration<KIND>(h, s))) return rc;
        if ((rc = count_running(h, s, &act))) return rc;
        if (act == 0) break;
        if ((rc = maybe_compact(h, s, act))) return rc;
    }
    if (act < 0 && (rc = count_running(h, s, &act))) return rc;
    if (act == 0 || h->it >= last) h->inner_done = true;
    *active = act;
    return 0;
}

template <int KIND>
int start_inner(trajopt_handle* h, cudaStream_t s, bool al_restart) {
    PhaseTimer t(h, s, PH_OTHER);
    const int bgrid = blocks_for(h->Bp, 128);
    if (al_restart) LAUNCH(k_reset_al_inner, bgrid, 128, 0, s, h->prm, h->w);
    else LAUNCH(k_reset, bgrid, 128, 0, s, h->prm, h->w);
    if (h->method == TRAJOPT_SS) {
        LAUNCH((k_init_ss<KIND>), h->Bp / kBlock, kBlock, 0, s, h->prm, h->w, (const int*)nullptr);
    } else {
        dim3 grid(bgrid, h->N + 1);
        LAUNCH((k_init_ms<KIND>), grid, 128, 0, s, h->prm, h->w, al_restart, (const int*)nullptr);
        if (h->prm.line_search) LAUNCH(k_fill_double, bgrid, 128, 0, s, h->Bp, h->prm.defect_mu0, h->d_dweight);
    }
    h->it = 0;
    h->inner_done = false;
    h->lin_ready = false;
    h->front = h->Bp;
    return 0;
}

template <int KIND>
int begin_impl(trajopt_handle* h, const double* d_x0, const double* d_us_init, int us_mode, cudaStream_t s) {
    h->w.us_init = d_us_init;
    h->w.us_mode = d_us_init ? us_mode : 0;
    {
        int rc0 = restore_caller_order(h, s);
        if (rc0) return rc0;
    }
    LAUNCH(k_identity, blocks_for(h->Bp, 128), 128, 0, s, h->Bp, h->w.orig);
    {   // histories: entries past a problem's last iteration read as zero
        const size_t Bp = (size_t)h->Bp, cap = (size_t)std::max(h->prm.max_iters, 1), cap1 = (size_t)h->prm.max_iters + 1;
        CUDA_OK(cudaMemsetAsync(h->w.Jhist, 0, cap * Bp * sizeof(double), s));
        CUDA_OK(cudaMemsetAsync(h->w.gradhist, 0, cap1 * Bp * sizeof(double), s));
        CUDA_OK(cudaMemsetAsync(h->w.defhist, 0, cap1 * Bp * sizeof(double), s));
        CUDA_OK(cudaMemsetAsync(h->w.alphahist, 0, cap * Bp * sizeof(int), s));
    }
    LAUNCH((k_load_x0<KIND>), blocks_for(h->Bp, 128), 128, 0, s, h->prm, h->w, d_x0);
    if (h->method == TRAJOPT_AL_MS) {
        dim3 grid(blocks_for(h->Bp, 128), h->N + 1);
        LAUNCH((k_al_init<KIND>), grid, 128, 0, s, h->prm, h->w, h->user.al_mu0);
    }
    h->al_outer = 0;
    h->al_finished = false;
    int rc = start_inner<KIND>(h, s, false);
    if (rc) return rc;
    h->al_inner_open = true;
    h->begun = true;
    return 0;
}

template <int KIND>
int iterate_impl(trajopt_handle* h, int n_iters, int* n_active_out, cudaStream_t s) {
    int rc, act = 0;
    if (h->method != TRAJOPT_AL_MS) {
        if (h->inner_done) {
            if ((rc = count_running(h, s, &act))) return rc;
        } else if ((rc = run_inner<KIND>(h, s, n_iters, &act))) return rc;
        if (n_active_out) *n_active_out = act;
        return 0;
    }
    // augmented Lagrangian: one unit = one outer iteration (:3231-3264)
    int remaining = 1;
    for (int j = 0; j < n_iters && !h->al_finished; ++j) {
        if (!h->al_inner_open) {
            if ((rc = start_inner<KIND>(h, s, true))) return rc;   // cold restart (:3237)
            h->al_inner_open = true;
        }
        if (!h->inner_done && (rc = run_inner<KIND>(h, s, h->prm.max_iters + 1, &act))) return rc;
        h->al_inner_open = false;
        {
            PhaseTimer t(h, s, PH_OTHER);
            CUDA_OK(cudaMemsetAsync(h->w.counters, 0, 4 * sizeof(int), s));
            const int bgrid = blocks_for(h->Bp, 128);
            const dim3 sgrid(bgrid, blocks_for(h->N + 1, kAlChunk));
            LAUNCH(k_al_viol_zero, bgrid, 128, 0, s, h->prm, h->w);
            LAUNCH((k_al_viol<KIND>), sgrid, 128, 0, s, h->prm, h->w);
            LAUNCH(k_al_decide, bgrid, 128, 0, s, h->prm, h->w, h->user.tol_constr, h->user.al_mu_scale, h->user.al_mu_max, h->al_outer);
            LAUNCH((k_al_apply<KIND>), sgrid, 128, 0, s, h->prm, h->w);
            LAUNCH(k_publish4, 1, 32, 0, s, (const int*)h->w.counters, (volatile int*)h->h_counters_dev);
            CUDA_OK(cudaStreamSynchronize(s));
        }
        remaining = h->h_counters[2];
        h->al_outer += 1;
        if (remaining == 0 || h->al_outer >= h->user.n_al_iters) h->al_finished = true;
    }
    if (n_active_out) *n_active_out = h->al_finished ? 0 : remaining;
    return 0;
}

// Augmented Lagrangian, inner iterations one at a time (trajopt_iterate_inner): up to n_iters iterations of the inner
// solve of the CURRENT outer iteration (started here if the previous one has been closed by trajopt_iterate).  The
// multiplier update stays with trajopt_iterate.  *n_active_out = problems whose inner solve is still running.
template <int KIND>
int iterate_inner_impl(trajopt_handle* h, int n_iters, int* n_active_out, cudaStream_t s) {
    if (h->method != TRAJOPT_AL_MS) return iterate_impl<KIND>(h, n_iters, n_active_out, s);
    int rc, act = 0;
    if (!h->al_finished) {
        if (!h->al_inner_open) {
            if ((rc = start_inner<KIND>(h, s, true))) return rc;
            h->al_inner_open = true;
        }
        if (h->inner_done) {
            if ((rc = count_running(h, s, &act))) return rc;
        } else if ((rc = run_inner<KIND>(h, s, n_iters, &act))) return rc;
        if (h->inner_done) act = 0;
    }
    if (n_active_out) *n_active_out = act;
    return 0;
}

template <int KIND>
int debug_linearize_impl(trajopt_handle* h, double* Fx, double* Fu, double* dd, double* L, double* Lx, double* Lxx,
                         double* Lu, cudaStream_t s) {
    int rc;
    if (h->method == TRAJOPT_SS) rc = run_linearize<KIND, false>(h, s);
    else rc = run_linearize<KIND, true>(h, s);
    if (rc) return rc;
    dim3 grid(blocks_for(h->Bp, 128), h->N + 1);
    LAUNCH((k_export_lin<KIND>), grid, 128, 0, s, h->prm, h->w, Fx, Fu, dd, L, Lx, Lxx, Lu);
    return 0;
}

template <int KIND>
int set_reference_batch_impl(trajopt_handle* h, const double* d_q, const double* d_xi, cudaStream_t s) {
    dim3 grid(blocks_for(h->Bp, 128), h->N + 1);
    LAUNCH((k_pack_ref_batch<KIND>), grid, 128, 0, s, h->B, h->Bp, h->N + 1, d_q, d_xi, h->d_ref_batch);
    return 0;
}

template <int KIND>
int debug_stage_impl(trajopt_handle* h, int i, int terminal, int n, const double* x, const double* u, double* f,
                     double* Fx, double* Fu, double* l, double* lx, double* lxx, double* lu, double* err, cudaStream_t s) {
    LAUNCH((k_debug_stage<KIND>), blocks_for(n, 64), 64, 0, s, h->prm, h->w.ref, i, terminal, n, x, u, f, Fx, Fu, l, lx,
           lxx, lu, err);
    return 0;
}


inline int ensure_hist(trajopt_handle* h) {
    const int cap = h->prm.max_iters;
    const int rows = (h->method == TRAJOPT_SS) ? h->prm.n_alphas : 2 * h->prm.n_alphas + 4;
    if (cap == h->hist_cap && rows == h->cand_rows) return 0;
    for (void*& p : h->hist_allocs) {
        if (p) cudaFree(p);
        p = nullptr;
    }
    const size_t Bp = (size_t)h->Bp;
    const size_t sz[5] = {(size_t)(cap > 0 ? cap : 1) * Bp * 8, (size_t)(cap + 1) * Bp * 8, (size_t)(cap + 1) * Bp * 8,
                          (size_t)(cap > 0 ? cap : 1) * Bp * 4, (size_t)rows * Bp * 8};
    for (int i = 0; i < 5; ++i) {
        CUDA_OK(cudaMalloc(&h->hist_allocs[i], sz[i]));
        CUDA_OK(cudaMemset(h->hist_allocs[i], 0, sz[i]));
    }
    h->w.Jhist = (double*)h->hist_allocs[0];
    h->w.gradhist = (double*)h->hist_allocs[1];
    h->w.defhist = (double*)h->hist_allocs[2];
    h->w.alphahist = (int*)h->hist_allocs[3];
    h->w.Jcand = (double*)h->hist_allocs[4];
    h->hist_cap = cap;
    h->cand_rows = rows;
    return 0;
}


// Candidate trajectory buffers of the one-launch line search: only where a line search runs (single shooting, multiple
// shooting with line_search), more than one step size, and a batch small enough that n_alphas rollouts side by side still
// leave the GPU under-filled (<= cand_max_batch slots; TRAJOPT_LS_ALL_BATCH overrides, 0 disables).
inline int ensure_cand(trajopt_handle* h) {
    static const int env_max = [] { const char* e = getenv("TRAJOPT_LS_ALL_BATCH"); return e ? atoi(e) : -1; }();
    const int max_batch = env_max >= 0 ? env_max : h->cand_max_batch;
    const bool ls = h->method == TRAJOPT_SS || (h->method != TRAJOPT_SS && h->prm.line_search);
    const int want = (ls && h->prm.n_alphas > 1 && h->Bp <= max_batch) ? h->prm.n_alphas : 0;
    if (want == h->cand_alphas) return 0;
    for (void*& p : h->cand_allocs) {
        if (p) cudaFree(p);
        p = nullptr;
    }
    h->w.Xc = h->w.Uc = nullptr;
    h->cand_alphas = 0;
    if (want) {
        const size_t Bp = (size_t)h->Bp;
        const size_t sz[2] = {(size_t)want * (h->N + 1) * h->NS * Bp * 8, (size_t)want * h->N * h->NU * Bp * 8};
        for (int i = 0; i < 2; ++i) {
            CUDA_OK(cudaMalloc(&h->cand_allocs[i], sz[i]));
            CUDA_OK(cudaMemset(h->cand_allocs[i], 0, sz[i]));
        }
        h->w.Xc = (double*)h->cand_allocs[0];
        h->w.Uc = (double*)h->cand_allocs[1];
        h->cand_alphas = want;
    }
    return 0;
}

}  // namespace trajopt_host

#include "stream.cuh"

// the entry points api.cu dispatches on the family
#define TRAJOPT_KIND_INSTANTIATE(K)                                                                                             \
    template int trajopt_host::set_reference_batch_impl<K>(trajopt_handle*, const double*, const double*, cudaStream_t);        \
    template int trajopt_host::begin_impl<K>(trajopt_handle*, const double*, const double*, int, cudaStream_t);                  \
    template int trajopt_host::iterate_impl<K>(trajopt_handle*, int, int*, cudaStream_t);                                        \
    template int trajopt_host::iterate_inner_impl<K>(trajopt_handle*, int, int*, cudaStream_t);                                  \
    template int trajopt_host::solve_stream_impl<K>(trajopt_handle*, const double*, int, double*, double*, double*, int*, int*, \
                                                    double*, double*, cudaStream_t);                                             \
    template int trajopt_host::debug_linearize_impl<K>(trajopt_handle*, double*, double*, double*, double*, double*, double*,   \
                                                       double*, cudaStream_t);                                                   \
    template int trajopt_host::debug_stage_impl<K>(trajopt_handle*, int, int, int, const double*, const double*, double*,       \
                                                   double*, double*, double*, double*, double*, double*, double*, cudaStream_t);
#define TRAJOPT_KIND_EXTERN(K)                                                                                                  \
    extern template int trajopt_host::set_reference_batch_impl<K>(trajopt_handle*, const double*, const double*, cudaStream_t); \
    extern template int trajopt_host::begin_impl<K>(trajopt_handle*, const double*, const double*, int, cudaStream_t);           \
    extern template int trajopt_host::iterate_impl<K>(trajopt_handle*, int, int*, cudaStream_t);                                 \
    extern template int trajopt_host::iterate_inner_impl<K>(trajopt_handle*, int, int*, cudaStream_t);                           \
    extern template int trajopt_host::solve_stream_impl<K>(trajopt_handle*, const double*, int, double*, double*, double*,      \
                                                           int*, int*, double*, double*, cudaStream_t);                          \
    extern template int trajopt_host::debug_linearize_impl<K>(trajopt_handle*, double*, double*, double*, double*, double*,     \
                                                              double*, double*, cudaStream_t);                                   \
    extern template int trajopt_host::debug_stage_impl<K>(trajopt_handle*, int, int, int, const double*, const double*,         \
                                                          double*, double*, double*, double*, double*, double*, double*,        \
                                                          double*, cudaStream_t);
