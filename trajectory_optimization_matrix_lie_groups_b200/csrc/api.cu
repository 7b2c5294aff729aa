// C ABI of the batched DDP/iLQR solver (see include/trajopt_b200.h for the contract and the
// reference call each entry point stands for).  Host side only orchestrates kernel launches on
// one stream; every number is produced by the kernels in kernels.cuh / kernels_fwd.cuh.
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <new>
#include <vector>

#include <cstdlib>

#include "kernels_fwd.cuh"
#include "debug.cuh"

using namespace trajopt;

// ------------------------------------------------------------------------------------------
// error reporting / launch accounting
// ------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

extern "C" int trajopt_set_error_(cudaError_t e, const char* file, int line) {
    snprintf(g_err, sizeof(g_err), "CUDA error %d (%s) at %s:%d", (int)e, cudaGetErrorString(e), file, line);
    return TRAJOPT_E_CUDA;
}
static int fail(int code, const char* msg) {
    snprintf(g_err, sizeof(g_err), "%s", msg);
    return code;
}

#define LAUNCH(kern, grid, block, smem, stream, ...)                 \
    do {                                                             \
        kern<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);    \
        g_launches.fetch_add(1, std::memory_order_relaxed);          \
        CUDA_OK(cudaPeekAtLastError());                              \
    } while (0)

enum { PH_LIN = 0, PH_BWD = 1, PH_FWD = 2, PH_OTHER = 3, PH_COUNT = 4 };

struct trajopt_handle {
    int kind = 0, method = 0, N = 0, B = 0, Bp = 0, device = 0;
    int NX = 0, NP = 0, NU = 0, NS = 0, LEN = 0, REFROW = 0;
    Params prm{};
    Work w{};
    trajopt_params user{};
    bool have_params = false, have_ref = false, begun = false;
    int it = 0;              // next inner iteration
    bool inner_done = true;  // the inner loop of the current fit() / AL outer iteration has ended
    int al_outer = 0;        // AL outer iterations completed
    bool al_finished = false;
    int hist_cap = -1, cand_rows = -1;
    std::vector<void*> allocs;
    void* hist_allocs[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    double* d_ref = nullptr;
    double* d_dweight = nullptr;
    int* h_counters = nullptr;   // pinned
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

namespace {

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

// host copies of the small algebra needed to pack parameters
void inv3(const double* A, double* Ai) {
    const double a = A[0], b = A[1], c = A[2], d = A[3], e = A[4], f = A[5], g = A[6], hh = A[7], i = A[8];
    const double det = a * (e * i - f * hh) - b * (d * i - f * g) + c * (d * hh - e * g);
    const double id = 1.0 / det;
    Ai[0] = (e * i - f * hh) * id; Ai[1] = (c * hh - b * i) * id; Ai[2] = (b * f - c * e) * id;
    Ai[3] = (f * g - d * i) * id;  Ai[4] = (a * i - c * g) * id;  Ai[5] = (c * d - a * f) * id;
    Ai[6] = (d * hh - e * g) * id; Ai[7] = (b * g - a * hh) * id; Ai[8] = (a * e - b * d) * id;
}
void host_quat_to_rot(const double* q, double* R) {
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
    LAUNCH((k_linearize<KIND, MS>), grid, 128, 0, s, h->prm, h->w);
    return 0;
}

// TRAJOPT_BACKWARD=1 selects the one-warp-per-group sweep for the 12-dimensional families too (A/B measurements)
static bool use_one_warp_sweep() {
    static const bool v = [] { const char* e = getenv("TRAJOPT_BACKWARD"); return e && e[0] == '1'; }();
    return v;
}

template <int KIND, bool MS>
int run_backward(trajopt_handle* h, cudaStream_t s, int it) {
    PhaseTimer t(h, s, PH_BWD);
    if constexpr (!on_so3(KIND)) {
        if (!use_one_warp_sweep()) {
            constexpr size_t smem3 = B3Smem<KIND>::BYTES;
            CUDA_OK(cudaFuncSetAttribute(k_backward3<KIND, MS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem3));
            LAUNCH((k_backward3<KIND, MS>), h->Bp / 32, kB3Threads, smem3, s, h->prm, h->w, it);
            return 0;
        }
    }
    constexpr size_t smem = (size_t)bwd_smem_doubles<KIND>() * kBlock * sizeof(double);
    CUDA_OK(cudaFuncSetAttribute(k_backward<KIND, MS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    LAUNCH((k_backward<KIND, MS>), h->Bp / kBlock, kBlock, smem, s, h->prm, h->w, it);
    return 0;
}

template <int KIND, bool MS, bool WRITE, bool COST>
int run_forward(trajopt_handle* h, cudaStream_t s, int a_lo, int a_cnt, int need, int from_state) {
    PhaseTimer t(h, s, PH_FWD);
    const int grid = blocks_for(h->Bp * a_cnt, kBlock);
    if (h->prm.rollout_linear)
        LAUNCH((k_forward<KIND, MS, true, WRITE, COST>), grid, kBlock, 0, s, h->prm, h->w, a_lo, a_cnt, need, from_state);
    else
        LAUNCH((k_forward<KIND, MS, false, WRITE, COST>), grid, kBlock, 0, s, h->prm, h->w, a_lo, a_cnt, need, from_state);
    return 0;
}

// number of problems still running -> host (synchronises the stream)
int count_running(trajopt_handle* h, cudaStream_t s, int* out) {
    PhaseTimer t(h, s, PH_OTHER);
    CUDA_OK(cudaMemsetAsync(h->w.counters, 0, 4 * sizeof(int), s));
    LAUNCH(k_count_running, blocks_for(h->Bp, 128), 128, 0, s, h->prm, h->w);
    CUDA_OK(cudaMemcpyAsync(h->h_counters, h->w.counters, 4 * sizeof(int), cudaMemcpyDeviceToHost, s));
    CUDA_OK(cudaStreamSynchronize(s));
    *out = h->h_counters[0];
    return 0;
}

// one pass of the `for iteration in range(n_iterations)` body for every running problem
template <int KIND>
int inner_iteration(trajopt_handle* h, cudaStream_t s) {
    const int it = h->it;
    const int na = h->prm.n_alphas;
    const int bgrid = blocks_for(h->Bp, 128);
    int rc;
    if (h->method == TRAJOPT_SS) {
        if ((rc = run_linearize<KIND, false>(h, s))) return rc;
        if ((rc = run_backward<KIND, false>(h, s, it))) return rc;
        // line search (:1972-1990): step size 0 first (accepted by nearly every problem), then
        // all remaining step sizes at once for the problems that rejected it
        if ((rc = run_forward<KIND, false, true, true>(h, s, 0, 1, -2, 0))) return rc;
        LAUNCH(k_ls_select_ss, bgrid, 128, 0, s, h->prm, h->w, it, 0, 1, na == 1 ? 1 : 0);
        if (na > 1) {
            if ((rc = run_forward<KIND, false, false, true>(h, s, 1, na - 1, -1, 0))) return rc;
            LAUNCH(k_ls_select_ss, bgrid, 128, 0, s, h->prm, h->w, it, 1, na - 1, 1);
            if ((rc = run_forward<KIND, false, true, false>(h, s, 1, 1, -3, 1))) return rc;
        }
        LAUNCH(k_ls_commit_ss, bgrid, 128, 0, s, h->prm, h->w, it);
    } else {
        if ((rc = run_linearize<KIND, true>(h, s))) return rc;
        if ((rc = run_backward<KIND, true>(h, s, it))) return rc;
        if (it < h->prm.max_iters) {
            if (h->prm.line_search) {
                {
                    PhaseTimer t(h, s, PH_FWD);
                    LAUNCH((k_ms_expected<KIND>), h->Bp / kBlock, kBlock, 0, s, h->prm, h->w, h->d_dweight);
                }
                if ((rc = run_forward<KIND, true, true, true>(h, s, 0, 1, -2, 0))) return rc;
                LAUNCH(k_ls_select_ms, bgrid, 128, 0, s, h->prm, h->w, it, 0, 1, na == 1 ? 1 : 0);
                if (na > 1) {
                    if ((rc = run_forward<KIND, true, false, true>(h, s, 1, na - 1, -1, 0))) return rc;
                    LAUNCH(k_ls_select_ms, bgrid, 128, 0, s, h->prm, h->w, it, 1, na - 1, 1);
                    if ((rc = run_forward<KIND, true, true, false>(h, s, 1, 1, -3, 1))) return rc;
                }
                LAUNCH(k_ls_commit_ms, bgrid, 128, 0, s, h->prm, h->w, it);
            } else {
                if (h->prm.rollout_linear) {
                    if ((rc = run_forward<KIND, true, true, false>(h, s, 0, 1, -2, 0))) return rc;
                } else {
                    PhaseTimer t(h, s, PH_FWD);
                    constexpr size_t fsmem = FwdSmem<KIND>::BYTES;
                    CUDA_OK(cudaFuncSetAttribute(k_forward_ms_full<KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsmem));
                    LAUNCH((k_forward_ms_full<KIND>), h->Bp / kBlock, kBlock, fsmem, s, h->prm, h->w);
                }
                LAUNCH(k_accept_all, bgrid, 128, 0, s, h->prm, h->w, it);
            }
        }
    }
    h->it = it + 1;
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
        LAUNCH((k_init_ss<KIND>), h->Bp / kBlock, kBlock, 0, s, h->prm, h->w);
    } else {
        dim3 grid(bgrid, h->N + 1);
        LAUNCH((k_init_ms<KIND>), grid, 128, 0, s, h->prm, h->w, al_restart);
        if (h->prm.line_search) {
            std::vector<double> dw((size_t)h->Bp, h->prm.defect_mu0);
            CUDA_OK(cudaMemcpyAsync(h->d_dweight, dw.data(), dw.size() * sizeof(double), cudaMemcpyHostToDevice, s));
            CUDA_OK(cudaStreamSynchronize(s));
        }
    }
    h->it = 0;
    h->inner_done = false;
    return 0;
}

template <int KIND>
int begin_impl(trajopt_handle* h, const double* d_x0, const double* d_us_init, int us_mode, cudaStream_t s) {
    h->w.us_init = d_us_init;
    h->w.us_mode = d_us_init ? us_mode : 0;
    LAUNCH((k_load_x0<KIND>), blocks_for(h->Bp, 128), 128, 0, s, h->prm, h->w, d_x0);
    if (h->method == TRAJOPT_AL_MS) {
        dim3 grid(blocks_for(h->Bp, 128), h->N + 1);
        LAUNCH((k_al_init<KIND>), grid, 128, 0, s, h->prm, h->w, h->user.al_mu0);
    }
    h->al_outer = 0;
    h->al_finished = false;
    int rc = start_inner<KIND>(h, s, false);
    if (rc) return rc;
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
        if (h->inner_done) {
            if ((rc = start_inner<KIND>(h, s, true))) return rc;   // cold restart (:3237)
        }
        if ((rc = run_inner<KIND>(h, s, h->prm.max_iters + 1, &act))) return rc;
        {
            PhaseTimer t(h, s, PH_OTHER);
            CUDA_OK(cudaMemsetAsync(h->w.counters, 0, 4 * sizeof(int), s));
            LAUNCH((k_al_update<KIND>), blocks_for(h->Bp, 128), 128, 0, s, h->prm, h->w, h->user.tol_constr,
                   h->user.al_mu_scale, h->user.al_mu_max, h->al_outer);
            CUDA_OK(cudaMemcpyAsync(h->h_counters, h->w.counters, 4 * sizeof(int), cudaMemcpyDeviceToHost, s));
            CUDA_OK(cudaStreamSynchronize(s));
        }
        remaining = h->h_counters[2];
        h->al_outer += 1;
        if (remaining == 0 || h->al_outer >= h->user.n_al_iters) h->al_finished = true;
    }
    if (n_active_out) *n_active_out = h->al_finished ? 0 : remaining;
    return 0;
}

template <int KIND>
int debug_linearize_impl(trajopt_handle* h, double* Fx, double* Fu, double* dd, double* L, double* Lx, double* Lxx,
                         double* Lu, cudaStream_t s) {
    int rc;
    if (h->method == TRAJOPT_SS) rc = run_linearize<KIND, false>(h, s);
    else rc = run_linearize<KIND, true>(h, s);
    if (rc) return rc;
    dim3 grid(blocks_for(h->B, 128), h->N + 1);
    LAUNCH((k_export_lin<KIND>), grid, 128, 0, s, h->prm, h->w, Fx, Fu, dd, L, Lx, Lxx, Lu);
    return 0;
}

template <int KIND>
int debug_stage_impl(trajopt_handle* h, int i, int terminal, int n, const double* x, const double* u, double* f,
                     double* Fx, double* Fu, double* l, double* lx, double* lxx, double* lu, double* err, cudaStream_t s) {
    LAUNCH((k_debug_stage<KIND>), blocks_for(n, 64), 64, 0, s, h->prm, h->d_ref, i, terminal, n, x, u, f, Fx, Fu, l, lx,
           lxx, lu, err);
    return 0;
}

#define DISPATCH_KIND(h, fn, ...)                                             \
    ((h)->kind == TRAJOPT_SO3   ? fn<TRAJOPT_SO3>(__VA_ARGS__)                \
     : (h)->kind == TRAJOPT_SE3 ? fn<TRAJOPT_SE3>(__VA_ARGS__)                \
     : (h)->kind == TRAJOPT_DRONE ? fn<TRAJOPT_DRONE>(__VA_ARGS__)            \
     : (h)->kind == TRAJOPT_RIGID ? fn<TRAJOPT_RIGID>(__VA_ARGS__)            \
                                  : fn<TRAJOPT_PEND>(__VA_ARGS__))

int ensure_hist(trajopt_handle* h) {
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

}  // namespace

// ------------------------------------------------------------------------------------------
// exported entry points
// ------------------------------------------------------------------------------------------
extern "C" {

const char* trajopt_last_error(void) { return g_err; }
int trajopt_version(void) { return 100; }

int64_t trajopt_launch_count(int reset) {
    const long long v = g_launches.load();
    if (reset) g_launches.store(0);
    return (int64_t)v;
}

int trajopt_create(int kind, int method, int N, int B, int device, trajopt_handle** out) {
    if (!out) return fail(TRAJOPT_E_INVALID, "trajopt_create: out is NULL");
    *out = nullptr;
    if (kind < TRAJOPT_SO3 || kind > TRAJOPT_PEND) return fail(TRAJOPT_E_INVALID, "trajopt_create: unknown problem kind");
    if (method < TRAJOPT_SS || method > TRAJOPT_AL_MS) return fail(TRAJOPT_E_INVALID, "trajopt_create: unknown method");
    if (N < 1 || N > 65534 || B < 1) return fail(TRAJOPT_E_INVALID, "trajopt_create: need 1 <= N <= 65534 and B >= 1");
    if (method == TRAJOPT_AL_MS && on_so3(kind))
        return fail(TRAJOPT_E_INVALID, "trajopt_create: the augmented-Lagrangian controller exists for SE3 problems only");
    int ndev = 0;
    CUDA_OK(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return fail(TRAJOPT_E_INVALID, "trajopt_create: no such CUDA device");
    DeviceGuard guard(device);
    trajopt_handle* h = new (std::nothrow) trajopt_handle();
    if (!h) return fail(TRAJOPT_E_INVALID, "trajopt_create: out of host memory");
    h->kind = kind; h->method = method; h->N = N; h->B = B; h->device = device;
    h->Bp = (B + kBlock - 1) / kBlock * kBlock;
    if (kind == TRAJOPT_SO3) { h->NX = 6; h->NP = 3; h->NU = 3; h->NS = 7; h->LEN = LinRec<TRAJOPT_SO3>::LEN; h->REFROW = RefRow<TRAJOPT_SO3>::N; }
    else if (kind == TRAJOPT_PEND) { h->NX = 6; h->NP = 3; h->NU = 3; h->NS = 7; h->LEN = LinRec<TRAJOPT_PEND>::LEN; h->REFROW = RefRow<TRAJOPT_PEND>::N; }
    else if (kind == TRAJOPT_SE3) { h->NX = 12; h->NP = 6; h->NU = 6; h->NS = 13; h->LEN = LinRec<TRAJOPT_SE3>::LEN; h->REFROW = RefRow<TRAJOPT_SE3>::N; }
    else if (kind == TRAJOPT_DRONE) { h->NX = 12; h->NP = 6; h->NU = 4; h->NS = 13; h->LEN = LinRec<TRAJOPT_DRONE>::LEN; h->REFROW = RefRow<TRAJOPT_DRONE>::N; }
    else { h->NX = 12; h->NP = 6; h->NU = 6; h->NS = 13; h->LEN = LinRec<TRAJOPT_RIGID>::LEN; h->REFROW = RefRow<TRAJOPT_RIGID>::N; }
    const size_t Bp = h->Bp, Np1 = (size_t)N + 1;
    Work& w = h->w;
    int rc = 0;
#define A_(ptr, count) if (!rc) rc = dalloc(h, &(ptr), (count))
    A_(w.X[0], Np1 * h->NS * Bp); A_(w.X[1], Np1 * h->NS * Bp);
    A_(w.U[0], (size_t)N * h->NU * Bp); A_(w.U[1], (size_t)N * h->NU * Bp);
    A_(w.sel, Bp);
    A_(h->d_ref, Np1 * h->REFROW);
    A_(w.lin, Np1 * h->LEN * Bp);
    A_(w.Lc, Np1 * Bp); A_(w.Dsq, (size_t)N * Bp);
    if (method != TRAJOPT_SS) {
        const int gl = on_so3(kind) ? GPre<TRAJOPT_SO3>::LEN : GPre<TRAJOPT_SE3>::LEN;
        A_(w.Gpre, (size_t)N * gl * Bp);
    }
    A_(w.kff, (size_t)N * h->NU * Bp); A_(w.Kfb, (size_t)N * h->NU * h->NX * Bp);
    A_(w.J, Bp); A_(w.grad, Bp); A_(w.dnorm, Bp); A_(w.mu, Bp); A_(w.delta, Bp);
    A_(w.iters, Bp); A_(w.status, Bp); A_(w.ls_state, Bp);
    A_(w.x0, (size_t)h->NS * Bp);
    A_(w.counters, 4);
    A_(h->d_dweight, Bp);
    if (method == TRAJOPT_AL_MS) {
        A_(w.lam, Np1 * 2 * h->NU * Bp); A_(w.imu, Np1 * 2 * h->NU * Bp);
        A_(w.al_mu, Bp); A_(w.al_outer, Bp); A_(w.al_viol, Bp); A_(w.al_done, Bp);
    }
#undef A_
    w.ref = h->d_ref;
    if (!rc && cudaMallocHost((void**)&h->h_counters, 4 * sizeof(int)) != cudaSuccess) rc = fail(TRAJOPT_E_CUDA, "cudaMallocHost failed");
    if (!rc && (cudaEventCreate(&h->ev[0]) != cudaSuccess || cudaEventCreate(&h->ev[1]) != cudaSuccess))
        rc = fail(TRAJOPT_E_CUDA, "cudaEventCreate failed");
    if (rc) {
        trajopt_destroy(h);
        return rc;
    }
    *out = h;
    return 0;
}

int trajopt_destroy(trajopt_handle* h) {
    if (!h) return 0;
    DeviceGuard guard(h->device);
    for (void* p : h->allocs) cudaFree(p);
    for (void* p : h->hist_allocs) if (p) cudaFree(p);
    void* stage[] = {h->s_x0, h->s_us0, h->s_xs, h->s_us, h->s_J, h->s_grad, h->s_def, h->s_iters, h->s_status};
    for (void* p : stage) if (p) cudaFree(p);
    if (h->h_counters) cudaFreeHost(h->h_counters);
    if (h->ev[0]) cudaEventDestroy(h->ev[0]);
    if (h->ev[1]) cudaEventDestroy(h->ev[1]);
    delete h;
    return 0;
}

int trajopt_set_params(trajopt_handle* h, const trajopt_params* p) {
    if (!h || !p) return fail(TRAJOPT_E_INVALID, "trajopt_set_params: NULL argument");
    if (!(p->dt > 0.0) || !(p->mass > 0.0)) return fail(TRAJOPT_E_INVALID, "trajopt_set_params: dt and mass must be positive");
    if (p->max_iters < 0) return fail(TRAJOPT_E_INVALID, "trajopt_set_params: max_iters < 0");
    if (h->method == TRAJOPT_AL_MS && !p->has_constraints)
        return fail(TRAJOPT_E_INVALID, "trajopt_set_params: the augmented-Lagrangian method needs input bounds");
    DeviceGuard guard(h->device);
    h->user = *p;
    Params& q = h->prm;
    const int NX = h->NX, NP = h->NP, NU = h->NU, NV = NX - NP;
    q.kind = h->kind; q.N = h->N; q.B = h->B; q.Bp = h->Bp; q.method = h->method;
    q.rollout_linear = p->rollout_linear ? 1 : 0;
    q.line_search = (h->method != TRAJOPT_SS && p->line_search) ? 1 : 0;
    q.n_alphas = p->n_alphas > 0 ? p->n_alphas : ((h->method == TRAJOPT_SS || on_so3(h->kind)) ? 13 : 20);
    if (q.n_alphas > 64) return fail(TRAJOPT_E_INVALID, "trajopt_set_params: n_alphas > 64");
    q.max_iters = p->max_iters;
    q.has_constraints = (h->method == TRAJOPT_AL_MS) ? 1 : 0;
    q.dt = p->dt;
    memcpy(q.Ib, p->Ib, sizeof(q.Ib));
    inv3(q.Ib, q.Ibinv);
    q.mass = p->mass;
    q.grav = p->gravity;
    q.length = p->length;
    if (h->kind == TRAJOPT_PEND && !(p->length > 0.0)) return fail(TRAJOPT_E_INVALID, "trajopt_set_params: the pendulum needs length > 0");
    memset(q.W1, 0, sizeof(q.W1)); memset(q.W2, 0, sizeof(q.W2));
    memset(q.P1, 0, sizeof(q.P1)); memset(q.P2, 0, sizeof(q.P2));
    memset(q.R, 0, sizeof(q.R)); memset(q.Bv, 0, sizeof(q.Bv));
    for (int r = 0; r < NP; ++r)
        for (int c = 0; c < NP; ++c) {
            q.W1[r * NP + c] = p->Q[r * NX + c];
            q.P1[r * NP + c] = p->P[r * NX + c];
        }
    for (int r = 0; r < NV; ++r)
        for (int c = 0; c < NV; ++c) {
            q.W2[r * NV + c] = p->Q[(NP + r) * NX + NP + c];
            q.P2[r * NV + c] = p->P[(NP + r) * NX + NP + c];
        }
    for (int r = 0; r < NU; ++r)
        for (int c = 0; c < NU; ++c) q.R[r * NU + c] = p->R[r * NU + c];
    // velocity rows of f_u = J^-1 Pu dt  (traopt_dynamics.py:311-313, 668-670, 1256-1258)
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) q.Bv[r * NU + c] = q.Ibinv[3 * r + c] * q.dt;
    if (h->kind == TRAJOPT_SE3 || h->kind == TRAJOPT_RIGID)
        for (int r = 0; r < 3; ++r) q.Bv[(3 + r) * NU + 3 + r] = q.dt / q.mass;
    if (h->kind == TRAJOPT_DRONE) q.Bv[5 * NU + 3] = q.dt / q.mass;
    memset(q.BtB, 0, sizeof(q.BtB));
    for (int a = 0; a < NU; ++a)
        for (int c = 0; c < NU; ++c) {
            double sacc = 0.0;
            for (int r = 0; r < NV; ++r) sacc = std::fma(q.Bv[r * NU + a], q.Bv[r * NU + c], sacc);
            q.BtB[a * NU + c] = sacc;
        }
    for (int j = 0; j < 6; ++j) { q.lb[j] = p->lb[j]; q.ub[j] = p->ub[j]; }
    q.tol_grad = p->tol_grad_norm;
    q.tol_defect = p->tol_d_norm;
    q.mu_min = 1e-6;          // traopt_controller.py:1863-1866
    q.mu_max = p->max_reg;
    q.delta0 = 2.0;
    q.defect_mu0 = 10.0; q.defect_rho = 0.5; q.defect_gamma = 0.05;   // :2406-2410
    q.defect_kappa = p->defect_kappa > 0.0 ? p->defect_kappa : (on_so3(h->kind) ? 1e-14 : 1e-12);
    q.so3_terminal_quirk = on_so3(h->kind) ? 1 : 0;
    int rc = ensure_hist(h);
    if (rc) return rc;
    h->have_params = true;
    h->begun = false;
    return 0;
}

int trajopt_set_reference(trajopt_handle* h, const double* h_q_ref, const double* h_xi_ref) {
    if (!h || !h_q_ref || !h_xi_ref) return fail(TRAJOPT_E_INVALID, "trajopt_set_reference: NULL argument");
    DeviceGuard guard(h->device);
    const int rr = h->REFROW;
    std::vector<double> rows((size_t)(h->N + 1) * rr);
    for (int i = 0; i <= h->N; ++i) {
        double* r = rows.data() + (size_t)i * rr;
        if (on_so3(h->kind)) {
            const double* q = h_q_ref + (size_t)i * 4;
            const double nq = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
            if (!(nq > 0.0)) return fail(TRAJOPT_E_INVALID, "trajopt_set_reference: zero quaternion");
            for (int j = 0; j < 4; ++j) r[j] = q[j] / nq;
            for (int j = 0; j < 3; ++j) r[4 + j] = h_xi_ref[(size_t)i * 3 + j];
            host_quat_to_rot(r, r + 7);
        } else {
            const double* q = h_q_ref + (size_t)i * 7;
            const double nq = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
            if (!(nq > 0.0)) return fail(TRAJOPT_E_INVALID, "trajopt_set_reference: zero quaternion");
            for (int j = 0; j < 4; ++j) r[j] = q[j] / nq;
            for (int j = 0; j < 3; ++j) r[4 + j] = q[4 + j];
            for (int j = 0; j < 6; ++j) r[7 + j] = h_xi_ref[(size_t)i * 6 + j];
            double* R = r + 13;
            host_quat_to_rot(r, R);
            const double* p = r + 4;   // [p]x R
            for (int j = 0; j < 3; ++j) {
                r[22 + j] = p[1] * R[6 + j] - p[2] * R[3 + j];
                r[25 + j] = p[2] * R[j] - p[0] * R[6 + j];
                r[28 + j] = p[0] * R[3 + j] - p[1] * R[j];
            }
        }
    }
    CUDA_OK(cudaMemcpy(h->d_ref, rows.data(), rows.size() * sizeof(double), cudaMemcpyHostToDevice));
    h->have_ref = true;
    h->begun = false;
    return 0;
}

int trajopt_begin(trajopt_handle* h, const double* d_x0, const double* d_us_init, int us_mode, void* stream) {
    if (!h || !d_x0) return fail(TRAJOPT_E_INVALID, "trajopt_begin: NULL argument");
    if (!h->have_params || !h->have_ref) return fail(TRAJOPT_E_STATE, "trajopt_begin: set_params and set_reference first");
    if (us_mode < 0 || us_mode > 2) return fail(TRAJOPT_E_INVALID, "trajopt_begin: us_mode must be 0, 1 or 2");
    if (us_mode != 0 && !d_us_init) return fail(TRAJOPT_E_INVALID, "trajopt_begin: us_mode != 0 needs d_us_init");
    DeviceGuard guard(h->device);
    return DISPATCH_KIND(h, begin_impl, h, d_x0, d_us_init, us_mode, (cudaStream_t)stream);
}

int trajopt_iterate(trajopt_handle* h, int n_iters, int* n_active_out, void* stream) {
    if (!h) return fail(TRAJOPT_E_INVALID, "trajopt_iterate: NULL handle");
    if (!h->begun) return fail(TRAJOPT_E_STATE, "trajopt_iterate: call trajopt_begin first");
    if (n_iters < 0) return fail(TRAJOPT_E_INVALID, "trajopt_iterate: n_iters < 0");
    DeviceGuard guard(h->device);
    return DISPATCH_KIND(h, iterate_impl, h, n_iters, n_active_out, (cudaStream_t)stream);
}

int trajopt_export(trajopt_handle* h, double* d_xs, double* d_us, double* d_J, int32_t* d_iters, int32_t* d_status,
                   double* d_grad, double* d_defect, void* stream) {
    if (!h) return fail(TRAJOPT_E_INVALID, "trajopt_export: NULL handle");
    if (!h->begun) return fail(TRAJOPT_E_STATE, "trajopt_export: nothing solved yet");
    DeviceGuard guard(h->device);
    cudaStream_t s = (cudaStream_t)stream;
    const Work& w = h->w;
    const int bg = blocks_for(h->B, 128);
    if (d_xs) LAUNCH(k_export_traj, dim3(bg, h->N + 1), 128, 0, s, h->B, h->Bp, h->NS, w.X[0], w.X[1], w.sel, d_xs, h->N + 1);
    if (d_us) LAUNCH(k_export_traj, dim3(bg, h->N), 128, 0, s, h->B, h->Bp, h->NU, w.U[0], w.U[1], w.sel, d_us, h->N);
    if (d_J) LAUNCH(k_export_rows<double>, dim3(bg, 1), 128, 0, s, h->B, h->Bp, 1, w.J, d_J);
    if (d_grad) LAUNCH(k_export_rows<double>, dim3(bg, 1), 128, 0, s, h->B, h->Bp, 1, w.grad, d_grad);
    if (d_defect) LAUNCH(k_export_rows<double>, dim3(bg, 1), 128, 0, s, h->B, h->Bp, 1, w.dnorm, d_defect);
    if (d_iters) LAUNCH(k_export_rows<int>, dim3(bg, 1), 128, 0, s, h->B, h->Bp, 1, w.iters, (int*)d_iters);
    if (d_status) LAUNCH(k_export_rows<int>, dim3(bg, 1), 128, 0, s, h->B, h->Bp, 1, w.status, (int*)d_status);
    return 0;
}

int trajopt_export_hist(trajopt_handle* h, double* d_J_hist, double* d_grad_hist, double* d_defect_hist,
                        int32_t* d_alpha_hist, void* stream) {
    if (!h) return fail(TRAJOPT_E_INVALID, "trajopt_export_hist: NULL handle");
    if (!h->begun) return fail(TRAJOPT_E_STATE, "trajopt_export_hist: nothing solved yet");
    DeviceGuard guard(h->device);
    cudaStream_t s = (cudaStream_t)stream;
    const Work& w = h->w;
    const int bg = blocks_for(h->B, 128), mi = h->prm.max_iters;
    if (mi == 0) return 0;
    if (d_J_hist) LAUNCH(k_export_rows<double>, dim3(bg, mi), 128, 0, s, h->B, h->Bp, mi, w.Jhist, d_J_hist);
    if (d_grad_hist) LAUNCH(k_export_rows<double>, dim3(bg, mi + 1), 128, 0, s, h->B, h->Bp, mi + 1, w.gradhist, d_grad_hist);
    if (d_defect_hist) LAUNCH(k_export_rows<double>, dim3(bg, mi + 1), 128, 0, s, h->B, h->Bp, mi + 1, w.defhist, d_defect_hist);
    if (d_alpha_hist) LAUNCH(k_export_rows<int>, dim3(bg, mi), 128, 0, s, h->B, h->Bp, mi, w.alphahist, (int*)d_alpha_hist);
    return 0;
}

int trajopt_export_reg(trajopt_handle* h, double* d_mu, double* d_delta, void* stream) {
    if (!h) return fail(TRAJOPT_E_INVALID, "trajopt_export_reg: NULL handle");
    if (!h->begun) return fail(TRAJOPT_E_STATE, "trajopt_export_reg: nothing solved yet");
    DeviceGuard guard(h->device);
    cudaStream_t s = (cudaStream_t)stream;
    const int bg = blocks_for(h->B, 128);
    if (d_mu) LAUNCH(k_export_rows<double>, dim3(bg, 1), 128, 0, s, h->B, h->Bp, 1, h->w.mu, d_mu);
    if (d_delta) LAUNCH(k_export_rows<double>, dim3(bg, 1), 128, 0, s, h->B, h->Bp, 1, h->w.delta, d_delta);
    return 0;
}

int trajopt_export_al(trajopt_handle* h, double* d_lmbd, double* d_imu, double* d_mu, int32_t* d_outer_iters,
                      double* d_violation, void* stream) {
    if (!h) return fail(TRAJOPT_E_INVALID, "trajopt_export_al: NULL handle");
    if (h->method != TRAJOPT_AL_MS) return fail(TRAJOPT_E_STATE, "trajopt_export_al: not an augmented-Lagrangian handle");
    if (!h->begun) return fail(TRAJOPT_E_STATE, "trajopt_export_al: nothing solved yet");
    DeviceGuard guard(h->device);
    cudaStream_t s = (cudaStream_t)stream;
    const Work& w = h->w;
    const int bg = blocks_for(h->B, 128);
    if (d_lmbd) LAUNCH(k_export_traj, dim3(bg, h->N + 1), 128, 0, s, h->B, h->Bp, 2 * h->NU, w.lam, w.lam, (const int*)nullptr, d_lmbd, h->N + 1);
    if (d_imu) LAUNCH(k_export_traj, dim3(bg, h->N + 1), 128, 0, s, h->B, h->Bp, 2 * h->NU, w.imu, w.imu, (const int*)nullptr, d_imu, h->N + 1);
    if (d_mu) LAUNCH(k_export_rows<double>, dim3(bg, 1), 128, 0, s, h->B, h->Bp, 1, w.al_mu, d_mu);
    if (d_violation) LAUNCH(k_export_rows<double>, dim3(bg, 1), 128, 0, s, h->B, h->Bp, 1, w.al_viol, d_violation);
    if (d_outer_iters) LAUNCH(k_export_rows<int>, dim3(bg, 1), 128, 0, s, h->B, h->Bp, 1, w.al_outer, (int*)d_outer_iters);
    return 0;
}

int trajopt_solve(trajopt_handle* h, const double* d_x0, const double* d_us_init, int us_mode, double* d_xs, double* d_us,
                  double* d_J, int32_t* d_iters, int32_t* d_status, double* d_grad, double* d_defect, void* stream) {
    int rc = trajopt_begin(h, d_x0, d_us_init, us_mode, stream);
    if (rc) return rc;
    const int units = (h->method == TRAJOPT_AL_MS) ? h->user.n_al_iters : h->prm.max_iters + 1;
    int active = 0;
    rc = trajopt_iterate(h, units, &active, stream);
    if (rc) return rc;
    return trajopt_export(h, d_xs, d_us, d_J, d_iters, d_status, d_grad, d_defect, stream);
}

int trajopt_solve_host(trajopt_handle* h, const double* h_x0, const double* h_us_init, int us_mode, double* h_xs,
                       double* h_us, double* h_J, int32_t* h_iters, int32_t* h_status, double* h_grad, double* h_defect,
                       void* stream) {
    if (!h || !h_x0) return fail(TRAJOPT_E_INVALID, "trajopt_solve_host: NULL argument");
    if (us_mode < 0 || us_mode > 2) return fail(TRAJOPT_E_INVALID, "trajopt_solve_host: us_mode must be 0, 1 or 2");
    if (us_mode != 0 && !h_us_init) return fail(TRAJOPT_E_INVALID, "trajopt_solve_host: us_mode != 0 needs h_us_init");
    DeviceGuard guard(h->device);
    cudaStream_t s = (cudaStream_t)stream;
    const size_t B = h->B, N = h->N;
    if (!h->s_x0) {
        CUDA_OK(cudaMalloc((void**)&h->s_x0, B * h->NS * 8));
        CUDA_OK(cudaMalloc((void**)&h->s_J, B * 8));
        CUDA_OK(cudaMalloc((void**)&h->s_grad, B * 8));
        CUDA_OK(cudaMalloc((void**)&h->s_def, B * 8));
        CUDA_OK(cudaMalloc((void**)&h->s_iters, B * 4));
        CUDA_OK(cudaMalloc((void**)&h->s_status, B * 4));
    }
    if (h_xs && !h->s_xs) CUDA_OK(cudaMalloc((void**)&h->s_xs, B * (N + 1) * h->NS * 8));
    if (h_us && !h->s_us) CUDA_OK(cudaMalloc((void**)&h->s_us, B * N * h->NU * 8));
    const size_t us0_bytes = (us_mode == 0) ? 0 : (us_mode == 1 ? N * h->NU * 8 : B * N * h->NU * 8);
    if (us0_bytes > h->s_us0_bytes) {
        if (h->s_us0) cudaFree(h->s_us0);
        h->s_us0 = nullptr;
        CUDA_OK(cudaMalloc((void**)&h->s_us0, us0_bytes));
        h->s_us0_bytes = us0_bytes;
    }
    CUDA_OK(cudaMemcpyAsync(h->s_x0, h_x0, B * h->NS * 8, cudaMemcpyHostToDevice, s));
    if (us0_bytes) CUDA_OK(cudaMemcpyAsync(h->s_us0, h_us_init, us0_bytes, cudaMemcpyHostToDevice, s));
    int rc = trajopt_solve(h, h->s_x0, us0_bytes ? h->s_us0 : nullptr, us_mode, h_xs ? h->s_xs : nullptr,
                           h_us ? h->s_us : nullptr, h->s_J, h->s_iters, h->s_status, h->s_grad, h->s_def, stream);
    if (rc) return rc;
    if (h_xs) CUDA_OK(cudaMemcpyAsync(h_xs, h->s_xs, B * (N + 1) * h->NS * 8, cudaMemcpyDeviceToHost, s));
    if (h_us) CUDA_OK(cudaMemcpyAsync(h_us, h->s_us, B * N * h->NU * 8, cudaMemcpyDeviceToHost, s));
    if (h_J) CUDA_OK(cudaMemcpyAsync(h_J, h->s_J, B * 8, cudaMemcpyDeviceToHost, s));
    if (h_grad) CUDA_OK(cudaMemcpyAsync(h_grad, h->s_grad, B * 8, cudaMemcpyDeviceToHost, s));
    if (h_defect) CUDA_OK(cudaMemcpyAsync(h_defect, h->s_def, B * 8, cudaMemcpyDeviceToHost, s));
    if (h_iters) CUDA_OK(cudaMemcpyAsync(h_iters, h->s_iters, B * 4, cudaMemcpyDeviceToHost, s));
    if (h_status) CUDA_OK(cudaMemcpyAsync(h_status, h->s_status, B * 4, cudaMemcpyDeviceToHost, s));
    CUDA_OK(cudaStreamSynchronize(s));
    return 0;
}

int trajopt_debug_linearize(trajopt_handle* h, double* d_Fx, double* d_Fu, double* d_defect, double* d_L, double* d_Lx,
                            double* d_Lxx, double* d_Lu, void* stream) {
    if (!h) return fail(TRAJOPT_E_INVALID, "trajopt_debug_linearize: NULL handle");
    if (!h->begun) return fail(TRAJOPT_E_STATE, "trajopt_debug_linearize: call trajopt_begin first");
    DeviceGuard guard(h->device);
    return DISPATCH_KIND(h, debug_linearize_impl, h, d_Fx, d_Fu, d_defect, d_L, d_Lx, d_Lxx, d_Lu, (cudaStream_t)stream);
}

int trajopt_debug_gains(trajopt_handle* h, double* d_k, double* d_K, void* stream) {
    if (!h) return fail(TRAJOPT_E_INVALID, "trajopt_debug_gains: NULL handle");
    if (!h->begun) return fail(TRAJOPT_E_STATE, "trajopt_debug_gains: call trajopt_begin first");
    DeviceGuard guard(h->device);
    cudaStream_t s = (cudaStream_t)stream;
    const int bg = blocks_for(h->B, 128);
    if (d_k) LAUNCH(k_export_traj, dim3(bg, h->N), 128, 0, s, h->B, h->Bp, h->NU, h->w.kff, h->w.kff, (const int*)nullptr, d_k, h->N);
    if (d_K) LAUNCH(k_export_traj, dim3(bg, h->N), 128, 0, s, h->B, h->Bp, h->NU * h->NX, h->w.Kfb, h->w.Kfb, (const int*)nullptr, d_K, h->N);
    return 0;
}

int trajopt_debug_linesearch_rows(trajopt_handle* h) {
    if (!h || !h->have_params) return fail(TRAJOPT_E_INVALID, "trajopt_debug_linesearch_rows: handle without parameters");
    return h->cand_rows;
}

int trajopt_debug_linesearch(trajopt_handle* h, double* d_table, void* stream) {
    if (!h || !d_table) return fail(TRAJOPT_E_INVALID, "trajopt_debug_linesearch: NULL argument");
    if (!h->begun) return fail(TRAJOPT_E_STATE, "trajopt_debug_linesearch: call trajopt_begin first");
    DeviceGuard guard(h->device);
    // source is [rows][Bp]; the export kernel writes out[b][r] for a [B][rows] target, so transpose by swapping roles:
    // k_export_rows(B, Bp, rows, src, out) gives out[b * rows + r]; the table is documented as [rows][B], hence a 2D copy
    CUDA_OK(cudaMemcpy2DAsync(d_table, (size_t)h->B * sizeof(double), h->w.Jcand, (size_t)h->Bp * sizeof(double),
                              (size_t)h->B * sizeof(double), (size_t)h->cand_rows, cudaMemcpyDeviceToDevice,
                              (cudaStream_t)stream));
    return 0;
}

int trajopt_debug_stage(trajopt_handle* h, int i, int terminal, int n, const double* d_x, const double* d_u, double* d_f,
                        double* d_Fx, double* d_Fu, double* d_l, double* d_lx, double* d_lxx, double* d_lu, double* d_err,
                        void* stream) {
    if (!h || !d_x) return fail(TRAJOPT_E_INVALID, "trajopt_debug_stage: NULL argument");
    if (!h->have_params || !h->have_ref) return fail(TRAJOPT_E_STATE, "trajopt_debug_stage: set_params and set_reference first");
    if (i < 0 || i > h->N || n < 0) return fail(TRAJOPT_E_INVALID, "trajopt_debug_stage: stage index out of range");
    if (!terminal && !d_u) return fail(TRAJOPT_E_INVALID, "trajopt_debug_stage: d_u is NULL for a non-terminal stage");
    if (n == 0) return 0;
    DeviceGuard guard(h->device);
    return DISPATCH_KIND(h, debug_stage_impl, h, i, terminal, n, d_x, d_u, d_f, d_Fx, d_Fu, d_l, d_lx, d_lxx, d_lu, d_err,
                         (cudaStream_t)stream);
}

int trajopt_debug_lie(int op, int n, const double* d_in, double* d_out, void* stream) {
    if (op < 0 || op >= LIE_OP_COUNT) return fail(TRAJOPT_E_INVALID, "trajopt_debug_lie: unknown op");
    if (n == 0) return 0;
    if (n < 0 || !d_in || !d_out) return fail(TRAJOPT_E_INVALID, "trajopt_debug_lie: bad argument");
    LAUNCH(k_debug_lie, blocks_for(n, 128), 128, 0, (cudaStream_t)stream, op, n, d_in, d_out);
    return 0;
}

int trajopt_debug_fp64_peak(double ms_target, double* out_tflops, void* stream) {
    if (!out_tflops) return fail(TRAJOPT_E_INVALID, "trajopt_debug_fp64_peak: NULL argument");
    cudaStream_t s = (cudaStream_t)stream;
    int dev = 0, sms = 0;
    CUDA_OK(cudaGetDevice(&dev));
    CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int blocks = sms * 8, threads = 256;
    double* sink = nullptr;
    CUDA_OK(cudaMalloc((void**)&sink, (size_t)blocks * threads * sizeof(double)));
    cudaEvent_t e0, e1;
    CUDA_OK(cudaEventCreate(&e0));
    CUDA_OK(cudaEventCreate(&e1));
    int iters = 2000;
    double best = 0.0;
    for (int rep = 0; rep < 6; ++rep) {
        CUDA_OK(cudaEventRecord(e0, s));
        LAUNCH(k_fp64_peak, blocks, threads, 0, s, iters, 1.0, sink);
        CUDA_OK(cudaEventRecord(e1, s));
        CUDA_OK(cudaEventSynchronize(e1));
        float ms = 0.f;
        CUDA_OK(cudaEventElapsedTime(&ms, e0, e1));
        const double flop = 2.0 * 64.0 * (double)iters * (double)blocks * threads;
        const double tf = flop / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;
        if (rep == 0 && ms > 0.f && ms_target > 0.0) {   // rescale the loop length to the requested duration
            const double scale = ms_target / ms;
            iters = (int)fmin(2.0e6, fmax(500.0, iters * scale));
        }
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(sink);
    *out_tflops = best;
    return 0;
}

int trajopt_phase_times(trajopt_handle* h, double* out_ms, int64_t* cnt, int reset) {
    if (!h) return fail(TRAJOPT_E_INVALID, "trajopt_phase_times: NULL handle");
    for (int i = 0; i < PH_COUNT; ++i) {
        if (out_ms) out_ms[i] = h->phase_ms[i];
        if (cnt) cnt[i] = h->phase_cnt[i];
        if (reset) { h->phase_ms[i] = 0.0; h->phase_cnt[i] = 0; }
    }
    return 0;
}

int trajopt_set_profiling(trajopt_handle* h, int enable) {
    if (!h) return fail(TRAJOPT_E_INVALID, "trajopt_set_profiling: NULL handle");
    h->profiling = enable != 0;
    return 0;
}

}  // extern "C"
