// C ABI of the batched DDP/iLQR solver (see include/trajopt_b200.h for the contract and the
// reference call each entry point stands for).  Host side only orchestrates kernel launches on
// one stream; every number is produced by the kernels.  The kernel sequences of each problem family
// are instantiated in kind_<family>.cu (host_impl.cuh), this file only dispatches.
#include "host_impl.cuh"

TRAJOPT_KIND_EXTERN(TRAJOPT_SO3)
TRAJOPT_KIND_EXTERN(TRAJOPT_SE3)
TRAJOPT_KIND_EXTERN(TRAJOPT_DRONE)
TRAJOPT_KIND_EXTERN(TRAJOPT_RIGID)
TRAJOPT_KIND_EXTERN(TRAJOPT_PEND)

using namespace trajopt_host;

// grid of k_export_traj for rows of `len` doubles
static inline dim3 export_grid(const trajopt_handle* h, int len) {
    return dim3((unsigned)((h->Bp + 3) / 4), (unsigned)((len + kExportChunk - 1) / kExportChunk));
}

// ------------------------------------------------------------------------------------------
// error reporting / launch accounting
// ------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
std::atomic<long long> g_trajopt_launches{0};

extern "C" int trajopt_set_error_(cudaError_t e, const char* file, int line) {
    snprintf(g_err, sizeof(g_err), "CUDA error %d (%s) at %s:%d", (int)e, cudaGetErrorString(e), file, line);
    return TRAJOPT_E_CUDA;
}
int trajopt_fail_(int code, const char* msg) {
    snprintf(g_err, sizeof(g_err), "%s", msg);
    return code;
}

#define DISPATCH_KIND(h, fn, ...)                                             \
    ((h)->kind == TRAJOPT_SO3   ? fn<TRAJOPT_SO3>(__VA_ARGS__)                \
     : (h)->kind == TRAJOPT_SE3 ? fn<TRAJOPT_SE3>(__VA_ARGS__)                \
     : (h)->kind == TRAJOPT_DRONE ? fn<TRAJOPT_DRONE>(__VA_ARGS__)            \
     : (h)->kind == TRAJOPT_RIGID ? fn<TRAJOPT_RIGID>(__VA_ARGS__)            \
                                  : fn<TRAJOPT_PEND>(__VA_ARGS__))


// ------------------------------------------------------------------------------------------
// exported entry points
// ------------------------------------------------------------------------------------------
extern "C" {

const char* trajopt_last_error(void) { return g_err; }
int trajopt_version(void) { return 100; }

int64_t trajopt_launch_count(int reset) {
    const long long v = g_trajopt_launches.load();
    if (reset) g_trajopt_launches.store(0);
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
    if (!guard.ok) return fail(TRAJOPT_E_CUDA, "trajopt_create: cannot select the CUDA device");
    trajopt_handle* h = new (std::nothrow) trajopt_handle();
    if (!h) return fail(TRAJOPT_E_INVALID, "trajopt_create: out of host memory");
    h->kind = kind; h->method = method; h->N = N; h->B = B; h->device = device;
    if (cudaDeviceGetAttribute(&h->sms, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || h->sms < 1) h->sms = 148;
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
    A_(w.gains, (size_t)N * (h->NU * h->NX + h->NU) * Bp);
    A_(w.J, Bp); A_(w.grad, Bp); A_(w.dnorm, Bp); A_(w.mu, Bp); A_(w.delta, Bp);
    A_(w.iters, Bp); A_(w.status, Bp); A_(w.ls_state, Bp);
    A_(w.x0, (size_t)h->NS * Bp);
    A_(w.counters, 4);
    A_(w.orig, Bp);
    A_(w.Nb, Bp);
    A_(w.slot_id, Bp); A_(w.fresh, Bp); A_(w.free_list, Bp); A_(w.done_list, Bp); A_(w.scnt, 4);
    A_(h->d_dweight, Bp);
    if (method == TRAJOPT_AL_MS) {
        A_(w.lam, Np1 * 2 * h->NU * Bp); A_(w.imu, Np1 * 2 * h->NU * Bp);
        A_(w.al_mu, Bp); A_(w.al_outer, Bp); A_(w.al_viol, Bp); A_(w.al_done, Bp);
    }
#undef A_
    w.ref = h->d_ref;
    w.ref_batch = nullptr;
    if (!rc && cudaHostAlloc((void**)&h->h_counters, 4 * sizeof(int), cudaHostAllocMapped) != cudaSuccess) rc = fail(TRAJOPT_E_CUDA, "cudaHostAlloc failed");
    if (!rc && cudaHostGetDevicePointer((void**)&h->h_counters_dev, h->h_counters, 0) != cudaSuccess) rc = fail(TRAJOPT_E_CUDA, "cudaHostGetDevicePointer failed");
    if (!rc && cudaHostAlloc((void**)&h->h_ints, (size_t)Bp * sizeof(int), cudaHostAllocMapped) != cudaSuccess) rc = fail(TRAJOPT_E_CUDA, "cudaHostAlloc failed");
    if (!rc && cudaHostGetDevicePointer((void**)&h->h_ints_dev, h->h_ints, 0) != cudaSuccess) rc = fail(TRAJOPT_E_CUDA, "cudaHostGetDevicePointer failed");
    if (!rc && (cudaEventCreate(&h->ev[0]) != cudaSuccess || cudaEventCreate(&h->ev[1]) != cudaSuccess))
        rc = fail(TRAJOPT_E_CUDA, "cudaEventCreate failed");
    if (!rc && method != TRAJOPT_SS) {
        int pr_least = 0, pr_greatest = 0;
        cudaDeviceGetStreamPriorityRange(&pr_least, &pr_greatest);
        if (cudaStreamCreateWithPriority(&h->s2, cudaStreamNonBlocking, pr_greatest) != cudaSuccess) rc = fail(TRAJOPT_E_CUDA, "cudaStreamCreate failed");
        for (int c = 0; c < kMaxChunks && !rc; ++c)
            if (cudaEventCreateWithFlags(&h->ev_chunk[c], cudaEventDisableTiming) != cudaSuccess) rc = fail(TRAJOPT_E_CUDA, "cudaEventCreate failed");
        if (!rc && cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming) != cudaSuccess) rc = fail(TRAJOPT_E_CUDA, "cudaEventCreate failed");
    }
    if (rc) {
        trajopt_destroy(h);
        return rc;
    }
    rc = [&]() -> int {   // a failure here must not leak the handle
        LAUNCH(k_set_horizons, blocks_for(h->Bp, 128), 128, 0, (cudaStream_t)0, h->B, h->Bp, h->N, (const int*)nullptr, h->w.Nb);
        CUDA_OK(cudaDeviceSynchronize());
        return 0;
    }();
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
    for (void* p : h->cand_allocs) if (p) cudaFree(p);
    void* stage[] = {h->s_x0, h->s_us0, h->s_xs, h->s_us, h->s_J, h->s_grad, h->s_def, h->s_iters, h->s_status, h->s_snap, h->d_late};
    for (void* p : stage) if (p) cudaFree(p);
    if (h->d_perm) cudaFree(h->d_perm);
    if (h->d_ref_long) cudaFree(h->d_ref_long);
    if (h->d_scratch) cudaFree(h->d_scratch);
    if (h->h_counters) cudaFreeHost(h->h_counters);
    if (h->h_ints) cudaFreeHost(h->h_ints);
    if (h->ev[0]) cudaEventDestroy(h->ev[0]);
    if (h->ev[1]) cudaEventDestroy(h->ev[1]);
    for (cudaEvent_t e : h->ev_chunk) if (e) cudaEventDestroy(e);
    if (h->ev_join) cudaEventDestroy(h->ev_join);
    if (h->ev_block) cudaEventDestroy(h->ev_block);
    if (h->s2) cudaStreamDestroy(h->s2);
    if (h->s_copy) cudaStreamDestroy(h->s_copy);
    for (cudaEvent_t e : h->ev_host) if (e) cudaEventDestroy(e);
    for (cudaEvent_t e : h->ev_done) if (e) cudaEventDestroy(e);
    for (void* p : h->so_buf) if (p) cudaFree(p);
    if (h->so_x0) cudaFree(h->so_x0);
    delete h;
    return 0;
}

int trajopt_set_params(trajopt_handle* h, const trajopt_params* p) {
    if (!h || !p) return fail(TRAJOPT_E_INVALID, "trajopt_set_params: NULL argument");
    if (!(p->dt > 0.0) || !(p->mass > 0.0)) return fail(TRAJOPT_E_INVALID, "trajopt_set_params: dt and mass must be positive");
    if (p->max_iters < 0 || p->max_iters > 65534)   // the history exports use one grid row per iteration
        return fail(TRAJOPT_E_INVALID, "trajopt_set_params: need 0 <= max_iters <= 65534");
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
    for (int j = 0; j < 6; ++j) { q.lb[j] = p->lb[j]; q.ub[j] = p->ub[j]; q.xlb[j] = p->xi_lb[j]; q.xub[j] = p->xi_ub[j]; }
    q.has_state_bounds = (h->method == TRAJOPT_AL_MS && p->has_state_bounds) ? 1 : 0;
    if (q.has_state_bounds && !h->w.lam_s) {
        const size_t Np1 = (size_t)h->N + 1, nv = (size_t)(h->NX - h->NP), Bp = (size_t)h->Bp;
        int rc2 = dalloc(h, &h->w.lam_s, Np1 * 2 * nv * Bp);
        if (!rc2) rc2 = dalloc(h, &h->w.imu_s, Np1 * 2 * nv * Bp);
        if (!rc2) rc2 = dalloc(h, &h->w.lxxv, Np1 * nv * Bp);
        if (rc2) return rc2;
    }
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
    if ((rc = ensure_cand(h))) return rc;
    h->have_params = true;
    h->begun = false;
    return 0;
}

// pack n_rows reference samples into the rows the cost reads (pose, twist, rotation matrix, [p]x R)
static int pack_reference_rows(const trajopt_handle* h, const double* h_q_ref, const double* h_xi_ref, size_t n_rows, std::vector<double>& rows) {
    const int rr = h->REFROW;
    rows.assign(n_rows * rr, 0.0);
    for (size_t i = 0; i < n_rows; ++i) {
        double* r = rows.data() + i * rr;
        if (on_so3(h->kind)) {
            const double* q = h_q_ref + i * 4;
            const double nq = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
            if (!(nq > 0.0)) return fail(TRAJOPT_E_INVALID, "trajopt_set_reference: zero quaternion");
            for (int j = 0; j < 4; ++j) r[j] = q[j] / nq;
            for (int j = 0; j < 3; ++j) r[4 + j] = h_xi_ref[i * 3 + j];
            host_quat_to_rot(r, r + 7);
        } else {
            const double* q = h_q_ref + i * 7;
            const double nq = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
            if (!(nq > 0.0)) return fail(TRAJOPT_E_INVALID, "trajopt_set_reference: zero quaternion");
            for (int j = 0; j < 4; ++j) r[j] = q[j] / nq;
            for (int j = 0; j < 3; ++j) r[4 + j] = q[4 + j];
            for (int j = 0; j < 6; ++j) r[7 + j] = h_xi_ref[i * 6 + j];
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
    return 0;
}

int trajopt_set_reference(trajopt_handle* h, const double* h_q_ref, const double* h_xi_ref) {
    if (!h || !h_q_ref || !h_xi_ref) return fail(TRAJOPT_E_INVALID, "trajopt_set_reference: NULL argument");
    DeviceGuard guard(h->device);
    std::vector<double> rows;
    int rc = pack_reference_rows(h, h_q_ref, h_xi_ref, (size_t)h->N + 1, rows);
    if (rc) return rc;
    CUDA_OK(cudaMemcpy(h->d_ref, rows.data(), rows.size() * sizeof(double), cudaMemcpyHostToDevice));
    h->w.ref = h->d_ref;
    h->ref_long_rows = 0;
    h->w.ref_batch = nullptr;      // back to the shared reference
    h->have_ref = true;
    h->begun = false;
    return 0;
}

int trajopt_set_reference_long(trajopt_handle* h, const double* h_q_ref, const double* h_xi_ref, int64_t n_rows) {
    if (!h || !h_q_ref || !h_xi_ref) return fail(TRAJOPT_E_INVALID, "trajopt_set_reference_long: NULL argument");
    if (n_rows < (int64_t)h->N + 1) return fail(TRAJOPT_E_INVALID, "trajopt_set_reference_long: need at least N + 1 samples");
    DeviceGuard guard(h->device);
    std::vector<double> rows;
    int rc = pack_reference_rows(h, h_q_ref, h_xi_ref, (size_t)n_rows, rows);
    if (rc) return rc;
    if (h->d_ref_long) cudaFree(h->d_ref_long);
    h->d_ref_long = nullptr;
    CUDA_OK(cudaMalloc((void**)&h->d_ref_long, rows.size() * sizeof(double)));
    CUDA_OK(cudaMemcpy(h->d_ref_long, rows.data(), rows.size() * sizeof(double), cudaMemcpyHostToDevice));
    h->ref_long_rows = n_rows;
    h->w.ref = h->d_ref_long;
    h->w.ref_batch = nullptr;
    h->have_ref = true;
    h->begun = false;
    return 0;
}

int trajopt_set_reference_offset(trajopt_handle* h, int64_t first_row) {
    if (!h) return fail(TRAJOPT_E_INVALID, "trajopt_set_reference_offset: NULL handle");
    if (!h->ref_long_rows) return fail(TRAJOPT_E_STATE, "trajopt_set_reference_offset: call trajopt_set_reference_long first");
    if (first_row < 0 || first_row + h->N + 1 > h->ref_long_rows) return fail(TRAJOPT_E_INVALID, "trajopt_set_reference_offset: window out of range");
    h->w.ref = h->d_ref_long + (size_t)first_row * h->REFROW;
    h->w.ref_batch = nullptr;
    h->begun = false;
    return 0;
}

int trajopt_set_reference_batch(trajopt_handle* h, const double* d_q_ref, const double* d_xi_ref, void* stream) {
    if (!h || !d_q_ref || !d_xi_ref) return fail(TRAJOPT_E_INVALID, "trajopt_set_reference_batch: NULL argument");
    DeviceGuard guard(h->device);
    if (!h->d_ref_batch) {
        void* q = nullptr;
        CUDA_OK(cudaMalloc(&q, ((size_t)h->N + 1) * h->REFROW * (size_t)h->Bp * sizeof(double)));
        h->allocs.push_back(q);
        h->d_ref_batch = (double*)q;
    }
    int rc = DISPATCH_KIND(h, set_reference_batch_impl, h, d_q_ref, d_xi_ref, (cudaStream_t)stream);
    if (rc) return rc;
    h->w.ref_batch = h->d_ref_batch;
    h->ref_permuted = false;       // freshly written in the caller's order
    h->have_ref = true;
    h->begun = false;
    return 0;
}

int trajopt_set_horizons(trajopt_handle* h, const int32_t* d_N, void* stream) {
    if (!h) return fail(TRAJOPT_E_INVALID, "trajopt_set_horizons: NULL handle");
    DeviceGuard guard(h->device);
    LAUNCH(k_set_horizons, blocks_for(h->Bp, 128), 128, 0, (cudaStream_t)stream, h->B, h->Bp, h->N, (const int*)d_N, h->w.Nb);
    h->var_horizons = d_N != nullptr;
    h->nb_permuted = false;        // freshly written in the caller's order
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

int trajopt_iterate_inner(trajopt_handle* h, int n_iters, int* n_active_out, void* stream) {
    if (!h) return fail(TRAJOPT_E_INVALID, "trajopt_iterate_inner: NULL handle");
    if (!h->begun) return fail(TRAJOPT_E_STATE, "trajopt_iterate_inner: call trajopt_begin first");
    if (n_iters < 0) return fail(TRAJOPT_E_INVALID, "trajopt_iterate_inner: n_iters < 0");
    DeviceGuard guard(h->device);
    return DISPATCH_KIND(h, iterate_inner_impl, h, n_iters, n_active_out, (cudaStream_t)stream);
}

int trajopt_export(trajopt_handle* h, double* d_xs, double* d_us, double* d_J, int32_t* d_iters, int32_t* d_status,
                   double* d_grad, double* d_defect, void* stream) {
    if (!h) return fail(TRAJOPT_E_INVALID, "trajopt_export: NULL handle");
    if (!h->begun) return fail(TRAJOPT_E_STATE, "trajopt_export: nothing solved yet");
    DeviceGuard guard(h->device);
    cudaStream_t s = (cudaStream_t)stream;
    const Work& w = h->w;
    const int bg = blocks_for(h->Bp, 128);
    if (d_xs) LAUNCH(k_export_traj, export_grid(h, (h->N + 1) * (h->NS)), 128, 0, s, h->B, h->Bp, h->NS, w.X[0], w.X[1], w.sel, (const int*)h->w.orig, d_xs, h->N + 1, h->var_horizons ? (const int*)w.Nb : (const int*)nullptr, 0);
    if (d_us) LAUNCH(k_export_traj, export_grid(h, (h->N) * (h->NU)), 128, 0, s, h->B, h->Bp, h->NU, w.U[0], w.U[1], w.sel, (const int*)h->w.orig, d_us, h->N, h->var_horizons ? (const int*)w.Nb : (const int*)nullptr, 1);
    if (d_J) LAUNCH(k_export_rows<double>, dim3(bg, 1), 128, 0, s, h->B, h->Bp, 1, w.J, (const int*)h->w.orig, d_J);
    if (d_grad) LAUNCH(k_export_rows<double>, dim3(bg, 1), 128, 0, s, h->B, h->Bp, 1, w.grad, (const int*)h->w.orig, d_grad);
    if (d_defect) LAUNCH(k_export_rows<double>, dim3(bg, 1), 128, 0, s, h->B, h->Bp, 1, w.dnorm, (const int*)h->w.orig, d_defect);
    if (d_iters) LAUNCH(k_export_rows<int>, dim3(bg, 1), 128, 0, s, h->B, h->Bp, 1, w.iters, (const int*)h->w.orig, (int*)d_iters);
    if (d_status) LAUNCH(k_export_rows<int>, dim3(bg, 1), 128, 0, s, h->B, h->Bp, 1, w.status, (const int*)h->w.orig, (int*)d_status);
    return 0;
}

int trajopt_export_hist(trajopt_handle* h, double* d_J_hist, double* d_grad_hist, double* d_defect_hist,
                        int32_t* d_alpha_hist, void* stream) {
    if (!h) return fail(TRAJOPT_E_INVALID, "trajopt_export_hist: NULL handle");
    if (!h->begun) return fail(TRAJOPT_E_STATE, "trajopt_export_hist: nothing solved yet");
    DeviceGuard guard(h->device);
    cudaStream_t s = (cudaStream_t)stream;
    const Work& w = h->w;
    const int bg = blocks_for(h->Bp, 128), mi = h->prm.max_iters;
    if (mi == 0) return 0;
    if (d_J_hist) LAUNCH(k_export_rows<double>, dim3(bg, mi), 128, 0, s, h->B, h->Bp, mi, w.Jhist, (const int*)h->w.orig, d_J_hist);
    if (d_grad_hist) LAUNCH(k_export_rows<double>, dim3(bg, mi + 1), 128, 0, s, h->B, h->Bp, mi + 1, w.gradhist, (const int*)h->w.orig, d_grad_hist);
    if (d_defect_hist) LAUNCH(k_export_rows<double>, dim3(bg, mi + 1), 128, 0, s, h->B, h->Bp, mi + 1, w.defhist, (const int*)h->w.orig, d_defect_hist);
    if (d_alpha_hist) LAUNCH(k_export_rows<int>, dim3(bg, mi), 128, 0, s, h->B, h->Bp, mi, w.alphahist, (const int*)h->w.orig, (int*)d_alpha_hist);
    return 0;
}

int trajopt_export_reg(trajopt_handle* h, double* d_mu, double* d_delta, void* stream) {
    if (!h) return fail(TRAJOPT_E_INVALID, "trajopt_export_reg: NULL handle");
    if (!h->begun) return fail(TRAJOPT_E_STATE, "trajopt_export_reg: nothing solved yet");
    DeviceGuard guard(h->device);
    cudaStream_t s = (cudaStream_t)stream;
    const int bg = blocks_for(h->Bp, 128);
    if (d_mu) LAUNCH(k_export_rows<double>, dim3(bg, 1), 128, 0, s, h->B, h->Bp, 1, h->w.mu, (const int*)h->w.orig, d_mu);
    if (d_delta) LAUNCH(k_export_rows<double>, dim3(bg, 1), 128, 0, s, h->B, h->Bp, 1, h->w.delta, (const int*)h->w.orig, d_delta);
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
    const int bg = blocks_for(h->Bp, 128);
    if (d_lmbd) LAUNCH(k_export_traj, export_grid(h, (h->N + 1) * (2 * h->NU)), 128, 0, s, h->B, h->Bp, 2 * h->NU, w.lam, w.lam, (const int*)nullptr, (const int*)h->w.orig, d_lmbd, h->N + 1, (const int*)nullptr, 0);
    if (d_imu) LAUNCH(k_export_traj, export_grid(h, (h->N + 1) * (2 * h->NU)), 128, 0, s, h->B, h->Bp, 2 * h->NU, w.imu, w.imu, (const int*)nullptr, (const int*)h->w.orig, d_imu, h->N + 1, (const int*)nullptr, 0);
    if (d_mu) LAUNCH(k_export_rows<double>, dim3(bg, 1), 128, 0, s, h->B, h->Bp, 1, w.al_mu, (const int*)h->w.orig, d_mu);
    if (d_violation) LAUNCH(k_export_rows<double>, dim3(bg, 1), 128, 0, s, h->B, h->Bp, 1, w.al_viol, (const int*)h->w.orig, d_violation);
    if (d_outer_iters) LAUNCH(k_export_rows<int>, dim3(bg, 1), 128, 0, s, h->B, h->Bp, 1, w.al_outer, (const int*)h->w.orig, (int*)d_outer_iters);
    return 0;
}

int trajopt_export_al_state(trajopt_handle* h, double* d_lmbd_state, double* d_imu_state, void* stream) {
    if (!h) return fail(TRAJOPT_E_INVALID, "trajopt_export_al_state: NULL handle");
    if (h->method != TRAJOPT_AL_MS || !h->prm.has_state_bounds) return fail(TRAJOPT_E_STATE, "trajopt_export_al_state: no velocity bounds on this handle");
    if (!h->begun) return fail(TRAJOPT_E_STATE, "trajopt_export_al_state: nothing solved yet");
    DeviceGuard guard(h->device);
    cudaStream_t s = (cudaStream_t)stream;
    const Work& w = h->w;
    const int bg = blocks_for(h->Bp, 128), cw = 2 * (h->NX - h->NP);
    if (d_lmbd_state) LAUNCH(k_export_traj, export_grid(h, (h->N + 1) * (cw)), 128, 0, s, h->B, h->Bp, cw, w.lam_s, w.lam_s, (const int*)nullptr, (const int*)h->w.orig, d_lmbd_state, h->N + 1, (const int*)nullptr, 0);
    if (d_imu_state) LAUNCH(k_export_traj, export_grid(h, (h->N + 1) * (cw)), 128, 0, s, h->B, h->Bp, cw, w.imu_s, w.imu_s, (const int*)nullptr, (const int*)h->w.orig, d_imu_state, h->N + 1, (const int*)nullptr, 0);
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

int trajopt_solve_stream(trajopt_handle* h, const double* d_x0, int64_t n_problems, const double* d_us_init, double* d_xs,
                         double* d_us, double* d_J, int32_t* d_iters, int32_t* d_status, double* d_grad, double* d_defect,
                         void* stream) {
    if (!h) return fail(TRAJOPT_E_INVALID, "trajopt_solve_stream: NULL handle");
    if (n_problems == 0) return 0;
    if (!d_x0) return fail(TRAJOPT_E_INVALID, "trajopt_solve_stream: NULL argument");
    if (!h->have_params || !h->have_ref) return fail(TRAJOPT_E_STATE, "trajopt_solve_stream: set parameters and reference first");
    if (n_problems < 0 || n_problems > 0x7fffffff) return fail(TRAJOPT_E_INVALID, "trajopt_solve_stream: n_problems out of range");
    if (h->method == TRAJOPT_AL_MS) return fail(TRAJOPT_E_INVALID, "trajopt_solve_stream: the augmented-Lagrangian outer loop is per batch; use trajopt_solve");
    if (h->w.ref_batch || h->var_horizons)
        return fail(TRAJOPT_E_INVALID, "trajopt_solve_stream: per-problem references / horizons belong to slots; use trajopt_solve");
    DeviceGuard guard(h->device);
    int rc = ensure_hist(h);
    if (rc) return rc;
    h->w.us_init = d_us_init;
    h->w.us_mode = d_us_init ? 1 : 0;
    h->begun = false;
    return DISPATCH_KIND(h, solve_stream_impl, h, d_x0, (int)n_problems, d_xs, d_us, d_J, (int*)d_iters, (int*)d_status, d_grad,
                         d_defect, (cudaStream_t)stream);
}

int trajopt_solve_stream_host(trajopt_handle* h, const double* h_x0, int64_t n_problems, const double* h_us_init, double* h_xs,
                              double* h_us, double* h_J, int32_t* h_iters, int32_t* h_status, double* h_grad, double* h_defect,
                              void* stream) {
    if (!h) return fail(TRAJOPT_E_INVALID, "trajopt_solve_stream_host: NULL handle");
    if (n_problems == 0) return 0;
    if (!h_x0) return fail(TRAJOPT_E_INVALID, "trajopt_solve_stream_host: NULL argument");
    if (n_problems < 0 || n_problems > 0x7fffffff) return fail(TRAJOPT_E_INVALID, "trajopt_solve_stream_host: n_problems out of range");
    DeviceGuard guard(h->device);
    cudaStream_t s = (cudaStream_t)stream;
    const size_t n = (size_t)n_problems, N = h->N;
    void* host[7] = {h_xs, h_us, h_J, h_iters, h_status, h_grad, h_defect};
    const size_t row[7] = {(N + 1) * h->NS * 8, N * h->NU * 8, 8, 4, 4, 8, 8};   // bytes per problem
    for (int k = 0; k < 7; ++k) {
        if (!host[k] || row[k] * n <= h->so_bytes[k]) continue;
        if (h->so_buf[k]) cudaFree(h->so_buf[k]);
        h->so_buf[k] = nullptr;
        h->so_bytes[k] = 0;
        CUDA_OK(cudaMalloc(&h->so_buf[k], row[k] * n));
        h->so_bytes[k] = row[k] * n;
    }
    const size_t x0_bytes = n * h->NS * 8, us0_bytes = h_us_init ? N * h->NU * 8 : 0;
    if (x0_bytes > h->so_x0_bytes) {
        if (h->so_x0) cudaFree(h->so_x0);
        h->so_x0 = nullptr;
        h->so_x0_bytes = 0;
        CUDA_OK(cudaMalloc((void**)&h->so_x0, x0_bytes));
        h->so_x0_bytes = x0_bytes;
    }
    if (us0_bytes > h->s_us0_bytes) {
        if (h->s_us0) cudaFree(h->s_us0);
        h->s_us0 = nullptr;
        CUDA_OK(cudaMalloc((void**)&h->s_us0, us0_bytes));
        h->s_us0_bytes = us0_bytes;
    }
    if (!h->s_copy) CUDA_OK(cudaStreamCreateWithFlags(&h->s_copy, cudaStreamNonBlocking));
    if (n) CUDA_OK(cudaMemcpyAsync(h->so_x0, h_x0, x0_bytes, cudaMemcpyHostToDevice, s));
    if (us0_bytes) CUDA_OK(cudaMemcpyAsync(h->s_us0, h_us_init, us0_bytes, cudaMemcpyHostToDevice, s));
    // rows of problems [copied, upto) leave for the host on the copy stream while the solve goes on; the caller of
    // stream_progress has just synchronised the solve's stream, so those rows are complete in the staging arrays
    size_t copied = 0;
    auto copy_out = [&](size_t upto) -> int {
        if (upto <= copied) return 0;
        for (int k = 0; k < 7; ++k)
            if (host[k])
                CUDA_OK(cudaMemcpyAsync((char*)host[k] + copied * row[k], (char*)h->so_buf[k] + copied * row[k], (upto - copied) * row[k],
                                        cudaMemcpyDeviceToHost, h->s_copy));
        copied = upto;
        return 0;
    };
    h->stream_progress = [&](int prefix) -> int { return copy_out((size_t)std::max(prefix, 0) > n ? n : (size_t)std::max(prefix, 0)); };
    auto dev = [&](int k) { return host[k] ? h->so_buf[k] : nullptr; };
    int rc = trajopt_solve_stream(h, h->so_x0, n_problems, us0_bytes ? h->s_us0 : nullptr, (double*)dev(0), (double*)dev(1),
                                  (double*)dev(2), (int32_t*)dev(3), (int32_t*)dev(4), (double*)dev(5), (double*)dev(6), stream);
    h->stream_progress = nullptr;
    if (rc) return rc;
    if ((rc = copy_out(n))) return rc;
    CUDA_OK(cudaStreamSynchronize(h->s_copy));
    return 0;
}

// TRAJOPT_HOST_EARLY_COPY=0: copy everything at the end of the solve (A/B measurements)
static bool host_early_copy() {
    static const bool v = [] { const char* e = getenv("TRAJOPT_HOST_EARLY_COPY"); return !e || atoi(e) != 0; }();
    return v;
}

// The host-buffer solve in two halves, so that the device->host copies of one solve can drain while the next one
// computes on the same handle:
//   trajopt_solve_host_begin  copies the inputs up, runs the whole solve, queues every device->host copy on the handle's
//                             copy stream and returns a ticket WITHOUT waiting for them; the caller's stream is left free
//                             (it never waits for a copy), so the next solve may begin at once;
//   trajopt_solve_host_wait   blocks until the host arrays of that ticket are complete.
// Copies and the solves share nothing but the staging arrays: a later solve waits (on the device) for the previous ticket's
// copies before it rewrites them, which on the headline workload is ~300 ms after they were queued.
// Order on the copy stream (everything that lands in the host arrays goes through it, in this order):
//   bulk copy of ALL rows at the moment three quarters of the batch have stopped (finished problems' rows are final; the
//   running ones' rows are stale and may even be half-rewritten by the final export, which runs beside the copy)
//   -> rows of the problems that were still running then, written by a kernel straight into the pinned host arrays
//   -> the per-problem summaries.
int trajopt_solve_host_begin(trajopt_handle* h, const double* h_x0, const double* h_us_init, int us_mode, double* h_xs,
                             double* h_us, double* h_J, int32_t* h_iters, int32_t* h_status, double* h_grad, double* h_defect,
                             void* stream, int* ticket_out) {
    if (!h || !h_x0 || !ticket_out) return fail(TRAJOPT_E_INVALID, "trajopt_solve_host: NULL argument");
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
        CUDA_OK(cudaMalloc((void**)&h->s_snap, B * 4));
        CUDA_OK(cudaMalloc((void**)&h->d_late, B * 4));
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
    if (!h->s_copy) CUDA_OK(cudaStreamCreateWithFlags(&h->s_copy, cudaStreamNonBlocking));
    if (!h->ev_host[0]) {
        CUDA_OK(cudaEventCreateWithFlags(&h->ev_host[0], cudaEventDisableTiming));
        CUDA_OK(cudaEventCreateWithFlags(&h->ev_host[1], cudaEventDisableTiming));
        for (cudaEvent_t& e : h->ev_done) CUDA_OK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming | cudaEventBlockingSync));   // the waiting thread sleeps
    }
    const int ticket = h->next_ticket;
    const cudaEvent_t prev_done = h->tickets_issued ? h->ev_done[(ticket + kHostTickets - 1) % kHostTickets] : nullptr;
    // the event of this ticket is about to be re-recorded: the ticket that used it kHostTickets solves ago must be complete
    if (h->tickets_issued >= kHostTickets) CUDA_OK(cudaEventSynchronize(h->ev_done[ticket]));
    CUDA_OK(cudaMemcpyAsync(h->s_x0, h_x0, B * h->NS * 8, cudaMemcpyHostToDevice, s));
    if (us0_bytes) CUDA_OK(cudaMemcpyAsync(h->s_us0, h_us_init, us0_bytes, cudaMemcpyHostToDevice, s));
    int rc = trajopt_begin(h, h->s_x0, us0_bytes ? h->s_us0 : nullptr, us_mode, stream);
    if (rc) return rc;
    const int units = (h->method == TRAJOPT_AL_MS) ? h->user.n_al_iters : h->prm.max_iters + 1;
    const bool can_early = (h_xs || h_us) && h->method != TRAJOPT_AL_MS && host_early_copy();
    bool early = false, staged_wait = false;
    std::vector<int32_t> snap;
    // before anything is written into the staging arrays: the previous ticket's copies have read them (device-side wait)
    auto staging_free = [&]() -> int {
        if (!staged_wait && prev_done) CUDA_OK(cudaStreamWaitEvent(s, prev_done, 0));
        staged_wait = true;
        return 0;
    };
    int active = 1;
    for (int u = 0; u < units && active != 0; ++u) {
        if ((rc = trajopt_iterate(h, 1, &active, stream))) return rc;
        if (can_early && !early && active > 0 && (size_t)active * 4 <= B) {
            if ((rc = staging_free())) return rc;
            if ((rc = trajopt_export(h, h_xs ? h->s_xs : nullptr, h_us ? h->s_us : nullptr, nullptr, nullptr, h->s_snap, nullptr,
                                     nullptr, stream)))
                return rc;
            LAUNCH(k_ints_to_host, blocks_for((int)B, 256), 256, 0, s, (const int*)h->s_snap, (volatile int*)h->h_ints_dev, (int)B);
            CUDA_OK(cudaEventRecord(h->ev_host[0], s));
            CUDA_OK(cudaStreamWaitEvent(h->s_copy, h->ev_host[0], 0));
            if (h_xs) CUDA_OK(cudaMemcpyAsync(h_xs, h->s_xs, B * (N + 1) * h->NS * 8, cudaMemcpyDeviceToHost, h->s_copy));
            if (h_us) CUDA_OK(cudaMemcpyAsync(h_us, h->s_us, B * N * h->NU * 8, cudaMemcpyDeviceToHost, h->s_copy));
            CUDA_OK(cudaStreamSynchronize(s));
            snap.assign(h->h_ints, h->h_ints + B);
            early = true;
        }
    }
    if ((rc = staging_free())) return rc;
    const Work& w = h->w;
    const int* nb = h->var_horizons ? (const int*)w.Nb : (const int*)nullptr;
    if (early) {   // only the stragglers' rows are new; s_snap holds the status snapshot of the early export
        if (h_xs) LAUNCH(k_export_traj_late, export_grid(h, (h->N + 1) * (h->NS)), 128, 0, s, h->B, h->Bp, h->NS, w.X[0], w.X[1], w.sel, (const int*)w.orig, (const int*)h->s_snap, h->s_xs, h->N + 1, nb, 0);
        if (h_us) LAUNCH(k_export_traj_late, export_grid(h, (h->N) * (h->NU)), 128, 0, s, h->B, h->Bp, h->NU, w.U[0], w.U[1], w.sel, (const int*)w.orig, (const int*)h->s_snap, h->s_us, h->N, nb, 1);
    }
    if ((rc = trajopt_export(h, (!early && h_xs) ? h->s_xs : nullptr, (!early && h_us) ? h->s_us : nullptr, h->s_J, h->s_iters,
                             h->s_status, h->s_grad, h->s_def, stream)))
        return rc;
    std::vector<int> late;
    if (early)
        for (size_t b = 0; b < B; ++b)
            if ((snap[b] & 15) == TRAJOPT_RUNNING) late.push_back((int)b);
    void *dx = nullptr, *du = nullptr;
    const bool mapped = early && !late.empty() && (!h_xs || cudaHostGetDevicePointer(&dx, h_xs, 0) == cudaSuccess) &&
                        (!h_us || cudaHostGetDevicePointer(&du, h_us, 0) == cudaSuccess);
    if (early && !late.empty() && !mapped) (void)cudaGetLastError();
    if (mapped) CUDA_OK(cudaMemcpyAsync(h->d_late, late.data(), late.size() * sizeof(int), cudaMemcpyHostToDevice, s));
    // everything the solve produced is in the staging arrays: the rest belongs to the copy stream
    CUDA_OK(cudaEventRecord(h->ev_host[1], s));
    CUDA_OK(cudaStreamWaitEvent(h->s_copy, h->ev_host[1], 0));
    cudaStream_t c = h->s_copy;
    if (!early) {
        if (h_xs) CUDA_OK(cudaMemcpyAsync(h_xs, h->s_xs, B * (N + 1) * h->NS * 8, cudaMemcpyDeviceToHost, c));
        if (h_us) CUDA_OK(cudaMemcpyAsync(h_us, h->s_us, B * N * h->NU * 8, cudaMemcpyDeviceToHost, c));
    } else if (mapped) {          // pinned host arrays: the kernel writes the stragglers' rows over PCIe itself
        if (h_xs) LAUNCH(k_rows_to_host, (unsigned)late.size(), 256, 0, c, (const int*)h->d_late, (size_t)(N + 1) * h->NS, (const double*)h->s_xs, (double*)dx);
        if (h_us) LAUNCH(k_rows_to_host, (unsigned)late.size(), 256, 0, c, (const int*)h->d_late, (size_t)N * h->NU, (const double*)h->s_us, (double*)du);
    } else {
        const size_t rx = (N + 1) * h->NS * 8, ru = N * h->NU * 8;
        for (size_t i = 0; i < late.size();) {      // pageable host arrays: one copy per run of consecutive rows
            size_t e = i + 1;
            while (e < late.size() && late[e] == late[e - 1] + 1) ++e;
            const size_t b = (size_t)late[i], n = e - i;
            if (h_xs) CUDA_OK(cudaMemcpyAsync((char*)h_xs + b * rx, (char*)h->s_xs + b * rx, n * rx, cudaMemcpyDeviceToHost, c));
            if (h_us) CUDA_OK(cudaMemcpyAsync((char*)h_us + b * ru, (char*)h->s_us + b * ru, n * ru, cudaMemcpyDeviceToHost, c));
            i = e;
        }
    }
    if (h_J) CUDA_OK(cudaMemcpyAsync(h_J, h->s_J, B * 8, cudaMemcpyDeviceToHost, c));
    if (h_grad) CUDA_OK(cudaMemcpyAsync(h_grad, h->s_grad, B * 8, cudaMemcpyDeviceToHost, c));
    if (h_defect) CUDA_OK(cudaMemcpyAsync(h_defect, h->s_def, B * 8, cudaMemcpyDeviceToHost, c));
    if (h_iters) CUDA_OK(cudaMemcpyAsync(h_iters, h->s_iters, B * 4, cudaMemcpyDeviceToHost, c));
    if (h_status) CUDA_OK(cudaMemcpyAsync(h_status, h->s_status, B * 4, cudaMemcpyDeviceToHost, c));
    CUDA_OK(cudaEventRecord(h->ev_done[ticket], c));
    h->next_ticket = (ticket + 1) % kHostTickets;
    h->tickets_issued += 1;
    *ticket_out = ticket;
    return 0;
}

int trajopt_solve_host_wait(trajopt_handle* h, int ticket) {
    if (!h) return fail(TRAJOPT_E_INVALID, "trajopt_solve_host_wait: NULL handle");
    if (ticket < 0 || ticket >= kHostTickets || !h->ev_done[ticket]) return fail(TRAJOPT_E_INVALID, "trajopt_solve_host_wait: no such ticket");
    DeviceGuard guard(h->device);
    CUDA_OK(cudaEventSynchronize(h->ev_done[ticket]));
    return 0;
}

int trajopt_solve_host(trajopt_handle* h, const double* h_x0, const double* h_us_init, int us_mode, double* h_xs,
                       double* h_us, double* h_J, int32_t* h_iters, int32_t* h_status, double* h_grad, double* h_defect,
                       void* stream) {
    int ticket = -1;
    int rc = trajopt_solve_host_begin(h, h_x0, h_us_init, us_mode, h_xs, h_us, h_J, h_iters, h_status, h_grad, h_defect, stream, &ticket);
    if (rc) return rc;
    return trajopt_solve_host_wait(h, ticket);
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
    const int bg = blocks_for(h->Bp, 128);
    const int glen = h->NU * h->NX + h->NU;
    const dim3 gg(bg, h->N);
    if (d_k) LAUNCH(k_export_gains, gg, 128, 0, s, h->B, h->Bp, h->N, glen, h->NU * h->NX, h->NU, (const double*)h->w.gains, (const int*)h->w.orig, d_k);
    if (d_K) LAUNCH(k_export_gains, gg, 128, 0, s, h->B, h->Bp, h->N, glen, 0, h->NU * h->NX, (const double*)h->w.gains, (const int*)h->w.orig, d_K);
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

int trajopt_set_compaction(trajopt_handle* h, int min_batch, int ratio) {
    if (!h) return fail(TRAJOPT_E_INVALID, "trajopt_set_compaction: NULL handle");
    if (ratio < 1) return fail(TRAJOPT_E_INVALID, "trajopt_set_compaction: ratio must be >= 1");
    h->compact_min_batch = min_batch;
    h->compact_ratio = ratio;
    return 0;
}

int trajopt_set_sweep(trajopt_handle* h, int variant, int lanes) {
    if (!h) return fail(TRAJOPT_E_INVALID, "trajopt_set_sweep: NULL handle");
    if ((variant != 0 && variant != 2 && variant != 4 && variant != 6) || lanes < 1) return fail(TRAJOPT_E_INVALID, "trajopt_set_sweep: variant must be 0, 2, 4 or 6 and lanes >= 1");
    h->sweep_variant = variant;
    h->sweep_lanes = lanes;
    return 0;
}

int trajopt_set_line_search_batch(trajopt_handle* h, int max_batch) {
    if (!h) return fail(TRAJOPT_E_INVALID, "trajopt_set_line_search_batch: NULL handle");
    if (max_batch < 0) return fail(TRAJOPT_E_INVALID, "trajopt_set_line_search_batch: max_batch must be >= 0");
    DeviceGuard guard(h->device);
    h->cand_max_batch = max_batch;
    return h->have_params ? ensure_cand(h) : 0;
}

int trajopt_set_profiling(trajopt_handle* h, int enable) {
    if (!h) return fail(TRAJOPT_E_INVALID, "trajopt_set_profiling: NULL handle");
    h->profiling = enable != 0;
    return 0;
}

}  // extern "C"
