// cabi.cu -- extern "C" entry points of libcambb200.so (include/cambb200.h): context and
// device-buffer management, host<->device marshalling, launch orchestration and timing.
// No torch types cross this boundary; there is no CPU fallback (every compute entry point
// fails with CB_ERR_NO_DEVICE when no CUDA device is usable).
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "dmodel.cuh"

// evolve.cu
void cb_mode_size(const DModel& M, double q, int transfer_only, int* nvar, int* maxl, int* lmax3);
cudaError_t cb_evolve_launch(const DModel*, const DModeItem*, int n_items, int nv, int maxl, int lanes_per_mode, int lean,
                             unsigned int*, int*, int n_sm, cudaStream_t, int* grid_out, int block_cap);
cudaError_t cb_debug_derivs_launch(const DModel* d_model, double q, double tau, const int* flags4, const double* d_y, int n, int nv,
                                   int maxl, int lanes, int lean, double* d_yprime, double* d_sources, int* d_status, cudaStream_t stream);
// evolve_reg.cu
int cb_reg_class(int lmaxg, int lmaxgpol, int lmaxnr);
cudaError_t cb_evolve_reg_launch(int reg_class, const DModel*, const DModeItem*, int n_items, unsigned int*, int*, int n_sm, cudaStream_t,
                                 int* grid_out);
// tmem_probe.cu
cudaError_t cb_tmem_probe_launch(int n_blocks, int* d_mismatches, long long* d_cycles, cudaStream_t stream);
// evolve_mpl.cu
int cb_mpl_class(int lmaxg, int lmaxgpol, int lmaxnr);
cudaError_t cb_evolve_mpl_launch(const DModel* d_models, const DModeItem* d_items_a, int n_a, const DModeItem* d_items_c, int n_c,
                                 unsigned int* d_counter_a, unsigned int* d_counter_c, int* d_error, int n_sm, cudaStream_t stream,
                                 int* grid_out);
cudaError_t cb_debug_derivs_mpl_launch(const DModel* d_model, double q, double tau, const int* flags4, const int* lmax3, const double* d_y,
                                       int n, double* d_yprime, double* d_sources, cudaStream_t stream);
#ifdef CB_TRACE
void cb_trace_mpl_begin(int q_ix, int cap); void cb_trace_mpl_end(const char* path);
void cb_trace_reg_begin(int q_ix, int cap); void cb_trace_reg_end(const char* path);
#endif
// evolve_tensor.cu
void cb_tensor_mode_size(const DModel& M, int* nvart, int* maxl);
cudaError_t cb_evolve_tensor_launch(const DModel*, const DModeItem*, int n_items, int nv, int maxl, unsigned int*, int*, int n_sm,
                                    cudaStream_t, int* grid_out);
// project.cu
void cb_spline_k_launch(const DModel*, int, int, double* const*, cudaStream_t);
size_t cb_los_smem(int, int, int, int);
cudaError_t cb_los_launch(const DModel*, int, const DBessel&, int, int, int, int, cudaStream_t);
void cb_bessel_transpose_launch(const double*, double*, int, int, int, cudaStream_t);
// collective.cu (NCCL bound at run time)
bool cb_nccl_unique_id(char* id128, std::string* err);
bool cb_nccl_init(void** comm, int nranks, int rank, const char* id128, std::string* err);
void cb_nccl_destroy(void* comm);
bool cb_nccl_gather_rows(void* comm, int nranks, double* d_src, int n_k, int ncol, const int* d_idx, int n_act, int n_pack, const int* d_owner,
                         const int* d_pos, double* d_packed, double* d_gathered, cudaStream_t st, std::string* err);
bool cb_nccl_allreduce_sum(void* comm, double* d_buf, size_t n, cudaStream_t st, std::string* err);
struct ThermoPackArgs {   // thermo_pack.cu
    int n;
    double dlntau;
    const double *cs2, *dotmu, *emmu, *adot, *sf;
    double *ddotmu, *dddotmu, *ddddotmu, *demmu, *dadot, *dcs2, *g, *horner;
};
cudaError_t cb_thermo_pack_launch(const ThermoPackArgs* d_args, int n_models, int max_n, cudaStream_t st);
cudaError_t cb_bessel_build_launch(const double* d_x, int nx, const int* d_l, int n_l, int n_l_pad, const cb_range* d_regions, int n_regions,
                                   double* d_work, double* d_dend, int* d_min_ix, double* d_h, double* d_ht, cudaStream_t st);   // bessel_table.cu
cudaError_t cb_nonlinear_sources_launch(const DModel* d_model, int n_k, const double* d_ratio, int n_z, double min_kh, cudaStream_t st);
cudaError_t cb_los_win_launch(const DModel*, int, const DBessel&, int, int, cudaStream_t);
void cb_cl_launch(const DModel*, int, const DBessel&, cudaStream_t);
void cb_cl_array_launch(const DModel*, int, const DBessel&, double* const*, cudaStream_t);
#include "lensing.cuh"
void cb_limber_launch(const DModel*, int, const DBessel&, int, double* const*, cudaStream_t);
void cb_cl_windows_launch(const DModel*, int, const DBessel&, double* const*, double* const*, double* const*, cudaStream_t);
// peak.cu
double cb_fp64_peak_measure(int n_sm, cudaStream_t st, int reps);

// Modes are binned by the size of their shared-memory slab (worst-case number of equations AND deepest multipole:
// the low-k modes carry a 32..64-deep photon hierarchy that would otherwise inflate everyone's coefficient
// tables), so that every launch sizes its per-mode slab for its own class.  The limits are the slab sizes at
// which 6, 5, 4 and 3 blocks of eight modes fit one SM's 227 KB (1 KB per block is reserved by the system).
// The last class holds tensor modes (their own kernel, evolve_tensor.cu).
// Classes 6, 7, 8 are the three static layouts of the register-resident kernel (evolve_reg.cu), 9 and 10 the two classes
// of the mode-per-lane kernel (evolve_mpl.cu, one launch for both): lean launches only.
#define CB_N_CLASSES 11
#define CB_MPL_CLASS_A 9
#define CB_MPL_CLASS_C 10
#define CB_TENSOR_CLASS 5
#define CB_REG_CLASS0 6
static const size_t kClassSlabLimit[5] = {4714, 5683, 7136, 9557, ~size_t(0)};
size_t cb_evolve_slab_bytes(int nv, int maxl, int lean);   // evolve.cu

struct cb_dmodel {
    DModel h;                      // host copy with device pointers
    std::vector<void*> owned;      // table allocations freed with the model
    // per-call buffers (grown on demand)
    double *d_q_src = nullptr, *d_src = nullptr, *d_dd = nullptr, *d_scratch = nullptr, *d_delta = nullptr, *d_icl = nullptr;
    double *d_q_int = nullptr, *d_dq_int = nullptr, *d_pows = nullptr, *d_icl_array = nullptr;
    size_t cap_icl_array = 0;
    double *d_limb = nullptr, *d_lpow = nullptr, *d_icl_w = nullptr;   // source windows: Limber sources, Limber P(k), iCl (n_l_all, 6)
    size_t cap_limb = 0, cap_lpow = 0, cap_icl_w = 0;
    float* d_transfer = nullptr;
    long long* d_stats = nullptr;
    size_t cap_src = 0, cap_k = 0, cap_q = 0, cap_delta = 0, cap_tr = 0, cap_icl = 0, cap_stats = 0;
    int n_cl = 3;
    // thermo tables derived on the device (cb_model.thermo_horner == NULL): packed by the next cb_evolve together with the
    // other models of its batch, one launch for all of them
    bool needs_thermo_pack = false;
    ThermoPackArgs pack_args{};
    std::vector<void*> pack_tmp;   // work arrays of the packing, released after it
};

struct cb_ctx {
    int device = 0, n_sm = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, evt0 = nullptr, evt1 = nullptr;
    std::string err;
    DBessel bess{};                // the table in use (a copy of one cache entry)
    struct BessEntry { unsigned long long key; DBessel B; std::vector<void*> owned; long long stamp; };
    std::vector<BessEntry> bess_cache;   // resident tables, keyed by content: a batch whose sets have different l samplings
    long long bess_clock = 0;            // switches between them without re-uploading (26 MB each; evicted oldest-first)
    bool have_bessel = false;
    double ms[4] = {0, 0, 0, 0};
    long long launches = 0;
    DModel* d_models = nullptr;    size_t cap_models = 0;
    DModeItem* d_items = nullptr;  size_t cap_items = 0;
    double** d_scratch_ptrs = nullptr; size_t cap_scr = 0;
    unsigned int* d_counter = nullptr;   // one work-queue counter per size class
    int* d_error = nullptr;
    cudaStream_t class_stream[CB_N_CLASSES] = {};
    cudaEvent_t class_done[CB_N_CLASSES] = {}, fork = nullptr;
    int lanes_per_mode = 0;              // 0 = automatic
    int class_lanes[CB_N_CLASSES] = {};  // per slab-size class override of the lane width (0 = lanes_per_mode)
    int async_model_upload = 0;          // cb_model_create returns without waiting for its host->device copies
    // lean launches: 0 = the general kernel only; 1 (default) = register kernel for every mode that fits a static layout;
    // 2 = the same but the 16-lane layout's modes through the general kernel; 3 = mode-per-lane Tensor-Memory kernel for the
    // 48-element classes, register kernel for the deep low-k hierarchies (equal throughput at >= 192 sets per call,
    // slower below: DESIGN.md section 4.1)
    int use_reg_kernel = 1;
    int reg_min_modes = -1;   // launches with fewer lean modes take the general kernel (one mode per warp: latency); -1 = 8 per SM
    // asynchronous result copies (option "async_outputs"): Delta_p_l_k goes to the host on its own stream, behind the
    // kernel that produced it, while the next kernels run; cb_wait_outputs is the point where the host may read it
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_los_done = nullptr, ev_copy_done = nullptr;
    bool async_outputs = false, copies_pending = false;
    // "async_calls": cb_evolve / cb_project return once their work is queued when the call has no host output to fill; the
    // staging buffers are protected by ev_staged (recorded behind the call's host->device copies), an integrator failure
    // surfaces at the next blocking call through the page-locked h_error
    bool async_calls = false, staged_pending = false, error_pending = false;
    cudaEvent_t ev_staged = nullptr;
    int* h_error = nullptr;
    struct DeferredCopy { void* dst; const void* src; size_t bytes; };
    std::vector<DeferredCopy> deferred;   // queued by cb_project, issued by the next cb_evolve / cb_wait_outputs (see flush_outputs)
    void* comm = nullptr; int comm_ranks = 1, comm_rank = 0;   // NCCL communicator of a k-sharded run (cb_comm_init)
    int los_kernel = 1;            // lanes on the multipoles: 1 rows by coalesced loads (default), 2 rows by bulk copies (TMA) into per-warp rings (measured slower); 0: the gather kernel
    cudaEvent_t stage_ev[4][2] = {};     // per-stage kernel brackets of the last calls (evolve, spline-k, LOS, C_l)
    DModel* h_models = nullptr;   size_t cap_h_models = 0;   // page-locked staging
    DModeItem* h_items = nullptr; size_t cap_h_items = 0;
    int last_grid[CB_N_CLASSES] = {};
    // many small host arrays travel as one page-locked block and one PCIe copy each way (pieces move device-to-device)
    double *h_in = nullptr, *d_in = nullptr, *h_out = nullptr, *d_out = nullptr;
    size_t cap_in = 0, cap_out = 0, used_in = 0;
    LensTab* d_lens_tab = nullptr; LensTab2* d_lens_tab2 = nullptr; int lens_tab_n = 0;       // per-l constants of the lensing kernels, grown on demand
    cudaEvent_t lens_ev[2] = {};
    bool lens_timed = false;
};

#define CK(call)                                                                                  \
    do {                                                                                          \
        cudaError_t e_ = (call);                                                                  \
        if (e_ != cudaSuccess) {                                                                  \
            ctx->err = std::string(#call) + ": " + cudaGetErrorString(e_);                        \
            return CB_ERR_CUDA;                                                                   \
        }                                                                                         \
    } while (0)

// Device memory comes from the stream-ordered pool (release threshold = never): creating and
// destroying models every call costs no cudaMalloc/cudaFree round trips or device-wide syncs.
static cudaError_t dev_alloc(cb_ctx* ctx, void** p, size_t bytes) { return cudaMallocAsync(p, bytes ? bytes : 8, ctx->stream); }
static void dev_free(cb_ctx* ctx, void* p) { if (p) cudaFreeAsync(p, ctx->stream); }

template <class T>
static int upload(cb_ctx* ctx, const T* src, size_t n, T** dst, std::vector<void*>* owned) {
    *dst = nullptr;
    if (!src || n == 0) return 0;
    CK(dev_alloc(ctx, (void**)dst, n * sizeof(T)));
    if (owned) owned->push_back(*dst);
    CK(cudaMemcpyAsync(*dst, src, n * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
    return 0;
}

template <class T>
static int ensure(cb_ctx* ctx, T** p, size_t* cap, size_t n) {
    if (n <= *cap && *p) return 0;
    if (*p) dev_free(ctx, *p);
    *p = nullptr;
    CK(dev_alloc(ctx, (void**)p, n * sizeof(T)));
    *cap = n;
    return 0;
}

// Staging of small arrays.  Safe to reuse between entry points because each one ends with a stream synchronisation.
static int stage_reserve(cb_ctx* ctx, size_t n_in, size_t n_out) {
    if (n_in > ctx->cap_in) {
        if (ctx->h_in) cudaFreeHost(ctx->h_in);
        if (ctx->d_in) cudaFree(ctx->d_in);
        ctx->h_in = ctx->d_in = nullptr; ctx->cap_in = 0;
        CK(cudaHostAlloc((void**)&ctx->h_in, sizeof(double) * n_in, cudaHostAllocDefault));
        CK(cudaMalloc((void**)&ctx->d_in, sizeof(double) * n_in));
        ctx->cap_in = n_in;
    }
    if (n_out > ctx->cap_out) {
        if (ctx->h_out) cudaFreeHost(ctx->h_out);
        if (ctx->d_out) cudaFree(ctx->d_out);
        ctx->h_out = ctx->d_out = nullptr; ctx->cap_out = 0;
        CK(cudaHostAlloc((void**)&ctx->h_out, sizeof(double) * n_out, cudaHostAllocDefault));
        CK(cudaMalloc((void**)&ctx->d_out, sizeof(double) * n_out));
        ctx->cap_out = n_out;
    }
    ctx->used_in = 0;
    return 0;
}
// queue src (host) -> dst (device): packed now, moved by stage_upload
struct StagePiece { double* dst; size_t off, n; };
static void stage_add(cb_ctx* ctx, std::vector<StagePiece>& pieces, double* dst, const double* src, size_t n) {
    memcpy(ctx->h_in + ctx->used_in, src, sizeof(double) * n);
    pieces.push_back({dst, ctx->used_in, n});
    ctx->used_in += n;
}
static int stage_upload(cb_ctx* ctx, const std::vector<StagePiece>& pieces) {
    if (pieces.empty()) return 0;
    CK(cudaMemcpyAsync(ctx->d_in, ctx->h_in, sizeof(double) * ctx->used_in, cudaMemcpyHostToDevice, ctx->stream));
    for (const StagePiece& p : pieces)
        CK(cudaMemcpyAsync(p.dst, ctx->d_in + p.off, sizeof(double) * p.n, cudaMemcpyDeviceToDevice, ctx->stream));
    return 0;
}

// "async_outputs": the big result copies are handed to the copy engine when the NEXT batch's evolve kernels are about to
// start (or when the host asks for them), not right behind the kernel that produced them -- issued there, the small
// device->host copies of cb_cl / cb_lens_contrib would queue behind them on the same engine.
// Entry points that write the staging buffers wait for the previous call's host->device copies (not for its kernels).
static int begin_call(cb_ctx* ctx) {
    if (ctx->staged_pending) { CK(cudaEventSynchronize(ctx->ev_staged)); ctx->staged_pending = false; }
    return 0;
}
// After a stream synchronisation: the error flag of an earlier asynchronous cb_evolve has landed in h_error.
static int deferred_error(cb_ctx* ctx) {
    if (!ctx->error_pending) return 0;
    ctx->error_pending = false;
    const int herr = ctx->h_error ? *ctx->h_error : 0;
    if (herr) { *ctx->h_error = 0; ctx->err = "RungeKuttaDP45 error: step size underflow in the device integrator (reported by an earlier cb_evolve)"; return herr; }
    return 0;
}

static int flush_outputs(cb_ctx* ctx) {
    if (ctx->deferred.empty()) return 0;
    CK(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_los_done, 0));
    for (const cb_ctx::DeferredCopy& c : ctx->deferred)
        CK(cudaMemcpyAsync(c.dst, c.src, c.bytes, cudaMemcpyDeviceToHost, ctx->copy_stream));
    ctx->deferred.clear();
    CK(cudaEventRecord(ctx->ev_copy_done, ctx->copy_stream));
    ctx->copies_pending = true;
    return 0;
}

extern "C" {

int cb_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

int cb_create(int device, cb_ctx** out) {
    *out = nullptr;
    int n = cb_device_count();
    if (n <= 0 || device >= n) return CB_ERR_NO_DEVICE;
    cb_ctx* ctx = new cb_ctx();
    ctx->device = device;
    if (cudaSetDevice(device) != cudaSuccess) { delete ctx; return CB_ERR_NO_DEVICE; }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { delete ctx; return CB_ERR_NO_DEVICE; }
    ctx->n_sm = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return CB_ERR_CUDA; }
    cudaEventCreate(&ctx->ev0);
    cudaEventCreate(&ctx->ev1);
    cudaEventCreate(&ctx->evt0);
    cudaEventCreate(&ctx->evt1);
    for (int c = 0; c < CB_N_CLASSES; c++) {
        cudaStreamCreateWithFlags(&ctx->class_stream[c], cudaStreamNonBlocking);
        cudaEventCreateWithFlags(&ctx->class_done[c], cudaEventDisableTiming);
    }
    cudaEventCreateWithFlags(&ctx->fork, cudaEventDisableTiming);
    cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking);
    cudaEventCreateWithFlags(&ctx->ev_staged, cudaEventDisableTiming);
    cudaHostAlloc((void**)&ctx->h_error, sizeof(int), cudaHostAllocDefault);
    if (ctx->h_error) *ctx->h_error = 0;
    cudaEventCreateWithFlags(&ctx->ev_los_done, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&ctx->ev_copy_done, cudaEventDisableTiming);
    if (const char* e = getenv("CAMB_B200_LANES")) ctx->lanes_per_mode = atoi(e);
    if (const char* e = getenv("CAMB_B200_REG")) ctx->use_reg_kernel = atoi(e);
    if (const char* e = getenv("CAMB_B200_REG_MIN_MODES")) ctx->reg_min_modes = atoi(e);
    if (const char* e = getenv("CAMB_B200_LOS")) ctx->los_kernel = atoi(e);
    if (const char* e = getenv("CAMB_B200_CLASS_LANES")) {   // tuning runs: lane width per slab-size class, e.g. "8,8,16,32,32"
        int i = 0;
        for (const char* p = e; *p && i < CB_TENSOR_CLASS; i++) { ctx->class_lanes[i] = atoi(p); while (*p && *p != ',') p++; if (*p == ',') p++; }
    }
    {
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
            unsigned long long thr = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
        }
    }
    for (int i = 0; i < 4; i++) { cudaEventCreate(&ctx->stage_ev[i][0]); cudaEventCreate(&ctx->stage_ev[i][1]); }
    if (cudaMalloc((void**)&ctx->d_counter, sizeof(unsigned int) * CB_N_CLASSES) != cudaSuccess ||
        cudaMalloc((void**)&ctx->d_error, sizeof(int)) != cudaSuccess) { delete ctx; return CB_ERR_CUDA; }
    *out = ctx;
    return 0;
}

void cb_destroy(cb_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (ctx->copy_stream) { cudaStreamSynchronize(ctx->copy_stream); cudaStreamDestroy(ctx->copy_stream); }
    if (ctx->ev_staged) cudaEventDestroy(ctx->ev_staged);
    if (ctx->h_error) cudaFreeHost(ctx->h_error);
    if (ctx->ev_los_done) cudaEventDestroy(ctx->ev_los_done);
    if (ctx->ev_copy_done) cudaEventDestroy(ctx->ev_copy_done);
    if (ctx->comm) cb_nccl_destroy(ctx->comm);
    for (auto& e : ctx->bess_cache) for (void* p : e.owned) cudaFree(p);
    for (int i = 0; i < 4; i++) { cudaEventDestroy(ctx->stage_ev[i][0]); cudaEventDestroy(ctx->stage_ev[i][1]); }
    if (ctx->h_models) cudaFreeHost(ctx->h_models);
    if (ctx->h_items) cudaFreeHost(ctx->h_items);
    if (ctx->d_models) cudaFree(ctx->d_models);
    if (ctx->d_items) cudaFree(ctx->d_items);
    if (ctx->d_scratch_ptrs) cudaFree(ctx->d_scratch_ptrs);
    if (ctx->d_counter) cudaFree(ctx->d_counter);
    if (ctx->d_error) cudaFree(ctx->d_error);
    if (ctx->h_in) cudaFreeHost(ctx->h_in);
    if (ctx->h_out) cudaFreeHost(ctx->h_out);
    if (ctx->d_in) cudaFree(ctx->d_in);
    if (ctx->d_out) cudaFree(ctx->d_out);
    if (ctx->d_lens_tab) cudaFree(ctx->d_lens_tab);
    if (ctx->d_lens_tab2) cudaFree(ctx->d_lens_tab2);
    for (cudaEvent_t e : ctx->lens_ev) if (e) cudaEventDestroy(e);
    for (int c = 0; c < CB_N_CLASSES; c++) {
        if (ctx->class_stream[c]) cudaStreamDestroy(ctx->class_stream[c]);
        if (ctx->class_done[c]) cudaEventDestroy(ctx->class_done[c]);
    }
    if (ctx->fork) cudaEventDestroy(ctx->fork);
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    if (ctx->evt0) cudaEventDestroy(ctx->evt0);
    if (ctx->evt1) cudaEventDestroy(ctx->evt1);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

const char* cb_last_error(const cb_ctx* ctx) { return ctx ? ctx->err.c_str() : "no context"; }

// Content key of a Bessel table: sizes, the whole l list, the region bounds, the x grid ends and a stride sample of the
// spline coefficients (the table is a pure function of (l list, x grid), bessels.f90:140-349).
static unsigned long long bessel_key(const cb_bessel* b, bool with_table = true) {
    unsigned long long h = 1469598103934665603ull;
    auto mix = [&h](const void* p, size_t n) { const unsigned char* c = (const unsigned char*)p; for (size_t i = 0; i < n; i++) { h ^= c[i]; h *= 1099511628211ull; } };
    const int n_l_all = b->n_l_all > 0 ? b->n_l_all : b->n_l;
    mix(&b->n_l, sizeof(int)); mix(&n_l_all, sizeof(int)); mix(&b->nx, sizeof(int)); mix(&b->n_regions, sizeof(int));
    mix(b->l, sizeof(int) * (size_t)n_l_all);
    mix(b->regions, sizeof(cb_range) * (size_t)b->n_regions);
    mix(b->x, sizeof(double)); mix(b->x + b->nx - 1, sizeof(double));
    if (!with_table) {   // a table generated on the device is a pure function of what has been mixed in so far
        const unsigned long long tag = 0x6275696c74ull;
        mix(&tag, sizeof(tag));
        mix(b->x, sizeof(double) * (size_t)b->nx);
        return h;
    }
    const size_t tot = (size_t)4 * (b->nx - 1) * b->n_l, step = tot / 509 + 1;
    for (size_t i = 0; i < tot; i += step) mix(b->horner + i, sizeof(double));
    return h;
}

int cb_set_bessel(cb_ctx* ctx, const cb_bessel* b) {
    if (!ctx) return CB_ERR_NO_DEVICE;
    CK(cudaSetDevice(ctx->device));
    if (b->n_regions > 8) { ctx->err = "too many Bessel x regions"; return CB_ERR_UNSUPPORTED_PARAMS; }
    const unsigned long long key = bessel_key(b);
    for (auto& e : ctx->bess_cache)
        if (e.key == key) { e.stamp = ++ctx->bess_clock; ctx->bess = e.B; ctx->have_bessel = true; return 0; }
    if (ctx->bess_cache.size() >= 8) {   // evict the least recently used table (no call is in flight: entry points are blocking)
        size_t old = 0;
        for (size_t i = 1; i < ctx->bess_cache.size(); i++) if (ctx->bess_cache[i].stamp < ctx->bess_cache[old].stamp) old = i;
        CK(cudaStreamSynchronize(ctx->stream));
        for (void* p : ctx->bess_cache[old].owned) dev_free(ctx, p);
        ctx->bess_cache.erase(ctx->bess_cache.begin() + old);
    }
    ctx->bess_cache.emplace_back();
    cb_ctx::BessEntry& E = ctx->bess_cache.back();
    E.key = key; E.stamp = ++ctx->bess_clock;
    DBessel& B = E.B;
    B.n_l = b->n_l; B.nx = b->nx; B.n_regions = b->n_regions;
    B.n_l_all = b->n_l_all > 0 ? b->n_l_all : b->n_l;
    for (int i = 0; i < b->n_regions; i++) B.regions[i] = b->regions[i];
    B.xmax = b->x[b->nx - 1];
    int* dl; double *dx, *dh;
    int e = upload(ctx, b->l, (size_t)B.n_l_all, &dl, &E.owned);
    if (!e) e = upload(ctx, b->x, (size_t)b->nx, &dx, &E.owned);
    if (!e) e = upload(ctx, b->horner, (size_t)4 * (b->nx - 1) * b->n_l, &dh, &E.owned);
    if (e) { for (void* p : E.owned) dev_free(ctx, p); ctx->bess_cache.pop_back(); return e; }
    B.l = dl; B.x = dx; B.horner = dh;
    {   // the transposed copy for the line-of-sight kernel (rows of all l at one x interval, padded to whole 128-byte lines)
        B.n_l_pad = (b->n_l + 3) & ~3;
        double* dht = nullptr;
        const size_t nt = (size_t)4 * (b->nx - 1) * B.n_l_pad;
        if (dev_alloc(ctx, (void**)&dht, sizeof(double) * nt) != cudaSuccess) {
            for (void* p : E.owned) dev_free(ctx, p);
            ctx->bess_cache.pop_back();
            ctx->err = "out of device memory for the transposed Bessel table";
            return CB_ERR_CUDA;
        }
        E.owned.push_back(dht);
        cb_bessel_transpose_launch(dh, dht, b->nx - 1, b->n_l, B.n_l_pad, ctx->stream);
        B.horner_t = dht;
    }
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->bess = B;
    ctx->have_bessel = true;
    return 0;
}

// GenerateBessels on the device (bessels.f90:140-228; bessel_table.cu): b->horner is not read.  The table becomes the one
// in use, exactly as after cb_set_bessel; horner_out (may be NULL) receives bessel_horner(4, nx-1, n_l) for the host.
int cb_build_bessel(cb_ctx* ctx, const cb_bessel* b, double* horner_out) {
    if (!ctx) return CB_ERR_NO_DEVICE;
    CK(cudaSetDevice(ctx->device));
    if (b->n_regions > 8) { ctx->err = "too many Bessel x regions"; return CB_ERR_UNSUPPORTED_PARAMS; }
    if (b->n_l < 1 || b->nx < 3) { ctx->err = "cb_build_bessel: empty table"; return CB_ERR_UNSUPPORTED_PARAMS; }
    if (b->x[b->nx - 1] < b->l[b->n_l - 1] + 1) { ctx->err = "cb_build_bessel: x grid ends below the largest l"; return CB_ERR_UNSUPPORTED_PARAMS; }
    const unsigned long long key = bessel_key(b, false);
    const size_t n_tab = (size_t)4 * (b->nx - 1) * b->n_l;
    for (auto& e : ctx->bess_cache)
        if (e.key == key) {
            e.stamp = ++ctx->bess_clock; ctx->bess = e.B; ctx->have_bessel = true;
            if (horner_out) {
                CK(cudaMemcpyAsync(horner_out, e.B.horner, sizeof(double) * n_tab, cudaMemcpyDeviceToHost, ctx->stream));
                CK(cudaStreamSynchronize(ctx->stream));
            }
            return 0;
        }
    if (ctx->bess_cache.size() >= 8) {
        size_t old = 0;
        for (size_t i = 1; i < ctx->bess_cache.size(); i++) if (ctx->bess_cache[i].stamp < ctx->bess_cache[old].stamp) old = i;
        CK(cudaStreamSynchronize(ctx->stream));
        for (void* p : ctx->bess_cache[old].owned) dev_free(ctx, p);
        ctx->bess_cache.erase(ctx->bess_cache.begin() + old);
    }
    ctx->bess_cache.emplace_back();
    cb_ctx::BessEntry& E = ctx->bess_cache.back();
    E.key = key; E.stamp = ++ctx->bess_clock;
    DBessel& B = E.B;
    B.n_l = b->n_l; B.nx = b->nx; B.n_regions = b->n_regions;
    B.n_l_all = b->n_l_all > 0 ? b->n_l_all : b->n_l;
    B.n_l_pad = (b->n_l + 3) & ~3;
    for (int i = 0; i < b->n_regions; i++) B.regions[i] = b->regions[i];
    B.xmax = b->x[b->nx - 1];
    int* dl = nullptr; double *dx = nullptr, *dh = nullptr, *dht = nullptr, *dwork = nullptr, *ddend = nullptr; int* dmin = nullptr; cb_range* dreg = nullptr;
    std::vector<void*> tmp;   // work buffers, released after the build
    int e = upload(ctx, b->l, (size_t)B.n_l_all, &dl, &E.owned);
    if (!e) e = upload(ctx, b->x, (size_t)b->nx, &dx, &E.owned);
    if (!e) e = upload(ctx, b->regions, (size_t)b->n_regions, &dreg, &tmp);
    auto alloc = [&](void** p, size_t bytes, std::vector<void*>* own) {
        if (e) return;
        if (dev_alloc(ctx, p, bytes) != cudaSuccess) { ctx->err = "out of device memory building the Bessel table"; e = CB_ERR_CUDA; return; }
        own->push_back(*p);
    };
    alloc((void**)&dh, sizeof(double) * n_tab, &E.owned);
    alloc((void**)&dht, sizeof(double) * 4 * (size_t)(b->nx - 1) * B.n_l_pad, &E.owned);
    alloc((void**)&dwork, sizeof(double) * 3 * (size_t)b->nx * b->n_l, &tmp);
    alloc((void**)&ddend, sizeof(double) * (size_t)b->n_l, &tmp);
    alloc((void**)&dmin, sizeof(int) * (size_t)b->n_l, &tmp);
    if (!e && cb_bessel_build_launch(dx, b->nx, dl, b->n_l, B.n_l_pad, dreg, b->n_regions, dwork, ddend, dmin, dh, dht, ctx->stream) != cudaSuccess) {
        ctx->err = "cb_build_bessel: launch failed"; e = CB_ERR_CUDA;
    }
    ctx->launches += 3;
    for (void* p : tmp) dev_free(ctx, p);
    if (e) { for (void* p : E.owned) dev_free(ctx, p); ctx->bess_cache.pop_back(); return e; }
    B.l = dl; B.x = dx; B.horner = dh; B.horner_t = dht;
    if (horner_out) CK(cudaMemcpyAsync(horner_out, dh, sizeof(double) * n_tab, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->bess = B;
    ctx->have_bessel = true;
    return 0;
}

int cb_model_create(cb_ctx* ctx, const cb_model* m, cb_dmodel** out) {
    *out = nullptr;
    if (!ctx) return CB_ERR_NO_DEVICE;
    CK(cudaSetDevice(ctx->device));
    // things the device path does not implement are errors, never a fallback
    if (m->grhok != 0) { ctx->err = "non-flat models are not supported by the GPU path"; return CB_ERR_UNSUPPORTED_PARAMS; }
    if (m->Scalar_initial_condition < 0 || m->Scalar_initial_condition > 5) { ctx->err = "Invalid initial condition for scalar modes"; return CB_ERR_UNSUPPORTED_PARAMS; }
    if (m->Scalar_initial_condition != 1 && m->omch2 == 0) { ctx->err = "Isocurvature initial conditions assume non-zero dark matter"; return CB_ERR_UNSUPPORTED_PARAMS; }
    if (m->n_tau_regions > 8) { ctx->err = "too many time-step regions"; return CB_ERR_UNSUPPORTED_PARAMS; }
    const int nwin = m->WantScalars ? m->num_redshiftwindows : 0;
    if (nwin < 0 || nwin > CB_MAX_WINDOWS || (nwin > 0 && !m->windows)) { ctx->err = "bad source-window count"; return CB_ERR_UNSUPPORTED_PARAMS; }
    if (m->WantScalars && m->n_src > 3 && m->n_src != 3 + nwin + m->num_extra_redshiftwindows) { ctx->err = "custom sources are not supported"; return CB_ERR_UNSUPPORTED_PARAMS; }
    for (int i = 0; i < nwin; i++)
        if (m->windows[i].kind != CB_WINDOW_COUNTS && m->windows[i].kind != CB_WINDOW_LENSING) { ctx->err = "21cm source windows are not supported"; return CB_ERR_UNSUPPORTED_PARAMS; }
    const bool is_tensor = !m->WantScalars && m->WantTensors;
    if (m->WantScalars && m->WantTensors) { ctx->err = "cmbmain cannot generate tensors and scalars in one pass (cmbmain.f90:142)"; return CB_ERR_UNSUPPORTED_PARAMS; }
    if (!m->WantScalars && !is_tensor) { ctx->err = "vector modes are not supported"; return CB_ERR_UNSUPPORTED_PARAMS; }
    if (is_tensor && (m->n_src != 3 || !m->adot || (m->thermo_horner && !m->dadot))) { ctx->err = "tensor models need n_src = 3 and the adot/dadot tables"; return CB_ERR_UNSUPPORTED_PARAMS; }
    cb_dmodel* dm = new cb_dmodel();
    DModel& D = dm->h;
    memset(&D, 0, sizeof(D));
    D.grhog = m->grhog; D.grhornomass = m->grhornomass; D.grhob = m->grhob; D.grhoc = m->grhoc; D.grhov = m->grhov; D.grhok = m->grhok;
    D.adotrad = m->adotrad; D.tau0 = m->tau0; D.tau_maxvis = m->tau_maxvis; D.taurst = m->taurst; D.taurend = m->taurend; D.H0 = m->H0;
    D.Nu_mass_eigenstates = m->Nu_mass_eigenstates; D.Num_Nu_massive = m->Num_Nu_massive;
    for (int i = 0; i < CB_MAX_NU; i++) { D.grhormass[i] = m->grhormass[i]; D.nu_masses[i] = m->nu_masses[i]; }
    D.WantCls = m->WantCls; D.WantTransfer = m->WantTransfer; D.Want_CMB = m->Want_CMB; D.WantTensors = m->WantTensors;
    D.DoLateRadTruncation = m->DoLateRadTruncation; D.MassiveNuMethod = m->MassiveNuMethod;
    D.transfer_high_precision = m->transfer_high_precision; D.transfer_accurate_massive_neutrinos = m->transfer_accurate_massive_neutrinos;
    D.AccuratePolarization = m->AccuratePolarization; D.AccurateReionization = m->AccurateReionization; D.AccuracyTarget = m->AccuracyTarget;
    D.AccuracyBoost = m->AccuracyBoost; D.lAccuracyBoost = m->lAccuracyBoost; D.IntTolBoost = m->IntTolBoost;
    D.TimeSwitchBoost = m->TimeSwitchBoost; D.BessIntBoost = m->BessIntBoost; D.LimberBoost = m->LimberBoost; D.Alens = m->Alens;
    D.de_is_cc = m->de_is_cosmological_constant; D.de_model = m->de_model; D.de_nperturb = m->de_num_perturb_equations;
    D.de_w = m->de_w; D.de_wa = m->de_wa; D.de_cs2 = m->de_cs2; D.de_c_gamma_ppf = m->de_c_gamma_ppf;
    D.nthermo = m->nthermo; D.tauminn = m->tauminn; D.dlntau = m->dlntau; D.dlntau_inv = 1.0 / m->dlntau;
    D.matter_verydom_tau = m->matter_verydom_tau;
    D.n_tau = m->n_tau; D.n_tau_regions = m->n_tau_regions;
    for (int i = 0; i < m->n_tau_regions; i++) D.tau_regions[i] = m->tau_regions[i];
    D.tau_lowest = m->tau_lowest; D.tau_highest = m->tau_highest;
    D.nqmax = m->nqmax;
    for (int i = 0; i < CB_MAX_NQ; i++) { D.nu_q[i] = m->nu_q[i]; D.nu_int_kernel[i] = m->nu_int_kernel[i]; }
    for (int i = 0; i < CB_MAX_NQ * CB_MAX_NU; i++) D.nu_tau_notmassless[i] = m->nu_tau_notmassless[i];
    for (int i = 0; i < CB_MAX_NU; i++) {
        D.nu_tau_nonrelativistic[i] = m->nu_tau_nonrelativistic[i];
        D.nu_tau_nonrelativistic_physical[i] = m->nu_tau_nonrelativistic_physical[i];
    }
    D.num_transfer_redshifts = m->WantTransfer ? m->num_transfer_redshifts : 0;
    D.max_etak_scalar = m->max_etak_scalar; D.max_etak_tensor = m->max_etak_tensor;
    D.WantLateTime = m->WantLateTime; D.n_src = m->n_src;
    D.is_tensor = is_tensor ? 1 : 0;
    D.Scalar_initial_condition = m->Scalar_initial_condition; D.omch2 = m->omch2; D.ombh2 = m->ombh2;
    for (int i = 0; i < 5; i++) D.InitialConditionVector[i] = m->InitialConditionVector[i];
    dm->n_cl = is_tensor ? 4 : (m->n_src > 2 ? 6 : 3);
    int e = 0;
    double* p;
    size_t nt = (size_t)m->nthermo;
#define UP(field, src, n) do { if ((e = upload(ctx, (src), (n), &p, &dm->owned))) { cb_model_destroy(ctx, dm); return e; } D.field = p; } while (0)
    if (m->thermo_horner) {
        UP(thermo_horner, m->thermo_horner, nt * 16);
        UP(cs2, m->cs2, nt); UP(dotmu, m->dotmu, nt); UP(ddotmu, m->ddotmu, nt); UP(dddotmu, m->dddotmu, nt);
        UP(ddddotmu, m->ddddotmu, nt); UP(emmu, m->emmu, nt); UP(demmu, m->demmu, nt);
        if (is_tensor) { UP(adot, m->adot, nt); UP(dadot, m->dadot, nt); }
    } else {
        // results.f90:2430-2494 on the device: five base arrays up, the derivative tables and the Horner records are
        // computed by thermo_pack.cu before the first kernel that reads them
        if (!m->cs2 || !m->dotmu || !m->emmu || !m->adot || !m->scale_factor || nt < 4) {
            cb_model_destroy(ctx, dm);
            ctx->err = "thermo_horner is NULL: cs2, dotmu, emmu, adot and scale_factor are needed to build the tables on the device";
            return CB_ERR_UNSUPPORTED_PARAMS;
        }
        UP(cs2, m->cs2, nt); UP(dotmu, m->dotmu, nt); UP(emmu, m->emmu, nt); UP(adot, m->adot, nt);
        ThermoPackArgs& A = dm->pack_args;
        A.n = (int)nt; A.dlntau = m->dlntau;
        A.cs2 = D.cs2; A.dotmu = D.dotmu; A.emmu = D.emmu; A.adot = D.adot;
        if ((e = upload(ctx, m->scale_factor, nt, &p, &dm->pack_tmp))) { cb_model_destroy(ctx, dm); return e; }
        A.sf = p;
        auto dalloc = [&](double** q, size_t n, std::vector<void*>* own) {
            if (e) return;
            if (dev_alloc(ctx, (void**)q, sizeof(double) * n) != cudaSuccess) { ctx->err = "out of device memory for the thermo tables"; e = CB_ERR_CUDA; return; }
            own->push_back(*q);
        };
        dalloc(&A.ddotmu, nt, &dm->owned); dalloc(&A.dddotmu, nt, &dm->owned); dalloc(&A.ddddotmu, nt, &dm->owned);
        dalloc(&A.demmu, nt, &dm->owned); dalloc(&A.dadot, nt, &dm->owned); dalloc(&A.horner, nt * 16, &dm->owned);
        dalloc(&A.dcs2, nt, &dm->pack_tmp); dalloc(&A.g, nt * 4, &dm->pack_tmp);
        if (e) { cb_model_destroy(ctx, dm); return e; }
        D.thermo_horner = A.horner; D.ddotmu = A.ddotmu; D.dddotmu = A.dddotmu; D.ddddotmu = A.ddddotmu; D.demmu = A.demmu; D.dadot = A.dadot;
        dm->needs_thermo_pack = true;
    }
    UP(tau, m->tau, (size_t)m->n_tau); UP(dtau, m->dtau, (size_t)m->n_tau);
    if (D.num_transfer_redshifts > 0) UP(transfer_times, m->transfer_times, (size_t)D.num_transfer_redshifts);
#undef UP
    D.num_redshiftwindows = nwin; D.num_extra_redshiftwindows = nwin > 0 ? m->num_extra_redshiftwindows : 0;
    D.SourceTerms = m->SourceTerms; D.tau_start_redshiftwindows = m->tau_start_redshiftwindows; D.qmax = m->qmax;
    D.Want_CMB_lensing = m->Want_CMB_lensing; D.max_bessels_l_index = m->max_bessels_l_index;
    for (int i = 0; i < 3 + CB_MAX_WINDOWS; i++) D.Limber_l_min[i] = nwin > 0 || m->SourceTerms.limber_windows ? m->Limber_l_min[i] : 0;
    for (int i = 0; i < nwin; i++) {
        cb_window w = m->windows[i];
        const double** fields[] = {&w.wing, &w.wing2, &w.dwing2, &w.ddwing2, &w.wingtau, &w.dwingtau, &w.winV, &w.winF, &w.win_lens, &w.comoving_density_ev};
        for (const double** f : fields) {
            double* dp;
            if ((e = upload(ctx, *f, (size_t)m->n_tau, &dp, &dm->owned))) { cb_model_destroy(ctx, dm); return e; }
            *f = dp;
        }
        D.windows[i] = w;
    }
    // the copies are stream-ordered before every later call; waiting here only frees the caller's arrays for reuse
    if (!ctx->async_model_upload) CK(cudaStreamSynchronize(ctx->stream));
    *out = dm;
    return 0;
}

void cb_model_destroy(cb_ctx* ctx, cb_dmodel* dm) {
    if (!dm) return;
    if (ctx) cudaSetDevice(ctx->device);
    if (!ctx) { delete dm; return; }
    flush_outputs(ctx);
    if (ctx->copies_pending) { cudaEventSynchronize(ctx->ev_copy_done); ctx->copies_pending = false; }   // a result copy may still read this model's buffers
    for (void* p : dm->owned) dev_free(ctx, p);
    for (void* p : dm->pack_tmp) dev_free(ctx, p);
    void* bufs[] = {dm->d_q_src, dm->d_src, dm->d_dd, dm->d_scratch, dm->d_delta, dm->d_icl, dm->d_q_int, dm->d_dq_int,
                    dm->d_pows, dm->d_transfer, dm->d_stats, dm->d_icl_array, dm->d_limb, dm->d_lpow, dm->d_icl_w};
    for (void* p : bufs) dev_free(ctx, p);
    delete dm;
}

static int push_models(cb_ctx* ctx, int n_models, cb_dmodel* const* dm) {
    if (int e = ensure(ctx, &ctx->d_models, &ctx->cap_models, (size_t)n_models)) return e;
    if ((size_t)n_models > ctx->cap_h_models) {
        // the previous contents are not in flight: every entry point ends with a stream synchronisation
        if (ctx->h_models) cudaFreeHost(ctx->h_models);
        ctx->h_models = nullptr;
        CK(cudaHostAlloc((void**)&ctx->h_models, sizeof(DModel) * n_models, cudaHostAllocDefault));
        ctx->cap_h_models = n_models;
    }
    for (int i = 0; i < n_models; i++) ctx->h_models[i] = dm[i]->h;
    CK(cudaMemcpyAsync(ctx->d_models, ctx->h_models, sizeof(DModel) * n_models, cudaMemcpyHostToDevice, ctx->stream));
    return 0;
}

}  // extern "C"

// The models of this call whose thermo tables are still to be derived (cb_model.thermo_horner == NULL): one launch for all.
static int pack_thermo_tables(cb_ctx* ctx, int n_models, cb_dmodel* const* dm) {
    std::vector<ThermoPackArgs> args;
    int max_n = 0;
    for (int m = 0; m < n_models; m++)
        if (dm[m]->needs_thermo_pack) { args.push_back(dm[m]->pack_args); max_n = std::max(max_n, dm[m]->pack_args.n); }
    if (args.empty()) return 0;
    ThermoPackArgs* d_args = nullptr;
    CK(dev_alloc(ctx, (void**)&d_args, sizeof(ThermoPackArgs) * args.size()));
    CK(cudaMemcpyAsync(d_args, args.data(), sizeof(ThermoPackArgs) * args.size(), cudaMemcpyHostToDevice, ctx->stream));   // pageable: staged before return
    CK(cb_thermo_pack_launch(d_args, (int)args.size(), max_n, ctx->stream));
    ctx->launches += 2;
    dev_free(ctx, d_args);
    for (int m = 0; m < n_models; m++)
        if (dm[m]->needs_thermo_pack) {
            for (void* p : dm[m]->pack_tmp) dev_free(ctx, p);
            dm[m]->pack_tmp.clear();
            dm[m]->needs_thermo_pack = false;
        }
    return 0;
}

extern "C" {

int cb_evolve(cb_ctx* ctx, int n_models, cb_dmodel* const* dm, const cb_modes* modes, double* const* LinearSrc,
              float* const* TransferData, long long* const* stats) {
    if (!ctx) return CB_ERR_NO_DEVICE;
    CK(cudaSetDevice(ctx->device));
    if (int e = begin_call(ctx)) return e;
    if (int e = flush_outputs(ctx)) return e;   // the previous batch's results travel to the host under this batch's kernels
    if (int e = pack_thermo_tables(ctx, n_models, dm)) return e;
    std::vector<DModeItem> cls_items[CB_N_CLASSES], all_items;
    int lean = 1;   // every model massless-nu LCDM: the kernel variants without massive-nu / dark-energy code
    for (int m = 0; m < n_models; m++) if (dm[m]->h.Num_Nu_massive != 0 || !dm[m]->h.de_is_cc || dm[m]->h.num_redshiftwindows > 0 || dm[m]->h.is_tensor) lean = 0;
    int cls_nv[CB_N_CLASSES] = {}, cls_maxl[CB_N_CLASSES] = {};
    std::vector<StagePiece> pieces;
    std::vector<size_t> item_off((size_t)n_models + 1, 0);
    std::vector<int> src_nk((size_t)n_models, -1);
    {
        size_t tot = 0;
        for (int m = 0; m < n_models; m++) tot += (size_t)std::max(modes[m].n_k, 0);
        if (int e = stage_reserve(ctx, tot, 0)) return e;
    }
    for (int m = 0; m < n_models; m++) {
        cb_dmodel* d = dm[m];
        const cb_modes& md = modes[m];
        DModel& D = d->h;
        const int n_k = md.n_k, nz = D.num_transfer_redshifts;
        if (!md.transfer_only) {
            if (int e = ensure(ctx, &d->d_q_src, &d->cap_k, (size_t)n_k)) return e;
            size_t ns = (size_t)n_k * D.n_src * D.n_tau;
            if (ns > d->cap_src) {
                size_t c1 = 0, c2 = 0, c3 = 0;
                if (d->d_src) { dev_free(ctx, d->d_src); d->d_src = nullptr; }
                if (d->d_dd) { dev_free(ctx, d->d_dd); d->d_dd = nullptr; }
                if (d->d_scratch) { dev_free(ctx, d->d_scratch); d->d_scratch = nullptr; }
                if (int e = ensure(ctx, &d->d_src, &c1, ns)) return e;
                if (int e = ensure(ctx, &d->d_dd, &c2, ns)) return e;
                if (int e = ensure(ctx, &d->d_scratch, &c3, ns)) return e;
                d->cap_src = ns;
            }
            stage_add(ctx, pieces, d->d_q_src, md.q, (size_t)n_k);
            // row tau_1 is never written by the reference (allocated zero, cmbmain.f90:777-778) and the rows
            // after a mode's k*tau early-out are zero (:1034-1037): the kernel writes only live entries
            CK(cudaMemsetAsync(d->d_src, 0, sizeof(double) * ns, ctx->stream));
            D.n_k = n_k; D.q_src = d->d_q_src; D.LinearSrc = d->d_src; D.ddSrc = d->d_dd;
        } else {
            // loop A' (GetTransfer) on a model that already holds sources: the kernels of THIS call index TransferData by the
            // transfer-only mode count; the model's own n_k (the source wavenumbers loop B splines over) comes back after
            // the launch descriptors have been staged
            src_nk[m] = d->d_src ? D.n_k : -1;
            D.n_k = n_k;
        }
        if (nz > 0) {
            size_t ntr = (size_t)13 * n_k * nz;
            if (ntr > d->cap_tr) { if (d->d_transfer) dev_free(ctx, d->d_transfer); d->d_transfer = nullptr; d->cap_tr = 0; }
            if (int e = ensure(ctx, &d->d_transfer, &d->cap_tr, ntr)) return e;
            D.TransferData = d->d_transfer;
        } else D.TransferData = nullptr;
        if (int e = ensure(ctx, &d->d_stats, &d->cap_stats, (size_t)4 * n_k)) return e;
        D.stats = d->d_stats;
        // k-sharding (cb_modes.active): this call evolves a subset; rows of the other modes stay zero
        const bool subset = md.n_active > 0 && md.active;
        const int n_act = md.n_active < 0 ? 0 : (subset ? md.n_active : n_k);   // n_active < 0: a rank with an empty share evolves nothing
        if (md.n_active < 0) {
            CK(cudaMemsetAsync(d->d_stats, 0, sizeof(long long) * 4 * (size_t)n_k, ctx->stream));
            if (nz > 0) CK(cudaMemsetAsync(d->d_transfer, 0, sizeof(float) * (size_t)13 * n_k * nz, ctx->stream));
        }
        if (subset) {
            CK(cudaMemsetAsync(d->d_stats, 0, sizeof(long long) * 4 * (size_t)n_k, ctx->stream));
            if (nz > 0) CK(cudaMemsetAsync(d->d_transfer, 0, sizeof(float) * (size_t)13 * n_k * nz, ctx->stream));
        }
        for (int a = 0; a < n_act; a++) {
            const int i = subset ? md.active[a] : a;
            if (i < 0 || i >= n_k) { ctx->err = "cb_evolve: cb_modes.active index out of range"; return CB_ERR_UNSUPPORTED_PARAMS; }
        }
        item_off[m + 1] = item_off[m] + (size_t)n_act;
    }
    // The work items (GetNumEqns per mode, equations.f90:860-1070, on the host: a few hundred ns each, 14 000 of them per
    // 64 parameter sets) are filled by the host's cores model by model.
    all_items.resize(item_off[n_models]);
    int size_err = 0;
#pragma omp parallel for schedule(dynamic, 4) reduction(|| : size_err)
    for (int m = 0; m < n_models; m++) {
        const cb_modes& md = modes[m];
        const DModel& D = dm[m]->h;
        const bool subset = md.n_active > 0 && md.active;
        const int n_act = (int)(item_off[m + 1] - item_off[m]);
        for (int a = 0; a < n_act; a++) {
            const int i = subset ? md.active[a] : a;
            DModeItem& it = all_items[item_off[m] + a];
            memset(&it, 0, sizeof(it));
            it.model = m; it.q_ix = i; it.transfer_only = md.transfer_only; it.q = md.q[i];
            it.tight_switchoff = md.tight_switchoff[i];
            for (int j = 0; j < CB_MAX_NU; j++) it.nu_massive_approx_time[j] = md.nu_massive_approx_time[(size_t)i * CB_MAX_NU + j];
            int ml, lm3[3] = {0, 0, 0};
            if (D.is_tensor) cb_tensor_mode_size(D, &it.nvar, &ml);
            else cb_mode_size(D, it.q, it.transfer_only, &it.nvar, &ml, lm3);
            it.lmaxg = lm3[0]; it.lmaxgpol = lm3[1]; it.lmaxnr = lm3[2];
            if (ml > CBD_MAX_L_EVOLVE) size_err = 1;
            it.transfer_only |= ml << 8;   // deepest multipole rides along until the binning pass below
        }
    }
    if (size_err) { ctx->err = "Need to increase max_l_evolve"; return CB_ERR_UNSUPPORTED_PARAMS; }
    // Binning pass, most expensive modes first (largest k: cmbmain.f90:205-206 runs them first too): a mode joins the first
    // class whose slab, grown to hold it, stays inside that class's limit -- so the high-k bulk (45-ish equations, 14
    // multipoles) keeps the smallest slab and the few deep low-k hierarchies land in the larger classes.
    {   // stable order on a light key array, then one gather of the (large) items
        std::vector<unsigned> ord(all_items.size());
        for (size_t i = 0; i < ord.size(); i++) ord[i] = (unsigned)i;
        std::stable_sort(ord.begin(), ord.end(), [&](unsigned a, unsigned b) { return all_items[a].q > all_items[b].q; });
        std::vector<DModeItem> sorted(all_items.size());
        for (size_t i = 0; i < ord.size(); i++) sorted[i] = all_items[ord[i]];
        all_items.swap(sorted);
    }
    // A launch of few modes is a latency problem, not a throughput one: the register kernel shares a warp between eight modes
    // (31 ms for one cosmology), the general kernel gives every mode a warp of its own there (22 ms).
    const bool few_modes = all_items.size() < (size_t)(ctx->reg_min_modes >= 0 ? ctx->reg_min_modes : 8 * ctx->n_sm);
    for (DModeItem& it : all_items) {
        const int ml = it.transfer_only >> 8;
        it.transfer_only &= 0xff;
        int c = 0;
        int rc = (lean && ctx->use_reg_kernel && !few_modes) ? cb_reg_class(it.lmaxg, it.lmaxgpol, it.lmaxnr) : -1;
        if (rc == 1 && ctx->use_reg_kernel == 2) rc = -1;   // experiment: the 16-lane layout's modes through the general kernel
        if (lean && ctx->use_reg_kernel == 3) {
            const int mc = cb_mpl_class(it.lmaxg, it.lmaxgpol, it.lmaxnr);
            if (mc >= 0) { cls_items[mc == 0 ? CB_MPL_CLASS_A : CB_MPL_CLASS_C].push_back(it); continue; }
        }
        if (dm[it.model]->h.is_tensor) c = CB_TENSOR_CLASS;
        else if (rc >= 0) { cls_items[CB_REG_CLASS0 + rc].push_back(it); continue; }
        else
            for (;; c++) {
                const size_t grown = cb_evolve_slab_bytes(std::max(std::max(cls_nv[c], it.nvar), 32), std::max(cls_maxl[c], ml), lean);
                if (grown <= kClassSlabLimit[c]) break;
            }
        cls_nv[c] = std::max(cls_nv[c], it.nvar);
        cls_maxl[c] = std::max(cls_maxl[c], ml);
        cls_items[c].push_back(it);
    }
    int n_items = 0;
    for (int c = 0; c < CB_N_CLASSES; c++) n_items += (int)cls_items[c].size();
    if (n_items == 0) {
        for (int m = 0; m < n_models; m++) if (src_nk[m] >= 0) dm[m]->h.n_k = src_nk[m];
        return 0;
    }
    if (int e = stage_upload(ctx, pieces)) return e;
    std::vector<DModeItem> items;
    items.reserve(n_items);
    int cls_off[CB_N_CLASSES];
    for (int c = 0; c < CB_N_CLASSES; c++) {
        // already largest k first (cmbmain.f90:205-206: descending q with dynamic scheduling): the binning pass kept the order
        cls_off[c] = (int)items.size();
        items.insert(items.end(), cls_items[c].begin(), cls_items[c].end());
    }
    if (int e = push_models(ctx, n_models, dm)) return e;
    for (int m = 0; m < n_models; m++) if (src_nk[m] >= 0) dm[m]->h.n_k = src_nk[m];
    if (int e = ensure(ctx, &ctx->d_items, &ctx->cap_items, (size_t)n_items)) return e;
    if ((size_t)n_items > ctx->cap_h_items) {
        if (ctx->h_items) cudaFreeHost(ctx->h_items);
        ctx->h_items = nullptr;
        CK(cudaHostAlloc((void**)&ctx->h_items, sizeof(DModeItem) * n_items, cudaHostAllocDefault));
        ctx->cap_h_items = n_items;
    }
    memcpy(ctx->h_items, items.data(), sizeof(DModeItem) * n_items);
    CK(cudaMemcpyAsync(ctx->d_items, ctx->h_items, sizeof(DModeItem) * n_items, cudaMemcpyHostToDevice, ctx->stream));
#ifdef CB_TRACE
    if (const char* tq = getenv("CAMB_B200_TRACE")) { cb_trace_mpl_begin(atoi(tq), 4096); cb_trace_reg_begin(atoi(tq), 4096); }
#endif
    CK(cudaMemsetAsync(ctx->d_counter, 0, sizeof(unsigned int) * CB_N_CLASSES, ctx->stream));
    if (!ctx->error_pending) CK(cudaMemsetAsync(ctx->d_error, 0, sizeof(int), ctx->stream));   // an unreported failure stays set
    int lanes = ctx->lanes_per_mode;
    if (lanes != 4 && lanes != 8 && lanes != 16 && lanes != 32) lanes = n_items >= 16 * ctx->n_sm ? 8 : 32;
    CK(cudaEventRecord(ctx->ev_staged, ctx->stream));   // every host->device copy of this call is queued before this point
    ctx->staged_pending = true;
    CK(cudaEventRecord(ctx->stage_ev[0][0], ctx->stream));
    CK(cudaEventRecord(ctx->fork, ctx->stream));
    // Launch order and block caps of the slab classes.  The scheduling experiments of DESIGN.md 4.1 (profiles/r01o, r01p)
    // are compiled in only with -DCB_TUNING_KNOBS: CAMB_B200_ORDER, CAMB_B200_CAPS, CAMB_B200_ONLY_CLASS.
    int order_desc = 0, caps[CB_N_CLASSES] = {};
#ifdef CB_TUNING_KNOBS
    if (const char* o = getenv("CAMB_B200_ORDER")) order_desc = atoi(o);
    if (const char* cp = getenv("CAMB_B200_CAPS")) {
        int i = 0;
        for (const char* p = cp; *p && i < CB_N_CLASSES; i++) { caps[i] = atoi(p); while (*p && *p != ',') p++; if (*p == ',') p++; }
    }
#endif
    for (int cc = 0; cc < CB_N_CLASSES; cc++) ctx->last_grid[cc] = 0;
    for (int cc = 0; cc < CB_N_CLASSES; cc++) {
        // the register classes go first: their modes are the longest serial chains
        // register layouts: the lowest-k class, the high-k bulk, the deep low-k hierarchies last (each launch is one block per SM
        // now, so the classes follow one another on an SM: of the six orders this one is 3 % ahead, profiles/r02zzv_class_order.log)
        static const int kOrder[CB_N_CLASSES] = {CB_MPL_CLASS_A, CB_MPL_CLASS_C, CB_REG_CLASS0 + 2, CB_REG_CLASS0, CB_REG_CLASS0 + 1, 0, 1, 2, 3, 4, CB_TENSOR_CLASS};
        int c = order_desc ? CB_N_CLASSES - 1 - cc : kOrder[cc];
#ifdef CB_TUNING_KNOBS
        if (const char* ro = getenv("CAMB_B200_REG_ORDER"))   // launch order of the three register layouts, e.g. "0,2,1"
            if (cc >= 2 && cc <= 4 && strlen(ro) >= 5) c = CB_REG_CLASS0 + (ro[2 * (cc - 2)] - '0');
#endif
        int n_c = (int)cls_items[c].size();
        if (c == CB_MPL_CLASS_C) continue;                                   // launched together with class A
        if (c == CB_MPL_CLASS_A) n_c += (int)cls_items[CB_MPL_CLASS_C].size();
        if (n_c == 0) continue;
#ifdef CB_TUNING_KNOBS
        if (const char* only = getenv("CAMB_B200_ONLY_CLASS")) if (atoi(only) != c) continue;   // timing experiments: one slab class alone
#endif
        // the scratch EvolutionVars of a layout switch aliases two stage vectors (evolve.cu)
        const int nv = std::max(cls_nv[c], 32);
        // Lane width of a general-kernel class (measured on the massive-neutrino batches, profiles/r02z_lanes.log): a class with
        // few modes for the machine, or one whose slab leaves room for a single eight-mode block per SM, runs a mode per warp
        // (the elements of a deep hierarchy spread over 32 lanes, 5 two-mode blocks per SM instead of one eight-mode block)
        int lanes_c = ctx->class_lanes[c] ? ctx->class_lanes[c] : lanes;
        if (!ctx->class_lanes[c] && ctx->lanes_per_mode == 0 && c < CB_TENSOR_CLASS &&
            (n_c < 24 * ctx->n_sm || cb_evolve_slab_bytes(nv, cls_maxl[c], lean) > 16384))
            lanes_c = 32;
        if (getenv("CAMB_B200_VERBOSE"))
            fprintf(stderr, "cb_evolve: class %d: %d modes, nv %d, maxl %d, slab %zu bytes\n", c, n_c, nv, cls_maxl[c],
                    c < CB_TENSOR_CLASS ? cb_evolve_slab_bytes(nv, cls_maxl[c], lean) : (size_t)0);
        CK(cudaStreamWaitEvent(ctx->class_stream[c], ctx->fork, 0));
        cudaError_t le = c == CB_MPL_CLASS_A
                             ? cb_evolve_mpl_launch(ctx->d_models, ctx->d_items + cls_off[CB_MPL_CLASS_A], (int)cls_items[CB_MPL_CLASS_A].size(),
                                                    ctx->d_items + cls_off[CB_MPL_CLASS_C], (int)cls_items[CB_MPL_CLASS_C].size(),
                                                    ctx->d_counter + CB_MPL_CLASS_A, ctx->d_counter + CB_MPL_CLASS_C, ctx->d_error, ctx->n_sm,
                                                    ctx->class_stream[c], &ctx->last_grid[c])
                         : c >= CB_REG_CLASS0
                             ? cb_evolve_reg_launch(c - CB_REG_CLASS0, ctx->d_models, ctx->d_items + cls_off[c], n_c, ctx->d_counter + c, ctx->d_error,
                                                    ctx->n_sm, ctx->class_stream[c], &ctx->last_grid[c])
                         : c == CB_TENSOR_CLASS
                             ? cb_evolve_tensor_launch(ctx->d_models, ctx->d_items + cls_off[c], n_c, cls_nv[c], cls_maxl[c], ctx->d_counter + c,
                                                       ctx->d_error, ctx->n_sm, ctx->class_stream[c], &ctx->last_grid[c])
                             : cb_evolve_launch(ctx->d_models, ctx->d_items + cls_off[c], n_c, nv, cls_maxl[c], lanes_c, lean, ctx->d_counter + c,
                                                ctx->d_error, ctx->n_sm, ctx->class_stream[c], &ctx->last_grid[c], caps[c]);
        if (le != cudaSuccess) {
            for (int j = 0; j < CB_N_CLASSES; j++) cudaStreamSynchronize(ctx->class_stream[j]);   // classes already launched finish first
            ctx->err = std::string("evolve launch: ") + cudaGetErrorString(le);
            return CB_ERR_CUDA;
        }
        ctx->launches++;
        CK(cudaEventRecord(ctx->class_done[c], ctx->class_stream[c]));
        CK(cudaStreamWaitEvent(ctx->stream, ctx->class_done[c], 0));
    }
    CK(cudaEventRecord(ctx->stage_ev[0][1], ctx->stream));
    bool host_out = false;
    for (int m = 0; m < n_models; m++)
        host_out = host_out || (LinearSrc && LinearSrc[m]) || (TransferData && TransferData[m]) || (stats && stats[m]);
    if (ctx->async_calls && !host_out && ctx->h_error) {   // nothing for the host to read: queued is done
        CK(cudaMemcpyAsync(ctx->h_error, ctx->d_error, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        ctx->error_pending = true;
        return 0;
    }
    int herr = 0;
    CK(cudaMemcpyAsync(&herr, ctx->d_error, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    for (int m = 0; m < n_models; m++) {
        cb_dmodel* d = dm[m];
        const DModel& D = d->h;
        const size_t call_nk = (size_t)modes[m].n_k;   // the mode count of THIS call (a transfer-only call leaves D.n_k at the source count)
        if (LinearSrc && LinearSrc[m] && !modes[m].transfer_only)
            CK(cudaMemcpyAsync(LinearSrc[m], d->d_src, sizeof(double) * call_nk * D.n_src * D.n_tau, cudaMemcpyDeviceToHost, ctx->stream));
        if (TransferData && TransferData[m] && D.num_transfer_redshifts > 0)
            CK(cudaMemcpyAsync(TransferData[m], d->d_transfer, sizeof(float) * (size_t)13 * call_nk * D.num_transfer_redshifts, cudaMemcpyDeviceToHost, ctx->stream));
        if (stats && stats[m])
            CK(cudaMemcpyAsync(stats[m], d->d_stats, sizeof(long long) * 4 * call_nk, cudaMemcpyDeviceToHost, ctx->stream));
    }
    CK(cudaStreamSynchronize(ctx->stream));
#ifdef CB_TRACE
    if (getenv("CAMB_B200_TRACE")) {
        cb_trace_mpl_end(ctx->use_reg_kernel == 3 ? "gpurun_out/trace_mpl.bin" : "gpurun_out/trace_mpl_unused.bin");
        cb_trace_reg_end(ctx->use_reg_kernel == 1 ? "gpurun_out/trace_reg.bin" : "gpurun_out/trace_reg_unused.bin");
    }
#endif
    ctx->staged_pending = false;
    if (herr) { ctx->error_pending = false; ctx->err = "RungeKuttaDP45 error: step size underflow in the device integrator"; return herr; }
    return deferred_error(ctx);
}

// MakeNonlinearSources (cmbmain.f90:1190-1247) on the device-resident sources of the last cb_evolve, in place.
// nonlin_ratio: CAMB_PK%nonlin_ratio(n_k, n_z), wavenumber fastest, n_z = num_transfer_redshifts of the model.
int cb_make_nonlinear_sources(cb_ctx* ctx, cb_dmodel* dm, int n_k, int n_z, const double* nonlin_ratio, double Min_kh_nonlinear) {
    if (!ctx || !dm) return CB_ERR_NO_DEVICE;
    CK(cudaSetDevice(ctx->device));
    if (int e = begin_call(ctx)) return e;
    DModel& D = dm->h;
    if (!dm->d_src || n_k != D.n_k) { ctx->err = "cb_make_nonlinear_sources: call cb_evolve on this model first"; return CB_ERR_UNSUPPORTED_PARAMS; }
    if (D.is_tensor || n_z < 2 || n_z != D.num_transfer_redshifts || !D.transfer_times) {
        ctx->err = "cb_make_nonlinear_sources: the model has no transfer output times to interpolate the ratios over";
        return CB_ERR_UNSUPPORTED_PARAMS;
    }
    if (D.n_src < 3) return 0;   // nothing to scale (cmbmain.f90:1240 scales sources 3..)
    double* d_ratio = nullptr;
    DModel* d_model = nullptr;
    CK(dev_alloc(ctx, (void**)&d_ratio, sizeof(double) * (size_t)n_k * n_z));
    CK(dev_alloc(ctx, (void**)&d_model, sizeof(DModel)));
    CK(cudaMemcpyAsync(d_ratio, nonlin_ratio, sizeof(double) * (size_t)n_k * n_z, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(d_model, &D, sizeof(DModel), cudaMemcpyHostToDevice, ctx->stream));
    cudaError_t le = cb_nonlinear_sources_launch(d_model, n_k, d_ratio, n_z, Min_kh_nonlinear, ctx->stream);
    ctx->launches++;
    dev_free(ctx, d_ratio); dev_free(ctx, d_model);
    if (le != cudaSuccess) { ctx->err = std::string("cb_make_nonlinear_sources: ") + cudaGetErrorString(le); return CB_ERR_CUDA; }
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->staged_pending = false;
    if (int e = deferred_error(ctx)) return e;
    return 0;
}

int cb_set_sources(cb_ctx* ctx, cb_dmodel* d, int n_k, const double* q, const double* ScaledSrc) {
    if (!ctx) return CB_ERR_NO_DEVICE;
    CK(cudaSetDevice(ctx->device));
    if (int e = begin_call(ctx)) return e;
    DModel& D = d->h;
    size_t ns = (size_t)n_k * D.n_src * D.n_tau;
    if (ns > d->cap_src || !d->d_src) {
        size_t c1 = 0, c2 = 0, c3 = 0;
        if (d->d_src) { dev_free(ctx, d->d_src); d->d_src = nullptr; }
        if (d->d_dd) { dev_free(ctx, d->d_dd); d->d_dd = nullptr; }
        if (d->d_scratch) { dev_free(ctx, d->d_scratch); d->d_scratch = nullptr; }
        if (int e = ensure(ctx, &d->d_src, &c1, ns)) return e;
        if (int e = ensure(ctx, &d->d_dd, &c2, ns)) return e;
        if (int e = ensure(ctx, &d->d_scratch, &c3, ns)) return e;
        d->cap_src = ns;
    }
    if (int e = ensure(ctx, &d->d_q_src, &d->cap_k, (size_t)n_k)) return e;
    CK(cudaMemcpyAsync(d->d_q_src, q, sizeof(double) * n_k, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(d->d_src, ScaledSrc, sizeof(double) * ns, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->staged_pending = false;
    if (int e = deferred_error(ctx)) return e;
    D.n_k = n_k; D.q_src = d->d_q_src; D.LinearSrc = d->d_src; D.ddSrc = d->d_dd;
    return 0;
}

int cb_project(cb_ctx* ctx, int n_models, cb_dmodel* const* dm, const int* n_q, const double* const* q_int,
               const double* const* dq, double* const* Delta_p_l_k) {
    if (!ctx) return CB_ERR_NO_DEVICE;
    CK(cudaSetDevice(ctx->device));
    if (!ctx->have_bessel) { ctx->err = "cb_set_bessel must be called before cb_project"; return CB_ERR_UNSUPPORTED_PARAMS; }
    if (int e = begin_call(ctx)) return e;
    // a Delta buffer may be rewritten below: the copies of an earlier call on these models go first
    for (int m = 0; m < n_models && !ctx->deferred.empty(); m++)
        for (const cb_ctx::DeferredCopy& c : ctx->deferred)
            if (c.src == dm[m]->d_delta) { if (int e = flush_outputs(ctx)) return e; break; }
    if (ctx->copies_pending) CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_copy_done, 0));
    const int n_l = ctx->bess.n_l;
    int max_cols = 0, max_nq = 0, max_ntau = 0, max_nsrc = 0;
    std::vector<double*> scr(n_models);
    std::vector<StagePiece> pieces;
    {
        size_t tot = 0;
        for (int m = 0; m < n_models; m++) tot += 2 * (size_t)std::max(n_q[m], 0);
        if (int e = stage_reserve(ctx, tot, 0)) return e;
    }
    for (int m = 0; m < n_models; m++) {
        cb_dmodel* d = dm[m];
        DModel& D = d->h;
        if (!d->d_src || D.n_k <= 0) { ctx->err = "cb_project: no sources (call cb_evolve or cb_set_sources first)"; return CB_ERR_UNSUPPORTED_PARAMS; }
        const int nq = n_q[m];
        if ((size_t)nq > d->cap_q) {
            dev_free(ctx, d->d_q_int);
            dev_free(ctx, d->d_dq_int);
            dev_free(ctx, d->d_pows);
            d->d_q_int = d->d_dq_int = d->d_pows = nullptr;
            size_t c1 = 0, c2 = 0, c3 = 0;
            if (int e = ensure(ctx, &d->d_q_int, &c1, (size_t)nq)) return e;
            if (int e = ensure(ctx, &d->d_dq_int, &c2, (size_t)nq)) return e;
            if (int e = ensure(ctx, &d->d_pows, &c3, (size_t)nq)) return e;
            d->cap_q = nq;
        }
        size_t nd = (size_t)D.n_src * n_l * nq;
        if (nd > d->cap_delta) { if (d->d_delta) dev_free(ctx, d->d_delta); d->d_delta = nullptr; d->cap_delta = 0; }
        if (int e = ensure(ctx, &d->d_delta, &d->cap_delta, nd)) return e;
        stage_add(ctx, pieces, d->d_q_int, q_int[m], (size_t)nq);
        stage_add(ctx, pieces, d->d_dq_int, dq[m], (size_t)nq);
        D.n_q = nq; D.q_int = d->d_q_int; D.dq_int = d->d_dq_int; D.pows = d->d_pows; D.Delta = d->d_delta;
        max_cols = std::max(max_cols, D.n_src * D.n_tau);
        max_nq = std::max(max_nq, nq); max_ntau = std::max(max_ntau, D.n_tau); max_nsrc = std::max(max_nsrc, D.n_src);
        scr[m] = d->d_scratch;
    }
    if (int e = stage_upload(ctx, pieces)) return e;
    if (int e = push_models(ctx, n_models, dm)) return e;
    if (int e = ensure(ctx, &ctx->d_scratch_ptrs, &ctx->cap_scr, (size_t)n_models)) return e;
    CK(cudaMemcpyAsync(ctx->d_scratch_ptrs, scr.data(), sizeof(double*) * n_models, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaEventRecord(ctx->ev_staged, ctx->stream));
    ctx->staged_pending = true;
    CK(cudaEventRecord(ctx->stage_ev[1][0], ctx->stream));
    cb_spline_k_launch(ctx->d_models, n_models, max_cols, ctx->d_scratch_ptrs, ctx->stream);
    ctx->launches++;
    CK(cudaGetLastError());
    CK(cudaEventRecord(ctx->stage_ev[1][1], ctx->stream));
    bool any_windows = false;
    for (int m = 0; m < n_models; m++) any_windows = any_windows || dm[m]->h.num_redshiftwindows > 0;
    const size_t smem = cb_los_smem(max_nsrc, max_ntau, n_l, 1);   // (the staged kernels; the windows kernel sizes its own)
    CK(cudaEventRecord(ctx->stage_ev[2][0], ctx->stream));
    if (max_nq == 0) {
        // an empty slice of the integration grid (a rank with no share of a q-sharded run): nothing to project
    } else if (any_windows || smem > 227 * 1024 || n_l > 192 /* LOS_MAX_L */) {
        // source windows (or a time grid too long to stage the sources): the general kernel
        for (int m = 0; m < n_models; m++)
            if (dm[m]->h.is_tensor) { ctx->err = "LOS: tensor models cannot share a call with source-window models"; return CB_ERR_UNSUPPORTED_PARAMS; }
        cudaError_t le = cb_los_win_launch(ctx->d_models, n_models, ctx->bess, max_nq, max_ntau, ctx->stream);
        if (le != cudaSuccess) { ctx->err = std::string("LOS (windows): ") + cudaGetErrorString(le); return le == cudaErrorInvalidValue ? CB_ERR_UNSUPPORTED_PARAMS : CB_ERR_CUDA; }
    } else {
        cudaError_t le = cb_los_launch(ctx->d_models, n_models, ctx->bess, max_nq, max_nsrc, max_ntau, ctx->los_kernel, ctx->stream);
        if (le != cudaSuccess) { ctx->err = std::string("LOS: ") + cudaGetErrorString(le); return CB_ERR_CUDA; }
    }
    ctx->launches++;
    CK(cudaGetLastError());
    CK(cudaEventRecord(ctx->stage_ev[2][1], ctx->stream));
    {
        bool any = false;
        for (int m = 0; m < n_models; m++) any = any || (Delta_p_l_k && Delta_p_l_k[m]);
        const bool defer = any && ctx->async_outputs;   // queued now, issued by flush_outputs
        for (int m = 0; m < n_models; m++) {
            const DModel& D = dm[m]->h;
            if (!(Delta_p_l_k && Delta_p_l_k[m])) continue;
            const size_t bytes = sizeof(double) * (size_t)D.n_src * n_l * D.n_q;
            if (defer) ctx->deferred.push_back({Delta_p_l_k[m], dm[m]->d_delta, bytes});
            else CK(cudaMemcpyAsync(Delta_p_l_k[m], dm[m]->d_delta, bytes, cudaMemcpyDeviceToHost, ctx->stream));
        }
        if (defer) CK(cudaEventRecord(ctx->ev_los_done, ctx->stream));
        if (ctx->async_calls && (!any || defer)) return 0;   // no host array is filled by this call itself: queued is done
    }
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->staged_pending = false;
    return deferred_error(ctx);
}

int cb_cl(cb_ctx* ctx, int n_models, cb_dmodel* const* dm, int tensor, const double* const* pows, double* const* iCl) {
    if (!ctx) return CB_ERR_NO_DEVICE;
    CK(cudaSetDevice(ctx->device));
    if (int e = begin_call(ctx)) return e;
    const int n_l = ctx->bess.n_l;
    std::vector<StagePiece> pieces;
    {
        size_t tot_in = 0, tot_out = 0;
        for (int m = 0; m < n_models; m++) { tot_in += (size_t)std::max(dm[m]->h.n_q, 0); tot_out += (size_t)n_l * dm[m]->n_cl; }
        if (int e = stage_reserve(ctx, tot_in, tot_out)) return e;
    }
    for (int m = 0; m < n_models; m++) {
        cb_dmodel* d = dm[m];
        DModel& D = d->h;
        if (D.num_redshiftwindows > 0) { ctx->err = "cb_cl: models with source windows go through cb_cl_windows"; return CB_ERR_UNSUPPORTED_PARAMS; }
        if ((tensor != 0) != (D.is_tensor != 0)) { ctx->err = "cb_cl: `tensor` does not match the model (CalcScalCls vs CalcTensCls)"; return CB_ERR_UNSUPPORTED_PARAMS; }
        if (!d->d_delta || D.n_q <= 0) { ctx->err = "cb_cl: no transfer functions (call cb_project first)"; return CB_ERR_UNSUPPORTED_PARAMS; }
        size_t nc = (size_t)n_l * d->n_cl;
        if (nc > d->cap_icl) { if (d->d_icl) dev_free(ctx, d->d_icl); d->d_icl = nullptr; d->cap_icl = 0; }
        if (int e = ensure(ctx, &d->d_icl, &d->cap_icl, nc)) return e;
        stage_add(ctx, pieces, d->d_pows, pows[m], (size_t)D.n_q);
        D.iCl = d->d_icl;
    }
    if (int e = stage_upload(ctx, pieces)) return e;
    if (int e = push_models(ctx, n_models, dm)) return e;
    CK(cudaEventRecord(ctx->stage_ev[3][0], ctx->stream));
    cb_cl_launch(ctx->d_models, n_models, ctx->bess, ctx->stream);
    ctx->launches++;
    CK(cudaGetLastError());
    CK(cudaEventRecord(ctx->stage_ev[3][1], ctx->stream));
    size_t off = 0;
    for (int m = 0; m < n_models; m++) {
        const size_t n = (size_t)n_l * dm[m]->n_cl;
        if (iCl && iCl[m]) CK(cudaMemcpyAsync(ctx->d_out + off, dm[m]->d_icl, sizeof(double) * n, cudaMemcpyDeviceToDevice, ctx->stream));
        off += n;
    }
    if (iCl && off) CK(cudaMemcpyAsync(ctx->h_out, ctx->d_out, sizeof(double) * off, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->staged_pending = false;
    if (int e = deferred_error(ctx)) return e;
    off = 0;
    for (int m = 0; m < n_models; m++) {
        const size_t n = (size_t)n_l * dm[m]->n_cl;
        if (iCl && iCl[m]) memcpy(iCl[m], ctx->h_out + off, sizeof(double) * n);
        off += n;
    }
    return 0;
}

int cb_cl_array(cb_ctx* ctx, int n_models, cb_dmodel* const* dm, const double* const* pows, double* const* iCl_Array) {
    if (!ctx) return CB_ERR_NO_DEVICE;
    CK(cudaSetDevice(ctx->device));
    if (int e = begin_call(ctx)) return e;
    const int n_l = ctx->bess.n_l;
    std::vector<double*> outs(n_models);
    for (int m = 0; m < n_models; m++) {
        cb_dmodel* d = dm[m];
        DModel& D = d->h;
        if (D.is_tensor || D.n_src != 3) { ctx->err = "cb_cl_array: scalar models with n_src = 3 only (Want_cl_2D_Array needs NumSources > 2)"; return CB_ERR_UNSUPPORTED_PARAMS; }
        if (!d->d_delta || D.n_q <= 0) { ctx->err = "cb_cl_array: no transfer functions (call cb_project first)"; return CB_ERR_UNSUPPORTED_PARAMS; }
        const size_t nc = (size_t)n_l * D.n_src * D.n_src;
        if (nc > d->cap_icl_array) { if (d->d_icl_array) dev_free(ctx, d->d_icl_array); d->d_icl_array = nullptr; d->cap_icl_array = 0; }
        if (int e = ensure(ctx, &d->d_icl_array, &d->cap_icl_array, nc)) return e;
        if (pows && pows[m]) CK(cudaMemcpyAsync(d->d_pows, pows[m], sizeof(double) * D.n_q, cudaMemcpyHostToDevice, ctx->stream));
        outs[m] = d->d_icl_array;
    }
    if (int e = push_models(ctx, n_models, dm)) return e;
    if (int e = ensure(ctx, &ctx->d_scratch_ptrs, &ctx->cap_scr, (size_t)n_models)) return e;
    CK(cudaMemcpyAsync(ctx->d_scratch_ptrs, outs.data(), sizeof(double*) * n_models, cudaMemcpyHostToDevice, ctx->stream));
    cb_cl_array_launch(ctx->d_models, n_models, ctx->bess, ctx->d_scratch_ptrs, ctx->stream);
    ctx->launches++;
    CK(cudaGetLastError());
    for (int m = 0; m < n_models; m++)
        if (iCl_Array && iCl_Array[m])
            CK(cudaMemcpyAsync(iCl_Array[m], dm[m]->d_icl_array, sizeof(double) * (size_t)n_l * dm[m]->h.n_src * dm[m]->h.n_src, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->staged_pending = false;
    if (int e = deferred_error(ctx)) return e;
    return 0;
}

int cb_cl_windows(cb_ctx* ctx, int n_models, cb_dmodel* const* dm, const double* const* pows, const double* const* limber_pows,
                  double* const* iCl, double* const* iCl_Array) {
    if (!ctx) return CB_ERR_NO_DEVICE;
    CK(cudaSetDevice(ctx->device));
    if (int e = begin_call(ctx)) return e;
    const int n_l = ctx->bess.n_l_all;
    std::vector<double*> limb(n_models), lpow(n_models), outs(2 * (size_t)n_models);
    int max_nw = 0;
    for (int m = 0; m < n_models; m++) {
        cb_dmodel* d = dm[m];
        DModel& D = d->h;
        if (D.is_tensor || D.n_src < 3) { ctx->err = "cb_cl_windows: scalar models with n_src >= 3 only"; return CB_ERR_UNSUPPORTED_PARAMS; }
        if (!d->d_delta || D.n_q <= 0 || !d->d_dd) { ctx->err = "cb_cl_windows: no transfer functions (call cb_project first)"; return CB_ERR_UNSUPPORTED_PARAMS; }
        const int nw = D.num_redshiftwindows, n = 3 + nw;
        max_nw = std::max(max_nw, nw);
        if (int e = ensure(ctx, &d->d_limb, &d->cap_limb, (size_t)(1 + nw) * n_l * D.n_tau)) return e;
        if (int e = ensure(ctx, &d->d_lpow, &d->cap_lpow, (size_t)n_l * D.n_tau)) return e;
        if (int e = ensure(ctx, &d->d_icl_w, &d->cap_icl_w, (size_t)n_l * 6)) return e;
        const size_t nc = (size_t)n_l * n * n;
        if (nc > d->cap_icl_array) { if (d->d_icl_array) dev_free(ctx, d->d_icl_array); d->d_icl_array = nullptr; d->cap_icl_array = 0; }
        if (int e = ensure(ctx, &d->d_icl_array, &d->cap_icl_array, nc)) return e;
        if (pows && pows[m]) CK(cudaMemcpyAsync(d->d_pows, pows[m], sizeof(double) * D.n_q, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemcpyAsync(d->d_lpow, limber_pows[m], sizeof(double) * (size_t)n_l * D.n_tau, cudaMemcpyHostToDevice, ctx->stream));
        limb[m] = d->d_limb; lpow[m] = d->d_lpow; outs[2 * m] = d->d_icl_w; outs[2 * m + 1] = d->d_icl_array;
    }
    if (int e = push_models(ctx, n_models, dm)) return e;
    if (int e = ensure(ctx, &ctx->d_scratch_ptrs, &ctx->cap_scr, 4 * (size_t)n_models)) return e;
    double** dp = ctx->d_scratch_ptrs;
    CK(cudaMemcpyAsync(dp, limb.data(), sizeof(double*) * n_models, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(dp + n_models, lpow.data(), sizeof(double*) * n_models, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(dp + 2 * n_models, outs.data(), sizeof(double*) * 2 * n_models, cudaMemcpyHostToDevice, ctx->stream));
    cb_limber_launch(ctx->d_models, n_models, ctx->bess, max_nw, dp, ctx->stream);
    CK(cudaGetLastError());
    cb_cl_windows_launch(ctx->d_models, n_models, ctx->bess, dp, dp + n_models, dp + 2 * n_models, ctx->stream);
    CK(cudaGetLastError());
    ctx->launches += 2;
    for (int m = 0; m < n_models; m++) {
        const int n = 3 + dm[m]->h.num_redshiftwindows;
        if (iCl && iCl[m]) CK(cudaMemcpyAsync(iCl[m], dm[m]->d_icl_w, sizeof(double) * (size_t)n_l * 6, cudaMemcpyDeviceToHost, ctx->stream));
        if (iCl_Array && iCl_Array[m])
            CK(cudaMemcpyAsync(iCl_Array[m], dm[m]->d_icl_array, sizeof(double) * (size_t)n_l * n * n, cudaMemcpyDeviceToHost, ctx->stream));
    }
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->staged_pending = false;
    if (int e = deferred_error(ctx)) return e;
    return 0;
}

int cb_lens_contrib(cb_ctx* ctx, int n_sets, const cb_lens_plan* P, const double* const* spectra, double* const* lens_contrib) {
    if (!ctx) return CB_ERR_NO_DEVICE;
    CK(cudaSetDevice(ctx->device));
    if (int e = begin_call(ctx)) return e;
    if (!P || n_sets <= 0 || !spectra || !lens_contrib) { ctx->err = "cb_lens_contrib: null argument"; return CB_ERR_UNSUPPORTED_PARAMS; }
    if ((P->variant != CB_LENS_APODIZED && P->variant != CB_LENS_FULL) || (P->lmin != 1 && P->lmin != 2) || P->lmax < 4 ||
        P->lmax_lensed < P->lmin || P->lmax_lensed > P->lmax || P->npoints < 2) {
        ctx->err = "cb_lens_contrib: inconsistent plan (variant, lmin in {1,2}, lmin <= lmax_lensed <= lmax, npoints)";
        return CB_ERR_UNSUPPORTED_PARAMS;
    }
    LensArgs A{};
    A.variant = P->variant; A.lmin = P->lmin; A.lmax = P->lmax; A.lmax_lensed = P->lmax_lensed; A.npoints = P->npoints;
    std::vector<void*> owned;
    auto release = [&]() { for (void* p : owned) dev_free(ctx, p); };
    if (P->variant == CB_LENS_APODIZED) {
        if (!(P->dtheta > 0) || P->interp_fac < 1 || (P->short_integral_range && P->apodize_point_width < 1)) {
            ctx->err = "cb_lens_contrib: apodized plan needs dtheta > 0, interp_fac >= 1, apodize_point_width >= 1";
            return CB_ERR_UNSUPPORTED_PARAMS;
        }
        A.first = 1; A.last = P->npoints - 1;
        A.dtheta = P->dtheta; A.apodize_point_width = P->apodize_point_width; A.interp_fac = P->interp_fac;
        A.short_integral_range = P->short_integral_range;
    } else {
        if (!P->xvals || !P->weights || P->imin < 1 || P->imin > P->npoints) {
            ctx->err = "cb_lens_contrib: full plan needs xvals, weights and 1 <= imin <= npoints";
            return CB_ERR_UNSUPPORTED_PARAMS;
        }
        A.first = P->imin; A.last = P->npoints;
        double *dx = nullptr, *dw = nullptr;
        if (int e = upload(ctx, P->xvals, (size_t)P->npoints, &dx, &owned)) { release(); return e; }
        if (int e = upload(ctx, P->weights, (size_t)P->npoints, &dw, &owned)) { release(); return e; }
        A.xvals = dx; A.weights = dw;
        A.theta_max = P->theta_max; A.apodize_theta_width = P->apodize_theta_width; A.xtaper_start = P->xtaper_start;
    }
    if (P->theta_last > 0) {        // a theta slice of the loop (multi-GPU sharding)
        A.first = std::max(A.first, P->theta_first);
        A.last = std::min(A.last, P->theta_last);
    }
    if (A.last < A.first) {         // this rank's slice holds no angle of the method's range
        for (int m = 0; m < n_sets; m++) memset(lens_contrib[m], 0, sizeof(double) * 4 * ((size_t)P->lmax_lensed + 1));
        release();
        CK(cudaStreamSynchronize(ctx->stream));
        return 0;
    }
    if (ctx->lens_tab_n < P->lmax + 1) {
        if (ctx->d_lens_tab) dev_free(ctx, ctx->d_lens_tab);
        if (ctx->d_lens_tab2) dev_free(ctx, ctx->d_lens_tab2);
        ctx->d_lens_tab = nullptr; ctx->d_lens_tab2 = nullptr; ctx->lens_tab_n = 0;
        CK(dev_alloc(ctx, (void**)&ctx->d_lens_tab, sizeof(LensTab) * (size_t)(P->lmax + 1)));
        CK(dev_alloc(ctx, (void**)&ctx->d_lens_tab2, sizeof(LensTab2) * (size_t)(P->lmax + 1)));
        ctx->lens_tab_n = P->lmax + 1;
        cb_lens_tab_launch(ctx->d_lens_tab, ctx->d_lens_tab2, ctx->lens_tab_n, ctx->stream);
        ctx->launches++;
    }
    A.tab = ctx->d_lens_tab; A.tab2 = ctx->d_lens_tab2;
    const size_t ld = 4 * ((size_t)P->lmax + 1), n_out = 4 * ((size_t)P->lmax_lensed + 1);
    const int n_warps = cb_lens_warps(A);
    double *d_spec = nullptr, *d_part = nullptr, *d_out = nullptr;
    cudaError_t ce = dev_alloc(ctx, (void**)&d_spec, sizeof(double) * ld * n_sets);
    if (ce == cudaSuccess) { owned.push_back(d_spec); ce = dev_alloc(ctx, (void**)&d_part, sizeof(double) * n_out * n_warps * n_sets); }
    if (ce == cudaSuccess) { owned.push_back(d_part); ce = dev_alloc(ctx, (void**)&d_out, sizeof(double) * n_out * n_sets); }
    if (ce != cudaSuccess) { release(); ctx->err = std::string("cb_lens_contrib: ") + cudaGetErrorString(ce); return CB_ERR_CUDA; }
    owned.push_back(d_out);
    for (int m = 0; m < n_sets; m++) {
        ce = cudaMemcpyAsync(d_spec + ld * m, spectra[m], sizeof(double) * ld, cudaMemcpyHostToDevice, ctx->stream);
        if (ce != cudaSuccess) { release(); ctx->err = std::string("cb_lens_contrib: ") + cudaGetErrorString(ce); return CB_ERR_CUDA; }
    }
    A.spectra = d_spec; A.partial = d_part;
    if (!ctx->lens_ev[0]) { CK(cudaEventCreate(&ctx->lens_ev[0])); CK(cudaEventCreate(&ctx->lens_ev[1])); }
    cudaEventRecord(ctx->lens_ev[0], ctx->stream);
    cb_lens_launch(A, n_sets, d_out, ctx->stream);
    cudaEventRecord(ctx->lens_ev[1], ctx->stream);
    ctx->lens_timed = true;
    ctx->launches += 2;
    ce = cudaGetLastError();
    for (int m = 0; m < n_sets && ce == cudaSuccess; m++)
        ce = cudaMemcpyAsync(lens_contrib[m], d_out + n_out * m, sizeof(double) * n_out, cudaMemcpyDeviceToHost, ctx->stream);
    release();
    if (ce != cudaSuccess) { ctx->err = std::string("cb_lens_contrib: ") + cudaGetErrorString(ce); return CB_ERR_CUDA; }
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->staged_pending = false;
    if (int e = deferred_error(ctx)) return e;
    return 0;
}

int cb_last_lens_timing(cb_ctx* ctx, double* ms) {
    if (!ctx || !ms) return CB_ERR_NO_DEVICE;
    *ms = 0;
    if (!ctx->lens_timed) return 0;
    float f = 0;
    CK(cudaEventElapsedTime(&f, ctx->lens_ev[0], ctx->lens_ev[1]));
    *ms = f;
    return 0;
}

int cb_last_timings(const cb_ctx* ctx, double ms[4]) {
    if (!ctx) return CB_ERR_NO_DEVICE;
    cudaStreamSynchronize(ctx->stream);
    for (int i = 0; i < 4; i++) {
        float f = 0;
        ms[i] = cudaEventElapsedTime(&f, ctx->stage_ev[i][0], ctx->stage_ev[i][1]) == cudaSuccess ? f : 0.0;
    }
    cudaGetLastError();   // a stage that has not run yet leaves an unrecorded event: not an error
    return 0;
}

long long cb_launch_count(const cb_ctx* ctx) { return ctx ? ctx->launches : 0; }

// With the option "async_outputs" cb_project returns while Delta_p_l_k is still on its way to the host: this is the point
// after which the host arrays of every earlier call are complete.
int cb_wait_outputs(cb_ctx* ctx) {
    if (!ctx) return CB_ERR_NO_DEVICE;
    if (int e = flush_outputs(ctx)) return e;
    if (ctx->copies_pending) { CK(cudaEventSynchronize(ctx->ev_copy_done)); ctx->copies_pending = false; }
    return 0;
}

int cb_set_option(cb_ctx* ctx, const char* name, double value) {
    if (!ctx) return CB_ERR_NO_DEVICE;
    if (!strncmp(name, "lanes_class", 11) && name[11] >= '0' && name[11] < '0' + CB_TENSOR_CLASS && !name[12]) {
        const int v = (int)value;
        if (v != 0 && v != 4 && v != 8 && v != 16 && v != 32) { ctx->err = "lanes_classN must be 0 (follow lanes_per_mode), 4, 8, 16 or 32"; return CB_ERR_UNSUPPORTED_PARAMS; }
        ctx->class_lanes[name[11] - '0'] = v;
        return 0;
    }
    if (!strcmp(name, "lanes_per_mode")) {
        const int v = (int)value;
        if (v != 0 && v != 4 && v != 8 && v != 16 && v != 32) { ctx->err = "lanes_per_mode must be 0 (automatic), 4, 8, 16 or 32"; return CB_ERR_UNSUPPORTED_PARAMS; }
        ctx->lanes_per_mode = v;
        return 0;
    }
    if (!strcmp(name, "async_model_upload")) { ctx->async_model_upload = value != 0; return 0; }
    if (!strcmp(name, "use_reg_kernel")) { ctx->use_reg_kernel = (int)value; return 0; }
    if (!strcmp(name, "reg_min_modes")) { ctx->reg_min_modes = (int)value; return 0; }
    if (!strcmp(name, "los_kernel")) { ctx->los_kernel = (int)value; return 0; }
    if (!strcmp(name, "async_calls")) {
        if (ctx->async_calls && value == 0) { CK(cudaStreamSynchronize(ctx->stream)); ctx->staged_pending = false; if (int e = deferred_error(ctx)) return e; }
        ctx->async_calls = value != 0;
        return 0;
    }
    if (!strcmp(name, "async_outputs")) {
        flush_outputs(ctx);
        if (ctx->copies_pending) { cudaEventSynchronize(ctx->ev_copy_done); ctx->copies_pending = false; }
        ctx->async_outputs = value != 0;
        return 0;
    }
    ctx->err = std::string("unknown option ") + name;
    return CB_ERR_UNSUPPORTED_PARAMS;
}

int cb_timer_start(cb_ctx* ctx) {
    if (!ctx) return CB_ERR_NO_DEVICE;
    CK(cudaSetDevice(ctx->device));
    CK(cudaEventRecord(ctx->evt0, ctx->stream));
    return 0;
}

int cb_timer_stop(cb_ctx* ctx, double* ms) {
    if (!ctx) return CB_ERR_NO_DEVICE;
    CK(cudaEventRecord(ctx->evt1, ctx->stream));
    CK(cudaEventSynchronize(ctx->evt1));
    float f = 0;
    CK(cudaEventElapsedTime(&f, ctx->evt0, ctx->evt1));
    *ms = f;
    return 0;
}

int cb_fp64_peak(cb_ctx* ctx, double* tflops) {
    if (!ctx) return CB_ERR_NO_DEVICE;
    CK(cudaSetDevice(ctx->device));
    *tflops = cb_fp64_peak_measure(ctx->n_sm, ctx->stream, 5);
    ctx->launches += 6;
    return *tflops > 0 ? 0 : CB_ERR_CUDA;
}

void* cb_host_alloc(size_t bytes) {
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) return nullptr;
    return p;
}

int cb_host_register(void* p, size_t bytes) {
    return cudaHostRegister(p, bytes, cudaHostRegisterDefault) == cudaSuccess ? 0 : CB_ERR_CUDA;
}

void cb_host_unregister(void* p) {
    if (p) cudaHostUnregister(p);
}

void cb_host_free(void* p) {
    if (p) cudaFreeHost(p);
}

// ---- one cosmology sharded over the GPUs of a box: the two exchange steps, stream-ordered (collective.cu) ----
int cb_comm_unique_id(cb_ctx* ctx, char* id128) {
    if (!ctx) return CB_ERR_NO_DEVICE;
    if (!cb_nccl_unique_id(id128, &ctx->err)) return CB_ERR_CUDA;
    return 0;
}

int cb_comm_init(cb_ctx* ctx, int nranks, int rank, const char* id128) {
    if (!ctx) return CB_ERR_NO_DEVICE;
    CK(cudaSetDevice(ctx->device));
    if (nranks < 1 || rank < 0 || rank >= nranks) { ctx->err = "cb_comm_init: bad rank"; return CB_ERR_UNSUPPORTED_PARAMS; }
    if (ctx->comm) { cb_nccl_destroy(ctx->comm); ctx->comm = nullptr; }
    if (!cb_nccl_init(&ctx->comm, nranks, rank, id128, &ctx->err)) return CB_ERR_CUDA;
    ctx->comm_ranks = nranks; ctx->comm_rank = rank;
    return 0;
}

// owner[k] (0 .. nranks-1): the rank whose cb_evolve call had mode k in cb_modes.active.  Returns with the exchange
// queued on the library's stream (no host synchronisation): the next entry point is ordered after it.
int cb_gather_sources(cb_ctx* ctx, cb_dmodel* dm, int n_k, const int* owner) {
    if (!ctx || !dm) return CB_ERR_NO_DEVICE;
    CK(cudaSetDevice(ctx->device));
    if (!ctx->comm) { ctx->err = "cb_gather_sources: no communicator (cb_comm_init)"; return CB_ERR_UNSUPPORTED_PARAMS; }
    const DModel& D = dm->h;
    if (n_k != D.n_k || !dm->d_src) { ctx->err = "cb_gather_sources: call cb_evolve on this model first"; return CB_ERR_UNSUPPORTED_PARAMS; }
    const int R = ctx->comm_ranks, me = ctx->comm_rank, ncol = D.n_src * D.n_tau;
    std::vector<int> cnt(R, 0), pos(n_k), idx;
    for (int k = 0; k < n_k; k++) {
        if (owner[k] < 0 || owner[k] >= R) { ctx->err = "cb_gather_sources: owner out of range"; return CB_ERR_UNSUPPORTED_PARAMS; }
        pos[k] = cnt[owner[k]]++;
        if (owner[k] == me) idx.push_back(k);
    }
    int n_pack = 1;
    for (int r = 0; r < R; r++) n_pack = std::max(n_pack, cnt[r]);
    // small index arrays: pageable cudaMemcpyAsync stages them before returning
    int *d_owner = nullptr, *d_pos = nullptr, *d_idx = nullptr;
    double *d_packed = nullptr, *d_gathered = nullptr;
    CK(dev_alloc(ctx, (void**)&d_owner, sizeof(int) * n_k));
    CK(dev_alloc(ctx, (void**)&d_pos, sizeof(int) * n_k));
    CK(dev_alloc(ctx, (void**)&d_idx, sizeof(int) * std::max<size_t>(idx.size(), 1)));
    CK(dev_alloc(ctx, (void**)&d_packed, sizeof(double) * (size_t)n_pack * ncol));
    CK(dev_alloc(ctx, (void**)&d_gathered, sizeof(double) * (size_t)R * n_pack * ncol));
    CK(cudaMemcpyAsync(d_owner, owner, sizeof(int) * n_k, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(d_pos, pos.data(), sizeof(int) * n_k, cudaMemcpyHostToDevice, ctx->stream));
    if (!idx.empty()) CK(cudaMemcpyAsync(d_idx, idx.data(), sizeof(int) * idx.size(), cudaMemcpyHostToDevice, ctx->stream));
    const bool ok = cb_nccl_gather_rows(ctx->comm, R, dm->d_src, n_k, ncol, d_idx, (int)idx.size(), n_pack, d_owner, d_pos, d_packed, d_gathered,
                                        ctx->stream, &ctx->err);
    ctx->launches += 2;
    dev_free(ctx, d_owner); dev_free(ctx, d_pos); dev_free(ctx, d_idx); dev_free(ctx, d_packed); dev_free(ctx, d_gathered);
    return ok ? 0 : CB_ERR_CUDA;
}

// Sum of the partial iCl (this rank's slice of the integration wavenumbers) over the ranks; iCl (may be NULL) receives it.
int cb_allreduce_cl(cb_ctx* ctx, cb_dmodel* dm, double* iCl) {
    if (!ctx || !dm) return CB_ERR_NO_DEVICE;
    CK(cudaSetDevice(ctx->device));
    if (!ctx->comm) { ctx->err = "cb_allreduce_cl: no communicator (cb_comm_init)"; return CB_ERR_UNSUPPORTED_PARAMS; }
    if (!dm->d_icl) { ctx->err = "cb_allreduce_cl: call cb_cl on this model first"; return CB_ERR_UNSUPPORTED_PARAMS; }
    const size_t n = (size_t)ctx->bess.n_l * dm->n_cl;
    if (!cb_nccl_allreduce_sum(ctx->comm, dm->d_icl, n, ctx->stream, &ctx->err)) return CB_ERR_CUDA;
    if (iCl) CK(cudaMemcpyAsync(iCl, dm->d_icl, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

int cb_device_buffer(cb_ctx* ctx, cb_dmodel* dm, int which, void** ptr, size_t* bytes) {
    if (!ctx || !dm) return CB_ERR_NO_DEVICE;
    const DModel& D = dm->h;
    const size_t n_l = (size_t)ctx->bess.n_l;
    switch (which) {
        case CB_BUF_LINEARSRC: *ptr = dm->d_src; *bytes = sizeof(double) * (size_t)D.n_k * D.n_src * D.n_tau; return 0;
        case CB_BUF_DELTA: *ptr = dm->d_delta; *bytes = sizeof(double) * (size_t)D.n_src * n_l * D.n_q; return 0;
        case CB_BUF_ICL: *ptr = dm->d_icl; *bytes = sizeof(double) * n_l * dm->n_cl; return 0;
        case CB_BUF_TRANSFER: *ptr = dm->d_transfer; *bytes = sizeof(float) * (size_t)13 * D.n_k * D.num_transfer_redshifts; return 0;
    }
    ctx->err = "cb_device_buffer: unknown buffer id";
    return CB_ERR_UNSUPPORTED_PARAMS;
}

// One evaluation of the device `derivs` + source output at a given state (parity tests against the reference-authored
// per-evaluation fixture).  lanes: 8 or 32; lean: the massless-nu LCDM template variant (refused for other models).
int cb_debug_derivs(cb_ctx* ctx, cb_dmodel* dm, double q, double tau, const int* flags4, const double* y, int n, int lanes, int lean,
                    double* yprime, double* sources) {
    if (!ctx || !dm) return CB_ERR_NO_DEVICE;
    CK(cudaSetDevice(ctx->device));
    const DModel& D = dm->h;
    if (D.is_tensor) { ctx->err = "cb_debug_derivs: scalar models only"; return CB_ERR_UNSUPPORTED_PARAMS; }
    if (int e = pack_thermo_tables(ctx, 1, &dm)) return e;
    if (lean && (D.Num_Nu_massive != 0 || !D.de_is_cc || D.num_redshiftwindows > 0)) {
        ctx->err = "cb_debug_derivs: the lean variant is for massless-neutrino LCDM models";
        return CB_ERR_UNSUPPORTED_PARAMS;
    }
    int nv = 0, maxl = 0, lmax3[3] = {0, 0, 0};
    cb_mode_size(D, q, 0, &nv, &maxl, lmax3);
    if (lanes == 1 && (!lean || cb_mpl_class(lmax3[0], lmax3[1], lmax3[2]) < 0)) {
        ctx->err = "cb_debug_derivs: the mode-per-lane kernel takes lean models and modes of its two size classes";
        return CB_ERR_UNSUPPORTED_PARAMS;
    }
    if (n < 1 || n > nv || maxl > CBD_MAX_L_EVOLVE) { ctx->err = "cb_debug_derivs: bad state size"; return CB_ERR_UNSUPPORTED_PARAMS; }
    DModel* d_model = nullptr;
    double *d_y = nullptr, *d_out = nullptr;
    int* d_status = nullptr;
    CK(cudaMalloc(&d_model, sizeof(DModel)));
    CK(cudaMalloc(&d_y, sizeof(double) * nv));
    CK(cudaMalloc(&d_out, sizeof(double) * (nv + 3)));
    CK(cudaMalloc(&d_status, sizeof(int) * 2));
    int rc = 0, status[2] = {0, 1};
    cudaError_t e = cudaMemcpyAsync(d_model, &D, sizeof(DModel), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_y, y, sizeof(double) * n, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess && lanes == 1) {
        e = cudaMemsetAsync(d_out, 0, sizeof(double) * (nv + 3), ctx->stream);
        if (e == cudaSuccess) e = cb_debug_derivs_mpl_launch(d_model, q, tau, flags4, lmax3, d_y, n, d_out, d_out + nv, ctx->stream);
        status[1] = 0;
        if (e == cudaSuccess) e = cudaMemsetAsync(d_status, 0, sizeof(int) * 2, ctx->stream);
    } else if (e == cudaSuccess) e = cb_debug_derivs_launch(d_model, q, tau, flags4, d_y, n, nv, maxl, lanes, lean, d_out, d_out + nv, d_status, ctx->stream);
    if (e == cudaSuccess) { ctx->launches++; e = cudaMemcpyAsync(status, d_status, sizeof(status), cudaMemcpyDeviceToHost, ctx->stream); }
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e == cudaSuccess && status[1] == 0) {
        e = cudaMemcpy(yprime, d_out, sizeof(double) * n, cudaMemcpyDeviceToHost);
        if (e == cudaSuccess) e = cudaMemcpy(sources, d_out + nv, sizeof(double) * 3, cudaMemcpyDeviceToHost);
    }
    cudaFree(d_model); cudaFree(d_y); cudaFree(d_out); cudaFree(d_status);
    if (e != cudaSuccess) { ctx->err = std::string("cb_debug_derivs: ") + cudaGetErrorString(e); return CB_ERR_CUDA; }
    if (status[1] != 0) { ctx->err = "cb_debug_derivs: n does not match the layout the flags select"; rc = CB_ERR_UNSUPPORTED_PARAMS; }
    return rc;
}

// Self-test of the Tensor-Memory vector storage (tmem.cuh) the mode-per-lane evolve kernel relies on: pattern write /
// read-back through every column of every lane on n_blocks SMs; cycles[0..1] = SM cycles per 8-double load / store.
int cb_tmem_probe(cb_ctx* ctx, int n_blocks, int* mismatches, long long* cycles) {
    if (!ctx) return CB_ERR_NO_DEVICE;
    CK(cudaSetDevice(ctx->device));
    int* d_bad = nullptr;
    long long* d_cyc = nullptr;
    CK(cudaMalloc(&d_bad, sizeof(int)));
    CK(cudaMalloc(&d_cyc, 3 * sizeof(long long)));
    CK(cudaMemsetAsync(d_bad, 0, sizeof(int), ctx->stream));
    CK(cudaMemsetAsync(d_cyc, 0, 3 * sizeof(long long), ctx->stream));
    cudaError_t e = cb_tmem_probe_launch(n_blocks > 0 ? n_blocks : ctx->n_sm, d_bad, d_cyc, ctx->stream);
    if (e == cudaSuccess) { ctx->launches++; e = cudaStreamSynchronize(ctx->stream); }
    long long cyc[3] = {0, 0, 0};
    if (e == cudaSuccess) e = cudaMemcpy(mismatches, d_bad, sizeof(int), cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(cyc, d_cyc, sizeof(cyc), cudaMemcpyDeviceToHost);
    cudaFree(d_bad); cudaFree(d_cyc);
    if (e != cudaSuccess) { ctx->err = std::string("cb_tmem_probe: ") + cudaGetErrorString(e); return CB_ERR_CUDA; }
    if (cycles) { cycles[0] = cyc[0]; cycles[1] = cyc[1]; }
    return 0;
}

size_t cb_abi_sizeof(int which) {
    switch (which) {
        case 0: return sizeof(cb_params);
        case 1: return sizeof(cb_model);
        case 2: return sizeof(cb_modes);
        case 3: return sizeof(cb_bessel);
        case 4: return sizeof(cb_range);
        case 5: return sizeof(cb_accuracy);
        case 6: return sizeof(cb_lens_plan);
        case 7: return sizeof(cb_window);
    }
    return 0;
}

}  // extern "C"
