// api.cu -- the C ABI (include/slam/cuda/slamcu.h) over the sm_100a kernels.  No torch, no STL types
// across the boundary, no CPU fallback: a call either enqueues CUDA work or returns an error status.
#include <cstdarg>
#include <cstdio>
#include <climits>
#include <cstring>
#include <cmath>
#include <algorithm>
#include <cstdlib>
#include <new>
#include <random>
#include <utility>
#include <string>
#include <vector>

#include <cudaTypedefs.h>  // PFN_cuTensorMapEncodeTiled

#include "common.cuh"
#include "orb.cuh"
#include "match_tc.cuh"

using namespace slamcu;

namespace slamcu {
void init_sortnms_attributes(int smem_optin);
int launch_desc_or(const uint32_t* desc, const int* n_dev, int desc_words, uint32_t* out, cudaStream_t st, int sets = 1,
                   size_t set_stride = 0);
long long launch_popc_peak(unsigned* scratch, int blocks, int iters, cudaStream_t st);
void launch_trip_bound(cudaStream_t st);
void init_orb_attributes(int smem_optin);
}

namespace slamcu {
thread_local Profiler* g_prof = nullptr;
}

namespace {
constexpr int8_t kOrbBitPattern31[1024] = {
#include "orb_pattern.inc"
};

struct EventProfiler : Profiler {
    struct Rec { const char* name; cudaEvent_t a, b; };
    struct Acc { const char* name; double ms; long long count; };
    std::vector<Rec> pending;
    std::vector<cudaEvent_t> pool;
    std::vector<Acc> acc;
    cudaEvent_t cur_a = nullptr;
    const char* cur_name = nullptr;
    cudaEvent_t get() {
        if (!pool.empty()) { cudaEvent_t e = pool.back(); pool.pop_back(); return e; }
        cudaEvent_t e; cudaEventCreate(&e); return e;
    }
    void begin(const char* name, cudaStream_t st) override {
        cur_a = get(); cur_name = name; cudaEventRecord(cur_a, st);
    }
    void end(cudaStream_t st) override {
        cudaEvent_t b = get(); cudaEventRecord(b, st);
        pending.push_back({cur_name, cur_a, b});
    }
    void collect() {  // caller has synchronised the stream
        for (auto& r : pending) {
            float ms = 0.f;
            cudaEventElapsedTime(&ms, r.a, r.b);
            bool found = false;
            for (auto& a : acc) if (!strcmp(a.name, r.name)) { a.ms += ms; a.count++; found = true; break; }
            if (!found) acc.push_back({r.name, ms, 1});
            pool.push_back(r.a); pool.push_back(r.b);
        }
        pending.clear();
    }
    void reset() { collect(); acc.clear(); }
    ~EventProfiler() { for (auto e : pool) cudaEventDestroy(e); }
};
}  // namespace

struct slamcu_context {
    EventProfiler prof;
    bool profiling = false;
    int device = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    int smem_optin = 0;
    int64_t launches = 0;
    std::string err;
    // scratch reused by the preparation / ransac entry points
    void* scratch = nullptr;
    size_t scratch_bytes = 0;
    // copy engines of the pipelined sequence path: H2D and D2H run on their own streams
    cudaStream_t s_in = nullptr, s_out = nullptr;
    cudaStream_t s_dl = nullptr;  // exact-size downloads of dense outputs (issued by the host once the totals are known)
    std::vector<struct slamcu_sequence*> pending_dl;  // sequences whose dense outputs are still on the device
    std::vector<cudaEvent_t> events;
    // auxiliary compute stream: kernels that are independent inside one extract call overlap with the main chain
    cudaStream_t s_aux = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    // second compute lane (its own main + auxiliary stream and fork / join events): consecutive chunks of a sequence alternate
    // between the two lanes, so the ALU-bound matcher of chunk c overlaps the extraction of chunk c + 1 on the same SMs
    cudaStream_t lane1 = nullptr, lane1_aux = nullptr;
    cudaEvent_t lane1_fork = nullptr, lane1_join = nullptr, ev_lane = nullptr;
    bool two_lanes = true;  // SLAMCU_ONE_LANE=1 turns the second lane off (A/B measurements)
};

namespace {

constexpr int kMaxMatchSlices = 32;

int fail(slamcu_context* ctx, int status, const char* fmt, ...) {
    if (ctx) {
        char buf[512];
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(buf, sizeof buf, fmt, ap);
        va_end(ap);
        ctx->err = buf;
    }
    return status;
}

#define CU(ctx, call)                                                                                  \
    do {                                                                                               \
        cudaError_t e__ = (call);                                                                      \
        if (e__ != cudaSuccess)                                                                        \
            return fail(ctx, SLAMCU_CUDA_ERROR, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), \
                        __FILE__, __LINE__);                                                           \
    } while (0)

int check_launch(slamcu_context* ctx, const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(ctx, SLAMCU_CUDA_ERROR, "%s: %s", what, cudaGetErrorString(e));
    return SLAMCU_OK;
}

int ensure_scratch(slamcu_context* ctx, size_t bytes) {
    if (ctx->scratch_bytes >= bytes) return SLAMCU_OK;
    if (ctx->scratch) cudaFree(ctx->scratch);
    ctx->scratch = nullptr;
    ctx->scratch_bytes = 0;
    CU(ctx, cudaMalloc(&ctx->scratch, bytes));
    ctx->scratch_bytes = bytes;
    return SLAMCU_OK;
}

template <class T>
int dev_alloc(slamcu_context* ctx, T** p, size_t count, std::vector<void*>& owned, bool zero = false) {
    void* q = nullptr;
    const size_t bytes = (count ? count : 1) * sizeof(T);
    CU(ctx, cudaMalloc(&q, bytes));
    owned.push_back(q);
    if (zero) CU(ctx, cudaMemsetAsync(q, 0, bytes, ctx->stream));
    *p = static_cast<T*>(q);
    return SLAMCU_OK;
}

int round_up(int v, int m) { return (v + m - 1) / m * m; }

// Camera::undistortImage's per-pixel source index (common.hpp:143-162) with the host libm the reference itself calls
// (Eigen's r.pow(4) is std::pow(r, 4.0)); -1 = outside the image
int undistort_index_host(int i, int j, int rows, int cols, const CamParams& c) {
    const double x = ((double)j - c.cx) / c.fx;
    const double y = ((double)i - c.cy) / c.fy;
    const double r = std::sqrt(x * x + y * y);
    const double r2 = r * r;
    const double r4 = std::pow(r, 4.0);
    const double xd = x * (1 + c.k1 * r2 + c.k2 * r4) + 2 * c.p1 * x * y + c.p2 * (r2 + 2 * (x * x));
    const double yd = y * (1 + c.k1 * r2 + c.k2 * r4) + 2 * c.p2 * x * y + c.p1 * (r2 + 2 * (y * y));
    const double ud = c.fx * xd + c.cx;
    const double vd = c.fy * yd + c.cy;
    const int u = static_cast<int>(std::round(ud)), v = static_cast<int>(std::round(vd));
    return (u >= 0 && v >= 0 && u < cols && v < rows) ? v * cols + u : -1;
}

// Device map + host fix-up of the pixels that sit on a rounding boundary (prep.cu: undistort_map_kernel).  `fix` is a
// device scratch of 1 + kUndistFixCap ints.  Synchronises the stream (constructor-time work: once per camera).
constexpr int kUndistFixCap = 1 << 16;
int build_undistort_map(slamcu_context* ctx, int rows, int cols, const CamParams& cam, int* d_map, int* d_fix) {
    ctx->launches += launch_undistort_map(rows, cols, cam, d_map, d_fix, kUndistFixCap, ctx->stream);
    int n_fix = 0;
    CU(ctx, cudaMemcpyAsync(&n_fix, d_fix, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    if (n_fix == 0) return SLAMCU_OK;
    if (n_fix > kUndistFixCap) {  // degenerate calibration (e.g. NaNs): the whole map on the host
        std::vector<int> all((size_t)rows * cols);
        for (int i = 0; i < rows; i++)
            for (int j = 0; j < cols; j++) all[(size_t)i * cols + j] = undistort_index_host(i, j, rows, cols, cam);
        CU(ctx, cudaMemcpyAsync(d_map, all.data(), all.size() * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
        CU(ctx, cudaStreamSynchronize(ctx->stream));
        return SLAMCU_OK;
    }
    std::vector<int> idx((size_t)n_fix);
    CU(ctx, cudaMemcpyAsync(idx.data(), d_fix + 1, (size_t)n_fix * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    for (int k = 0; k < n_fix; k++) {
        const int v = undistort_index_host(idx[k] / cols, idx[k] % cols, rows, cols, cam);
        CU(ctx, cudaMemcpyAsync(d_map + idx[k], &v, sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
        CU(ctx, cudaStreamSynchronize(ctx->stream));  // v lives on this stack frame
    }
    return SLAMCU_OK;
}

// every failure leaves its own message: slamcu_last_error() never reports an earlier, unrelated call
int bad_args(slamcu_context* ctx, const char* fn) {
    return fail(ctx, SLAMCU_INVALID_ARGUMENT, "%s: null or mismatched handle / pointer argument", fn);
}

}  // namespace

struct ProfGuard {
    explicit ProfGuard(slamcu_context* ctx) { slamcu::g_prof = ctx->profiling ? &ctx->prof : nullptr; }
    ~ProfGuard() { slamcu::g_prof = nullptr; }
};

struct slamcu_sequence {
    slamcu_context* ctx = nullptr;
    SeqView v{};
    int max_frames = 0;
    unsigned long long* sort_keys = nullptr;  // [F][cap_kp]
    int* h_counts = nullptr;                  // pinned [F][4]
    int* h_status = nullptr;                  // pinned [F + 1]: status words of the latest slamcu_sequence_process call (+ dense overflow bits)
    int* d_dense = nullptr;                   // [2][F + 1] running offsets of the dense outputs (keypoints, matches) + 1 overflow word
    // dense outputs are compacted on the device and cross the link through the copy engine with their exact sizes, once the
    // host has the totals (slamcu_sequence_wait / slamcu_synchronize): SM-written stores to mapped host memory slowed the
    // concurrent compute kernels down by ~10 %
    void *dd_kps = nullptr, *dd_desc = nullptr, *dd_matches = nullptr;
    long long dd_kp_cap = 0, dd_m_cap = 0;
    int* h_tot = nullptr;                     // pinned [2]: total keypoints, total matches of the latest dense call
    void *dl_kps = nullptr, *dl_desc = nullptr, *dl_matches = nullptr;  // the caller's host buffers of that call
    bool dl_active = false;
    int h_status_n = 0;                       // ... and how many frames it covered
    uint8_t* stage = nullptr;                 // [F][rows][cols] dense landing zone of linear H2D copies (lazy)
    uint8_t* prep_stage = nullptr;            // landing zone of slamcu_sequence_prepare (gray or BGR host frames; lazy)
    size_t prep_stage_bytes = 0;
    int* undist_map = nullptr;                // per-camera gather map of Camera::undistortImage (lazy, keyed by K, D)
    double undist_key[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    bool has_undist = false;
    cudaEvent_t ev_compute_done = nullptr;    // last kernel of the latest slamcu_sequence_process call
    cudaEvent_t ev_out_done = nullptr;        // last download of the latest slamcu_sequence_process call
    EssentialJob ess{};                       // per-pair two-view RANSAC working set (lazy)
    bool has_ess = false;
    std::vector<void*> ess_owned;
    void* ess_work = nullptr;                 // per-launch RANSAC working set (essential_work_bytes_per_pair)
    size_t ess_work_bytes = 0;
    std::vector<void*> owned;
    // ORB-mode working set, allocated on first use and keyed by the detector parameters
    bool has_orb = false;
    OrbView orb{};
    int orb_levels = 0, orb_features = 0, orb_fast = 0;
    float orb_scale = 0.f;
    std::vector<void*> orb_owned;
    OrbTmaps tmaps{};  // TMA descriptors of the pyramid levels (host copies; passed to the kernels as __grid_constant__)
    // tensor-core matcher operands (lazy): every frame's descriptors widened to one byte per bit + key constants
    uint8_t* tc_x8 = nullptr;   // [F][tc_rows][256]
    uint32_t* tc_ck = nullptr;  // [F][tc_rows]
    int tc_rows = 0;            // cap_kp rounded up to 128
    int tc_state = 0;           // 0 = not tried, 1 = ready, -1 = unavailable (integer-pipe kernels)
    MatchTc tc{};
};

struct slamcu_detector {
    slamcu_context* ctx = nullptr;
    DetParams p{};
    int mode = 0;
    int n_levels = 8, max_features = 2000, fast_threshold = 20;
    float scale_factor = 1.2f;
    int8_t* d_orb_pattern = nullptr;  // [512][2]
    float2* d_orb_patf = nullptr;     // [512] (x, y) as floats
    int* d_pattern = nullptr;
    slamcu_sequence* one = nullptr;  // cached single-frame workspace
};

struct slamcu_matcher {
    slamcu_context* ctx = nullptr;
    MatchParams p{};
    int distance_type = 0;
    // single-call workspace
    uint32_t *d1 = nullptr, *d2 = nullptr;
    slamcu_keypoint *k1 = nullptr, *k2 = nullptr;
    int4* cand = nullptr;
    slamcu_dmatch* matches = nullptr;
    unsigned long long* keys = nullptr;
    uint32_t* ors = nullptr;
    int* counts = nullptr;  // nq, nt, n_match, status
    int cap1 = 0, cap2 = 0, cap_words = 0;
    int forced_slices = 0;  // slamcu_matcher_set_train_slices
    // tensor-core operands of the single-call path (lazy)
    uint8_t *x1 = nullptr, *x2 = nullptr;
    uint32_t *ck1 = nullptr, *ck2 = nullptr;
    int tc_rows1 = 0, tc_rows2 = 0;
    MatchTc tc{};
};

extern "C" {

int slamcu_abi_version(void) { return SLAMCU_ABI_VERSION; }

const char* slamcu_status_string(int status) {
    switch (status) {
        case SLAMCU_OK: return "ok";
        case SLAMCU_INVALID_ARGUMENT: return "invalid argument";
        case SLAMCU_EMPTY_INPUT: return "Empty descriptors provided.";
        case SLAMCU_SIZE_MISMATCH: return "size mismatch";
        case SLAMCU_CAPACITY: return "capacity exceeded";
        case SLAMCU_CUDA_ERROR: return "CUDA error";
        case SLAMCU_UNSUPPORTED: return "unsupported";
        default: return "unknown status";
    }
}

int slamcu_device_count(int* count) {
    if (!count) return SLAMCU_INVALID_ARGUMENT;
    cudaError_t e = cudaGetDeviceCount(count);
    if (e != cudaSuccess) {
        *count = 0;
        return SLAMCU_CUDA_ERROR;
    }
    return SLAMCU_OK;
}

int slamcu_create(int device_id, slamcu_context** out) {
    if (!out) return SLAMCU_INVALID_ARGUMENT;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0) return SLAMCU_CUDA_ERROR;  // no CPU fallback
    if (device_id < 0 || device_id >= count) return SLAMCU_INVALID_ARGUMENT;
    slamcu_context* ctx = new (std::nothrow) slamcu_context();
    if (!ctx) return SLAMCU_CUDA_ERROR;
    ctx->device = device_id;
    if (cudaSetDevice(device_id) != cudaSuccess ||
        cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete ctx;
        return SLAMCU_CUDA_ERROR;
    }
    ctx->stream = ctx->own_stream;
    cudaDeviceGetAttribute(&ctx->smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device_id);
    init_sortnms_attributes(ctx->smem_optin);
    init_orb_attributes(ctx->smem_optin);
    init_essential_attributes();
    init_match_attributes();
    init_match_tc_attributes();
    *out = ctx;
    return SLAMCU_OK;
}

void slamcu_destroy(slamcu_context* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (ctx->scratch) cudaFree(ctx->scratch);
    if (ctx->s_in) { cudaStreamSynchronize(ctx->s_in); cudaStreamDestroy(ctx->s_in); }
    if (ctx->s_out) { cudaStreamSynchronize(ctx->s_out); cudaStreamDestroy(ctx->s_out); }
    if (ctx->s_dl) { cudaStreamSynchronize(ctx->s_dl); cudaStreamDestroy(ctx->s_dl); }
    for (cudaEvent_t e : ctx->events) cudaEventDestroy(e);
    if (ctx->s_aux) { cudaStreamSynchronize(ctx->s_aux); cudaStreamDestroy(ctx->s_aux); }
    if (ctx->lane1) { cudaStreamSynchronize(ctx->lane1); cudaStreamDestroy(ctx->lane1); }
    if (ctx->lane1_aux) { cudaStreamSynchronize(ctx->lane1_aux); cudaStreamDestroy(ctx->lane1_aux); }
    for (cudaEvent_t e : {ctx->lane1_fork, ctx->lane1_join, ctx->ev_lane})
        if (e) cudaEventDestroy(e);
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    delete ctx;
}

const char* slamcu_last_error(const slamcu_context* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int slamcu_set_stream(slamcu_context* ctx, void* cuda_stream) {
    if (!ctx) return bad_args(ctx, __func__);
    ctx->stream = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : ctx->own_stream;
    return SLAMCU_OK;
}
void* slamcu_get_stream(slamcu_context* ctx) { return ctx ? ctx->stream : nullptr; }

static int seq_dense_flush(slamcu_sequence* s);
int slamcu_synchronize(slamcu_context* ctx) {
    if (!ctx) return bad_args(ctx, __func__);
    while (!ctx->pending_dl.empty()) {  // dense outputs still on the device: download them now
        const int rc = seq_dense_flush(ctx->pending_dl.back());
        if (rc != SLAMCU_OK) return rc;
    }
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->lane1) CU(ctx, cudaStreamSynchronize(ctx->lane1));
    if (ctx->s_in) CU(ctx, cudaStreamSynchronize(ctx->s_in));
    if (ctx->s_out) CU(ctx, cudaStreamSynchronize(ctx->s_out));
    if (ctx->s_dl) CU(ctx, cudaStreamSynchronize(ctx->s_dl));
    return SLAMCU_OK;
}
int64_t slamcu_launch_count(const slamcu_context* ctx) { return ctx ? ctx->launches : 0; }

int slamcu_alloc_pinned(size_t bytes, void** out) {
    if (!out) return SLAMCU_INVALID_ARGUMENT;
    *out = nullptr;
    return cudaMallocHost(out, bytes ? bytes : 1) == cudaSuccess ? SLAMCU_OK : SLAMCU_CUDA_ERROR;
}
void slamcu_free_pinned(void* p) {
    if (p) cudaFreeHost(p);
}

int slamcu_popc_peak(slamcu_context* ctx, double* gpopc_per_s) {
    if (!ctx || !gpopc_per_s) return bad_args(ctx, __func__);
    CU(ctx, cudaSetDevice(ctx->device));
    int rc = ensure_scratch(ctx, 4096);
    if (rc != SLAMCU_OK) return rc;
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device);
    cudaEvent_t a, b;
    CU(ctx, cudaEventCreate(&a));
    CU(ctx, cudaEventCreate(&b));
    double best = 0.0;
    for (int rep = 0; rep < 5; rep++) {
        CU(ctx, cudaEventRecord(a, ctx->stream));
        const long long ops = launch_popc_peak(static_cast<unsigned*>(ctx->scratch), sms * 8, 1 << 14, ctx->stream);
        ctx->launches++;
        CU(ctx, cudaEventRecord(b, ctx->stream));
        CU(ctx, cudaEventSynchronize(b));
        float ms = 0.f;
        cudaEventElapsedTime(&ms, a, b);
        if (ms > 0.f) best = std::max(best, (double)ops / (ms * 1e-3) / 1e9);
    }
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    *gpopc_per_s = best;
    return check_launch(ctx, "popc_peak");
}

int slamcu_debug_trip_bound(slamcu_context* ctx) {
    if (!ctx) return SLAMCU_INVALID_ARGUMENT;
    launch_trip_bound(ctx->stream);
    return check_launch(ctx, "trip_bound");
}

int slamcu_profile_enable(slamcu_context* ctx, int on) {
    if (!ctx) return bad_args(ctx, __func__);
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->prof.reset();
    ctx->profiling = on != 0;
    return SLAMCU_OK;
}

int slamcu_profile_read(slamcu_context* ctx, int index, char* name, int name_cap, double* total_ms, int64_t* launches) {
    if (!ctx) return bad_args(ctx, __func__);
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->prof.collect();
    if (index < 0 || index >= (int)ctx->prof.acc.size()) return SLAMCU_INVALID_ARGUMENT;  // end of list
    const auto& a = ctx->prof.acc[index];
    if (name && name_cap > 0) { strncpy(name, a.name, name_cap - 1); name[name_cap - 1] = 0; }
    if (total_ms) *total_ms = a.ms;
    if (launches) *launches = a.count;
    return SLAMCU_OK;
}

/* ---------------------------------------------------------------------------------------------- */
/* sequences                                                                                       */
/* ---------------------------------------------------------------------------------------------- */
int slamcu_sequence_create(slamcu_context* ctx, int rows, int cols, int max_frames, int max_raw, int max_kp,
                           int desc_bytes, slamcu_sequence** out) {
    if (!ctx || !out) return bad_args(ctx, __func__);
    *out = nullptr;
    // frames / frame pairs index grid dimensions y and z of the batched kernels: at most 65535 per sequence
    if (rows <= 0 || cols <= 0 || rows > 65535 || cols > 65535 || max_frames <= 0 || max_frames > 65535 || desc_bytes <= 0 || desc_bytes > 256)
        return fail(ctx, SLAMCU_INVALID_ARGUMENT, "bad sequence geometry %dx%d x%d desc %d", rows, cols, max_frames,
                    desc_bytes);
    CU(ctx, cudaSetDevice(ctx->device));
    slamcu_sequence* s = new (std::nothrow) slamcu_sequence();
    if (!s) return SLAMCU_CUDA_ERROR;
    s->ctx = ctx;
    s->max_frames = max_frames;
    SeqView& v = s->v;
    v.rows = rows;
    v.cols = cols;
    v.pitch = round_up(cols, 128);
    v.mwords = (cols + 31) / 32;
    const long long px = (long long)rows * cols;
    if (max_raw <= 0) max_raw = (int)std::min<long long>(std::max<long long>(px / 20, 8192), kMaxRawCap);
    max_raw = round_up(max_raw, 32);  // keeps every per-frame sub-array 16-byte aligned
    if (max_raw > kMaxRawCap) max_raw = kMaxRawCap;
    if (max_kp <= 0) max_kp = (int)std::min<long long>(std::max<long long>(px / 100, 1024), 65536);
    v.cap_raw = max_raw;
    v.cap_kp = max_kp;
    v.desc_bytes = desc_bytes;
    v.desc_words = (desc_bytes + 3) / 4;
    v.qcap = round_up(max_raw / 16 + 2, 4);
    v.frame_bytes = (size_t)rows * v.pitch;
    v.scratch_words = (2 * (size_t)max_raw + 4 * (size_t)v.qcap + (size_t)max_raw / 32 + 2 + 3) / 4 * 4;
    const size_t F = (size_t)max_frames;
    int rc = SLAMCU_OK;
    auto A = [&](auto** p, size_t count, bool zero) {
        if (rc == SLAMCU_OK) rc = dev_alloc(ctx, p, count, s->owned, zero);
    };
    A(&v.img, F * v.frame_bytes, true);
    A(&v.blur, F * v.frame_bytes, true);
    A(&v.mask, F * rows * v.mwords, false);
    A(&v.raw_xy, F * max_raw, false);
    A(&v.keys, F * max_raw, false);
    A(&v.sort_scratch, F * v.scratch_words, false);
    const size_t bm_bytes = (size_t)rows * v.mwords * 4;
    if (bm_bytes > (size_t)(ctx->smem_optin - 1024)) A(&v.nms_bitmap, F * rows * v.mwords, false);
    A(&v.n_raw, F, true);
    A(&v.kps, F * max_kp, false);
    A(&v.n_kp, F, true);
    A(&v.desc, F * max_kp * v.desc_words, true);
    A(&v.desc_or, F * v.desc_words, true);
    A(&v.cand, F * max_kp, false);
    A(&v.matches, F * max_kp, false);
    A(&v.n_match, F, true);
    A(&v.status, F, true);
    A(&s->sort_keys, F * max_kp, false);
    if (rc == SLAMCU_OK && cudaMallocHost(reinterpret_cast<void**>(&s->h_counts), F * 4 * sizeof(int)) != cudaSuccess)
        rc = fail(ctx, SLAMCU_CUDA_ERROR, "cudaMallocHost failed");
    if (rc == SLAMCU_OK && (cudaEventCreateWithFlags(&s->ev_compute_done, cudaEventDisableTiming) != cudaSuccess ||
                            cudaEventCreateWithFlags(&s->ev_out_done, cudaEventDisableTiming) != cudaSuccess))
        rc = fail(ctx, SLAMCU_CUDA_ERROR, "cudaEventCreate failed");
    if (rc != SLAMCU_OK) {
        slamcu_sequence_destroy(s);
        return rc;
    }
    *out = s;
    return SLAMCU_OK;
}

void slamcu_sequence_destroy(slamcu_sequence* s) {
    if (!s) return;
    cudaSetDevice(s->ctx->device);
    cudaStreamSynchronize(s->ctx->stream);
    if (s->ctx->s_in) cudaStreamSynchronize(s->ctx->s_in);
    if (s->ctx->s_out) cudaStreamSynchronize(s->ctx->s_out);
    if (s->ev_compute_done) cudaEventDestroy(s->ev_compute_done);
    if (s->ev_out_done) cudaEventDestroy(s->ev_out_done);
    for (void* p : s->owned) cudaFree(p);
    for (void* p : s->orb_owned) cudaFree(p);
    if (s->stage) cudaFree(s->stage);
    if (s->prep_stage) cudaFree(s->prep_stage);
    if (s->undist_map) cudaFree(s->undist_map);
    for (void* p : s->ess_owned) cudaFree(p);
    if (s->ess_work) cudaFree(s->ess_work);
    if (s->h_counts) cudaFreeHost(s->h_counts);
    if (s->h_status) cudaFreeHost(s->h_status);
    if (s->d_dense) cudaFree(s->d_dense);
    if (s->dd_kps) cudaFree(s->dd_kps);
    if (s->dd_desc) cudaFree(s->dd_desc);
    if (s->dd_matches) cudaFree(s->dd_matches);
    if (s->h_tot) cudaFreeHost(s->h_tot);
    {
        auto& pd = s->ctx->pending_dl;
        pd.erase(std::remove(pd.begin(), pd.end(), s), pd.end());
    }
    if (s->tc_x8) cudaFree(s->tc_x8);
    if (s->tc_ck) cudaFree(s->tc_ck);
    delete s;
}

// Host frames -> the pitched frame store.  Dense frames (stride == cols) take one linear copy into a staging
// block plus a re-pitch kernel; a strided cudaMemcpy2D is kept only for other strides.  `copy_stream` carries
// the copy, `kernel_stream` the re-pitch kernel (the caller orders the two).
static int seq_upload_async(slamcu_sequence* s, int first, int n, const uint8_t* host, int stride, cudaStream_t copy_stream,
                            bool* needs_repitch) {
    slamcu_context* ctx = s->ctx;
    const SeqView& v = s->v;
    *needs_repitch = false;
    if (stride == v.pitch) {
        CU(ctx, cudaMemcpyAsync(v.img + (size_t)first * v.frame_bytes, host, (size_t)n * v.frame_bytes, cudaMemcpyHostToDevice,
                                copy_stream));
        return SLAMCU_OK;
    }
    if (stride == v.cols) {
        if (!s->stage) CU(ctx, cudaMalloc(reinterpret_cast<void**>(&s->stage), (size_t)s->max_frames * v.rows * v.cols));
        const size_t fb = (size_t)v.rows * v.cols;
        CU(ctx, cudaMemcpyAsync(s->stage + (size_t)first * fb, host, (size_t)n * fb, cudaMemcpyHostToDevice, copy_stream));
        *needs_repitch = true;
        return SLAMCU_OK;
    }
    CU(ctx, cudaMemcpy2DAsync(v.img + (size_t)first * v.frame_bytes, v.pitch, host, stride, v.cols, (size_t)v.rows * n,
                              cudaMemcpyHostToDevice, copy_stream));
    return SLAMCU_OK;
}

static void seq_repitch(slamcu_sequence* s, int first, int n, cudaStream_t st) {
    const SeqView& v = s->v;
    s->ctx->launches += launch_repitch(s->stage + (size_t)first * v.rows * v.cols, v.cols, v.img + (size_t)first * v.frame_bytes,
                                       v.pitch, v.cols, (long long)n * v.rows, st);
}

int slamcu_sequence_upload(slamcu_sequence* s, int first, int n, const uint8_t* host, int stride) {
    if (!s || !host) return bad_args((s ? s->ctx : nullptr), __func__);
    slamcu_context* ctx = s->ctx;
    if (first < 0 || n < 0 || first + n > s->max_frames || stride < s->v.cols)
        return fail(ctx, SLAMCU_INVALID_ARGUMENT, "upload range [%d,%d) / stride %d invalid", first, first + n, stride);
    if (n == 0) return SLAMCU_OK;
    CU(ctx, cudaSetDevice(ctx->device));
    bool rp = false;
    int rc = seq_upload_async(s, first, n, host, stride, ctx->stream, &rp);
    if (rc != SLAMCU_OK) return rc;
    if (rp) {
        ProfGuard pg(ctx);
        seq_repitch(s, first, n, ctx->stream);
        return check_launch(ctx, "repitch");
    }
    return SLAMCU_OK;
}

int slamcu_sequence_prepare(slamcu_sequence* s, int first, int n, const uint8_t* host, int channels, int stride, const double* K4,
                            const double* D4) {
    if (!s || !host) return bad_args((s ? s->ctx : nullptr), __func__);
    slamcu_context* ctx = s->ctx;
    const SeqView& v = s->v;
    if (first < 0 || n < 0 || first + n > s->max_frames || (channels != 1 && channels != 3) || stride < v.cols * channels)
        return fail(ctx, SLAMCU_INVALID_ARGUMENT, "prepare: bad range / channels / stride");
    if ((K4 == nullptr) != (D4 == nullptr)) return fail(ctx, SLAMCU_INVALID_ARGUMENT, "prepare: K4 and D4 go together");
    if (n == 0) return SLAMCU_OK;
    CU(ctx, cudaSetDevice(ctx->device));
    const size_t fb = (size_t)v.rows * stride, need = (size_t)n * fb;
    if (s->prep_stage_bytes < need) {
        CU(ctx, cudaStreamSynchronize(ctx->stream));
        if (s->prep_stage) cudaFree(s->prep_stage);
        s->prep_stage = nullptr;
        s->prep_stage_bytes = 0;
        CU(ctx, cudaMalloc(reinterpret_cast<void**>(&s->prep_stage), need));
        s->prep_stage_bytes = need;
    }
    ProfGuard pg(ctx);
    if (K4) {
        const double key[8] = {K4[0], K4[1], K4[2], K4[3], D4[0], D4[1], D4[2], D4[3]};
        if (!s->has_undist || memcmp(key, s->undist_key, sizeof key) != 0) {
            if ((long long)v.rows * v.cols > INT_MAX) return fail(ctx, SLAMCU_UNSUPPORTED, "undistortion map: rows * cols exceeds INT_MAX");
            if (!s->undist_map)
                CU(ctx, cudaMalloc(reinterpret_cast<void**>(&s->undist_map), ((size_t)v.rows * v.cols + 1 + kUndistFixCap) * sizeof(int)));
            CamParams cam{K4[0], K4[1], K4[2], K4[3], D4[0], D4[1], D4[2], D4[3]};
            int rcm = build_undistort_map(ctx, v.rows, v.cols, cam, s->undist_map, s->undist_map + (size_t)v.rows * v.cols);
            if (rcm != SLAMCU_OK) return rcm;
            memcpy(s->undist_key, key, sizeof key);
            s->has_undist = true;
        }
    }
    CU(ctx, cudaMemcpyAsync(s->prep_stage, host, need, cudaMemcpyHostToDevice, ctx->stream));
    ctx->launches += launch_prepare(s->prep_stage, channels, stride, fb, K4 ? s->undist_map : nullptr, v.img + (size_t)first * v.frame_bytes,
                                    v.pitch, v.frame_bytes, v.rows, v.cols, n, ctx->stream);
    return check_launch(ctx, "prepare kernels");
}

int slamcu_sequence_image(slamcu_sequence* s, int f, uint8_t* out, int out_stride) {
    if (!s || !out) return bad_args((s ? s->ctx : nullptr), __func__);
    slamcu_context* ctx = s->ctx;
    if (f < 0 || f >= s->max_frames || out_stride < s->v.cols) return fail(ctx, SLAMCU_INVALID_ARGUMENT, "bad frame index / stride");
    CU(ctx, cudaSetDevice(ctx->device));
    CU(ctx, cudaMemcpy2DAsync(out, out_stride, s->v.img + (size_t)f * s->v.frame_bytes, s->v.pitch, s->v.cols, s->v.rows,
                              cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return SLAMCU_OK;
}

int slamcu_sequence_frames_device(slamcu_sequence* s, void** dptr, int* pitch, int64_t* frame_bytes) {
    if (!s) return bad_args((s ? s->ctx : nullptr), __func__);
    if (dptr) *dptr = s->v.img;
    if (pitch) *pitch = s->v.pitch;
    if (frame_bytes) *frame_bytes = (int64_t)s->v.frame_bytes;
    return SLAMCU_OK;
}

static int seq_detect(slamcu_sequence* s, slamcu_detector* det, int first, int n, bool raw_probe) {
    slamcu_context* ctx = s->ctx;
    ProfGuard pg(ctx);
    ctx->launches += launch_fast_corners(s->v, first, n, det->p, ctx->stream);
    if (raw_probe) ctx->launches += launch_raster_keypoints(s->v, first, n, true, ctx->stream);
    else if (det->p.nms) ctx->launches += launch_sort_nms(s->v, first, n, det->p, ctx->smem_optin, ctx->stream);
    else ctx->launches += launch_raster_keypoints(s->v, first, n, false, ctx->stream);
    return check_launch(ctx, "detect kernels");
}

static int seq_compute(slamcu_sequence* s, slamcu_detector* det, int first, int n) {
    slamcu_context* ctx = s->ctx;
    ProfGuard pg(ctx);
    ctx->launches += launch_blur(s->v, first, n, det->p, ctx->stream);
    ctx->launches += launch_describe(s->v, first, n, det->p, det->d_pattern, ctx->stream);
    return check_launch(ctx, "compute kernels");
}

// INTER_LINEAR_EXACT coefficient tables (OpenCV resize.cpp bit-exact path): first source index and the 8.8
// fixed-point weight of the second tap, computed in double exactly like OpenCV, packed (index << 8) | weight.
static void resize_table(int dst, int src, std::vector<uint32_t>& tab) {
    tab.resize(dst);
    const double scale = 1.0 / ((double)dst / (double)src);
    for (int d = 0; d < dst; d++) {
        const double f = scale * (d + 0.5) - 0.5;
        int i = (int)std::floor(f);
        int w = (int)std::nearbyint((f - i) * 256.0);  // cvRound: half to even
        if (i < 0) { i = 0; w = 0; }
        if (i >= src - 1) { i = src - 1; w = 0; }
        if (w == 256) { i += 1; w = 0; }  // (256 - 256) * s[i] + 256 * s[i + 1]  ==  256 * s[i + 1] + 0 * s[i + 2]
        tab[d] = ((uint32_t)i << 8) | (uint32_t)w;
    }
}

static int seq_ensure_orb(slamcu_sequence* s, slamcu_detector* det) {
    slamcu_context* ctx = s->ctx;
    if (s->has_orb && s->orb_levels == det->n_levels && s->orb_features == det->max_features &&
        s->orb_scale == det->scale_factor) {
        s->orb.fast_threshold = det->fast_threshold;
        s->orb.pattern = det->d_orb_pattern;
        s->orb.patf = det->d_orb_patf;
        return SLAMCU_OK;
    }
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    for (void* p : s->orb_owned) cudaFree(p);
    s->orb_owned.clear();
    s->has_orb = false;
    OrbView& o = s->orb;
    o = OrbView{};
    const SeqView& v = s->v;
    const int L = det->n_levels;
    o.nlevels = L;
    o.fast_threshold = det->fast_threshold;
    o.pattern = det->d_orb_pattern;
    o.patf = det->d_orb_patf;
    // ORB_Impl: scale = (float)pow(scaleFactor, level); inv_scale = 1.0f / scale; size = cvRound(dim * inv_scale) -- the
    // product, not the quotient: they differ for 81 (size, level) pairs below 2200 px; quotas by geometric series
    const double sf = (double)det->scale_factor;
    const float factor = (float)(1.0 / sf);
    float ndesired = det->max_features * (1 - factor) / (1 - (float)std::pow((double)factor, (double)L));
    int sum = 0;
    size_t pyr = 0, mw = 0, ct = 0, sb = 0;
    for (int l = 0; l < L; l++) {
        OrbLevel& lv = o.lv[l];
        lv.scale = (float)std::pow(sf, (double)l);
        const float inv_scale = 1.0f / lv.scale;
        lv.cols = (int)std::lrintf((float)v.cols * inv_scale);
        lv.rows = (int)std::lrintf((float)v.rows * inv_scale);
        if (lv.rows < 1 || lv.cols < 1) return fail(ctx, SLAMCU_INVALID_ARGUMENT, "pyramid level %d is %dx%d: too small", l, lv.cols, lv.rows);
        if (lv.rows < 8 || lv.cols < 8) {
            // OpenCV still builds such a level, but nothing in it (nor in the smaller ones after it) survives the 31-px
            // border filter and the per-level quotas are fixed up front: the remaining levels are simply not processed
            if (l == 0) return fail(ctx, SLAMCU_INVALID_ARGUMENT, "image is %dx%d: too small", lv.cols, lv.rows);
            o.nlevels = l;
            break;
        }
        lv.pitch = round_up(lv.cols, 128);
        lv.mwords = (lv.cols + 31) / 32;
        if (l == 0) lv.off = 0;
        else { lv.off = pyr; pyr += (size_t)lv.rows * lv.pitch; }
        lv.moff = mw;
        mw += (size_t)lv.rows * lv.mwords;
        lv.soff = sb;
        sb += (size_t)lv.rows * lv.pitch;
        if (l < L - 1) { lv.quota = (int)std::lrintf(ndesired); sum += lv.quota; ndesired *= factor; }
        else lv.quota = std::max(det->max_features - sum, 0);
        // FAST survivors of the strict 3x3 NMS are never adjacent: at most a quarter of the pixels.  A one-frame
        // workspace (the single-image calls) holds that bound; sequences keep 1/16 (status bit on overflow)
        lv.capc = round_up(std::max(2048, lv.rows * lv.cols / (s->max_frames == 1 ? 4 : 16) + 64), 32);
        lv.coff = ct;
        ct += lv.capc;
    }
    o.pyr_bytes = std::max<size_t>(pyr, 16);
    o.mask_words = mw;
    o.cand_total = ct;
    o.score_bytes = sb;
    const size_t F = (size_t)s->max_frames;
    int rc = SLAMCU_OK;
    auto A = [&](auto** p, size_t count, bool zero) {
        if (rc == SLAMCU_OK) rc = dev_alloc(ctx, p, count, s->orb_owned, zero);
    };
    A(&o.pyr, F * o.pyr_bytes, true);
    A(&o.pyrb, F * o.pyr_bytes, true);
    A(&o.mask, F * o.mask_words, false);
    A(&o.fscore, F * o.score_bytes, false);
    A(&o.cxy, F * ct, false);
    A(&o.cscore, F * ct, false);
    A(&o.sxy, F * ct, false);
    A(&o.sresp, F * ct, false);
    A(&o.fxy, F * ct, false);
    A(&o.fresp, F * ct, false);
    A(&o.n_cand, F * kMaxLevels, true);
    A(&o.n_sel, F * kMaxLevels, true);
    A(&o.n_fin, F * kMaxLevels, true);
    A(&o.octave, F * v.cap_kp, true);
    A(&o.lxy, F * v.cap_kp, true);
    for (int l = 1; l < o.nlevels && rc == SLAMCU_OK; l++) {
        std::vector<uint32_t> tx, ty;
        resize_table(o.lv[l].cols, o.lv[l - 1].cols, tx);
        resize_table(o.lv[l].rows, o.lv[l - 1].rows, ty);
        uint32_t *dx = nullptr, *dy = nullptr;
        A(&dx, tx.size(), false);
        A(&dy, ty.size(), false);
        if (rc == SLAMCU_OK && (cudaMemcpy(dx, tx.data(), tx.size() * 4, cudaMemcpyHostToDevice) != cudaSuccess ||
                                cudaMemcpy(dy, ty.data(), ty.size() * 4, cudaMemcpyHostToDevice) != cudaSuccess))
            rc = fail(ctx, SLAMCU_CUDA_ERROR, "table upload failed");
        o.lv[l].xt = dx;
        o.lv[l].yt = dy;
    }
    if (rc != SLAMCU_OK) return rc;
    // TMA descriptors: rank-3 uint8 tensors (x = level width, y = level height, z = frame), zero fill outside.
    // cuTensorMapEncodeTiled is a driver entry point; fetch it through the runtime so libcuda is not a link dependency.
    s->tmaps.valid = false;
    {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess && fn &&
            qres == cudaDriverEntryPointSuccess) {
            auto encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled>(fn);
            bool ok = true;
            for (int l = 0; l < o.nlevels && ok; l++) {
                const OrbLevel& lv = o.lv[l];
                void* base = l == 0 ? (void*)v.img : (void*)(o.pyr + lv.off);
                const cuuint64_t dims[3] = {(cuuint64_t)lv.cols, (cuuint64_t)lv.rows, (cuuint64_t)s->max_frames};
                const cuuint64_t strides[2] = {(cuuint64_t)lv.pitch, (cuuint64_t)(l == 0 ? v.frame_bytes : o.pyr_bytes)};
                const cuuint32_t box[3] = {(cuuint32_t)kFastBoxW, (cuuint32_t)kFastBoxH, 1};
                const cuuint32_t estr[3] = {1, 1, 1};
                ok = encode(&s->tmaps.fast[l], CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, base, dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
                const cuuint32_t bbox[3] = {(cuuint32_t)kBlurBoxW, (cuuint32_t)kBlurBoxH, 1};
                ok = ok && encode(&s->tmaps.blur[l], CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, base, dims, strides, bbox, estr,
                                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
            }
            s->tmaps.valid = ok;
        }
        cudaGetLastError();  // a failed query must not poison later launch checks
    }
    s->has_orb = true;
    s->orb_levels = L;
    s->orb_features = det->max_features;
    s->orb_scale = det->scale_factor;
    return SLAMCU_OK;
}

static int seq_extract_orb(slamcu_sequence* s, slamcu_detector* det, int first, int n) {
    slamcu_context* ctx = s->ctx;
    if (s->v.desc_bytes != 32) return fail(ctx, SLAMCU_SIZE_MISMATCH, "ORB mode needs 32-byte descriptors");
    int rc = seq_ensure_orb(s, det);
    if (rc != SLAMCU_OK) return rc;
    ProfGuard pg(ctx);
    CU(ctx, cudaMemsetAsync(s->v.status + first, 0, (size_t)n * sizeof(int), ctx->stream));
    if (!ctx->s_aux) {
        CU(ctx, cudaStreamCreateWithFlags(&ctx->s_aux, cudaStreamNonBlocking));
        CU(ctx, cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming));
        CU(ctx, cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming));
    }
    ctx->launches += launch_orb_extract(s->v, s->orb, first, n, ctx->stream, ctx->s_aux, ctx->ev_fork, ctx->ev_join, &s->tmaps);
    {
        const SeqView& v = s->v;
        ctx->launches += launch_desc_or(v.desc + (size_t)first * v.cap_kp * v.desc_words, v.n_kp + first, v.desc_words,
                                        v.desc_or + (size_t)first * v.desc_words, ctx->stream, n,
                                        (size_t)v.cap_kp * v.desc_words);
    }
    return check_launch(ctx, "orb kernels");
}

int slamcu_sequence_octaves(slamcu_sequence* s, int f, int32_t* octaves, int capacity) {
    if (!s || !octaves) return bad_args((s ? s->ctx : nullptr), __func__);
    slamcu_context* ctx = s->ctx;
    if (!s->has_orb) return fail(ctx, SLAMCU_UNSUPPORTED, "sequence was not extracted in ORB mode");
    if (f < 0 || f >= s->max_frames) return fail(ctx, SLAMCU_INVALID_ARGUMENT, "bad frame index");
    int32_t c[4];
    int rc = slamcu_sequence_counts(s, f, 1, c);
    if (rc != SLAMCU_OK) return rc;
    if (c[0] > capacity) return fail(ctx, SLAMCU_CAPACITY, "need room for %d octaves", c[0]);
    if (c[0] > 0) {
        CU(ctx, cudaMemcpyAsync(octaves, s->orb.octave + (size_t)f * s->v.cap_kp, (size_t)c[0] * sizeof(int),
                                cudaMemcpyDeviceToHost, ctx->stream));
        CU(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return SLAMCU_OK;
}

int slamcu_sequence_extract(slamcu_sequence* s, slamcu_detector* det, int first, int n) {
    if (!s || !det || s->ctx != det->ctx) return bad_args((s ? s->ctx : nullptr), __func__);
    slamcu_context* ctx = s->ctx;
    if (first < 0 || n < 0 || first + n > s->max_frames) return fail(ctx, SLAMCU_INVALID_ARGUMENT, "bad frame range");
    if (det->mode == SLAMCU_MODE_ORB) return n == 0 ? SLAMCU_OK : seq_extract_orb(s, det, first, n);
    if (det->p.pairs / 8 != s->v.desc_bytes)
        return fail(ctx, SLAMCU_SIZE_MISMATCH, "sequence desc_bytes %d != NumBRIEFPairs/8 = %d", s->v.desc_bytes,
                    det->p.pairs / 8);
    if (n == 0) return SLAMCU_OK;
    int rc = seq_detect(s, det, first, n, false);
    if (rc != SLAMCU_OK) return rc;
    return seq_compute(s, det, first, n);
}

// ---- tensor-core matcher plumbing --------------------------------------------------------------------------------------
// SLAMCU_MATCH_TC=0 keeps the integer-pipe kernels (A/B measurements); the tensor-core path also needs the TMA encoder.
static bool match_tc_enabled() {
    static const bool on = [] { const char* e = getenv("SLAMCU_MATCH_TC"); return !(e && e[0] == '0'); }();
    return on;
}
// rank-2 u8 tensor {256 bytes of K, rows}, box {128, 128}, 128-byte swizzle (the UMMA K-major operand layout), zero fill
static bool encode_tc_map(CUtensorMap* map, void* base, size_t rows) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    bool ok = false;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess && fn &&
        qres == cudaDriverEntryPointSuccess) {
        auto encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled>(fn);
        const cuuint64_t dims[2] = {(cuuint64_t)kTcRowBytes, (cuuint64_t)rows};
        const cuuint64_t strides[1] = {(cuuint64_t)kTcRowBytes};
        const cuuint32_t box[2] = {128, 128};
        const cuuint32_t estr[2] = {1, 1};
        ok = encode(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
    }
    cudaGetLastError();
    return ok;
}
static bool seq_ensure_tc(slamcu_sequence* s) {
    if (s->tc_state != 0) return s->tc_state > 0;
    s->tc_state = -1;
    if (!match_tc_enabled() || s->v.desc_words != 8) return false;
    const int rows = (s->v.cap_kp + kTcRowAlign - 1) / kTcRowAlign * kTcRowAlign;
    const size_t total = (size_t)s->max_frames * rows;
    if (cudaMalloc(reinterpret_cast<void**>(&s->tc_x8), total * kTcRowBytes) != cudaSuccess ||
        cudaMalloc(reinterpret_cast<void**>(&s->tc_ck), total * sizeof(uint32_t)) != cudaSuccess ||
        !encode_tc_map(&s->tc.map_q, s->tc_x8, total)) {
        cudaGetLastError();  // out of memory or no encoder: the integer-pipe kernels take over
        if (s->tc_x8) cudaFree(s->tc_x8);
        if (s->tc_ck) cudaFree(s->tc_ck);
        s->tc_x8 = nullptr;
        s->tc_ck = nullptr;
        return false;
    }
    s->tc.map_t = s->tc.map_q;
    s->tc_rows = rows;
    s->tc_state = 1;
    return true;
}

int slamcu_sequence_match(slamcu_sequence* s, slamcu_matcher* m, int first, int n_pairs, int with_kp) {
    if (!s || !m || s->ctx != m->ctx) return bad_args((s ? s->ctx : nullptr), __func__);
    slamcu_context* ctx = s->ctx;
    if (first < 0 || n_pairs < 0 || first + n_pairs + 1 > s->max_frames)
        return fail(ctx, SLAMCU_INVALID_ARGUMENT, "bad pair range");
    if (m->distance_type != SLAMCU_DISTANCE_HAMMING)
        return fail(ctx, SLAMCU_UNSUPPORTED, "L2 distance requires float descriptors. Use the float overload.");
    if (n_pairs == 0) return SLAMCU_OK;
    ProfGuard pg(ctx);
    const SeqView& v = s->v;
    MatchJob j{};
    const size_t dstride = (size_t)v.cap_kp * v.desc_words;
    j.dq = v.desc + (size_t)first * dstride;
    j.dt = j.dq + dstride;
    j.kq = v.kps + (size_t)first * v.cap_kp;
    j.kt = j.kq + v.cap_kp;
    j.nq = v.n_kp + first;
    j.nt = v.n_kp + first + 1;
    j.orq = v.desc_or + (size_t)first * v.desc_words;
    j.ort = j.orq + v.desc_words;
    j.cand = v.cand + (size_t)first * v.cap_kp;
    j.matches = v.matches + (size_t)first * v.cap_kp;
    j.n_match = v.n_match + first;
    j.status = v.status + first;
    j.desc_pair_stride = dstride;
    j.kp_pair_stride = v.cap_kp;
    j.count_stride = 1;
    j.cand_pair_stride = v.cap_kp;
    j.or_stride = v.desc_words;
    j.desc_words = v.desc_words;
    j.max_q = v.cap_kp;
    j.max_t = v.cap_kp;
    j.cap_out = v.cap_kp;
    const MatchTc* tc = nullptr;
    if (!with_kp && seq_ensure_tc(s)) {
        // every frame of the range once (a frame is the train set of one pair and the query set of the next)
        ctx->launches += launch_expand_bits(j.dq, dstride, j.nq, 1, n_pairs + 1, s->tc_rows, s->tc_x8 + (size_t)first * s->tc_rows * kTcRowBytes,
                                            s->tc_ck + (size_t)first * s->tc_rows, ctx->stream);
        s->tc.view = MatchTcView{s->tc_ck, s->tc_ck, first * s->tc_rows, (first + 1) * s->tc_rows, s->tc_rows, s->tc_rows};
        tc = &s->tc;
    }
    ctx->launches += launch_match(j, n_pairs, m->p, true, with_kp ? 1 : 0, s->sort_keys + (size_t)first * v.cap_kp,
                                  ctx->stream, 1, 0, 0, tc);
    return check_launch(ctx, "match kernels");
}

static int ctx_pipeline_resources(slamcu_context* ctx, size_t n_events);

// ---- two compute lanes ------------------------------------------------------------------------------------------------
static int ctx_lanes(slamcu_context* ctx) {
    if (ctx->lane1) return SLAMCU_OK;
    const char* one = getenv("SLAMCU_ONE_LANE");
    ctx->two_lanes = !(one && one[0] == '1');
    CU(ctx, cudaStreamCreateWithFlags(&ctx->lane1, cudaStreamNonBlocking));
    CU(ctx, cudaStreamCreateWithFlags(&ctx->lane1_aux, cudaStreamNonBlocking));
    CU(ctx, cudaEventCreateWithFlags(&ctx->lane1_fork, cudaEventDisableTiming));
    CU(ctx, cudaEventCreateWithFlags(&ctx->lane1_join, cudaEventDisableTiming));
    CU(ctx, cudaEventCreateWithFlags(&ctx->ev_lane, cudaEventDisableTiming));
    if (!ctx->s_aux) {
        CU(ctx, cudaStreamCreateWithFlags(&ctx->s_aux, cudaStreamNonBlocking));
        CU(ctx, cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming));
        CU(ctx, cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming));
    }
    return SLAMCU_OK;
}

// While alive, the context's entry points enqueue on lane 1 instead of lane 0 (every launcher reads ctx->stream / s_aux / ev_*).
struct LaneScope {
    slamcu_context* ctx;
    bool on;
    LaneScope(slamcu_context* c, int lane) : ctx(c), on(lane == 1) {
        if (on) swap();
    }
    ~LaneScope() {
        if (on) swap();
    }
    void swap() {
        std::swap(ctx->stream, ctx->lane1);
        std::swap(ctx->s_aux, ctx->lane1_aux);
        std::swap(ctx->ev_fork, ctx->lane1_fork);
        std::swap(ctx->ev_join, ctx->lane1_join);
    }
};

// detectAndCompute on frames [first, first + n) and match(f, f + 1) on the pairs they complete, in chunks that alternate between
// the two compute lanes.  chunk c's matches need the last frame of chunk c - 1 (other lane): one event per chunk.  On return
// everything is ordered before later work on the context's stream.  While per-kernel timing is on, one lane is used (overlapped
// kernels would each be charged the other's time).
int slamcu_sequence_extract_match(slamcu_sequence* s, slamcu_detector* det, slamcu_matcher* m, int first, int n, int with_keypoints, int chunk) {
    if (!s || !det || !m || s->ctx != det->ctx || s->ctx != m->ctx) return bad_args((s ? s->ctx : nullptr), __func__);
    slamcu_context* ctx = s->ctx;
    if (first < 0 || n < 0 || first + n > s->max_frames) return fail(ctx, SLAMCU_INVALID_ARGUMENT, "bad frame range");
    if (n == 0) return SLAMCU_OK;
    CU(ctx, cudaSetDevice(ctx->device));
    int rc = ctx_lanes(ctx);
    if (rc != SLAMCU_OK) return rc;
    // measured on B200 (1000 frames, ORB mode): with inputs resident, one lane is 1.5 % FASTER than two (14.27 vs 14.49 ms) --
    // the kernels of both steps are issue-bound, co-residency buys nothing and chunking costs tails -- so chunk <= 0 = one lane;
    // the end-to-end path, where lanes also hide copy latencies, gains 3 % from alternating lanes (slamcu_sequence_process)
    const bool pipelined = chunk > 0 && ctx->two_lanes && !ctx->profiling && n > chunk;
    if (!pipelined) {
        rc = slamcu_sequence_extract(s, det, first, n);
        if (rc == SLAMCU_OK && n > 1) rc = slamcu_sequence_match(s, m, first, n - 1, with_keypoints);
        return rc;
    }
    const int n_chunks = (n + chunk - 1) / chunk;
    rc = ctx_pipeline_resources(ctx, (size_t)2 * n_chunks + 2);
    if (rc != SLAMCU_OK) return rc;
    CU(ctx, cudaEventRecord(ctx->ev_lane, ctx->stream));  // lane 1 starts after everything already queued on the context
    CU(ctx, cudaStreamWaitEvent(ctx->lane1, ctx->ev_lane, 0));
    for (int c = 0; c < n_chunks; c++) {
        const int f0 = first + c * chunk, cnt = std::min(chunk, first + n - f0);
        LaneScope lane(ctx, c & 1);
        rc = slamcu_sequence_extract(s, det, f0, cnt);
        if (rc != SLAMCU_OK) return rc;
        CU(ctx, cudaEventRecord(ctx->events[c], ctx->stream));
        const int p0 = std::max(f0 - 1, first), np = f0 + cnt - 1 - p0;
        if (np > 0) {
            if (c > 0) CU(ctx, cudaStreamWaitEvent(ctx->stream, ctx->events[c - 1], 0));
            rc = slamcu_sequence_match(s, m, p0, np, with_keypoints);
            if (rc != SLAMCU_OK) return rc;
        }
    }
    CU(ctx, cudaEventRecord(ctx->ev_lane, ctx->lane1));
    CU(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_lane, 0));
    return SLAMCU_OK;
}

int slamcu_sequence_counts(slamcu_sequence* s, int first, int n, int32_t* counts4) {
    if (!s || !counts4) return bad_args((s ? s->ctx : nullptr), __func__);
    slamcu_context* ctx = s->ctx;
    if (first < 0 || n < 0 || first + n > s->max_frames) return fail(ctx, SLAMCU_INVALID_ARGUMENT, "bad frame range");
    if (n == 0) return SLAMCU_OK;
    int* h = s->h_counts;
    const size_t F = (size_t)s->max_frames;
    CU(ctx, cudaMemcpyAsync(h, s->v.n_kp + first, n * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaMemcpyAsync(h + F, s->v.n_match + first, n * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaMemcpyAsync(h + 2 * F, s->v.n_raw + first, n * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaMemcpyAsync(h + 3 * F, s->v.status + first, n * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < n; i++) {
        counts4[4 * i + 0] = h[i];
        counts4[4 * i + 1] = h[F + i];
        counts4[4 * i + 2] = h[2 * F + i];
        counts4[4 * i + 3] = h[3 * F + i];
    }
    return SLAMCU_OK;
}

int slamcu_sequence_frame(slamcu_sequence* s, int f, slamcu_keypoint* kps, uint8_t* desc, int desc_stride, int capacity,
                          int* n_out) {
    if (!s || !n_out) return bad_args((s ? s->ctx : nullptr), __func__);
    slamcu_context* ctx = s->ctx;
    if (f < 0 || f >= s->max_frames) return fail(ctx, SLAMCU_INVALID_ARGUMENT, "bad frame index");
    int32_t c[4];
    int rc = slamcu_sequence_counts(s, f, 1, c);
    if (rc != SLAMCU_OK) return rc;
    *n_out = c[0];
    if (c[3] & (kStRawOverflow | kStKpOverflow))
        return fail(ctx, SLAMCU_CAPACITY, "frame %d overflowed a device list (status %d): raise max_raw_corners / max_keypoints",
                    f, c[3]);
    if (c[0] > capacity) return fail(ctx, SLAMCU_CAPACITY, "need room for %d keypoints, capacity %d", c[0], capacity);
    if (c[0] == 0) return SLAMCU_OK;
    const SeqView& v = s->v;
    if (kps)
        CU(ctx, cudaMemcpyAsync(kps, v.kps + (size_t)f * v.cap_kp, c[0] * sizeof(slamcu_keypoint), cudaMemcpyDeviceToHost,
                                ctx->stream));
    if (desc) {
        if (desc_stride < v.desc_bytes) return fail(ctx, SLAMCU_INVALID_ARGUMENT, "desc_stride too small");
        CU(ctx, cudaMemcpy2DAsync(desc, desc_stride, v.desc + (size_t)f * v.cap_kp * v.desc_words, v.desc_words * 4,
                                  v.desc_bytes, c[0], cudaMemcpyDeviceToHost, ctx->stream));
    }
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return SLAMCU_OK;
}

int slamcu_sequence_matches(slamcu_sequence* s, int f, slamcu_dmatch* matches, int capacity, int* n_out) {
    if (!s || !n_out) return bad_args((s ? s->ctx : nullptr), __func__);
    slamcu_context* ctx = s->ctx;
    if (f < 0 || f >= s->max_frames) return fail(ctx, SLAMCU_INVALID_ARGUMENT, "bad frame index");
    int32_t c[4];
    int rc = slamcu_sequence_counts(s, f, 1, c);
    if (rc != SLAMCU_OK) return rc;
    *n_out = c[1];
    if (c[3] & kStMatchOverflow) return fail(ctx, SLAMCU_CAPACITY, "match list overflow on pair %d", f);
    if (c[1] > capacity) return fail(ctx, SLAMCU_CAPACITY, "need room for %d matches, capacity %d", c[1], capacity);
    if (c[1] && matches) {
        CU(ctx, cudaMemcpyAsync(matches, s->v.matches + (size_t)f * s->v.cap_kp, c[1] * sizeof(slamcu_dmatch),
                                cudaMemcpyDeviceToHost, ctx->stream));
        CU(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return SLAMCU_OK;
}

// descriptor block of frames [first, first+n) -> host [n][cap_kp][desc_bytes]; one linear copy when the HBM rows
// are not padded (a 2-D copy of 32-byte rows runs at less than half the link rate)
static int seq_download_desc(slamcu_sequence* s, int first, int n, uint8_t* desc, cudaStream_t st) {
    slamcu_context* ctx = s->ctx;
    const SeqView& v = s->v;
    const uint32_t* src = v.desc + (size_t)first * v.cap_kp * v.desc_words;
    if (v.desc_bytes == v.desc_words * 4)
        CU(ctx, cudaMemcpyAsync(desc, src, (size_t)n * v.cap_kp * v.desc_bytes, cudaMemcpyDeviceToHost, st));
    else
        CU(ctx, cudaMemcpy2DAsync(desc, v.desc_bytes, src, v.desc_words * 4, v.desc_bytes, (size_t)n * v.cap_kp,
                                  cudaMemcpyDeviceToHost, st));
    return SLAMCU_OK;
}

int slamcu_sequence_download(slamcu_sequence* s, int first, int n, slamcu_keypoint* kps, uint8_t* desc, slamcu_dmatch* matches,
                             int32_t* counts4) {
    if (!s) return bad_args((s ? s->ctx : nullptr), __func__);
    slamcu_context* ctx = s->ctx;
    if (first < 0 || n < 0 || first + n > s->max_frames) return fail(ctx, SLAMCU_INVALID_ARGUMENT, "bad frame range");
    if (n == 0) return SLAMCU_OK;
    const SeqView& v = s->v;
    if (kps)
        CU(ctx, cudaMemcpyAsync(kps, v.kps + (size_t)first * v.cap_kp, (size_t)n * v.cap_kp * sizeof(slamcu_keypoint),
                                cudaMemcpyDeviceToHost, ctx->stream));
    if (desc) {
        int rc = seq_download_desc(s, first, n, desc, ctx->stream);
        if (rc != SLAMCU_OK) return rc;
    }
    if (matches)
        CU(ctx, cudaMemcpyAsync(matches, v.matches + (size_t)first * v.cap_kp, (size_t)n * v.cap_kp * sizeof(slamcu_dmatch),
                                cudaMemcpyDeviceToHost, ctx->stream));
    if (counts4) {
        // strided D2H of the four count arrays into [n][4]
        CU(ctx, cudaMemcpy2DAsync(counts4 + 0, 16, v.n_kp + first, 4, 4, n, cudaMemcpyDeviceToHost, ctx->stream));
        CU(ctx, cudaMemcpy2DAsync(counts4 + 1, 16, v.n_match + first, 4, 4, n, cudaMemcpyDeviceToHost, ctx->stream));
        CU(ctx, cudaMemcpy2DAsync(counts4 + 2, 16, v.n_raw + first, 4, 4, n, cudaMemcpyDeviceToHost, ctx->stream));
        CU(ctx, cudaMemcpy2DAsync(counts4 + 3, 16, v.status + first, 4, 4, n, cudaMemcpyDeviceToHost, ctx->stream));
    }
    return SLAMCU_OK;
}

// The deferred half of slamcu_sequence_process_dense: the rows are dense in device memory; with the totals on the host the three
// arrays cross the link as exact-size copy-engine transfers.
static int seq_dense_flush(slamcu_sequence* s) {
    if (!s->dl_active) return SLAMCU_OK;
    slamcu_context* ctx = s->ctx;
    s->dl_active = false;
    {
        auto& pd = ctx->pending_dl;
        pd.erase(std::remove(pd.begin(), pd.end(), s), pd.end());
    }
    CU(ctx, cudaEventSynchronize(s->ev_out_done));  // compaction, totals and the overflow word are in
    if (s->h_status && s->h_status[s->max_frames] != 0) return SLAMCU_OK;  // overflow: slamcu_sequence_wait reports it
    const size_t n_kp = (size_t)std::max(s->h_tot[0], 0), n_m = (size_t)std::max(s->h_tot[1], 0);
    if (s->dl_kps && n_kp) CU(ctx, cudaMemcpyAsync(s->dl_kps, s->dd_kps, n_kp * sizeof(slamcu_keypoint), cudaMemcpyDeviceToHost, ctx->s_dl));
    if (s->dl_desc && n_kp) CU(ctx, cudaMemcpyAsync(s->dl_desc, s->dd_desc, n_kp * s->v.desc_words * 4, cudaMemcpyDeviceToHost, ctx->s_dl));
    if (s->dl_matches && n_m) CU(ctx, cudaMemcpyAsync(s->dl_matches, s->dd_matches, n_m * sizeof(slamcu_dmatch), cudaMemcpyDeviceToHost, ctx->s_dl));
    CU(ctx, cudaStreamSynchronize(ctx->s_dl));
    return SLAMCU_OK;
}

int slamcu_sequence_wait(slamcu_sequence* s) {
    if (!s) return bad_args((s ? s->ctx : nullptr), __func__);
    slamcu_context* ctx = s->ctx;
    CU(ctx, cudaEventSynchronize(s->ev_compute_done));
    CU(ctx, cudaEventSynchronize(s->ev_out_done));
    {
        const int rc = seq_dense_flush(s);
        if (rc != SLAMCU_OK) return rc;
    }
    // a frame that overflowed one of its device lists holds truncated results: say so instead of returning them as OK
    for (int f = 0; f < s->h_status_n; f++)
        if (s->h_status[f] != 0) {
            const int st = s->h_status[f];
            s->h_status_n = 0;
            return fail(ctx, SLAMCU_CAPACITY, "frame %d overflowed a device list (status %d: 1 raw corners / candidates, 2 keypoints, 4 matches): "
                        "raise max_raw_corners / max_keypoints", f, st);
        }
    const int dense_over = s->h_status ? s->h_status[s->max_frames] : 0;
    if (s->h_status) s->h_status[s->max_frames] = 0;
    s->h_status_n = 0;
    if (dense_over)
        return fail(ctx, SLAMCU_CAPACITY, "dense output overflow (%s): raise kp_capacity / match_capacity",
                    dense_over == 1 ? "keypoints" : dense_over == 2 ? "matches" : "keypoints and matches");
    return SLAMCU_OK;
}

static int ctx_pipeline_resources(slamcu_context* ctx, size_t n_events) {
    if (!ctx->s_in) CU(ctx, cudaStreamCreateWithFlags(&ctx->s_in, cudaStreamNonBlocking));
    if (!ctx->s_out) CU(ctx, cudaStreamCreateWithFlags(&ctx->s_out, cudaStreamNonBlocking));
    if (!ctx->s_dl) CU(ctx, cudaStreamCreateWithFlags(&ctx->s_dl, cudaStreamNonBlocking));
    while (ctx->events.size() < n_events) {
        cudaEvent_t e;
        CU(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        ctx->events.push_back(e);
    }
    return SLAMCU_OK;
}

// The intended SLAMModel loop of the reference (include/slam/model/model.hpp:20-27: Preprocessor -> FeatureDetector
// -> FeatureMatcher), for n frames held in host memory: upload, detectAndCompute, match(f, f+1), download --
// software-pipelined by chunks of `chunk` frames over three streams (H2D copy engine, compute, D2H copy engine) so
// the PCIe transfers hide behind the kernels.
static int seq_process(slamcu_sequence* s, slamcu_detector* det, slamcu_matcher* m, const uint8_t* host_frames,
                       int stride, int n, int chunk, int with_keypoints, slamcu_keypoint* kps, uint8_t* desc,
                       slamcu_dmatch* matches, int32_t* counts4, bool dense, long long kp_capacity, long long match_capacity) {
    if (!s || !det || !m || !host_frames || s->ctx != det->ctx || s->ctx != m->ctx) return bad_args((s ? s->ctx : nullptr), __func__);
    slamcu_context* ctx = s->ctx;
    if (n < 0 || n > s->max_frames || stride < s->v.cols) return fail(ctx, SLAMCU_INVALID_ARGUMENT, "bad frame count / stride");
    if (m->distance_type != SLAMCU_DISTANCE_HAMMING)
        return fail(ctx, SLAMCU_UNSUPPORTED, "L2 distance requires float descriptors. Use the float overload.");
    if (n == 0) return SLAMCU_OK;
    if (chunk <= 0) chunk = 64;
    CU(ctx, cudaSetDevice(ctx->device));
    std::vector<int> sizes;
    for (int left = n; left > 0; left -= chunk) sizes.push_back(std::min(chunk, left));
    const int n_chunks = (int)sizes.size();
    int rc = SLAMCU_OK;
    const SeqView& v = s->v;
    cudaStream_t cs = ctx->stream;
    if (!s->h_status) {
        if (cudaMallocHost(reinterpret_cast<void**>(&s->h_status), ((size_t)s->max_frames + 1) * sizeof(int)) != cudaSuccess)
            return fail(ctx, SLAMCU_CUDA_ERROR, "cudaMallocHost failed");
        s->h_status[s->max_frames] = 0;
    }
    // dense outputs: compacted into device buffers of the caller's capacities, running offsets on the device
    void *dk = nullptr, *dd = nullptr, *dm = nullptr;
    int *kp_off = nullptr, *m_off = nullptr, *d_over = nullptr;
    if (dense) {
        if (kp_capacity < 0 || match_capacity < 0 || kp_capacity > INT_MAX || match_capacity > INT_MAX)
            return fail(ctx, SLAMCU_INVALID_ARGUMENT, "dense capacities out of range");
        void* probe = nullptr;  // asynchronous exact-size downloads need page-locked destinations
        if ((kps && cudaHostGetDevicePointer(&probe, kps, 0) != cudaSuccess) || (desc && cudaHostGetDevicePointer(&probe, desc, 0) != cudaSuccess) ||
            (matches && cudaHostGetDevicePointer(&probe, matches, 0) != cudaSuccess)) {
            cudaGetLastError();
            return fail(ctx, SLAMCU_INVALID_ARGUMENT, "dense outputs must be page-locked, device-mapped host memory (slamcu_alloc_pinned)");
        }
        rc = seq_dense_flush(s);  // an earlier dense call on this sequence that nobody waited for
        if (rc != SLAMCU_OK) return rc;
        if (kp_capacity > s->dd_kp_cap || match_capacity > s->dd_m_cap) {
            CU(ctx, cudaDeviceSynchronize());
            if (s->dd_kps) cudaFree(s->dd_kps);
            if (s->dd_desc) cudaFree(s->dd_desc);
            if (s->dd_matches) cudaFree(s->dd_matches);
            s->dd_kps = s->dd_desc = s->dd_matches = nullptr;
            s->dd_kp_cap = s->dd_m_cap = 0;
            const long long kc = std::max(kp_capacity, s->dd_kp_cap), mc = std::max(match_capacity, s->dd_m_cap);
            CU(ctx, cudaMalloc(&s->dd_kps, (size_t)std::max(kc, 1LL) * sizeof(slamcu_keypoint)));
            CU(ctx, cudaMalloc(&s->dd_desc, (size_t)std::max(kc, 1LL) * v.desc_words * 4));
            CU(ctx, cudaMalloc(&s->dd_matches, (size_t)std::max(mc, 1LL) * sizeof(slamcu_dmatch)));
            s->dd_kp_cap = kc;
            s->dd_m_cap = mc;
        }
        if (!s->h_tot && cudaMallocHost(reinterpret_cast<void**>(&s->h_tot), 2 * sizeof(int)) != cudaSuccess)
            return fail(ctx, SLAMCU_CUDA_ERROR, "cudaMallocHost failed");
        dk = kps ? s->dd_kps : nullptr;
        dd = desc ? s->dd_desc : nullptr;
        dm = matches ? s->dd_matches : nullptr;
        const size_t F1 = (size_t)s->max_frames + 1;
        if (!s->d_dense) CU(ctx, cudaMalloc(reinterpret_cast<void**>(&s->d_dense), (2 * F1 + 1) * sizeof(int)));
        kp_off = s->d_dense;
        m_off = s->d_dense + F1;
        d_over = s->d_dense + 2 * F1;
    }
    // hazards are per sequence: the new frames may overwrite this sequence's stores only after ITS previous kernels,
    // and its kernels may overwrite the result arrays only after ITS previous downloads.  A call on another sequence
    // (double buffering) therefore overlaps its copies with this one's kernels.
    CU(ctx, cudaStreamWaitEvent(ctx->s_in, s->ev_compute_done, 0));
    CU(ctx, cudaStreamWaitEvent(cs, s->ev_out_done, 0));
    if (dense) {  // the previous call's compaction (on s_out) is covered by ev_out_done, which s_out itself orders
        CU(ctx, cudaMemsetAsync(kp_off, 0, sizeof(int), ctx->s_out));
        CU(ctx, cudaMemsetAsync(m_off, 0, sizeof(int), ctx->s_out));
        CU(ctx, cudaMemsetAsync(d_over, 0, sizeof(int), ctx->s_out));
    }
    rc = ctx_lanes(ctx);
    if (rc != SLAMCU_OK) return rc;
    const bool lanes = ctx->two_lanes && !ctx->profiling && n_chunks > 1;
    if (lanes) {  // lane 1 inherits lane 0's ordering (previous downloads of this sequence, earlier work on the context)
        CU(ctx, cudaEventRecord(ctx->ev_lane, cs));
        CU(ctx, cudaStreamWaitEvent(ctx->lane1, ctx->ev_lane, 0));
    }
    std::vector<cudaEvent_t> ev_ext((size_t)n_chunks);
    rc = ctx_pipeline_resources(ctx, (size_t)3 * n_chunks + 2);
    if (rc != SLAMCU_OK) return rc;
    int f0 = 0;
    for (int c = 0; c < n_chunks; c++) {
        const int cnt = sizes[c];
        bool rp = false;
        rc = seq_upload_async(s, f0, cnt, host_frames + (size_t)f0 * v.rows * stride, stride, ctx->s_in, &rp);
        if (rc != SLAMCU_OK) return rc;
        CU(ctx, cudaEventRecord(ctx->events[2 * c], ctx->s_in));
        LaneScope lane(ctx, lanes ? (c & 1) : 0);  // consecutive chunks alternate between the two compute lanes
        cs = ctx->stream;
        CU(ctx, cudaStreamWaitEvent(cs, ctx->events[2 * c], 0));
        if (rp) {
            ProfGuard pg(ctx);
            seq_repitch(s, f0, cnt, cs);
        }
        rc = slamcu_sequence_extract(s, det, f0, cnt);
        if (rc != SLAMCU_OK) return rc;
        ev_ext[(size_t)c] = ctx->events[(size_t)2 * n_chunks + c];
        CU(ctx, cudaEventRecord(ev_ext[(size_t)c], cs));
        const int p0 = std::max(f0 - 1, 0), np = f0 + cnt - 1 - p0;  // pairs (p, p+1) completed by this chunk
        if (np > 0) {
            if (c > 0) CU(ctx, cudaStreamWaitEvent(cs, ev_ext[(size_t)c - 1], 0));  // frame f0 - 1 was extracted on the other lane
            rc = slamcu_sequence_match(s, m, p0, np, with_keypoints);
            if (rc != SLAMCU_OK) return rc;
        }
        CU(ctx, cudaEventRecord(ctx->events[2 * c + 1], cs));
        CU(ctx, cudaStreamWaitEvent(ctx->s_out, ctx->events[2 * c + 1], 0));
        if (dense) {
            ProfGuard pg(ctx);
            ctx->launches += launch_dense_scan(v.n_kp, f0, cnt, kp_off, ctx->s_out);
            ctx->launches += launch_dense_scan(v.n_match, p0, np, m_off, ctx->s_out);
            ctx->launches += launch_dense_copy(v, f0, cnt, p0, np, kp_off, m_off, dk, dd, dm, (int)kp_capacity, (int)match_capacity, d_over, ctx->s_out);
            f0 += cnt;
            continue;
        }
        if (kps)
            CU(ctx, cudaMemcpyAsync(kps + (size_t)f0 * v.cap_kp, v.kps + (size_t)f0 * v.cap_kp,
                                    (size_t)cnt * v.cap_kp * sizeof(slamcu_keypoint), cudaMemcpyDeviceToHost, ctx->s_out));
        if (desc) {
            rc = seq_download_desc(s, f0, cnt, desc + (size_t)f0 * v.cap_kp * v.desc_bytes, ctx->s_out);
            if (rc != SLAMCU_OK) return rc;
        }
        if (matches && np > 0)
            CU(ctx, cudaMemcpyAsync(matches + (size_t)p0 * v.cap_kp, v.matches + (size_t)p0 * v.cap_kp,
                                    (size_t)np * v.cap_kp * sizeof(slamcu_dmatch), cudaMemcpyDeviceToHost, ctx->s_out));
        f0 += cnt;
    }
    if (counts4) {
        CU(ctx, cudaMemcpy2DAsync(counts4 + 0, 16, v.n_kp, 4, 4, n, cudaMemcpyDeviceToHost, ctx->s_out));
        CU(ctx, cudaMemcpy2DAsync(counts4 + 1, 16, v.n_match, 4, 4, n, cudaMemcpyDeviceToHost, ctx->s_out));
        CU(ctx, cudaMemcpy2DAsync(counts4 + 2, 16, v.n_raw, 4, 4, n, cudaMemcpyDeviceToHost, ctx->s_out));
        CU(ctx, cudaMemcpy2DAsync(counts4 + 3, 16, v.status, 4, 4, n, cudaMemcpyDeviceToHost, ctx->s_out));
    }
    cs = ctx->stream;  // lane 0 again (every LaneScope is gone)
    if (lanes) {
        CU(ctx, cudaEventRecord(ctx->ev_lane, ctx->lane1));
        CU(ctx, cudaStreamWaitEvent(cs, ctx->ev_lane, 0));
    }
    if (dense) {
        CU(ctx, cudaMemcpyAsync(s->h_status + s->max_frames, d_over, sizeof(int), cudaMemcpyDeviceToHost, ctx->s_out));
        CU(ctx, cudaMemcpyAsync(s->h_tot + 0, kp_off + n, sizeof(int), cudaMemcpyDeviceToHost, ctx->s_out));
        CU(ctx, cudaMemcpyAsync(s->h_tot + 1, m_off + std::max(n - 1, 0), sizeof(int), cudaMemcpyDeviceToHost, ctx->s_out));
        s->dl_kps = kps;
        s->dl_desc = desc;
        s->dl_matches = matches;
        s->dl_active = true;
        ctx->pending_dl.push_back(s);
    }
    CU(ctx, cudaMemcpyAsync(s->h_status, v.status, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, ctx->s_out));
    s->h_status_n = n;
    CU(ctx, cudaEventRecord(s->ev_compute_done, cs));
    CU(ctx, cudaEventRecord(s->ev_out_done, ctx->s_out));  // slamcu_sequence_wait() / slamcu_synchronize() cover it
    return SLAMCU_OK;
}

int slamcu_sequence_process(slamcu_sequence* s, slamcu_detector* det, slamcu_matcher* m, const uint8_t* host_frames,
                            int stride, int n, int chunk, int with_keypoints, slamcu_keypoint* kps, uint8_t* desc,
                            slamcu_dmatch* matches, int32_t* counts4) {
    return seq_process(s, det, m, host_frames, stride, n, chunk, with_keypoints, kps, desc, matches, counts4, false, 0, 0);
}

int slamcu_sequence_process_dense(slamcu_sequence* s, slamcu_detector* det, slamcu_matcher* m, const uint8_t* host_frames,
                                  int stride, int n, int chunk, int with_keypoints, slamcu_keypoint* kps, uint8_t* desc,
                                  slamcu_dmatch* matches, int32_t* counts4, int64_t kp_capacity, int64_t match_capacity) {
    return seq_process(s, det, m, host_frames, stride, n, chunk, with_keypoints, kps, desc, matches, counts4, true, kp_capacity, match_capacity);
}

int slamcu_sequence_counts_device(slamcu_sequence* s, int first, int n, int32_t* device_counts4) {
    if (!s || !device_counts4) return bad_args((s ? s->ctx : nullptr), __func__);
    slamcu_context* ctx = s->ctx;
    if (first < 0 || n < 0 || first + n > s->max_frames) return fail(ctx, SLAMCU_INVALID_ARGUMENT, "bad frame range");
    ProfGuard pg(ctx);
    ctx->launches += launch_pack_counts(s->v, first, n, device_counts4, ctx->stream);
    return check_launch(ctx, "pack_counts");
}

/* ---------------------------------------------------------------------------------------------- */
/* detector                                                                                        */
/* ---------------------------------------------------------------------------------------------- */
// Constructor-time host tables.  They are functions of the *host* C++ library the reference is built
// against -- libstdc++'s minstd_rand0 + Marsaglia-polar normal_distribution<float>, libm's exp -- so
// they are produced on the host with the very same calls (feature_detector.cpp:286-313, :321-335).
int slamcu_default_brief_pattern(int patch_size, int num_pairs, int32_t* pattern4, int capacity_pairs, int* n_out) {
    if (!pattern4 || !n_out || patch_size <= 0 || num_pairs <= 0) return SLAMCU_INVALID_ARGUMENT;
    const float scale = static_cast<float>(patch_size) / 2.0F;
    std::default_random_engine gen;
    std::normal_distribution<float> dist(0.0F, 1.0F);
    int n = 0;
    for (int i = 0; i < num_pairs; i++) {
        const float x1 = dist(gen) * scale;
        const float y1 = dist(gen) * scale;
        const float x2 = dist(gen) * scale;
        const float y2 = dist(gen) * scale;
        if (std::abs(x1) < scale && std::abs(y1) < scale && std::abs(x2) < scale && std::abs(y2) < scale) {
            if (n >= capacity_pairs) return SLAMCU_CAPACITY;
            pattern4[4 * n + 0] = static_cast<int>(x1);
            pattern4[4 * n + 1] = static_cast<int>(y1);
            pattern4[4 * n + 2] = static_cast<int>(x2);
            pattern4[4 * n + 3] = static_cast<int>(y2);
            n++;
        }
    }
    *n_out = n;
    return SLAMCU_OK;
}

int slamcu_default_blur_weights(double* weights25) {
    if (!weights25) return SLAMCU_INVALID_ARGUMENT;
    const double sigma = 1.0;
    double sum = 0.0;
    for (int i = -2; i <= 2; i++)
        for (int j = -2; j <= 2; j++) {
            const double v = std::exp(-((i * i) + (j * j)) / (2 * sigma * sigma));
            weights25[(i + 2) * 5 + (j + 2)] = v;
            sum += v;
        }
    for (int k = 0; k < 25; k++) weights25[k] /= sum;
    return SLAMCU_OK;
}

int slamcu_detector_create(slamcu_context* ctx, const slamcu_detector_config* cfg, slamcu_detector** out) {
    if (!ctx || !cfg || !out) return bad_args(ctx, __func__);
    *out = nullptr;
    // same range checks (and messages) as the reference constructor, feature_detector.hpp:60-93
    if (cfg->intensity_threshold < 0 || cfg->intensity_threshold > 255)
        return fail(ctx, SLAMCU_INVALID_ARGUMENT, "Intensity threshold must be in the range [0, 255].");
    if (cfg->contiguous_pixels_threshold < 0 || cfg->contiguous_pixels_threshold > 16)
        return fail(ctx, SLAMCU_INVALID_ARGUMENT, "Contiguous pixels threshold must be in the range [0, 16].");
    if (cfg->non_max_suppression != 0 && cfg->non_max_suppression != 1)
        return fail(ctx, SLAMCU_INVALID_ARGUMENT, "Non-max suppression must be either 0 (false) or 1 (true).");
    if (cfg->suppression_window_size <= 0)
        return fail(ctx, SLAMCU_INVALID_ARGUMENT, "Suppression window size must be a positive integer.");
    if (cfg->patch_size <= 0 || cfg->patch_size % 2 == 0)
        return fail(ctx, SLAMCU_INVALID_ARGUMENT, "Patch size must be a positive odd integer.");
    if (cfg->num_brief_pairs <= 0 || cfg->num_brief_pairs % 8 != 0)
        return fail(ctx, SLAMCU_INVALID_ARGUMENT, "Number of BRIEF pairs must be a positive multiple of 8.");
    if (cfg->num_brief_pairs > 2048) return fail(ctx, SLAMCU_UNSUPPORTED, "NumBRIEFPairs > 2048 is not supported");
    if (cfg->suppression_window_size > 4096) return fail(ctx, SLAMCU_UNSUPPORTED, "SuppressionWindowSize > 4096");
    if (cfg->n_pattern < 0 || cfg->n_pattern > cfg->num_brief_pairs || (cfg->n_pattern > 0 && !cfg->pattern))
        return fail(ctx, SLAMCU_INVALID_ARGUMENT, "BRIEF pattern missing or longer than NumBRIEFPairs");
    if (!cfg->blur_weights) return fail(ctx, SLAMCU_INVALID_ARGUMENT, "blur weights missing");
    if (cfg->mode != SLAMCU_MODE_REFERENCE && cfg->mode != SLAMCU_MODE_ORB)
        return fail(ctx, SLAMCU_UNSUPPORTED, "detector mode %d unknown", cfg->mode);
    if (cfg->mode == SLAMCU_MODE_ORB) {
        if (cfg->n_levels < 1 || cfg->n_levels > kMaxLevels) return fail(ctx, SLAMCU_INVALID_ARGUMENT, "NumLevels must be in [1, %d]", kMaxLevels);
        if (!(cfg->scale_factor > 1.0f)) return fail(ctx, SLAMCU_INVALID_ARGUMENT, "ScaleFactor must be > 1");
        if (cfg->max_features <= 0) return fail(ctx, SLAMCU_INVALID_ARGUMENT, "MaxFeatures must be positive");
        if (cfg->num_brief_pairs != 256) return fail(ctx, SLAMCU_INVALID_ARGUMENT, "ORB mode descriptors are 256 bits");
    }
    CU(ctx, cudaSetDevice(ctx->device));
    slamcu_detector* d = new (std::nothrow) slamcu_detector();
    if (!d) return SLAMCU_CUDA_ERROR;
    d->ctx = ctx;
    d->mode = cfg->mode;
    d->p.thr = cfg->intensity_threshold;
    d->p.arc = cfg->contiguous_pixels_threshold;
    d->p.nms = cfg->non_max_suppression;
    d->p.window = cfg->suppression_window_size;
    d->p.patch = cfg->patch_size;
    d->p.pairs = cfg->num_brief_pairs;
    d->p.n_pattern = cfg->n_pattern;
    for (int i = 0; i < 25; i++) d->p.blur_w[i] = cfg->blur_weights[i];
    if (cfg->mode == SLAMCU_MODE_ORB) {
        d->n_levels = cfg->n_levels;
        d->scale_factor = cfg->scale_factor;
        d->max_features = cfg->max_features;
        d->fast_threshold = cfg->fast_threshold > 0 ? cfg->fast_threshold : cfg->intensity_threshold;
        int8_t h[1024];  // NULL = OpenCV's own bit_pattern_31_, built into the library
        for (int i = 0; i < 1024; i++) h[i] = cfg->orb_pattern ? (int8_t)cfg->orb_pattern[i] : kOrbBitPattern31[i];
        float hf[1024];
        for (int i = 0; i < 1024; i++) hf[i] = (float)h[i];
        if (cudaMalloc(reinterpret_cast<void**>(&d->d_orb_pattern), 1024) != cudaSuccess ||
            cudaMemcpy(d->d_orb_pattern, h, 1024, cudaMemcpyHostToDevice) != cudaSuccess ||
            cudaMalloc(reinterpret_cast<void**>(&d->d_orb_patf), sizeof hf) != cudaSuccess ||
            cudaMemcpy(d->d_orb_patf, hf, sizeof hf, cudaMemcpyHostToDevice) != cudaSuccess) {
            delete d;
            return fail(ctx, SLAMCU_CUDA_ERROR, "cudaMalloc(orb pattern) failed");
        }
    }
    const size_t pbytes = (size_t)std::max(cfg->n_pattern, 1) * 4 * sizeof(int);
    if (cudaMalloc(reinterpret_cast<void**>(&d->d_pattern), pbytes) != cudaSuccess) {
        delete d;
        return fail(ctx, SLAMCU_CUDA_ERROR, "cudaMalloc(pattern) failed");
    }
    if (cfg->n_pattern > 0)
        cudaMemcpyAsync(d->d_pattern, cfg->pattern, (size_t)cfg->n_pattern * 4 * sizeof(int), cudaMemcpyHostToDevice,
                        ctx->stream);
    cudaStreamSynchronize(ctx->stream);
    *out = d;
    return SLAMCU_OK;
}

void slamcu_detector_destroy(slamcu_detector* d) {
    if (!d) return;
    cudaSetDevice(d->ctx->device);
    if (d->one) slamcu_sequence_destroy(d->one);
    if (d->d_pattern) cudaFree(d->d_pattern);
    if (d->d_orb_pattern) cudaFree(d->d_orb_pattern);
    if (d->d_orb_patf) cudaFree(d->d_orb_patf);
    delete d;
}

// single-frame workspace, re-created when the image geometry (or needed capacity) changes
static int detector_workspace(slamcu_detector* d, int rows, int cols, int min_kp) {
    slamcu_context* ctx = d->ctx;
    const int desc_bytes = d->p.pairs / 8;
    if (d->one && d->one->v.rows == rows && d->one->v.cols == cols && d->one->v.cap_kp >= min_kp) return SLAMCU_OK;
    if (d->one) {
        slamcu_sequence_destroy(d->one);
        d->one = nullptr;
    }
    const long long px = (long long)rows * cols;
    // single calls favour safety over footprint: every pixel may be a raw corner (noise, low thresholds); after the
    // greedy NMS the survivors are pairwise >= window apart (densest packing 2 / (sqrt(3) w^2) per pixel, + margin)
    int cap_raw = (int)std::min<long long>(std::max<long long>(px, 8192), kMaxRawCap);
    const long long w2 = std::max(1LL, (long long)d->p.window * d->p.window);
    int cap_kp = d->p.nms ? (int)std::min<long long>(std::max<long long>(px * 8 / (5 * w2) + 4096, 4096), cap_raw) : cap_raw;
    if (d->mode == SLAMCU_MODE_ORB) { cap_raw = 8192; cap_kp = std::max(4096, 2 * d->max_features); }
    cap_kp = std::max(cap_kp, min_kp);
    return slamcu_sequence_create(ctx, rows, cols, 1, cap_raw, cap_kp, desc_bytes, &d->one);
}

static int check_image(slamcu_context* ctx, const uint8_t* image, int rows, int cols, int stride) {
    if (!image || rows <= 0 || cols <= 0 || stride < cols) return fail(ctx, SLAMCU_INVALID_ARGUMENT, "bad image");
    return SLAMCU_OK;
}

static int detect_common(slamcu_detector* d, const uint8_t* image, int rows, int cols, int stride, slamcu_keypoint* kps,
                         uint8_t* desc, int desc_stride, int capacity, int* n_out, int what /*0 detect,1 both,2 raw*/) {
    if (!d || !n_out) return bad_args((d ? d->ctx : nullptr), __func__);
    slamcu_context* ctx = d->ctx;
    int rc = check_image(ctx, image, rows, cols, stride);
    if (rc != SLAMCU_OK) return rc;
    CU(ctx, cudaSetDevice(ctx->device));
    rc = detector_workspace(d, rows, cols, 0);
    if (rc != SLAMCU_OK) return rc;
    slamcu_sequence* s = d->one;
    rc = slamcu_sequence_upload(s, 0, 1, image, stride);
    if (rc != SLAMCU_OK) return rc;
    if (d->mode == SLAMCU_MODE_ORB) {
        if (what == 2) return fail(ctx, SLAMCU_UNSUPPORTED, "raw-corner probe is a reference-mode stage");
        rc = seq_extract_orb(s, d, 0, 1);
        if (rc != SLAMCU_OK) return rc;
        return slamcu_sequence_frame(s, 0, kps, what == 1 ? desc : nullptr, desc_stride, capacity, n_out);
    }
    rc = seq_detect(s, d, 0, 1, what == 2);
    if (rc != SLAMCU_OK) return rc;
    if (what == 1) {
        rc = seq_compute(s, d, 0, 1);
        if (rc != SLAMCU_OK) return rc;
    }
    return slamcu_sequence_frame(s, 0, kps, what == 1 ? desc : nullptr, desc_stride, capacity, n_out);
}

int slamcu_detect(slamcu_detector* d, const uint8_t* image, int rows, int cols, int stride, slamcu_keypoint* kps,
                  int capacity, int* n_out) {
    return detect_common(d, image, rows, cols, stride, kps, nullptr, 0, capacity, n_out, 0);
}

int slamcu_fast_corners(slamcu_detector* d, const uint8_t* image, int rows, int cols, int stride, slamcu_keypoint* kps,
                        int capacity, int* n_out) {
    if (!d) return bad_args((d ? d->ctx : nullptr), __func__);
    // raw probe: the keypoint list must be able to hold every raw corner
    slamcu_context* ctx = d->ctx;
    int rc = check_image(ctx, image, rows, cols, stride);
    if (rc != SLAMCU_OK) return rc;
    const long long px = (long long)rows * cols;
    rc = detector_workspace(d, rows, cols, (int)std::min<long long>(std::max<long long>(px / 4, 8192), kMaxRawCap));
    if (rc != SLAMCU_OK) return rc;
    return detect_common(d, image, rows, cols, stride, kps, nullptr, 0, capacity, n_out, 2);
}

int slamcu_detect_and_compute(slamcu_detector* d, const uint8_t* image, int rows, int cols, int stride,
                              slamcu_keypoint* kps, uint8_t* desc, int desc_stride, int capacity, int* n_out) {
    if (!desc) return bad_args(d ? d->ctx : nullptr, __func__);
    return detect_common(d, image, rows, cols, stride, kps, desc, desc_stride, capacity, n_out, 1);
}

int slamcu_compute(slamcu_detector* d, const uint8_t* image, int rows, int cols, int stride, slamcu_keypoint* kps, int n,
                   uint8_t* desc, int desc_stride) {
    if (!d) return bad_args((d ? d->ctx : nullptr), __func__);
    slamcu_context* ctx = d->ctx;
    if (n < 0) return fail(ctx, SLAMCU_INVALID_ARGUMENT, "negative keypoint count");
    if (n == 0) return SLAMCU_OK;  // reference: descriptors = DescriptorMatrix(0, 0) (feature_detector.cpp:22-25)
    if (d->mode == SLAMCU_MODE_ORB)
        return fail(ctx, SLAMCU_UNSUPPORTED, "compute() on caller-supplied keypoints is a reference-mode call; ORB mode extracts with detectAndCompute()");
    if (!kps || !desc) return fail(ctx, SLAMCU_INVALID_ARGUMENT, "null keypoints / descriptors");
    int rc = check_image(ctx, image, rows, cols, stride);
    if (rc != SLAMCU_OK) return rc;
    CU(ctx, cudaSetDevice(ctx->device));
    rc = detector_workspace(d, rows, cols, n);
    if (rc != SLAMCU_OK) return rc;
    slamcu_sequence* s = d->one;
    rc = slamcu_sequence_upload(s, 0, 1, image, stride);
    if (rc != SLAMCU_OK) return rc;
    CU(ctx, cudaMemcpyAsync(s->v.kps, kps, (size_t)n * sizeof(slamcu_keypoint), cudaMemcpyHostToDevice, ctx->stream));
    CU(ctx, cudaMemcpyAsync(s->v.n_kp, &n, sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    CU(ctx, cudaMemsetAsync(s->v.status, 0, sizeof(int), ctx->stream));
    rc = seq_compute(s, d, 0, 1);
    if (rc != SLAMCU_OK) return rc;
    int got = 0;
    return slamcu_sequence_frame(s, 0, kps, desc, desc_stride, n, &got);
}

int slamcu_detector_last_octaves(slamcu_detector* d, int32_t* octaves, int capacity) {
    if (!d || !d->one) return bad_args((d ? d->ctx : nullptr), __func__);
    return slamcu_sequence_octaves(d->one, 0, octaves, capacity);
}

int slamcu_orb_stage(slamcu_detector* d, int stage, int level, uint32_t* xy, float* value, int capacity, int* n_out) {
    if (!d || !n_out) return bad_args((d ? d->ctx : nullptr), __func__);
    slamcu_context* ctx = d->ctx;
    if (d->mode != SLAMCU_MODE_ORB || !d->one || !d->one->has_orb)
        return fail(ctx, SLAMCU_UNSUPPORTED, "no ORB-mode single-frame call to probe");
    const OrbView& o = d->one->orb;
    if (stage < 0 || stage > 2 || level < 0 || level >= o.nlevels) return fail(ctx, SLAMCU_INVALID_ARGUMENT, "bad stage / level");
    CU(ctx, cudaSetDevice(ctx->device));
    const int* counts = stage == 0 ? o.n_cand : stage == 1 ? o.n_sel : o.n_fin;
    int n = 0;
    CU(ctx, cudaMemcpyAsync(&n, counts + level, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    *n_out = n;
    if (n > capacity) return fail(ctx, SLAMCU_CAPACITY, "need room for %d entries", n);
    if (n == 0) return SLAMCU_OK;
    const size_t off = o.lv[level].coff;
    const uint32_t* sxy = stage == 0 ? o.cxy : stage == 1 ? o.sxy : o.fxy;
    if (xy) CU(ctx, cudaMemcpyAsync(xy, sxy + off, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (value) {
        if (stage == 0) {
            std::vector<int> tmp((size_t)n);
            CU(ctx, cudaMemcpyAsync(tmp.data(), o.cscore + off, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream));
            CU(ctx, cudaStreamSynchronize(ctx->stream));
            for (int i = 0; i < n; i++) value[i] = (float)tmp[i];
        } else {
            CU(ctx, cudaMemcpyAsync(value, (stage == 1 ? o.sresp : o.fresp) + off, (size_t)n * 4, cudaMemcpyDeviceToHost,
                                    ctx->stream));
        }
    }
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return SLAMCU_OK;
}

int slamcu_orb_level_image(slamcu_detector* d, int level, int blurred, uint8_t* out, int out_stride, int* rows, int* cols) {
    if (!d) return bad_args((d ? d->ctx : nullptr), __func__);
    slamcu_context* ctx = d->ctx;
    if (d->mode != SLAMCU_MODE_ORB || !d->one || !d->one->has_orb)
        return fail(ctx, SLAMCU_UNSUPPORTED, "no ORB-mode single-frame call to probe");
    const OrbView& o = d->one->orb;
    const SeqView& v = d->one->v;
    if (level < 0 || level >= o.nlevels) return fail(ctx, SLAMCU_INVALID_ARGUMENT, "bad level");
    const OrbLevel& L = o.lv[level];
    if (rows) *rows = L.rows;
    if (cols) *cols = L.cols;
    if (!out) return SLAMCU_OK;
    if (out_stride < L.cols) return fail(ctx, SLAMCU_INVALID_ARGUMENT, "out_stride too small");
    CU(ctx, cudaSetDevice(ctx->device));
    const uint8_t* src = level == 0 ? (blurred ? v.blur : v.img) : (blurred ? o.pyrb : o.pyr) + L.off;
    CU(ctx, cudaMemcpy2DAsync(out, out_stride, src, L.pitch, L.cols, L.rows, cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return SLAMCU_OK;
}

int slamcu_gaussian_blur(slamcu_detector* d, const uint8_t* image, int rows, int cols, int stride, uint8_t* out,
                         int out_stride) {
    if (!d || !out) return bad_args((d ? d->ctx : nullptr), __func__);
    slamcu_context* ctx = d->ctx;
    int rc = check_image(ctx, image, rows, cols, stride);
    if (rc != SLAMCU_OK) return rc;
    if (out_stride < cols) return fail(ctx, SLAMCU_INVALID_ARGUMENT, "out_stride too small");
    CU(ctx, cudaSetDevice(ctx->device));
    rc = detector_workspace(d, rows, cols, 0);
    if (rc != SLAMCU_OK) return rc;
    slamcu_sequence* s = d->one;
    rc = slamcu_sequence_upload(s, 0, 1, image, stride);
    if (rc != SLAMCU_OK) return rc;
    ctx->launches += launch_blur(s->v, 0, 1, d->p, ctx->stream);
    rc = check_launch(ctx, "blur kernel");
    if (rc != SLAMCU_OK) return rc;
    CU(ctx, cudaMemcpy2DAsync(out, out_stride, s->v.blur, s->v.pitch, cols, rows, cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return SLAMCU_OK;
}

/* ---------------------------------------------------------------------------------------------- */
/* matcher                                                                                         */
/* ---------------------------------------------------------------------------------------------- */
int slamcu_matcher_create(slamcu_context* ctx, const slamcu_matcher_config* cfg, slamcu_matcher** out) {
    if (!ctx || !cfg || !out) return bad_args(ctx, __func__);
    *out = nullptr;
    // feature_matcher.cpp:25-57
    if (cfg->distance_type != SLAMCU_DISTANCE_HAMMING && cfg->distance_type != SLAMCU_DISTANCE_L2)
        return fail(ctx, SLAMCU_INVALID_ARGUMENT, "Invalid distance type. Must be 'HAMMING' or 'L2'.");
    if (cfg->filter_matches != 0 && cfg->filter_matches != 1)
        return fail(ctx, SLAMCU_INVALID_ARGUMENT, "FilterMatches must be either 0 (false) or 1 (true).");
    if (cfg->filter_matches && cfg->good_matches_count <= 0)
        return fail(ctx, SLAMCU_INVALID_ARGUMENT, "GoodMatchesCount must be positive when filtering is enabled.");
    if (cfg->use_ratio_test != 0 && cfg->use_ratio_test != 1)
        return fail(ctx, SLAMCU_INVALID_ARGUMENT, "UseRatioTest must be either 0 (false) or 1 (true).");
    if (!(cfg->ratio_test_threshold >= 0.0f && cfg->ratio_test_threshold <= 1.0f))
        return fail(ctx, SLAMCU_INVALID_ARGUMENT, "RatioTestThreshold must be in the range [0, 1].");
    slamcu_matcher* m = new (std::nothrow) slamcu_matcher();
    if (!m) return SLAMCU_CUDA_ERROR;
    m->ctx = ctx;
    m->distance_type = cfg->distance_type;
    m->p.filter = cfg->filter_matches;
    m->p.good = cfg->good_matches_count;
    m->p.use_ratio = cfg->use_ratio_test;
    m->p.ratio = cfg->ratio_test_threshold;
    *out = m;
    return SLAMCU_OK;
}

int slamcu_matcher_set_train_slices(slamcu_matcher* m, int n_slices) {
    if (!m) return SLAMCU_INVALID_ARGUMENT;
    if (n_slices < 0 || n_slices > kMaxMatchSlices) return fail(m->ctx, SLAMCU_INVALID_ARGUMENT, "n_slices must be in [0, %d]", kMaxMatchSlices);
    m->forced_slices = n_slices;
    return SLAMCU_OK;
}

static void matcher_free(slamcu_matcher* m) {
    void* ptrs[] = {m->d1, m->d2, m->k1, m->k2, m->cand, m->matches, m->keys, m->ors, m->counts, m->x1, m->x2, m->ck1, m->ck2};
    for (void* p : ptrs)
        if (p) cudaFree(p);
    m->x1 = m->x2 = nullptr;
    m->ck1 = m->ck2 = nullptr;
    m->tc_rows1 = m->tc_rows2 = 0;
    m->d1 = m->d2 = nullptr;
    m->k1 = m->k2 = nullptr;
    m->cand = nullptr;
    m->matches = nullptr;
    m->keys = nullptr;
    m->ors = nullptr;
    m->counts = nullptr;
    m->cap1 = m->cap2 = m->cap_words = 0;
}

void slamcu_matcher_destroy(slamcu_matcher* m) {
    if (!m) return;
    cudaSetDevice(m->ctx->device);
    cudaStreamSynchronize(m->ctx->stream);
    matcher_free(m);
    delete m;
}

static int matcher_workspace(slamcu_matcher* m, int n1, int n2, int words) {
    slamcu_context* ctx = m->ctx;
    if (n1 <= m->cap1 && n2 <= m->cap2 && words <= m->cap_words) return SLAMCU_OK;
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    const int c1 = std::max(n1, m->cap1), c2 = std::max(n2, m->cap2), w = std::max(words, m->cap_words);
    matcher_free(m);
    CU(ctx, cudaMalloc(reinterpret_cast<void**>(&m->d1), (size_t)c1 * w * 4));
    CU(ctx, cudaMalloc(reinterpret_cast<void**>(&m->d2), (size_t)c2 * w * 4));
    CU(ctx, cudaMalloc(reinterpret_cast<void**>(&m->k1), (size_t)c1 * sizeof(slamcu_keypoint)));
    CU(ctx, cudaMalloc(reinterpret_cast<void**>(&m->k2), (size_t)c2 * sizeof(slamcu_keypoint)));
    CU(ctx, cudaMalloc(reinterpret_cast<void**>(&m->cand), (size_t)c1 * kMaxMatchSlices * sizeof(int4)));
    CU(ctx, cudaMalloc(reinterpret_cast<void**>(&m->matches), (size_t)c1 * sizeof(slamcu_dmatch)));
    CU(ctx, cudaMalloc(reinterpret_cast<void**>(&m->keys), (size_t)c1 * sizeof(unsigned long long)));
    CU(ctx, cudaMalloc(reinterpret_cast<void**>(&m->ors), (size_t)2 * w * 4));
    CU(ctx, cudaMalloc(reinterpret_cast<void**>(&m->counts), 4 * sizeof(int)));
    m->cap1 = c1;
    m->cap2 = c2;
    m->cap_words = w;
    return SLAMCU_OK;
}

static int match_common(slamcu_matcher* m, const uint8_t* d1, int n1, int width1, const uint8_t* d2, int n2, int width2,
                        const slamcu_keypoint* kp1, int nkp1, const slamcu_keypoint* kp2, int nkp2, bool emit) {
    slamcu_context* ctx = m->ctx;
    // validateInputs, feature_matcher.cpp:97-111 (same order of checks)
    if (n1 <= 0 || n2 <= 0 || !d1 || !d2) return fail(ctx, SLAMCU_EMPTY_INPUT, "Empty descriptors provided.");
    if (m->distance_type != SLAMCU_DISTANCE_HAMMING)
        return fail(ctx, SLAMCU_UNSUPPORTED, "DescriptorMatrix (uint8_t) requires HAMMING distance.");
    if (width1 != width2) return fail(ctx, SLAMCU_SIZE_MISMATCH, "Descriptor dimensions must match.");
    if (width1 <= 0 || width1 > 256) return fail(ctx, SLAMCU_INVALID_ARGUMENT, "descriptor width %d unsupported", width1);
    const bool with_kp = kp1 && kp2 && nkp1 > 0 && nkp2 > 0;  // feature_matcher.cpp:150
    if (with_kp && (nkp1 < n1 || nkp2 < n2))
        return fail(ctx, SLAMCU_SIZE_MISMATCH, "fewer keypoints than descriptors");
    CU(ctx, cudaSetDevice(ctx->device));
    ProfGuard pg(ctx);
    const int words = (width1 + 3) / 4;
    int rc = matcher_workspace(m, n1, n2, words);
    if (rc != SLAMCU_OK) return rc;
    if (width1 & 3) {
        CU(ctx, cudaMemsetAsync(m->d1, 0, (size_t)n1 * words * 4, ctx->stream));
        CU(ctx, cudaMemsetAsync(m->d2, 0, (size_t)n2 * words * 4, ctx->stream));
    }
    CU(ctx, cudaMemcpy2DAsync(m->d1, words * 4, d1, width1, width1, n1, cudaMemcpyHostToDevice, ctx->stream));
    CU(ctx, cudaMemcpy2DAsync(m->d2, words * 4, d2, width2, width2, n2, cudaMemcpyHostToDevice, ctx->stream));
    if (with_kp) {
        CU(ctx, cudaMemcpyAsync(m->k1, kp1, (size_t)n1 * sizeof(slamcu_keypoint), cudaMemcpyHostToDevice, ctx->stream));
        CU(ctx, cudaMemcpyAsync(m->k2, kp2, (size_t)n2 * sizeof(slamcu_keypoint), cudaMemcpyHostToDevice, ctx->stream));
    }
    const int h_counts[4] = {n1, n2, 0, 0};
    CU(ctx, cudaMemcpyAsync(m->counts, h_counts, sizeof h_counts, cudaMemcpyHostToDevice, ctx->stream));
    ctx->launches += launch_desc_or(m->d1, m->counts + 0, words, m->ors, ctx->stream);
    ctx->launches += launch_desc_or(m->d2, m->counts + 1, words, m->ors + words, ctx->stream);
    MatchJob j{};
    j.dq = m->d1;
    j.dt = m->d2;
    j.kq = m->k1;
    j.kt = m->k2;
    j.nq = m->counts + 0;
    j.nt = m->counts + 1;
    j.orq = m->ors;
    j.ort = m->ors + words;
    j.cand = m->cand;
    j.matches = m->matches;
    j.n_match = m->counts + 2;
    j.status = m->counts + 3;
    j.desc_words = words;
    j.max_q = n1;
    j.max_t = n2;
    j.cap_out = n1;
    // a single problem: slice the train set so that the grid fills the device (>= ~4 blocks per SM), 128-descriptor
    // tiles at least; the batched sequence path has thousands of blocks and does not need it
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device);
    // blocks hold 256 queries (128 for small problems, launch_match decides): aim at ~8 blocks per SM, slices of at least 128
    const MatchTc* tc = nullptr;
    if (!with_kp && words == 8 && match_tc_enabled()) {
        const int r1 = (n1 + kTcRowAlign - 1) / kTcRowAlign * kTcRowAlign, r2 = (n2 + kTcRowAlign - 1) / kTcRowAlign * kTcRowAlign;
        bool ok = true;
        if (r1 > m->tc_rows1 || r2 > m->tc_rows2) {
            CU(ctx, cudaStreamSynchronize(ctx->stream));
            const int c1 = std::max(r1, m->tc_rows1), c2 = std::max(r2, m->tc_rows2);
            void* old[] = {m->x1, m->x2, m->ck1, m->ck2};
            for (void* q : old)
                if (q) cudaFree(q);
            m->x1 = m->x2 = nullptr;
            m->ck1 = m->ck2 = nullptr;
            m->tc_rows1 = m->tc_rows2 = 0;
            ok = cudaMalloc(reinterpret_cast<void**>(&m->x1), (size_t)c1 * kTcRowBytes) == cudaSuccess &&
                 cudaMalloc(reinterpret_cast<void**>(&m->x2), (size_t)c2 * kTcRowBytes) == cudaSuccess &&
                 cudaMalloc(reinterpret_cast<void**>(&m->ck1), (size_t)c1 * sizeof(uint32_t)) == cudaSuccess &&
                 cudaMalloc(reinterpret_cast<void**>(&m->ck2), (size_t)c2 * sizeof(uint32_t)) == cudaSuccess &&
                 encode_tc_map(&m->tc.map_q, m->x1, (size_t)c1) && encode_tc_map(&m->tc.map_t, m->x2, (size_t)c2);
            if (ok) {
                m->tc_rows1 = c1;
                m->tc_rows2 = c2;
            } else {
                cudaGetLastError();
            }
        }
        if (ok) {
            ctx->launches += launch_expand_bits(m->d1, 0, m->counts + 0, 1, 1, m->tc_rows1, m->x1, m->ck1, ctx->stream);
            ctx->launches += launch_expand_bits(m->d2, 0, m->counts + 1, 1, 1, m->tc_rows2, m->x2, m->ck2, ctx->stream);
            m->tc.view = MatchTcView{m->ck1, m->ck2, 0, 0, 0, 0};
            tc = &m->tc;
        }
    }
    // integer-pipe blocks hold 256 queries, tensor-core CTAs 128
    const int q_blocks = tc ? (n1 + 127) / 128 : (n1 + 255) / 256;
    int n_seg = std::min(std::min(((tc ? 4 : 8) * sms + q_blocks - 1) / q_blocks, (n2 + 127) / 128), kMaxMatchSlices);
    if (m->forced_slices > 0) n_seg = std::min(std::min(m->forced_slices, (n2 + 127) / 128), kMaxMatchSlices);
    n_seg = std::max(n_seg, 1);
    const int seg_len = ((n2 + n_seg - 1) / n_seg + 127) / 128 * 128;
    n_seg = (n2 + seg_len - 1) / seg_len;
    ctx->launches += launch_match(j, 1, m->p, emit, with_kp ? 1 : 0, m->keys, ctx->stream, n_seg, seg_len, (size_t)m->cap1, tc);
    return check_launch(ctx, "match kernels");
}

int slamcu_match(slamcu_matcher* m, const uint8_t* d1, int n1, int width1, const uint8_t* d2, int n2, int width2,
                 const slamcu_keypoint* kp1, int nkp1, const slamcu_keypoint* kp2, int nkp2, slamcu_dmatch* matches,
                 int capacity, int* n_out) {
    if (!m || !n_out) return bad_args((m ? m->ctx : nullptr), __func__);
    slamcu_context* ctx = m->ctx;
    *n_out = 0;
    int rc = match_common(m, d1, n1, width1, d2, n2, width2, kp1, nkp1, kp2, nkp2, true);
    if (rc != SLAMCU_OK) return rc;
    int h[4];
    CU(ctx, cudaMemcpyAsync(h, m->counts, sizeof h, cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    *n_out = h[2];
    if (h[2] > capacity) return fail(ctx, SLAMCU_CAPACITY, "need room for %d matches, capacity %d", h[2], capacity);
    if (h[2] > 0) {
        if (!matches) return fail(ctx, SLAMCU_INVALID_ARGUMENT, "null output");
        CU(ctx, cudaMemcpyAsync(matches, m->matches, (size_t)h[2] * sizeof(slamcu_dmatch), cudaMemcpyDeviceToHost,
                                ctx->stream));
        CU(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return SLAMCU_OK;
}

int slamcu_knn2_hamming(slamcu_matcher* m, const uint8_t* d1, int n1, const uint8_t* d2, int n2, int width,
                        slamcu_knn2* out) {
    if (!m || !out) return bad_args((m ? m->ctx : nullptr), __func__);
    slamcu_context* ctx = m->ctx;
    int rc = match_common(m, d1, n1, width, d2, n2, width, nullptr, 0, nullptr, 0, false);
    if (rc != SLAMCU_OK) return rc;
    std::vector<int4> h((size_t)n1);
    CU(ctx, cudaMemcpyAsync(h.data(), m->cand, (size_t)n1 * sizeof(int4), cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < n1; i++) {
        out[i].trainIdx0 = h[i].x;
        out[i].distance0 = (float)h[i].y;
        out[i].trainIdx1 = h[i].w;
        out[i].distance1 = (h[i].w >= 0) ? (float)h[i].z : 0.0f;
    }
    return SLAMCU_OK;
}

/* ---------------------------------------------------------------------------------------------- */
/* image preparation                                                                               */
/* ---------------------------------------------------------------------------------------------- */
int slamcu_bgr_to_gray(slamcu_context* ctx, const uint8_t* bgr, int rows, int cols, int stride, uint8_t* gray,
                       int gray_stride) {
    if (!ctx) return bad_args(ctx, __func__);
    if (!bgr || !gray || rows <= 0 || cols <= 0 || stride < 3 * cols || gray_stride < cols)
        return fail(ctx, SLAMCU_INVALID_ARGUMENT, "bad bgr image");
    CU(ctx, cudaSetDevice(ctx->device));
    const size_t in_b = (size_t)rows * cols * 3, out_b = (size_t)rows * cols;
    int rc = ensure_scratch(ctx, in_b + out_b);
    if (rc != SLAMCU_OK) return rc;
    uint8_t* d_in = static_cast<uint8_t*>(ctx->scratch);
    uint8_t* d_out = d_in + in_b;
    CU(ctx, cudaMemcpy2DAsync(d_in, (size_t)cols * 3, bgr, stride, (size_t)cols * 3, rows, cudaMemcpyHostToDevice, ctx->stream));
    ctx->launches += launch_bgr2gray(d_in, rows, cols, cols * 3, d_out, cols, ctx->stream);
    rc = check_launch(ctx, "bgr2gray");
    if (rc != SLAMCU_OK) return rc;
    CU(ctx, cudaMemcpy2DAsync(gray, gray_stride, d_out, cols, cols, rows, cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return SLAMCU_OK;
}

int slamcu_undistort(slamcu_context* ctx, const uint8_t* gray, int rows, int cols, int stride, const double* K4,
                     const double* D4, uint8_t* out_u8, double* out_f64) {
    if (!ctx) return bad_args(ctx, __func__);
    if (!gray || rows <= 0 || cols <= 0) return fail(ctx, SLAMCU_EMPTY_INPUT, "Input image is empty.");  // common.hpp:130-132
    if (stride < cols || !K4 || !D4 || (!out_u8 && !out_f64)) return fail(ctx, SLAMCU_INVALID_ARGUMENT, "bad arguments");
    CU(ctx, cudaSetDevice(ctx->device));
    const size_t px = (size_t)rows * cols;
    if ((long long)rows * cols > INT_MAX) return fail(ctx, SLAMCU_UNSUPPORTED, "undistortion map: rows * cols exceeds INT_MAX");
    const size_t off_map = (px + 255) / 256 * 256, off_u8 = off_map + px * 4, off_f64 = (off_u8 + px + 255) / 256 * 256;
    const size_t off_fix = off_f64 + px * 8;
    int rc = ensure_scratch(ctx, off_fix + (1 + (size_t)kUndistFixCap) * sizeof(int));
    if (rc != SLAMCU_OK) return rc;
    uint8_t* base = static_cast<uint8_t*>(ctx->scratch);
    uint8_t* d_in = base;
    int* d_map = reinterpret_cast<int*>(base + off_map);
    uint8_t* d_u8 = base + off_u8;
    double* d_f64 = reinterpret_cast<double*>(base + off_f64);
    CU(ctx, cudaMemcpy2DAsync(d_in, cols, gray, stride, cols, rows, cudaMemcpyHostToDevice, ctx->stream));
    CamParams cam{K4[0], K4[1], K4[2], K4[3], D4[0], D4[1], D4[2], D4[3]};
    rc = build_undistort_map(ctx, rows, cols, cam, d_map, reinterpret_cast<int*>(base + off_fix));
    if (rc != SLAMCU_OK) return rc;
    ctx->launches += launch_remap(d_in, rows, cols, cols, d_map, out_u8 ? d_u8 : nullptr, out_f64 ? d_f64 : nullptr,
                                  ctx->stream);
    rc = check_launch(ctx, "undistort");
    if (rc != SLAMCU_OK) return rc;
    if (out_u8) CU(ctx, cudaMemcpyAsync(out_u8, d_u8, px, cudaMemcpyDeviceToHost, ctx->stream));
    if (out_f64) CU(ctx, cudaMemcpyAsync(out_f64, d_f64, px * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return SLAMCU_OK;
}

/* ---------------------------------------------------------------------------------------------- */
/* two-view geometry                                                                               */
/* ---------------------------------------------------------------------------------------------- */
int slamcu_ransac_score(slamcu_context* ctx, const double* models9, int n_models, const double* x1, const double* x2,
                        int n, double thr2, int32_t* counts, uint8_t* masks) {
    if (!ctx) return bad_args(ctx, __func__);
    if (!models9 || !x1 || !x2 || !counts || n_models <= 0 || n <= 0) return fail(ctx, SLAMCU_INVALID_ARGUMENT, "bad arguments");
    CU(ctx, cudaSetDevice(ctx->device));
    const size_t b_models = (size_t)n_models * 9 * 8, b_pts = (size_t)n * 2 * 8, b_cnt = ((size_t)n_models * 4 + 255) / 256 * 256;
    const size_t b_mask = masks ? (size_t)n_models * n : 0;
    int rc = ensure_scratch(ctx, b_models + 2 * b_pts + b_cnt + b_mask + 1024);
    if (rc != SLAMCU_OK) return rc;
    uint8_t* base = static_cast<uint8_t*>(ctx->scratch);
    double* d_models = reinterpret_cast<double*>(base);
    double* d_x1 = reinterpret_cast<double*>(base + b_models);
    double* d_x2 = reinterpret_cast<double*>(base + b_models + b_pts);
    int* d_cnt = reinterpret_cast<int*>(base + b_models + 2 * b_pts);
    uint8_t* d_mask = masks ? base + b_models + 2 * b_pts + b_cnt : nullptr;
    CU(ctx, cudaMemcpyAsync(d_models, models9, b_models, cudaMemcpyHostToDevice, ctx->stream));
    CU(ctx, cudaMemcpyAsync(d_x1, x1, b_pts, cudaMemcpyHostToDevice, ctx->stream));
    CU(ctx, cudaMemcpyAsync(d_x2, x2, b_pts, cudaMemcpyHostToDevice, ctx->stream));
    ctx->launches += launch_ransac_score(d_models, n_models, d_x1, d_x2, n, thr2, d_cnt, d_mask, ctx->stream);
    rc = check_launch(ctx, "ransac_score");
    if (rc != SLAMCU_OK) return rc;
    CU(ctx, cudaMemcpyAsync(counts, d_cnt, (size_t)n_models * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (masks) CU(ctx, cudaMemcpyAsync(masks, d_mask, b_mask, cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return SLAMCU_OK;
}

static void essential_params(EssentialJob& j, const double* K4, double prob, double threshold, int max_iters) {
    const double t = threshold / ((K4[0] + K4[1]) / 2.0);  // findEssentialMat: threshold /= (fx + fy) / 2
    j.thr2 = (float)(t * t);
    j.prob = prob;
    j.max_iters = max_iters;
}

int slamcu_find_essential(slamcu_context* ctx, const float* p1, const float* p2, int n, const double* K4, double prob,
                          double threshold, int max_iters, double* E9, uint8_t* mask, int* n_inliers) {
    if (!ctx) return bad_args(ctx, __func__);
    if (!p1 || !p2 || !K4 || !E9 || n < 0) return fail(ctx, SLAMCU_INVALID_ARGUMENT, "bad arguments");
    if (n < 6) return fail(ctx, SLAMCU_EMPTY_INPUT, "findEssentialMat needs more than 5 correspondences (got %d)", n);
    if (max_iters > kEssentialMaxIters) return fail(ctx, SLAMCU_INVALID_ARGUMENT, "maxIters above %d is not supported", kEssentialMaxIters);
    CU(ctx, cudaSetDevice(ctx->device));
    const size_t b_pts = ((size_t)n * 16 + 255) / 256 * 256, b_in = ((size_t)n * 8 + 255) / 256 * 256;
    const size_t b_mask = ((size_t)n + 255) / 256 * 256;
    int rc = ensure_scratch(ctx, 2 * b_pts + 2 * b_in + b_mask + 1024 + essential_work_bytes_per_pair(max_iters));
    if (rc != SLAMCU_OK) return rc;
    uint8_t* base = static_cast<uint8_t*>(ctx->scratch);
    EssentialJob j{};
    j.x1 = reinterpret_cast<double2*>(base);
    j.x2 = reinterpret_cast<double2*>(base + b_pts);
    j.work = base + 2 * b_pts + 2 * b_in + b_mask + 1024;
    float* d_p1 = reinterpret_cast<float*>(base + 2 * b_pts);
    float* d_p2 = reinterpret_cast<float*>(base + 2 * b_pts + b_in);
    j.mask = base + 2 * b_pts + 2 * b_in;
    uint8_t* tail = j.mask + b_mask;
    j.E = reinterpret_cast<double*>(tail);            // 72 bytes
    j.n_pts = reinterpret_cast<int*>(tail + 128);
    j.n_inliers = j.n_pts + 1;
    j.n_iters = j.n_pts + 2;
    j.pt_stride = n;
    essential_params(j, K4, prob, threshold, max_iters);
    CU(ctx, cudaMemcpyAsync(d_p1, p1, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
    CU(ctx, cudaMemcpyAsync(d_p2, p2, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
    {
        ProfGuard pg(ctx);
        ctx->launches += launch_essential_normalise(d_p1, d_p2, n, j, K4, ctx->stream);
        ctx->launches += launch_essential_ransac(j, 1, ctx->stream);
    }
    rc = check_launch(ctx, "essential kernels");
    if (rc != SLAMCU_OK) return rc;
    int h[3];
    CU(ctx, cudaMemcpyAsync(E9, j.E, 72, cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaMemcpyAsync(h, j.n_pts, sizeof h, cudaMemcpyDeviceToHost, ctx->stream));
    if (mask) CU(ctx, cudaMemcpyAsync(mask, j.mask, (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    if (n_inliers) *n_inliers = h[1];
    return SLAMCU_OK;
}

int slamcu_estimate_pose(slamcu_context* ctx, const float* p1, const float* p2, int n, const double* K4, double* E9, uint8_t* mask,
                         int* n_inliers, double* R9, double* t3, int32_t* front4) {
    if (!ctx) return bad_args(ctx, __func__);
    if (!p1 || !p2 || !K4 || !R9 || !t3 || n < 0) return fail(ctx, SLAMCU_INVALID_ARGUMENT, "bad arguments");
    if (n < 8) return fail(ctx, SLAMCU_EMPTY_INPUT, "Cannot estimate pose, not enough matches (%d). Required at least 8.", n);
    double E[9];
    int inl = 0;
    int rc = slamcu_find_essential(ctx, p1, p2, n, K4, 0.999, 1.0, 1000, E, mask, &inl);  // leaves x1/x2/E in the scratch block
    if (rc != SLAMCU_OK) return rc;
    if (E9) memcpy(E9, E, sizeof E);
    if (n_inliers) *n_inliers = inl;
    if (inl <= 0) return fail(ctx, SLAMCU_EMPTY_INPUT, "Essential Matrix could not be computed.");
    // rebuild the job view over the scratch block exactly as slamcu_find_essential laid it out
    const size_t b_pts = ((size_t)n * 16 + 255) / 256 * 256, b_in = ((size_t)n * 8 + 255) / 256 * 256;
    const size_t b_mask = ((size_t)n + 255) / 256 * 256;
    uint8_t* base = static_cast<uint8_t*>(ctx->scratch);
    EssentialJob j{};
    j.x1 = reinterpret_cast<double2*>(base);
    j.x2 = reinterpret_cast<double2*>(base + b_pts);
    j.mask = base + 2 * b_pts + 2 * b_in;
    uint8_t* tail = j.mask + b_mask;
    j.E = reinterpret_cast<double*>(tail);
    j.n_pts = reinterpret_cast<int*>(tail + 128);
    j.pt_stride = n;
    double* dR = reinterpret_cast<double*>(tail + 256);
    double* dt = dR + 9;
    int* dfront = reinterpret_cast<int*>(tail + 256 + 12 * 8);
    {
        ProfGuard pg(ctx);
        ctx->launches += launch_recover_pose(j, 1, K4, dR, dt, dfront, ctx->stream);
    }
    rc = check_launch(ctx, "recover_pose");
    if (rc != SLAMCU_OK) return rc;
    int hf[4];
    CU(ctx, cudaMemcpyAsync(R9, dR, 72, cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaMemcpyAsync(t3, dt, 24, cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaMemcpyAsync(hf, dfront, 16, cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    if (front4) memcpy(front4, hf, sizeof hf);
    return SLAMCU_OK;
}

// The sampling loop of LoopClosure::verifyGeometricConsistency (loop_closure.cpp:177-193) with the host C++ library's own
// std::mt19937 and std::uniform_int_distribution<int>(0, n - 1), seeded by the caller instead of std::random_device:
// six distinct indices per iteration, redrawn on duplicates.  (solvePnP never fails for six points, so no iteration is skipped.)
int slamcu_pnp_sample_indices(uint32_t seed, int n, int n_hypotheses, int32_t* samples6) {
    if (!samples6 || n < 6 || n_hypotheses < 0) return SLAMCU_INVALID_ARGUMENT;
    std::mt19937 rng(seed);
    std::uniform_int_distribution<int> dist(0, n - 1);
    for (int h = 0; h < n_hypotheses; h++) {
        int got = 0;
        while (got < 6) {
            const int idx = dist(rng);
            bool dup = false;
            for (int k = 0; k < got; k++) dup |= samples6[6 * h + k] == idx;
            if (!dup) samples6[6 * h + got++] = idx;
        }
    }
    return SLAMCU_OK;
}

int slamcu_triangulate(slamcu_context* ctx, const double* P1, const double* P2, const float* pts1, const float* pts2, int n, double* points4,
                       double* points3) {
    if (!ctx) return SLAMCU_INVALID_ARGUMENT;
    if (!P1 || !P2 || n < 0 || (n > 0 && (!pts1 || !pts2)) || (!points4 && !points3)) return fail(ctx, SLAMCU_INVALID_ARGUMENT, "bad arguments");
    if (n == 0) return SLAMCU_OK;
    CU(ctx, cudaSetDevice(ctx->device));
    const size_t b_in = ((size_t)n * 8 + 255) / 256 * 256, b4 = (size_t)n * 32, b3 = ((size_t)n * 24 + 255) / 256 * 256;
    int rc = ensure_scratch(ctx, 256 + 2 * b_in + b4 + b3);
    if (rc != SLAMCU_OK) return rc;
    uint8_t* base = static_cast<uint8_t*>(ctx->scratch);
    double* dP = reinterpret_cast<double*>(base);
    float* d1 = reinterpret_cast<float*>(base + 256);
    float* d2 = reinterpret_cast<float*>(base + 256 + b_in);
    double* d4 = reinterpret_cast<double*>(base + 256 + 2 * b_in);
    double* d3 = reinterpret_cast<double*>(base + 256 + 2 * b_in + b4);
    CU(ctx, cudaMemcpyAsync(dP, P1, 96, cudaMemcpyHostToDevice, ctx->stream));
    CU(ctx, cudaMemcpyAsync(dP + 12, P2, 96, cudaMemcpyHostToDevice, ctx->stream));
    CU(ctx, cudaMemcpyAsync(d1, pts1, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
    CU(ctx, cudaMemcpyAsync(d2, pts2, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
    {
        ProfGuard pg(ctx);
        ctx->launches += launch_triangulate(dP, d1, d2, n, d4, d3, ctx->stream);
    }
    rc = check_launch(ctx, "triangulate");
    if (rc != SLAMCU_OK) return rc;
    if (points4) CU(ctx, cudaMemcpyAsync(points4, d4, (size_t)n * 32, cudaMemcpyDeviceToHost, ctx->stream));
    if (points3) CU(ctx, cudaMemcpyAsync(points3, d3, (size_t)n * 24, cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return SLAMCU_OK;
}

int slamcu_pnp_ransac(slamcu_context* ctx, const double* points3d, const double* points2d, int n, const int32_t* samples6, int n_hypotheses,
                      const double* K9, double threshold, int32_t* counts2, double* Rt24) {
    if (!ctx) return SLAMCU_INVALID_ARGUMENT;
    if (!points3d || !points2d || !samples6 || !K9 || !counts2 || n < 6 || n_hypotheses <= 0)
        return fail(ctx, SLAMCU_INVALID_ARGUMENT, "pnp_ransac: need at least 6 correspondences and one hypothesis");
    for (long long i = 0; i < 6LL * n_hypotheses; i++)
        if (samples6[i] < 0 || samples6[i] >= n) return fail(ctx, SLAMCU_INVALID_ARGUMENT, "pnp_ransac: sample index %d out of range", samples6[i]);
    CU(ctx, cudaSetDevice(ctx->device));
    auto up = [](size_t b) { return (b + 255) / 256 * 256; };
    const size_t b3 = up((size_t)n * 24), b2 = up((size_t)n * 16), bs = up((size_t)n_hypotheses * 24), bc = up((size_t)n_hypotheses * 8),
                 br = (size_t)n_hypotheses * 2 * 96;
    int rc = ensure_scratch(ctx, 256 + b3 + b2 + bs + bc + br);
    if (rc != SLAMCU_OK) return rc;
    uint8_t* base = static_cast<uint8_t*>(ctx->scratch);
    double* dK = reinterpret_cast<double*>(base);
    double* d3 = reinterpret_cast<double*>(base + 256);
    double* d2 = reinterpret_cast<double*>(base + 256 + b3);
    int* ds = reinterpret_cast<int*>(base + 256 + b3 + b2);
    int* dc = reinterpret_cast<int*>(base + 256 + b3 + b2 + bs);
    double* dr = reinterpret_cast<double*>(base + 256 + b3 + b2 + bs + bc);
    CU(ctx, cudaMemcpyAsync(dK, K9, 72, cudaMemcpyHostToDevice, ctx->stream));
    CU(ctx, cudaMemcpyAsync(d3, points3d, (size_t)n * 24, cudaMemcpyHostToDevice, ctx->stream));
    CU(ctx, cudaMemcpyAsync(d2, points2d, (size_t)n * 16, cudaMemcpyHostToDevice, ctx->stream));
    CU(ctx, cudaMemcpyAsync(ds, samples6, (size_t)n_hypotheses * 24, cudaMemcpyHostToDevice, ctx->stream));
    {
        ProfGuard pg(ctx);
        ctx->launches += launch_pnp_ransac(d3, d2, n, ds, n_hypotheses, dK, threshold, dc, dr, ctx->stream);
    }
    rc = check_launch(ctx, "pnp_ransac");
    if (rc != SLAMCU_OK) return rc;
    CU(ctx, cudaMemcpyAsync(counts2, dc, (size_t)n_hypotheses * 8, cudaMemcpyDeviceToHost, ctx->stream));
    if (Rt24) CU(ctx, cudaMemcpyAsync(Rt24, dr, br, cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return SLAMCU_OK;
}

int slamcu_fivept_solve(slamcu_context* ctx, const double* x1, const double* x2, int n_samples, double* models, int32_t* counts) {
    if (!ctx) return bad_args(ctx, __func__);
    if (!x1 || !x2 || !models || !counts || n_samples <= 0) return fail(ctx, SLAMCU_INVALID_ARGUMENT, "bad arguments");
    CU(ctx, cudaSetDevice(ctx->device));
    const size_t b_pts = (size_t)n_samples * 10 * 8, b_mod = (size_t)n_samples * 90 * 8, b_cnt = (size_t)n_samples * 4;
    int rc = ensure_scratch(ctx, 2 * b_pts + b_mod + b_cnt + 256);
    if (rc != SLAMCU_OK) return rc;
    uint8_t* base = static_cast<uint8_t*>(ctx->scratch);
    double* d1 = reinterpret_cast<double*>(base);
    double* d2 = reinterpret_cast<double*>(base + b_pts);
    double* dm = reinterpret_cast<double*>(base + 2 * b_pts);
    int* dc = reinterpret_cast<int*>(base + 2 * b_pts + b_mod);
    CU(ctx, cudaMemcpyAsync(d1, x1, b_pts, cudaMemcpyHostToDevice, ctx->stream));
    CU(ctx, cudaMemcpyAsync(d2, x2, b_pts, cudaMemcpyHostToDevice, ctx->stream));
    CU(ctx, cudaMemsetAsync(dm, 0, b_mod, ctx->stream));
    ctx->launches += launch_fivept_probe(d1, d2, n_samples, dm, dc, ctx->stream);
    rc = check_launch(ctx, "fivept probe");
    if (rc != SLAMCU_OK) return rc;
    CU(ctx, cudaMemcpyAsync(models, dm, b_mod, cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaMemcpyAsync(counts, dc, b_cnt, cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return SLAMCU_OK;
}

int slamcu_sequence_essential(slamcu_sequence* s, int first, int n_pairs, const double* K4, double prob, double threshold,
                              int max_iters) {
    if (!s || !K4) return bad_args((s ? s->ctx : nullptr), __func__);
    slamcu_context* ctx = s->ctx;
    if (first < 0 || n_pairs < 0 || first + n_pairs + 1 > s->max_frames) return fail(ctx, SLAMCU_INVALID_ARGUMENT, "bad pair range");
    if (n_pairs == 0) return SLAMCU_OK;
    if (max_iters > kEssentialMaxIters) return fail(ctx, SLAMCU_INVALID_ARGUMENT, "maxIters above %d is not supported", kEssentialMaxIters);
    CU(ctx, cudaSetDevice(ctx->device));
    const SeqView& v = s->v;
    if (!s->has_ess) {
        const size_t F = (size_t)s->max_frames;
        int rc = SLAMCU_OK;
        auto A = [&](auto** p, size_t count) {
            if (rc == SLAMCU_OK) rc = dev_alloc(ctx, p, count, s->ess_owned, true);
        };
        EssentialJob& e = s->ess;
        A(&e.x1, F * v.cap_kp);
        A(&e.x2, F * v.cap_kp);
        A(&e.n_pts, F);
        A(&e.E, F * 9);
        A(&e.n_inliers, F);
        A(&e.n_iters, F);
        A(&e.mask, F * v.cap_kp);
        if (rc != SLAMCU_OK) return rc;
        e.pt_stride = v.cap_kp;
        s->has_ess = true;
    }
    const size_t work_bytes = (size_t)n_pairs * essential_work_bytes_per_pair(max_iters);
    if (work_bytes > s->ess_work_bytes) {  // RANSAC working set of one launch; grows with n_pairs * max_iters
        CU(ctx, cudaStreamSynchronize(ctx->stream));
        if (s->ess_work) cudaFree(s->ess_work);
        s->ess_work = nullptr;
        s->ess_work_bytes = 0;
        CU(ctx, cudaMalloc(&s->ess_work, work_bytes));
        s->ess_work_bytes = work_bytes;
    }
    EssentialJob j = s->ess;
    j.work = static_cast<unsigned char*>(s->ess_work);
    j.x1 += (size_t)first * v.cap_kp;
    j.x2 += (size_t)first * v.cap_kp;
    j.n_pts += first;
    j.E += (size_t)first * 9;
    j.n_inliers += first;
    j.n_iters += first;
    j.mask += (size_t)first * v.cap_kp;
    essential_params(j, K4, prob, threshold, max_iters);
    ProfGuard pg(ctx);
    ctx->launches += launch_essential_gather(v, first, n_pairs, j, K4, ctx->stream);
    ctx->launches += launch_essential_ransac(j, n_pairs, ctx->stream);
    return check_launch(ctx, "essential kernels");
}

int slamcu_sequence_essential_read(slamcu_sequence* s, int pair, double* E9, int* n_inliers, int* n_iters, uint8_t* mask,
                                   int capacity, int* n_points) {
    if (!s) return bad_args((s ? s->ctx : nullptr), __func__);
    slamcu_context* ctx = s->ctx;
    if (!s->has_ess) return fail(ctx, SLAMCU_UNSUPPORTED, "slamcu_sequence_essential has not run on this sequence");
    if (pair < 0 || pair + 1 >= s->max_frames) return fail(ctx, SLAMCU_INVALID_ARGUMENT, "bad pair index");
    const EssentialJob& e = s->ess;
    int h[3] = {0, 0, 0};
    CU(ctx, cudaMemcpyAsync(&h[0], e.n_pts + pair, 4, cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaMemcpyAsync(&h[1], e.n_inliers + pair, 4, cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaMemcpyAsync(&h[2], e.n_iters + pair, 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (E9) CU(ctx, cudaMemcpyAsync(E9, e.E + (size_t)pair * 9, 72, cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    if (n_points) *n_points = h[0];
    if (n_inliers) *n_inliers = h[1];
    if (n_iters) *n_iters = h[2];
    if (mask) {
        if (h[0] > capacity) return fail(ctx, SLAMCU_CAPACITY, "need room for %d mask bytes", h[0]);
        if (h[0] > 0) {
            CU(ctx, cudaMemcpyAsync(mask, e.mask + (size_t)pair * e.pt_stride, (size_t)h[0], cudaMemcpyDeviceToHost, ctx->stream));
            CU(ctx, cudaStreamSynchronize(ctx->stream));
        }
    }
    return SLAMCU_OK;
}

}  // extern "C"
