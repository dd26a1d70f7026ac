// common.cuh -- shared device-side views and launch declarations of the slamcu kernel library (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/slam/cuda/slamcu.h"

namespace slamcu {

constexpr int kKeyIdxBits = 20;                 // sort key = (score << 20) | raster index
constexpr uint32_t kKeyIdxMask = (1u << kKeyIdxBits) - 1;
constexpr int kMaxRawCap = 1 << kKeyIdxBits;    // raw-corner capacity per frame
constexpr int kMaxPattern = 1024;               // BRIEF pairs held in constant memory

// status bits accumulated per frame on the device
enum : int { kStRawOverflow = 1, kStKpOverflow = 2, kStMatchOverflow = 4 };

// Device-resident layout of a frame sequence (all pointers are HBM; `F` = max_frames).
struct SeqView {
    int rows, cols, pitch;   // pitch: bytes per image row, multiple of 128
    int mwords;              // 32-bit words per row of a 1-bit/pixel mask
    int cap_raw, cap_kp;     // capacities per frame
    int desc_bytes;          // bytes per descriptor as seen by the caller
    int desc_words;          // 32-bit words per descriptor row in HBM (zero padded)
    int qcap;                // segment-queue capacity (sort)
    size_t frame_bytes;      // rows * pitch
    uint8_t* img;            // [F][rows][pitch]     input frames
    uint8_t* blur;           // [F][rows][pitch]     5x5 gaussian of the frame
    uint32_t* mask;          // [F][rows][mwords]    FAST corner bitmask
    uint32_t* raw_xy;        // [F][cap_raw]         (y << 16) | x, raster order
    uint32_t* keys;          // [F][cap_raw]         (score << 20) | raster index; sorted in place
    uint32_t* sort_scratch;  // [F][scratch_words]   L/R stop lists, segment queues, boundary bitmap
    size_t scratch_words;
    uint32_t* nms_bitmap;    // [F][rows][mwords]    suppression bitmap when it does not fit in smem
    int* n_raw;              // [F]
    slamcu_keypoint* kps;    // [F][cap_kp]
    int* n_kp;               // [F]
    uint32_t* desc;          // [F][cap_kp][desc_words]
    uint32_t* desc_or;       // [F][desc_words]      OR of all descriptors of the frame
    int4* cand;              // [F][cap_kp]          per-query {bestIdx, bestDist, secondDist, secondIdx}
    slamcu_dmatch* matches;   // [F][cap_kp]          pair (f, f+1)
    int* n_match;            // [F]
    int* status;             // [F]
};

struct DetParams {
    int thr, arc, nms, window, patch, pairs, n_pattern;
    double blur_w[25];
};

struct MatchParams {
    int filter, good, use_ratio;
    float ratio;
};

// ---- launchers (each returns the number of kernels it enqueued) --------------------------------
int launch_fast_corners(const SeqView& s, int first, int n, const DetParams& p, cudaStream_t st);
int launch_sort_nms(const SeqView& s, int first, int n, const DetParams& p, int smem_optin, cudaStream_t st);
int launch_raster_keypoints(const SeqView& s, int first, int n, bool scored, cudaStream_t st);
int launch_blur(const SeqView& s, int first, int n, const DetParams& p, cudaStream_t st);
int launch_describe(const SeqView& s, int first, int n, const DetParams& p, const int* d_pattern, cudaStream_t st);
// generic matcher: nq x nt descriptors, optional keypoints; pair p: query = base + p*stride
struct MatchJob {
    const uint32_t* dq; const uint32_t* dt;           // descriptor rows (desc_words words each)
    const slamcu_keypoint* kq; const slamcu_keypoint* kt;  // may be null
    const int* nq; const int* nt;                      // device counts (one per pair, strided)
    const uint32_t* orq; const uint32_t* ort;          // per-set OR masks (may be null)
    int4* cand; slamcu_dmatch* matches; int* n_match; int* status;
    size_t desc_pair_stride;   // words between consecutive pairs' descriptor blocks
    size_t kp_pair_stride;     // keypoints between consecutive pairs
    int count_stride;          // ints between consecutive pairs' counts
    size_t cand_pair_stride;   // entries between pairs in cand / matches
    int or_stride;
    int desc_words, max_q, cap_out;
    int max_t;                 // upper bound of the train counts (sizes the shared-memory tile)
};
// sort_keys: [n_pairs][cand_pair_stride] 64-bit scratch for the top-K / sort stage
// n_seg > 1 splits the train set into n_seg slices of seg_len descriptors (extra grid dimension, for single small
// problems); job.cand must then hold n_seg blocks of seg_stride entries, slice 0 doubling as the merged result
void init_match_attributes();
struct MatchTc;  // match_tc.cuh: widened operands + tensor maps; non-null -> the 256-bit penalty-free search runs on the tensor cores
int launch_match(const MatchJob& job, int n_pairs, const MatchParams& p, bool emit_matches, int with_kp,
                 unsigned long long* sort_keys, cudaStream_t st, int n_seg = 1, int seg_len = 0, size_t seg_stride = 0,
                 const MatchTc* tc = nullptr);

int launch_prepare(const uint8_t* src, int channels, int src_stride, size_t src_frame_bytes, const int* map, uint8_t* dst, int pitch,
                   size_t dst_frame_bytes, int rows, int cols, int n, cudaStream_t st);
int launch_repitch(const uint8_t* src, int stride, uint8_t* dst, int pitch, int cols, long long rows_total, cudaStream_t st);
int launch_bgr2gray(const uint8_t* bgr, int rows, int cols, int stride, uint8_t* gray, int gstride, cudaStream_t st);
struct CamParams { double fx, fy, cx, cy, k1, k2, p1, p2; };
// fix: [1 + fix_cap] ints, fix[0] = number of pixels the host must recompute (those within 1e-6 of a rounding boundary)
int launch_undistort_map(int rows, int cols, const CamParams& cam, int* map, int* fix, int fix_cap, cudaStream_t st);
int launch_remap(const uint8_t* gray, int rows, int cols, int stride, const int* map, uint8_t* out_u8, double* out_f64,
                 cudaStream_t st);
int launch_ransac_score(const double* models9, int n_models, const double* x1, const double* x2, int n, double thr2,
                        int* counts, uint8_t* masks, cudaStream_t st);

// result compaction (pack.cu)
int launch_dense_scan(const int* counts, int first, int n, int* off, cudaStream_t st);
int launch_dense_copy(const SeqView& s, int first, int n_frames, int pair_first, int n_pairs, const int* kp_off, const int* m_off, void* h_kps,
                      void* h_desc, void* h_matches, int kp_cap, int m_cap, int* overflow, cudaStream_t st);
int launch_pack_counts(const SeqView& s, int first, int n, int* out4, cudaStream_t st);

// two-view geometry (essential.cu): one RANSAC problem per frame pair
struct EssentialJob {
    double2* x1; double2* x2;        // [pairs][pt_stride] correspondences normalised by K
    int* n_pts;                      // [pairs]
    double* E;                       // [pairs][9] row-major, |E|_F = 1 (zeros when no model was accepted)
    int* n_inliers; int* n_iters;    // [pairs]
    uint8_t* mask;                   // [pairs][pt_stride]
    unsigned char* work;             // [pairs in one launch][essential_work_bytes_per_pair(max_iters)] RANSAC working set
    int pt_stride, max_iters;
    double prob;
    float thr2;                      // (float)(t * t), t = threshold / ((fx + fy) / 2)
};
constexpr int kEssentialMaxIters = 50000;  // largest maxIters accepted (the replay keeps one word per iteration in smem)
void init_essential_attributes();
size_t essential_work_bytes_per_pair(int max_iters);
int launch_essential_gather(const SeqView& s, int first, int n_pairs, const EssentialJob& job, const double* K4, cudaStream_t st);
int launch_essential_normalise(const float* p1, const float* p2, int n, const EssentialJob& job, const double* K4, cudaStream_t st);
int launch_essential_ransac(const EssentialJob& job, int n_pairs, cudaStream_t st);
int launch_recover_pose(const EssentialJob& job, int n_pairs, const double* K4, double* R, double* t, int* front, cudaStream_t st);
// P24: the two row-major 3x4 projection matrices back to back (device); out4 / out3 may be null
int launch_triangulate(const double* P24, const float* p1, const float* p2, int n, double* out4, double* out3, cudaStream_t st);
// counts[2 * n_hyp], Rt[2 * n_hyp][12] (R row-major, then t): both signs of the DLT null vector per hypothesis
int launch_pnp_ransac(const double* X3, const double* x2, int n, const int* samples6, int n_hyp, const double* K9, double thr, int* counts,
                      double* Rt, cudaStream_t st);
int launch_fivept_probe(const double* x1, const double* x2, int n_samples, double* models, int* counts, cudaStream_t st);

// ---- optional per-kernel timing (CUDA events on the launching stream; bench.py's roofline input) ----
struct Profiler {
    virtual void begin(const char* name, cudaStream_t st) = 0;
    virtual void end(cudaStream_t st) = 0;
};
extern thread_local Profiler* g_prof;  // set by the API entry points while profiling is enabled
#define SLAM_KERNEL(name, st, ...)              \
    do {                                         \
        if (slamcu::g_prof) slamcu::g_prof->begin(name, st); \
        __VA_ARGS__;                             \
        if (slamcu::g_prof) slamcu::g_prof->end(st);         \
    } while (0)

// -DSLAMCU_DEBUG_BOUNDS (slam_cin0051_b200.build.build_debug -> libslamcu_dbg.so): every atomically indexed list write, tile
// offset and gather named in DESIGN.md traps when its index leaves [0, limit).  compute-sanitizer is not available on the GPU
// pool, so tests/test_gpu_debug_bounds.py runs the edge cases (noise, tiny images, list overflow) through this build instead.
#ifdef SLAMCU_DEBUG_BOUNDS
#define SLAMCU_BOUND(index, limit)                                                             \
    do {                                                                                       \
        if (!((long long)(index) >= 0 && (long long)(index) < (long long)(limit))) __trap();  \
    } while (0)
#else
#define SLAMCU_BOUND(index, limit) ((void)0)
#endif

__device__ __forceinline__ unsigned lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ unsigned lanemask_lt() {
    unsigned m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

}  // namespace slamcu
