/* slamcu.h -- C ABI of the B200-native (sm_100a) SLAM-frontend kernel library.
 *
 * This is the drop-in boundary for the reference's data-parallel hot path.  The reference has no FFI
 * layer: its boundary is four C++ classes.  Each entry point below names the reference interface it
 * replaces (paths relative to the reference repo):
 *
 *   slam::FeatureDetector   include/slam/frontend/feature_detector.hpp:47-192, src/frontend/feature_detector.cpp
 *   slam::FeatureMatcher    include/slam/frontend/feature_matcher.hpp:38-87,   src/frontend/feature_matcher.cpp
 *   slam::PoseEstimator     include/slam/frontend/pose_estimator.hpp:13-36,    src/frontend/pose_estimator.cpp:18-67
 *   slam::Camera / Preprocessor  include/slam/common/common.hpp:67-190, src/preprocessing/preprocessor.cpp:95-141
 *
 * Rules of the ABI: plain pointers and sizes only (no STL / Eigen / OpenCV / torch types), every
 * function returns an int status (0 = ok), caller allocates outputs and passes capacities, no
 * exceptions cross the boundary.  One context is single-threaded; different contexts may be driven
 * from different host threads / GPUs.  There is NO CPU fallback: every call either runs the CUDA path
 * or fails with a status.
 */
#ifndef SLAM_CUDA_SLAMCU_H_
#define SLAM_CUDA_SLAMCU_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SLAMCU_ABI_VERSION 1

/* ---- status codes -------------------------------------------------------------------------- */
enum {
    SLAMCU_OK = 0,
    SLAMCU_INVALID_ARGUMENT = 1, /* bad pointer / size / config value (adapters throw std::runtime_error)      */
    SLAMCU_EMPTY_INPUT = 2,      /* "Empty descriptors provided." -> std::invalid_argument (feature_matcher.cpp:99-102) */
    SLAMCU_SIZE_MISMATCH = 3,    /* "Descriptor dimensions must match." (feature_matcher.cpp:108-110), image size   */
    SLAMCU_CAPACITY = 4,         /* an output or internal list overflowed its capacity                          */
    SLAMCU_CUDA_ERROR = 5,       /* CUDA runtime failure; see slamcu_last_error                                 */
    SLAMCU_UNSUPPORTED = 6       /* e.g. DistanceType L2 on uint8 descriptors (feature_matcher.cpp:86,104-106)    */
};

/* ---- plain data mirrored from the reference ------------------------------------------------- */
/* slam::Keypoint, feature_detector.hpp:28-38: five floats, 20 bytes, AoS. */
typedef struct slamcu_keypoint {
    float x, y, size, angle, response;
} slamcu_keypoint;

/* slam::Match, feature_matcher.hpp:18-25. */
typedef struct slamcu_dmatch {
    int32_t queryIdx, trainIdx;
    float distance;
} slamcu_dmatch;

/* k = 2 nearest-neighbour record (cv::BFMatcher::knnMatch row), mode "orb" only. */
typedef struct slamcu_knn2 {
    int32_t trainIdx0, trainIdx1;
    float distance0, distance1;
} slamcu_knn2;

#define SLAMCU_MODE_REFERENCE 0 /* exactly slam::FeatureDetector / FeatureMatcher                     */
#define SLAMCU_MODE_ORB 1       /* OpenCV-ORB compatible: pyramid + FAST-9 + Harris + rBRIEF-256 (opt-in) */

/* The six YAML scalars of feature_detector.hpp:60-93, plus the two host-generated tables whose
 * provenance must stay on the host (libstdc++ <random> pattern, libm exp weights), plus the opt-in
 * ORB-mode keys (absent from the reference's YAML => mode 0). */
typedef struct slamcu_detector_config {
    int32_t intensity_threshold;         /* IntensityThreshold        [0,255]  */
    int32_t contiguous_pixels_threshold; /* ContiguousPixelsThreshold [0,16]   */
    int32_t non_max_suppression;         /* NonMaxSuppression         0|1      */
    int32_t suppression_window_size;     /* SuppressionWindowSize     > 0      */
    int32_t patch_size;                  /* PatchSize                 odd > 0  */
    int32_t num_brief_pairs;             /* NumBRIEFPairs             8k > 0   */
    int32_t n_pattern;                   /* number of surviving pairs from generateBRIEFPattern()           */
    const int32_t* pattern;              /* [n_pattern][4] = x1,y1,x2,y2  (feature_detector.cpp:286-313)    */
    const double* blur_weights;          /* [25] normalised 5x5 sigma=1 kernel (feature_detector.cpp:321-335) */
    /* --- ORB mode (SLAMCU_MODE_ORB) --- */
    int32_t mode;
    int32_t n_levels;      /* NumLevels   (8)    */
    float scale_factor;    /* ScaleFactor (1.2f) */
    int32_t max_features;  /* MaxFeatures (2000) */
    int32_t fast_threshold; /* = intensity_threshold by default */
    const int32_t* orb_pattern; /* [256][4] rBRIEF test pairs x1,y1,x2,y2; NULL = OpenCV's bit_pattern_31_ (built in) */
} slamcu_detector_config;

/* feature_matcher.cpp:25-57 */
#define SLAMCU_DISTANCE_HAMMING 0
#define SLAMCU_DISTANCE_L2 1
typedef struct slamcu_matcher_config {
    int32_t distance_type;      /* DistanceType */
    int32_t filter_matches;     /* FilterMatches 0|1 */
    int32_t good_matches_count; /* GoodMatchesCount  */
    int32_t use_ratio_test;     /* UseRatioTest 0|1  */
    float ratio_test_threshold; /* RatioTestThreshold [0,1] */
} slamcu_matcher_config;

typedef struct slamcu_context slamcu_context;
typedef struct slamcu_detector slamcu_detector;
typedef struct slamcu_matcher slamcu_matcher;
typedef struct slamcu_sequence slamcu_sequence;

/* ---- context -------------------------------------------------------------------------------- */
int slamcu_abi_version(void);
const char* slamcu_status_string(int status);
int slamcu_device_count(int* count);
/* One context per device; owns a stream unless slamcu_set_stream() hands it an external cudaStream_t. */
int slamcu_create(int device_id, slamcu_context** out);
void slamcu_destroy(slamcu_context* ctx);
const char* slamcu_last_error(const slamcu_context* ctx);
int slamcu_set_stream(slamcu_context* ctx, void* cuda_stream /* cudaStream_t, NULL = own stream */);
void* slamcu_get_stream(slamcu_context* ctx);
int slamcu_synchronize(slamcu_context* ctx);
/* Page-locked host memory for the asynchronous sequence calls (so host programs need not link the CUDA runtime). */
int slamcu_alloc_pinned(size_t bytes, void** out);
void slamcu_free_pinned(void* p);
/* number of kernels this library launched on the context since creation (bench.py gpu_launches) */
int64_t slamcu_launch_count(const slamcu_context* ctx);
/* Per-kernel device timing: when enabled, every kernel launch on the context is bracketed by CUDA
 * events on the launching stream.  slamcu_profile_read(index) synchronises and returns the accumulated
 * milliseconds and launch count of the index-th kernel name (SLAMCU_INVALID_ARGUMENT past the end). */
/* Measured POPC issue ceiling of the device (G popc/s): the matcher's roofline denominator. */
int slamcu_popc_peak(slamcu_context* ctx, double* gpopc_per_s);
/* Test hook of the -DSLAMCU_DEBUG_BOUNDS build (libslamcu_dbg.so): evaluates one out-of-range index through the range-check macro
 * the kernels use, which traps there; a no-op in the release build. */
int slamcu_debug_trip_bound(slamcu_context* ctx);
int slamcu_profile_enable(slamcu_context* ctx, int on);
int slamcu_profile_read(slamcu_context* ctx, int index, char* name, int name_cap, double* total_ms, int64_t* launches);

/* ---- FeatureDetector (feature_detector.hpp:53,114-135) ---------------------------------------- */
/* Constructor-time host tables, produced with the same host-library calls as the reference:
 * generateBRIEFPattern() (feature_detector.cpp:286-313; std::default_random_engine +
 * std::normal_distribution<float>) and the 5x5 sigma=1 kernel of gaussianBlur() (:321-335; std::exp). */
int slamcu_default_brief_pattern(int patch_size, int num_pairs, int32_t* pattern4, int capacity_pairs, int* n_out);
int slamcu_default_blur_weights(double* weights25);
int slamcu_detector_create(slamcu_context* ctx, const slamcu_detector_config* cfg, slamcu_detector** out);
void slamcu_detector_destroy(slamcu_detector* det);

/* FeatureDetector::detect (feature_detector.cpp:8-18).  image: row-major u8, `stride` bytes per row
 * (EigenGrayMatrix::data(), stride = cols).  *n_out receives the keypoint count; if it exceeds
 * `capacity` the call fails with SLAMCU_CAPACITY (and *n_out holds the needed size). */
int slamcu_detect(slamcu_detector* det, const uint8_t* image, int rows, int cols, int stride,
                  slamcu_keypoint* keypoints, int capacity, int* n_out);
/* FeatureDetector::compute (feature_detector.cpp:20-47): keypoints in/out (angle written);
 * descriptors: n x (num_brief_pairs/8) bytes, row stride desc_stride. n == 0 is ok (no output). */
int slamcu_compute(slamcu_detector* det, const uint8_t* image, int rows, int cols, int stride,
                   slamcu_keypoint* keypoints, int n, uint8_t* descriptors, int desc_stride);
/* FeatureDetector::detectAndCompute (feature_detector.cpp:49-54). */
int slamcu_detect_and_compute(slamcu_detector* det, const uint8_t* image, int rows, int cols, int stride,
                              slamcu_keypoint* keypoints, uint8_t* descriptors, int desc_stride, int capacity,
                              int* n_out);
/* Stage probes used by the parity tests (not part of the reference's public surface):
 * raster-order FAST corners with their SAD score in `response` (feature_detector.cpp:56-68,190-203) */
int slamcu_fast_corners(slamcu_detector* det, const uint8_t* image, int rows, int cols, int stride,
                        slamcu_keypoint* keypoints, int capacity, int* n_out);
/* ORB mode only: pyramid level (cv::KeyPoint::octave) of every keypoint of the last single-frame call. */
int slamcu_detector_last_octaves(slamcu_detector* det, int32_t* octaves, int capacity);
/* ORB mode stage probes (parity tests): lists of the last single-frame call at one pyramid level.
 * stage 0 = FAST-9 + 3x3 NMS + border filter survivors (value = FAST score), 1 = after retainBest(2*quota)
 * (value = Harris response), 2 = after retainBest(quota).  xy[i] = (y << 16) | x in level coordinates. */
int slamcu_orb_stage(slamcu_detector* det, int stage, int level, uint32_t* xy, float* value, int capacity, int* n_out);
/* ORB mode: pyramid level `level` of the last single-frame call (blurred != 0: the 7x7 sigma=2 blurred copy the
 * descriptors sample).  out may be NULL to query the size only. */
int slamcu_orb_level_image(slamcu_detector* det, int level, int blurred, uint8_t* out, int out_stride, int* rows,
                           int* cols);
/* FeatureDetector::gaussianBlur(image, 5, 1.0) (feature_detector.cpp:315-364) */
int slamcu_gaussian_blur(slamcu_detector* det, const uint8_t* image, int rows, int cols, int stride, uint8_t* out,
                         int out_stride);

/* ---- FeatureMatcher (feature_matcher.hpp:50,64-66) ------------------------------------------- */
int slamcu_matcher_create(slamcu_context* ctx, const slamcu_matcher_config* cfg, slamcu_matcher** out);
void slamcu_matcher_destroy(slamcu_matcher* m);
/* FeatureMatcher::match (feature_matcher.cpp:71-95).  d1: n1 x width bytes (row stride = width),
 * d2: n2 x width.  kp1/kp2 optional (NULL or n == 0 => no distance penalty, feature_matcher.cpp:150).
 * Output in the reference's order: query order, or the std::partial_sort / std::sort order when
 * FilterMatches = 1.  Errors: SLAMCU_EMPTY_INPUT, SLAMCU_SIZE_MISMATCH, SLAMCU_UNSUPPORTED. */
int slamcu_match(slamcu_matcher* m, const uint8_t* d1, int n1, int width1, const uint8_t* d2, int n2, int width2,
                 const slamcu_keypoint* kp1, int nkp1, const slamcu_keypoint* kp2, int nkp2, slamcu_dmatch* matches,
                 int capacity, int* n_out);
/* Brute-force k = 2 Hamming search for every query (no ratio/filter): best/second by (distance, index).
 * This is both the sweep benchmark kernel (BASELINE.json config 5) and cv::BFMatcher::knnMatch(k=2). */
int slamcu_knn2_hamming(slamcu_matcher* m, const uint8_t* d1, int n1, const uint8_t* d2, int n2, int width,
                        slamcu_knn2* out);
/* Tuning / test knob of the single-problem calls above: the train set is searched in `n_slices` slices (an extra grid
 * dimension; the per-slice top-2 lists are merged by (distance, index), which is exact).  0 = automatic (fill the
 * device), 1 = one slice, at most 32.  Results do not depend on it. */
int slamcu_matcher_set_train_slices(slamcu_matcher* m, int n_slices);

/* ---- device-resident sequences: the batched / asynchronous path -------------------------------
 * A sequence holds up to max_frames frames of one size in HBM together with every intermediate of
 * the frontend (corner masks, raw corner lists, keypoints, descriptors, consecutive-frame matches).
 * All calls enqueue on the context's stream and return immediately unless stated otherwise.
 * rows, cols and max_frames are at most 65535 (longer recordings are processed as several sequences). */
int slamcu_sequence_create(slamcu_context* ctx, int rows, int cols, int max_frames, int max_raw_corners,
                           int max_keypoints, int desc_bytes, slamcu_sequence** out);
void slamcu_sequence_destroy(slamcu_sequence* seq);
/* H2D copy of n frames from host memory (pinned for async) into slots [first, first+n). */
int slamcu_sequence_upload(slamcu_sequence* seq, int first, int n, const uint8_t* host_frames, int stride);
/* Preprocessor::yield for a batch (preprocessor.cpp:136-137), straight into slots [first, first+n): n host frames of
 * rows x stride bytes with 1 (gray) or 3 (BGR, as cv::imread(IMREAD_COLOR) gives) channels are uploaded with one linear
 * copy, converted with cv::cvtColor(BGR2GRAY)'s fixed-point formula and, when K4/D4 (fx,fy,cx,cy / k1,k2,p1,p2) are given,
 * undistorted with Camera::undistortImage's forward map + nearest-neighbour gather (outside -> 0).  The result is the
 * 8-bit image the detector consumes (the reference's value / 255.0 double image has no consumer).  Asynchronous. */
int slamcu_sequence_prepare(slamcu_sequence* seq, int first, int n, const uint8_t* host_frames, int channels, int stride,
                            const double* K4, const double* D4);
/* The 8-bit frame in slot f as the detector sees it (after upload / prepare); synchronises. */
int slamcu_sequence_image(slamcu_sequence* seq, int f, uint8_t* out, int out_stride);
/* Device pointer / pitch of the frame store, for producers that already live on the device. */
int slamcu_sequence_frames_device(slamcu_sequence* seq, void** dptr, int* pitch, int64_t* frame_bytes);
/* detectAndCompute on frames [first, first+n). */
int slamcu_sequence_extract(slamcu_sequence* seq, slamcu_detector* det, int first, int n);
/* match(frame f, frame f+1) for f in [first, first+n_pairs); with_keypoints selects the penalty path. */
int slamcu_sequence_match(slamcu_sequence* seq, slamcu_matcher* m, int first, int n_pairs, int with_keypoints);
/* Both steps for frames [first, first+n) -- detectAndCompute on every frame, match(f, f+1) on the n - 1 pairs inside the range --
 * chunk <= 0: one after the other on the context's stream (fastest with resident inputs: 14.27 vs 14.49 ms per 1000 frames
 * measured); chunk > 0: in chunks of that many frames alternating between the context's two compute lanes, the matcher of one
 * chunk next to the extraction kernels of the following one (what slamcu_sequence_process does, where it gains 3 %).  Same
 * results either way.  Asynchronous; ordered before later work on the context's stream. */
int slamcu_sequence_extract_match(slamcu_sequence* seq, slamcu_detector* det, slamcu_matcher* m, int first, int n,
                                  int with_keypoints, int chunk);
/* D2H of per-frame counts {n_keypoints, n_matches, n_raw_corners, status} (int32[n][4]); synchronises. */
int slamcu_sequence_counts(slamcu_sequence* seq, int first, int n, int32_t* counts4);
/* D2H of one frame's keypoints + descriptors / one pair's matches; synchronises. */
int slamcu_sequence_frame(slamcu_sequence* seq, int f, slamcu_keypoint* keypoints, uint8_t* descriptors,
                          int desc_stride, int capacity, int* n_out);
/* ORB mode only: octaves of frame f's keypoints. */
int slamcu_sequence_octaves(slamcu_sequence* seq, int f, int32_t* octaves, int capacity);
int slamcu_sequence_matches(slamcu_sequence* seq, int f, slamcu_dmatch* matches, int capacity, int* n_out);
/* Bulk asynchronous D2H of everything a consumer needs for frames [first, first+n): keypoints
 * [n][max_keypoints], descriptors [n][max_keypoints][desc_bytes], matches [n][max_keypoints],
 * counts [n][4], into pinned host buffers (any pointer may be NULL to skip it). */
int slamcu_sequence_download(slamcu_sequence* seq, int first, int n, slamcu_keypoint* keypoints, uint8_t* descriptors,
                             slamcu_dmatch* matches, int32_t* counts4);

/* The whole per-frame loop for n host frames (the SLAMModel wiring the reference sketches at
 * include/slam/model/model.hpp:20-27): upload -> detectAndCompute -> match(f, f+1) -> download, software-pipelined
 * in chunks of `chunk` frames (<= 0: 64) over separate H2D / compute / D2H streams.  host_frames: n frames of
 * rows x stride bytes (pinned for overlap); outputs as in slamcu_sequence_download (any may be NULL; pinned).
 * Asynchronous: slamcu_sequence_wait(seq) waits for this sequence's latest call (kernels and downloads),
 * slamcu_synchronize() for everything.  Dependencies are tracked per sequence, so calls that alternate between two
 * sequences double-buffer: the copies of one overlap the kernels of the other. */
int slamcu_sequence_process(slamcu_sequence* seq, slamcu_detector* det, slamcu_matcher* m, const uint8_t* host_frames,
                            int stride, int n, int chunk, int with_keypoints, slamcu_keypoint* keypoints,
                            uint8_t* descriptors, slamcu_dmatch* matches, int32_t* counts4);

/* The same loop with DENSE outputs: a device compaction kernel packs only the defined rows, frame after frame, and the three
 * arrays cross the link as exact-size copy-engine transfers into the caller's page-locked buffers (slamcu_alloc_pinned /
 * cudaHostAlloc), so no padding is moved.  The call itself stays asynchronous; the transfers are issued by
 * slamcu_sequence_wait(seq) / slamcu_synchronize() once the totals have reached the host, and the buffers are valid when
 * that call returns (the per-frame counts4 arrive as in slamcu_sequence_process):
 *   keypoints / descriptors of frame f at row kp_off[f] = counts4[0][0] + ... + counts4[f-1][0]
 *   matches of pair (f, f+1)           at row m_off[f]  = counts4[0][1] + ... + counts4[f-1][1]
 * kp_capacity / match_capacity: rows the buffers hold; slamcu_sequence_wait returns SLAMCU_CAPACITY when they were exceeded. */
int slamcu_sequence_process_dense(slamcu_sequence* seq, slamcu_detector* det, slamcu_matcher* m, const uint8_t* host_frames,
                                  int stride, int n, int chunk, int with_keypoints, slamcu_keypoint* keypoints,
                                  uint8_t* descriptors, slamcu_dmatch* matches, int32_t* counts4, int64_t kp_capacity,
                                  int64_t match_capacity);
/* Waits for the sequence's latest slamcu_sequence_process[_dense] call.  SLAMCU_CAPACITY if a frame overflowed one of its
 * device lists (its results are truncated) or the dense outputs overflowed their capacities. */
int slamcu_sequence_wait(slamcu_sequence* seq);
/* Per-frame counts {n_keypoints, n_matches, n_raw_corners, status} of frames [first, first+n) packed into a DEVICE array
 * int32[n][4] on the context's stream: the payload of the multi-GPU path's only collective (one all-gather of
 * per-frame counts over NCCL), with no host round trip.  Asynchronous. */
int slamcu_sequence_counts_device(slamcu_sequence* seq, int first, int n, int32_t* device_counts4);

/* ---- image preparation (src/preprocessing) ---------------------------------------------------- */
/* cv::cvtColor(BGR2GRAY) (preprocessor.cpp:136): gray = (3735 B + 19235 G + 9798 R + 16384) >> 15. */
int slamcu_bgr_to_gray(slamcu_context* ctx, const uint8_t* bgr, int rows, int cols, int stride, uint8_t* gray,
                       int gray_stride);
/* Camera::undistortImage (common.hpp:127-173).  K4 = fx,fy,cx,cy; D4 = k1,k2,p1,p2 (k3 is unused by the
 * reference).  out_u8 (rows x cols, may be NULL) receives the gathered bytes the detector consumes;
 * out_f64 (rows x cols row-major, may be NULL) receives the reference's value/255.0 image. */
int slamcu_undistort(slamcu_context* ctx, const uint8_t* gray, int rows, int cols, int stride, const double* K4,
                     const double* D4, uint8_t* out_u8, double* out_f64);

/* ---- two-view geometry (pose_estimator.cpp:18-67 -> cv::findEssentialMat(RANSAC)) ------------- */
/* Sampson-error scoring of n_models 3x3 essential-matrix hypotheses against n correspondences
 * (already normalised by K, double) with cv's threshold semantics: inlier iff (float)err <= (float)thr2.
 * counts[n_models] inlier counts; masks (may be NULL) n_models x n bytes. */
int slamcu_ransac_score(slamcu_context* ctx, const double* models9, int n_models, const double* x1, const double* x2,
                        int n, double thr2, int32_t* counts, uint8_t* masks);
/* cv::findEssentialMat(p1, p2, K, RANSAC, prob, threshold, maxIters) as called by PoseEstimator::estimate
 * (pose_estimator.cpp:42; defaults 0.999, 1.0, 1000): cv::RNG sampling sequence, 5-point minimal solver, Sampson
 * scoring, sequential accept / adaptive-iteration rule, all on the device.  p1/p2: n x 2 float pixels; K4 =
 * fx,fy,cx,cy.  E9 row-major with |E|_F = 1 (all zeros when no hypothesis was accepted, like cv's empty Mat);
 * mask n bytes (may be NULL).  n < 6 -> SLAMCU_EMPTY_INPUT (the reference itself returns early below 8 matches,
 * pose_estimator.cpp:22-26). */
int slamcu_find_essential(slamcu_context* ctx, const float* p1, const float* p2, int n, const double* K4, double prob,
                          double threshold, int max_iters, double* E9, uint8_t* mask, int* n_inliers);
/* PoseEstimator::estimate (pose_estimator.cpp:18-67) end to end: findEssentialMat as above (defaults 0.999, 1.0, 1000),
 * then simpleRecoverPose (simple_pose_recover.cpp:35-97): E -> R1, R2, +-t by SVD and the cheirality vote over all
 * correspondences (4x4 DLT per point and candidate).  R9 row-major, t3 unit norm, front4 = points in front of both
 * cameras for the candidates (R1,t), (R2,t), (R1,-t), (R2,-t).  n < 8 or no accepted hypothesis -> SLAMCU_EMPTY_INPUT
 * (the reference logs a warning and returns without touching R, t). */
int slamcu_estimate_pose(slamcu_context* ctx, const float* p1, const float* p2, int n, const double* K4, double* E9, uint8_t* mask,
                         int* n_inliers, double* R9, double* t3, int32_t* front4);
/* slam::triangulate (common.hpp:201-221) as PoseEstimator::triangulatePoints calls it (pose_estimator.cpp:69-104): per
 * correspondence, the null vector of the 4x4 DLT system of the projection matrices P1, P2 (row-major 3x4, e.g. K [I|0] and
 * K [R|t]) and the pixel coordinates pts1 / pts2 (n x 2 float).  points4 (n x 4, may be NULL): the homogeneous solution, unit
 * norm, last component >= 0; points3 (n x 3, may be NULL): x / x[3], what triangulatePoints returns.  (The reference writes
 * its solution through Mat::copyTo into a column view of a differently typed matrix -- common.hpp:217 -- and returns
 * indeterminate values; this is the result that code evidently means, checked against numpy's SVD.) */
int slamcu_triangulate(slamcu_context* ctx, const double* P1, const double* P2, const float* pts1, const float* pts2, int n,
                       double* points4, double* points3);
/* The RANSAC of LoopClosure::verifyGeometricConsistency (loop_closure.cpp:177-222): for every hypothesis h, solvePnP
 * (:238-274, the 6-point DLT with the reference's own vector -> matrix mapping reproduced literally: K is never removed and the
 * row-major null vector is read back column-major) on the correspondences samples6[h][0..5], then the reprojection test of
 * every correspondence (:201-215; threshold in pixels, points behind the camera skipped).  The caller draws the sample indices
 * (the reference seeds std::mt19937 from std::random_device: the host adapter takes a seed instead).  The sign of the DLT null
 * vector is an implementation detail of the SVD and changes the outcome, so both signs are evaluated: counts2[2 h + s] and
 * Rt24[(2 h + s) * 12 ..] = R (row-major 3x3) then t, s = 0: the null vector's largest component positive, s = 1: negated.
 * n >= 6; points3d n x 3, points2d n x 2 (double); K9 row-major 3x3. */
/* The sample draw of that loop (:177-193) with the host library's std::mt19937 / std::uniform_int_distribution<int>(0, n-1),
 * seeded by the caller: samples6[n_hypotheses][6], six distinct indices per hypothesis.  Host only. */
int slamcu_pnp_sample_indices(uint32_t seed, int n, int n_hypotheses, int32_t* samples6);
int slamcu_pnp_ransac(slamcu_context* ctx, const double* points3d, const double* points2d, int n, const int32_t* samples6,
                      int n_hypotheses, const double* K9, double threshold, int32_t* counts2, double* Rt24);
/* Stage probe: the 5-point minimal solver on n_samples independent samples.  x1/x2: [n_samples][5][2] normalised
 * coordinates; models: [n_samples][10][9] (row-major E, unit norm); counts[n_samples] = solutions per sample. */
int slamcu_fivept_solve(slamcu_context* ctx, const double* x1, const double* x2, int n_samples, double* models,
                        int32_t* counts);
/* Batched: findEssentialMat on the matches of pairs (f, f+1), f in [first, first + n_pairs), of a sequence
 * (keypoints of match.queryIdx / match.trainIdx, as PoseEstimator::estimate gathers them, pose_estimator.cpp:30-35).
 * Adaptive waves in one thread block per pair (one warp per 5-point sample); pairs that need more than 24 iterations are
 * finished speculatively by the whole grid and replayed exactly.  max_iters <= 50000.  Asynchronous. */
int slamcu_sequence_essential(slamcu_sequence* seq, int first, int n_pairs, const double* K4, double prob,
                              double threshold, int max_iters);
/* Result of one pair: E9, inlier count, RANSAC iterations run, inlier mask over the pair's matches; synchronises. */
int slamcu_sequence_essential_read(slamcu_sequence* seq, int pair, double* E9, int* n_inliers, int* n_iters, uint8_t* mask,
                                   int capacity, int* n_points);

#ifdef __cplusplus
}
#endif
#endif /* SLAM_CUDA_SLAMCU_H_ */
