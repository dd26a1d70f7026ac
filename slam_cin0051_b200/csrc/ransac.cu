// ransac.cu -- essential-matrix hypothesis scoring (row R1 of SURVEY.md section 8a).
//
// The reference calls cv::findEssentialMat(points1, points2, K, cv::RANSAC) (pose_estimator.cpp:42).
// OpenCV's scorer (EMEstimatorCallback::computeError, calib3d five-point.cpp, opencv 4.12 pinned in
// conanfile.txt:2 -- not vendored) evaluates, per correspondence, the Sampson-style error
//     err = (x2' E x1)^2 / (Ex1_0^2 + Ex1_1^2 + (E'x2)_0^2 + (E'x2)_1^2)   in double, narrowed to float,
// and counts inliers with err <= (float)threshold^2.  Matx products accumulate left to right from 0.
// This TU is compiled with -fmad=false: OpenCV's baseline build does not contract to FMA here.
//
// ransac_score_kernel: one warp per hypothesis, lanes stride the correspondences, inlier count via
// __popc(__ballot_sync); the per-model counts feed the sequential accept rule replayed on the host side
// of the ABI (or by ransac_replay in a later round).
#include "common.cuh"

namespace slamcu {
namespace {

__global__ void __launch_bounds__(128) ransac_score_kernel(const double* __restrict__ models9, int n_models,
                                                           const double* __restrict__ x1, const double* __restrict__ x2,
                                                           int n, float thr2, int* __restrict__ counts,
                                                           uint8_t* __restrict__ masks) {
    const int model = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (model >= n_models) return;
    const unsigned lane = lane_id();
    double E[9];
#pragma unroll
    for (int k = 0; k < 9; k++) E[k] = models9[(size_t)model * 9 + k];
    int good = 0;
    for (int base = 0; base < n; base += 32) {
        const int i = base + (int)lane;
        bool in = false;
        if (i < n) {
            const double ax = x1[2 * i], ay = x1[2 * i + 1];
            const double bx = x2[2 * i], by = x2[2 * i + 1];
            // Ex1 = E * (ax, ay, 1)
            const double e0 = (E[0] * ax + E[1] * ay) + E[2] * 1.0;
            const double e1 = (E[3] * ax + E[4] * ay) + E[5] * 1.0;
            const double e2 = (E[6] * ax + E[7] * ay) + E[8] * 1.0;
            // Etx2 = E' * (bx, by, 1)
            const double t0 = (E[0] * bx + E[3] * by) + E[6] * 1.0;
            const double t1 = (E[1] * bx + E[4] * by) + E[7] * 1.0;
            const double dot = (bx * e0 + by * e1) + 1.0 * e2;
            const double a = e0 * e0, b = e1 * e1, c = t0 * t0, d = t1 * t1;
            const float err = (float)(dot * dot / (a + b + c + d));
            in = err <= thr2;
            if (masks) masks[(size_t)model * n + i] = in ? 1 : 0;
        }
        good += __popc(__ballot_sync(0xffffffffu, in));
    }
    if (lane == 0) counts[model] = good;
}

}  // namespace

int launch_ransac_score(const double* models9, int n_models, const double* x1, const double* x2, int n, double thr2,
                        int* counts, uint8_t* masks, cudaStream_t st) {
    if (n_models <= 0) return 0;
    ransac_score_kernel<<<(n_models + 3) / 4, 128, 0, st>>>(models9, n_models, x1, x2, n, (float)thr2, counts, masks);
    return 1;
}

}  // namespace slamcu
