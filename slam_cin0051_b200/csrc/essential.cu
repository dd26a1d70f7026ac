// essential.cu -- cv::findEssentialMat(p1, p2, K, RANSAC, prob, threshold, maxIters) on the device
// (row R1 of SURVEY.md section 8a; the reference calls it at src/frontend/pose_estimator.cpp:42).
//
// The algorithm lives in OpenCV (calib3d five-point.cpp + ptsetreg.cpp, un-vendored, conanfile.txt:2).  What is
// reproduced EXACTLY: the cv::RNG multiply-with-carry stream seeded 0xFFFFFFFFFFFFFFFF, the 5-subset draw with
// redraw on duplicates, the Sampson-error inlier test (double arithmetic in Matx order, narrowed to float, compared
// with (float)t^2), the strictly-greater accept rule and RANSACUpdateNumIters.  The 5-point minimal solver is the
// same mathematics (Nister) with its own null-space basis and root finder: candidates agree with OpenCV's to
// rounding, their order inside one sample may differ (that only matters for exact ties at the maximum).
//
// essential_ransac_kernel: one block (RT threads) per frame pair, adaptive like the original loop but in waves that
// grow 16, 32, 64, 128, 256, 256, ... samples (most pairs finish inside the first wave: niters drops to ~10 after the
// first good model; a wave costs one solver latency whatever its size, so a pair that needs all 1000 iterations takes 7):
//   one thread per sample runs the minimal solver while an otherwise idle thread draws the NEXT wave's subsets from
//   the sequential RNG (drawing past the end of the loop is harmless: nothing consumes the stream afterwards);
//   one warp per hypothesis scores all correspondences (ballot + popc inlier count); thread 0 replays the
//   sequential accept / niters rule over the wave; the loop ends as soon as the replay reaches niters.
// Finally the block writes the winner's inlier mask.
#include "common.cuh"
#include "fivept.cuh"

namespace slamcu {
namespace {

constexpr int RT = 256;       // threads per block (scoring warps: RT / 32)
constexpr int WAVE = RT;      // most samples solved per wave (one thread each)
constexpr int WAVE0 = 16;     // first wave
constexpr int MAXM = kMaxModels;

struct CvRng {
    unsigned long long s;
    __device__ unsigned next() {
        s = (unsigned long long)(unsigned)s * 4164903690ULL + (s >> 32);
        return (unsigned)s;
    }
};

// EMEstimatorCallback::computeError for one correspondence; Matx products accumulate left to right (no FMA)
__device__ __forceinline__ bool sampson_inlier(const double* E, double ax, double ay, double bx, double by, float thr2) {
    const double e0 = (E[0] * ax + E[1] * ay) + E[2] * 1.0;
    const double e1 = (E[3] * ax + E[4] * ay) + E[5] * 1.0;
    const double e2 = (E[6] * ax + E[7] * ay) + E[8] * 1.0;
    const double t0 = (E[0] * bx + E[3] * by) + E[6] * 1.0;
    const double t1 = (E[1] * bx + E[4] * by) + E[7] * 1.0;
    const double dot = (bx * e0 + by * e1) + 1.0 * e2;
    const float err = (float)(dot * dot / (e0 * e0 + e1 * e1 + t0 * t0 + t1 * t1));
    return err <= thr2;
}

__device__ int update_num_iters(double p, double ep, int model_points, int max_iters) {  // RANSACUpdateNumIters
    p = fmin(fmax(p, 0.0), 1.0);
    ep = fmin(fmax(ep, 0.0), 1.0);
    double num = fmax(1.0 - p, 2.2250738585072014e-308);
    double denom = 1.0 - pow(1.0 - ep, (double)model_points);
    if (denom < 2.2250738585072014e-308) return 0;
    num = log(num);
    denom = log(denom);
    return (denom >= 0 || -num >= max_iters * (-denom)) ? max_iters : __double2int_rn(num / denom);
}

// getSubset for `count` samples: 5 distinct indices each, redraw on duplicates
__device__ void draw_subsets(CvRng& rng, int n, int count, int* sidx) {
    for (int s = 0; s < count; s++) {
        int* id = sidx + s * 5;
        for (int i = 0; i < 5;) {
            const int v = (int)(rng.next() % (unsigned)n);
            int j = 0;
            for (; j < i; j++)
                if (id[j] == v) break;
            if (j < i) continue;
            id[i++] = v;
        }
    }
}

__global__ void __launch_bounds__(RT, 2) essential_ransac_kernel(EssentialJob job) {
    __shared__ int nmod[WAVE];             // models per sample
    __shared__ int counts[WAVE * MAXM];    // inliers per hypothesis
    __shared__ int sidx[2 * WAVE * 5];     // subsets, double buffered
    double* models = job.models + (size_t)blockIdx.x * WAVE * MAXM * 9;  // [WAVE][MAXM][9] in HBM/L2
    __shared__ double bestE[9];
    __shared__ int sh_best, sh_niters, sh_done, sh_it;
    __shared__ CvRng rng;
    const int pair = blockIdx.x;
    const int n = job.n_pts[pair];
    const double2* x1 = job.x1 + (size_t)pair * job.pt_stride;
    const double2* x2 = job.x2 + (size_t)pair * job.pt_stride;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        sh_best = 0;
        sh_niters = job.max_iters;
        sh_done = (n < 6 || job.max_iters <= 0) ? 1 : 0;
        sh_it = 0;
        rng.s = 0xFFFFFFFFFFFFFFFFULL;
        for (int t = 0; t < 9; t++) bestE[t] = 0.0;
        if (!sh_done) draw_subsets(rng, n, WAVE0, sidx);
    }
    __syncthreads();
    int sched = WAVE0, buf = 0;
    while (!sh_done) {
        const int it0 = sh_it;
        const int wave = min(sched, sh_niters - it0);
        const int next_sched = min(2 * sched, WAVE);
        const int* cur = sidx + buf * WAVE * 5;
        if (tid < wave) {
            double a[5][2], b[5][2];
            for (int i = 0; i < 5; i++) {
                const double2 p = x1[cur[tid * 5 + i]], q = x2[cur[tid * 5 + i]];
                a[i][0] = p.x; a[i][1] = p.y; b[i][0] = q.x; b[i][1] = q.y;
            }
            nmod[tid] = five_point(a, b, models + (size_t)tid * MAXM * 9);
        }
        if (tid == RT - 1) {  // the last thread solves only in full waves; it draws the next wave's subsets afterwards
            CvRng r = rng;
            draw_subsets(r, n, next_sched, sidx + (buf ^ 1) * WAVE * 5);
            rng = r;
        }
        __syncthreads();
        // score: one warp per hypothesis
        for (int h = warp; h < wave * MAXM; h += RT / 32) {
            const int s = h / MAXM, k = h - s * MAXM;
            if (k >= nmod[s]) continue;
            const double* E = models + (size_t)h * 9;
            int good = 0;
            for (int base = 0; base < n; base += 32) {
                const int i = base + lane;
                bool in = false;
                if (i < n) {
                    const double2 p = x1[i], q = x2[i];
                    in = sampson_inlier(E, p.x, p.y, q.x, q.y, job.thr2);
                }
                good += __popc(__ballot_sync(0xffffffffu, in));
            }
            if (lane == 0) counts[h] = good;
        }
        __syncthreads();
        if (tid == 0) {  // sequential replay of RANSACPointSetRegistrator::run over this wave
            int it = it0, best = sh_best, niters = sh_niters;
            for (int s = 0; s < wave && it < niters; s++, it++) {
                for (int k = 0; k < nmod[s]; k++) {
                    const int good = counts[s * MAXM + k];
                    if (good > max(best, 4)) {
                        best = good;
                        for (int t = 0; t < 9; t++) bestE[t] = models[((size_t)s * MAXM + k) * 9 + t];
                        niters = update_num_iters(job.prob, (double)(n - good) / n, 5, niters);
                    }
                }
            }
            sh_best = best;
            sh_niters = niters;
            sh_it = it;
            sh_done = it >= niters ? 1 : 0;
        }
        sched = next_sched;
        buf ^= 1;
        __syncthreads();
    }
    // outputs: E (row-major, zeros if no model was accepted), inlier count, mask
    uint8_t* mask = job.mask + (size_t)pair * job.pt_stride;
    if (tid < 9) job.E[(size_t)pair * 9 + tid] = bestE[tid];
    if (tid == 0) {
        job.n_inliers[pair] = sh_best;
        job.n_iters[pair] = sh_it;
    }
    const bool have = sh_best > 0;
    for (int i = tid; i < n; i += RT) {
        bool in = false;
        if (have) {
            const double2 p = x1[i], q = x2[i];
            in = sampson_inlier(bestE, p.x, p.y, q.x, q.y, job.thr2);
        }
        mask[i] = in ? 1 : 0;
    }
}

// gather matched keypoint coordinates and normalise by K: x = ((double)px - cx) / fx  (findEssentialMat's preamble)
__global__ void __launch_bounds__(256) essential_gather_kernel(SeqView s, int first, EssentialJob job, double fx, double fy,
                                                               double cx, double cy) {
    const int pair = blockIdx.y;
    const int f = first + pair;
    const int n = min(s.n_match[f], job.pt_stride);
    if (blockIdx.x == 0 && threadIdx.x == 0) job.n_pts[pair] = n;
    const slamcu_dmatch* m = s.matches + (size_t)f * s.cap_kp;
    const slamcu_keypoint* k1 = s.kps + (size_t)f * s.cap_kp;
    const slamcu_keypoint* k2 = k1 + s.cap_kp;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const slamcu_keypoint a = k1[m[i].queryIdx], b = k2[m[i].trainIdx];
        job.x1[(size_t)pair * job.pt_stride + i] = make_double2(((double)a.x - cx) / fx, ((double)a.y - cy) / fy);
        job.x2[(size_t)pair * job.pt_stride + i] = make_double2(((double)b.x - cx) / fx, ((double)b.y - cy) / fy);
    }
}

__global__ void __launch_bounds__(256) essential_normalise_kernel(const float* p1, const float* p2, int n, EssentialJob job,
                                                                  double fx, double fy, double cx, double cy) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) job.n_pts[0] = n;
    if (i >= n) return;
    job.x1[i] = make_double2(((double)p1[2 * i] - cx) / fx, ((double)p1[2 * i + 1] - cy) / fy);
    job.x2[i] = make_double2(((double)p2[2 * i] - cx) / fx, ((double)p2[2 * i + 1] - cy) / fy);
}

__global__ void fivept_probe_kernel(const double* x1, const double* x2, int n_samples, double* models, int* counts) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_samples) return;
    double a[5][2], b[5][2];
    for (int i = 0; i < 5; i++) {
        a[i][0] = x1[(s * 5 + i) * 2]; a[i][1] = x1[(s * 5 + i) * 2 + 1];
        b[i][0] = x2[(s * 5 + i) * 2]; b[i][1] = x2[(s * 5 + i) * 2 + 1];
    }
    counts[s] = five_point(a, b, models + (size_t)s * MAXM * 9);
}


}  // namespace

void init_essential_attributes() {}
size_t essential_model_scratch_doubles() { return (size_t)WAVE * MAXM * 9; }

int launch_essential_gather(const SeqView& s, int first, int n_pairs, const EssentialJob& job, const double* K4, cudaStream_t st) {
    SLAM_KERNEL("essential_gather", st,
                essential_gather_kernel<<<dim3(4, n_pairs), 256, 0, st>>>(s, first, job, K4[0], K4[1], K4[2], K4[3]));
    return 1;
}

int launch_essential_normalise(const float* p1, const float* p2, int n, const EssentialJob& job, const double* K4, cudaStream_t st) {
    essential_normalise_kernel<<<(n + 255) / 256, 256, 0, st>>>(p1, p2, n, job, K4[0], K4[1], K4[2], K4[3]);
    return 1;
}

int launch_essential_ransac(const EssentialJob& job, int n_pairs, cudaStream_t st) {
    SLAM_KERNEL("essential_ransac", st, essential_ransac_kernel<<<n_pairs, RT, 0, st>>>(job));
    return 1;
}

int launch_fivept_probe(const double* x1, const double* x2, int n_samples, double* models, int* counts, cudaStream_t st) {
    fivept_probe_kernel<<<(n_samples + 31) / 32, 32, 0, st>>>(x1, x2, n_samples, models, counts);
    return 1;
}

}  // namespace slamcu

// ---- simpleRecoverPose (src/frontend/simple_pose_recover.cpp:35-97) as called by PoseEstimator::estimate ---------------
// E -> (R1, R2, +-t) through an SVD, then a cheirality vote: every correspondence is triangulated (4x4 DLT, smallest
// right singular vector) against each of the four candidates and the candidate with the most points in front of both
// cameras wins (first maximum).  Faithful to the reference including its quirk: the points are already normalised by K
// (rounded to float, pose_estimator.cpp:58-64) and are still projected with K * [R | t] (simple_pose_recover.cpp:61-65).
// cv::SVD is a one-sided Jacobi SVD; so is this one (sign / ordering conventions of the singular vectors do not change
// the set of four candidates nor the triangulated points).  Tolerance is stated in tests/test_gpu_essential.py.
namespace slamcu {
namespace {

// One-sided (Hestenes) Jacobi on the columns of an n x n matrix A (row-major, n <= 4); V accumulates the rotations.
// On return the columns of A are U * diag(w); w[k] = column norms.
template <int N>
__device__ void jacobi_svd(double (&A)[N][N], double (&V)[N][N], double (&w)[N]) {
    for (int i = 0; i < N; i++)
        for (int j = 0; j < N; j++) V[i][j] = i == j ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 30; sweep++) {
        bool changed = false;
        for (int p = 0; p < N - 1; p++)
            for (int q = p + 1; q < N; q++) {
                double a = 0, b = 0, g = 0;
                for (int k = 0; k < N; k++) { a += A[k][p] * A[k][p]; b += A[k][q] * A[k][q]; g += A[k][p] * A[k][q]; }
                if (fabs(g) <= 2.220446049250313e-16 * sqrt(a * b)) continue;
                changed = true;
                const double beta = a - b, gamma = hypot(2.0 * g, beta);
                double c, s;
                if (beta < 0) {
                    const double delta = (gamma - beta) * 0.5;
                    s = sqrt(delta / gamma);
                    c = g / (gamma * s);
                } else {
                    c = sqrt((gamma + beta) / (gamma * 2.0));
                    s = g / (gamma * c);
                }
                for (int k = 0; k < N; k++) {
                    const double t0 = c * A[k][p] + s * A[k][q], t1 = -s * A[k][p] + c * A[k][q];
                    A[k][p] = t0; A[k][q] = t1;
                    const double v0 = c * V[k][p] + s * V[k][q], v1 = -s * V[k][p] + c * V[k][q];
                    V[k][p] = v0; V[k][q] = v1;
                }
            }
        if (!changed) break;
    }
    for (int j = 0; j < N; j++) {
        double nn = 0;
        for (int k = 0; k < N; k++) nn += A[k][j] * A[k][j];
        w[j] = sqrt(nn);
    }
}

__device__ double det3(const double* m) {
    return m[0] * (m[4] * m[8] - m[5] * m[7]) - m[1] * (m[3] * m[8] - m[5] * m[6]) + m[2] * (m[3] * m[7] - m[4] * m[6]);
}

__global__ void __launch_bounds__(128) recover_pose_kernel(EssentialJob job, double fx, double fy, double cx, double cy,
                                                           double* Rout, double* tout, int* front_out) {
    __shared__ double P[4][12];   // K * [R | t] for the four candidates
    __shared__ double Rc[2][9], tc[3];
    __shared__ int front[4];
    const int pair = blockIdx.x;
    const int n = job.n_pts[pair];
    const double* E = job.E + (size_t)pair * 9;
    const int tid = threadIdx.x;
    if (tid < 4) front[tid] = 0;
    if (tid == 0) {
        double A[3][3], V[3][3], w[3];
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 3; j++) A[i][j] = E[3 * i + j];
        jacobi_svd<3>(A, V, w);
        // order the singular values descending (cv::SVD does); U columns = A columns / w
        int ord[3] = {0, 1, 2};
        for (int i = 0; i < 2; i++)
            for (int j = i + 1; j < 3; j++)
                if (w[ord[j]] > w[ord[i]]) { const int t = ord[i]; ord[i] = ord[j]; ord[j] = t; }
        double U[3][3], Vt[3][3];
        for (int c = 0; c < 3; c++) {
            const int s = ord[c];
            for (int k = 0; k < 3; k++) {
                U[k][c] = w[s] > 0 ? A[k][s] / w[s] : 0.0;
                Vt[c][k] = V[k][s];
            }
        }
        // the null direction has no defined U column from A / w: complete the basis with a cross product
        if (!(w[ord[2]] > 1e-12 * w[ord[0]])) {
            U[0][2] = U[1][0] * U[2][1] - U[2][0] * U[1][1];
            U[1][2] = U[2][0] * U[0][1] - U[0][0] * U[2][1];
            U[2][2] = U[0][0] * U[1][1] - U[1][0] * U[0][1];
        }
        const double W[3][3] = {{0, -1, 0}, {1, 0, 0}, {0, 0, 1}};
        for (int v = 0; v < 2; v++) {  // R1 = U W Vt, R2 = U W' Vt
            double UW[3][3];
            for (int i = 0; i < 3; i++)
                for (int j = 0; j < 3; j++) {
                    double acc = 0;
                    for (int k = 0; k < 3; k++) acc += U[i][k] * (v == 0 ? W[k][j] : W[j][k]);
                    UW[i][j] = acc;
                }
            for (int i = 0; i < 3; i++)
                for (int j = 0; j < 3; j++) {
                    double acc = 0;
                    for (int k = 0; k < 3; k++) acc += UW[i][k] * Vt[k][j];
                    Rc[v][3 * i + j] = acc;
                }
            if (det3(Rc[v]) < 0)
                for (int k = 0; k < 9; k++) Rc[v][k] = -Rc[v][k];
        }
        for (int k = 0; k < 3; k++) tc[k] = U[k][2];
        const double K[3][3] = {{fx, 0, cx}, {0, fy, cy}, {0, 0, 1}};
        for (int c = 0; c < 4; c++) {
            const double* R = Rc[c & 1];
            const double sgn = c < 2 ? 1.0 : -1.0;
            for (int i = 0; i < 3; i++)
                for (int j = 0; j < 4; j++) {
                    double acc = 0;
                    for (int k = 0; k < 3; k++) acc += K[i][k] * (j < 3 ? R[3 * k + j] : sgn * tc[k]);
                    P[c][4 * i + j] = acc;
                }
        }
    }
    __syncthreads();
    const double P0[12] = {fx, 0, cx, 0, 0, fy, cy, 0, 0, 0, 1, 0};  // K * [I | 0]
    const double2* x1 = job.x1 + (size_t)pair * job.pt_stride;
    const double2* x2 = job.x2 + (size_t)pair * job.pt_stride;
    int mine[4] = {0, 0, 0, 0};
    for (int j = tid; j < n; j += blockDim.x) {
        // Point2f((pt.x - cx) / fx, (pt.y - cy) / fy): the double quotient narrowed to float
        const double ax = (double)(float)x1[j].x, ay = (double)(float)x1[j].y;
        const double bx = (double)(float)x2[j].x, by = (double)(float)x2[j].y;
        for (int c = 0; c < 4; c++) {
            double A[4][4], V[4][4], w[4];
            for (int k = 0; k < 4; k++) {
                A[0][k] = ax * P0[8 + k] - P0[k];
                A[1][k] = ay * P0[8 + k] - P0[4 + k];
                A[2][k] = bx * P[c][8 + k] - P[c][k];
                A[3][k] = by * P[c][8 + k] - P[c][4 + k];
            }
            jacobi_svd<4>(A, V, w);
            int m = 0;
            for (int k = 1; k < 4; k++)
                if (w[k] < w[m]) m = k;
            const double X0 = V[0][m] / V[3][m], X1 = V[1][m] / V[3][m], X2 = V[2][m] / V[3][m];
            const double z1 = X2;
            const double z2 = ((P[c][8] * X0 + P[c][9] * X1) + P[c][10] * X2) + P[c][11] * 1.0;
            if (z1 > 0 && z2 > 0) mine[c]++;
        }
    }
    for (int c = 0; c < 4; c++)
        if (mine[c]) atomicAdd(&front[c], mine[c]);
    __syncthreads();
    if (tid == 0) {
        int best = 0, maxFront = -1;
        for (int c = 0; c < 4; c++)
            if (front[c] > maxFront) { maxFront = front[c]; best = c; }
        for (int k = 0; k < 9; k++) Rout[(size_t)pair * 9 + k] = Rc[best & 1][k];
        for (int k = 0; k < 3; k++) tout[(size_t)pair * 3 + k] = best < 2 ? tc[k] : -tc[k];
        for (int c = 0; c < 4; c++) front_out[(size_t)pair * 4 + c] = front[c];
    }
}

}  // namespace

int launch_recover_pose(const EssentialJob& job, int n_pairs, const double* K4, double* R, double* t, int* front, cudaStream_t st) {
    SLAM_KERNEL("recover_pose", st, recover_pose_kernel<<<n_pairs, 128, 0, st>>>(job, K4[0], K4[1], K4[2], K4[3], R, t, front));
    return 1;
}

}  // namespace slamcu
