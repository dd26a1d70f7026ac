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
// Two phases, both with ONE WARP PER SAMPLE running the minimal solver (fivept_warp.cuh):
//   essential_ransac_kernel   one block per frame pair, adaptive like the original loop in waves of 8 and 16 samples:
//       warps take samples from a shared counter (the last warp first draws the NEXT wave's subsets from the
//       sequential RNG), one warp per hypothesis scores all correspondences (ballot + popc inlier count), thread 0
//       replays the sequential accept / niters rule.  Most pairs end here (niters drops to ~10 after the first good
//       model).  A pair still running after 24 iterations draws the subsets of ALL remaining iterations and stops.
//   essential_spec_kernel     the whole grid works on the unfinished pairs: one warp solves one remaining sample and
//       scores its own hypotheses (speculative: iterations past the final niters are simply ignored later).
//   essential_finish_kernel   per unfinished pair: the sequential replay over the recorded inlier counts picks exactly
//       the model the original loop would have accepted last; one warp re-solves that sample; mask and E are written.
#include "common.cuh"
#include "fivept_warp.cuh"

namespace slamcu {
namespace {

constexpr int RT = 256;       // threads per block
constexpr int NW = RT / 32;   // warps: one sample / one hypothesis each at a time
constexpr int WAVE0 = 8;      // first wave of the adaptive phase
constexpr int WAVE = 16;      // its largest wave
constexpr int P1_ITERS = 24;  // the adaptive phase hands over to the speculative one after 8 + 16 iterations
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
    const double inv = 1.0 / (double)n;
    for (int s = 0; s < count; s++) {
        int* id = sidx + s * 5;
        for (int i = 0; i < 5;) {
            const unsigned x = rng.next();
            long long v = (long long)x - (long long)__double2uint_rz((double)x * inv) * n;  // x % n (quotient off by <= 1)
            if (v < 0) v += n;
            if (v >= n) v -= n;
            int j = 0;
            for (; j < i; j++)
                if (id[j] == (int)v) break;
            if (j < i) continue;
            id[i++] = (int)v;
        }
    }
}

// per-pair work area: [models of one wave][state][subsets][hypothesis counts][models per sample]
struct PairWork {
    double* models;          // [WAVE][MAXM][9]
    int* state;              // done, it, best, niters
    unsigned long long* rng; // RNG state after the adaptive phase
    int* subsets;            // [max_iters][5]   (entries from state.it on)
    int* cnt;                // [max_iters][MAXM]
    int* nmod;               // [max_iters]
};
__host__ __device__ inline size_t pair_work_bytes(int max_iters) {
    const size_t mi = (size_t)(max_iters > 0 ? max_iters : 0);
    return (size_t)WAVE * MAXM * 9 * 8 + 64 + mi * (5 + MAXM + 1) * 4 + 64;
}
__device__ inline PairWork pair_work(const EssentialJob& job, int pair) {
    unsigned char* base = job.work + (size_t)pair * pair_work_bytes(job.max_iters);
    const size_t mi = (size_t)max(job.max_iters, 0);
    PairWork w;
    w.models = reinterpret_cast<double*>(base);
    base += (size_t)WAVE * MAXM * 9 * 8;
    w.state = reinterpret_cast<int*>(base);
    w.rng = reinterpret_cast<unsigned long long*>(base + 32);
    base += 64;
    w.subsets = reinterpret_cast<int*>(base);
    w.cnt = w.subsets + mi * 5;
    w.nmod = w.cnt + mi * MAXM;
    return w;
}

// one warp counts the inliers of one hypothesis
__device__ __forceinline__ int score_model(const double* E9, const double2* x1, const double2* x2, int n, float thr2, int lane) {
    double E[9];
#pragma unroll
    for (int t = 0; t < 9; t++) E[t] = E9[t];
    int good = 0;
    for (int base = 0; base < n; base += 32) {
        const int i = base + lane;
        bool in = false;
        if (i < n) {
            const double2 p = x1[i], q = x2[i];
            in = sampson_inlier(E, p.x, p.y, q.x, q.y, thr2);
        }
        good += __popc(__ballot_sync(0xffffffffu, in));
    }
    return good;
}

__device__ void write_result(const EssentialJob& job, int pair, const double* bestE, int best, int it, int n, const double2* x1,
                             const double2* x2) {
    uint8_t* mask = job.mask + (size_t)pair * job.pt_stride;
    const int tid = threadIdx.x;
    if (tid < 9) job.E[(size_t)pair * 9 + tid] = bestE[tid];
    if (tid == 0) {
        job.n_inliers[pair] = best;
        job.n_iters[pair] = it;
    }
    const bool have = best > 0;
    for (int i = tid; i < n; i += blockDim.x) {
        bool in = false;
        if (have) {
            const double2 p = x1[i], q = x2[i];
            in = sampson_inlier(bestE, p.x, p.y, q.x, q.y, job.thr2);
        }
        mask[i] = in ? 1 : 0;
    }
}

__global__ void __launch_bounds__(RT, 2) essential_ransac_kernel(EssentialJob job) {
    extern __shared__ __align__(16) double fp_scratch[];  // [NW][kFiveptScratchDoubles]
    __shared__ int nmod[WAVE];             // models per sample
    __shared__ int counts[WAVE * MAXM];    // inliers per hypothesis
    __shared__ int sidx[2 * WAVE * 5];     // subsets, double buffered
    __shared__ double bestE[9];
    __shared__ int sh_best, sh_niters, sh_done, sh_it, sh_next;
    __shared__ CvRng rng;
    const int pair = blockIdx.x;
    const PairWork wk = pair_work(job, pair);
    double* models = wk.models;
    const int n = job.n_pts[pair];
    const double2* x1 = job.x1 + (size_t)pair * job.pt_stride;
    const double2* x2 = job.x2 + (size_t)pair * job.pt_stride;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        sh_best = 0;
        sh_niters = job.max_iters;
        sh_done = (n < 6 || job.max_iters <= 0) ? 1 : 0;
        sh_it = 0;
        sh_next = 0;
        rng.s = 0xFFFFFFFFFFFFFFFFULL;
        for (int t = 0; t < 9; t++) bestE[t] = 0.0;
        if (!sh_done) draw_subsets(rng, n, WAVE0, sidx);
    }
    __syncthreads();
    int sched = WAVE0, buf = 0;
    while (!sh_done && sh_it < P1_ITERS) {
        const int it0 = sh_it;
        const int wave = min(sched, sh_niters - it0);
        const int next_sched = min(2 * sched, WAVE);
        const int* cur = sidx + buf * WAVE * 5;
        if (warp == NW - 1 && it0 + wave < P1_ITERS) {  // the last warp joins the solve after drawing the next wave's subsets
            if (lane == 0) {
                CvRng r = rng;
                draw_subsets(r, n, next_sched, sidx + (buf ^ 1) * WAVE * 5);
                rng = r;
            }
            __syncwarp();
        }
        for (;;) {
            int s = 0;
            if (lane == 0) s = atomicAdd(&sh_next, 1);
            s = __shfl_sync(0xffffffffu, s, 0);
            if (s >= wave) break;
            const int cnt = five_point_warp(x1, x2, cur + s * 5, fp_scratch + warp * kFiveptScratchDoubles,
                                            models + (size_t)s * MAXM * 9);
            if (lane == 0) nmod[s] = cnt;
        }
        __syncthreads();
        for (int h = warp; h < wave * MAXM; h += NW) {  // score: one warp per hypothesis
            const int s = h / MAXM, k = h - s * MAXM;
            if (k >= nmod[s]) continue;
            const int good = score_model(models + (size_t)h * 9, x1, x2, n, job.thr2, lane);
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
            sh_next = 0;
        }
        sched = next_sched;
        buf ^= 1;
        __syncthreads();
    }
    if (sh_done) {
        if (tid == 0) wk.state[0] = 1;
        write_result(job, pair, bestE, sh_best, sh_it, n, x1, x2);
        return;
    }
    // hand over to the speculative phase: state + the subsets of every remaining iteration
    if (tid < 9) job.E[(size_t)pair * 9 + tid] = bestE[tid];
    if (tid == 0) {
        wk.state[0] = 0;
        wk.state[1] = sh_it;
        wk.state[2] = sh_best;
        wk.state[3] = sh_niters;
    }
    // getSubset for every remaining iteration.  Only the generator itself is sequential (~10 cycles per output), so
    // thread 0 fills a chunk of raw outputs, the block reduces them modulo n in parallel, and thread 0 strings the
    // values into 5-subsets (redraw on duplicates).  Outputs generated past the last subset are never consumed.
    unsigned* raw = reinterpret_cast<unsigned*>(fp_scratch);
    constexpr int CHUNK = NW * kFiveptScratchDoubles * 2;  // 32-bit words in the solver scratch
    __shared__ int sh_s, sh_i, sh_id[5];
    if (tid == 0) { sh_s = sh_it; sh_i = 0; }
    __syncthreads();
    const double inv = 1.0 / (double)n;
    while (sh_s < sh_niters) {
        const int want = min(CHUNK, (sh_niters - sh_s) * 5 + 64);
        if (tid == 0) {
            CvRng r = rng;
            for (int t = 0; t < want; t++) raw[t] = r.next();
            rng = r;
        }
        __syncthreads();
        for (int t = tid; t < want; t += RT) {
            const unsigned x = raw[t];
            long long v = (long long)x - (long long)__double2uint_rz((double)x * inv) * n;  // x % n (quotient off by <= 1)
            if (v < 0) v += n;
            if (v >= n) v -= n;
            raw[t] = (unsigned)v;
        }
        __syncthreads();
        if (tid == 0) {
            int s = sh_s, i = sh_i, id[5];
            for (int t = 0; t < 5; t++) id[t] = sh_id[t];
            const int last = sh_niters;
            for (int t = 0; t < want && s < last; t++) {
                const int v = (int)raw[t];
                bool dup = false;
                for (int j = 0; j < 5; j++) dup |= j < i && id[j] == v;
                if (dup) continue;
                id[i++] = v;
                if (i == 5) {
                    int* out = wk.subsets + (size_t)s * 5;
                    for (int j = 0; j < 5; j++) out[j] = id[j];
                    s++;
                    i = 0;
                }
            }
            sh_s = s;
            sh_i = i;
            for (int t = 0; t < 5; t++) sh_id[t] = id[t];
        }
        __syncthreads();
    }
    if (tid == 0) *wk.rng = rng.s;
}

// grid (ceil(max_iters / NW), pairs): warp -> one remaining sample of an unfinished pair
__global__ void __launch_bounds__(RT, 2) essential_spec_kernel(EssentialJob job) {
    extern __shared__ __align__(16) double fp_scratch[];  // [NW][kFiveptScratchDoubles + MAXM * 9]
    const int pair = blockIdx.y;
    const PairWork wk = pair_work(job, pair);
    if (wk.state[0]) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int s = wk.state[1] + blockIdx.x * NW + warp;
    if (s >= wk.state[3]) return;  // past niters as it stood after the adaptive phase (it can only shrink)
    const int n = job.n_pts[pair];
    const double2* x1 = job.x1 + (size_t)pair * job.pt_stride;
    const double2* x2 = job.x2 + (size_t)pair * job.pt_stride;
    double* S = fp_scratch + warp * (kFiveptScratchDoubles + MAXM * 9);
    double* models = S + kFiveptScratchDoubles;
    const int cnt = five_point_warp(x1, x2, wk.subsets + (size_t)s * 5, S, models);
    __syncwarp();
    for (int k = 0; k < cnt; k++) {
        const int good = score_model(models + k * 9, x1, x2, n, job.thr2, lane);
        if (lane == 0) wk.cnt[(size_t)s * MAXM + k] = good;
    }
    if (lane == 0) wk.nmod[s] = cnt;
}

__global__ void __launch_bounds__(RT) essential_finish_kernel(EssentialJob job) {
    extern __shared__ __align__(16) unsigned char fin_smem[];
    double* fp_scratch = reinterpret_cast<double*>(fin_smem);                          // [kFiveptScratchDoubles + MAXM * 9]
    int* mx = reinterpret_cast<int*>(fp_scratch + kFiveptScratchDoubles + MAXM * 9);   // [niters - it0] best count of the iteration
    __shared__ double bestE[9];
    __shared__ int sh_best, sh_it, sh_ws, sh_wk;
    const int pair = blockIdx.x;
    const PairWork wk = pair_work(job, pair);
    if (wk.state[0]) return;
    const int n = job.n_pts[pair];
    const double2* x1 = job.x1 + (size_t)pair * job.pt_stride;
    const double2* x2 = job.x2 + (size_t)pair * job.pt_stride;
    const int tid = threadIdx.x;
    const int it0 = wk.state[1], niters0 = wk.state[3];
    if (tid < 9) bestE[tid] = job.E[(size_t)pair * 9 + tid];
    // within one iteration the strictly-greater accept rule ends at the first hypothesis with the largest count, and
    // RANSACUpdateNumIters is monotone in the count, so the replay only needs (max count, its first index) per iteration
    for (int i = it0 + tid; i < niters0; i += RT) {
        const int nm = wk.nmod[i];
        int best = -1, bk = 0;
        for (int k = 0; k < nm; k++) {
            const int good = wk.cnt[(size_t)i * MAXM + k];
            if (good > best) { best = good; bk = k; }
        }
        mx[i - it0] = (best << 4) | bk;
    }
    __syncthreads();
    if (tid == 0) {
        int it = it0, best = wk.state[2], niters = niters0, ws = -1, wkk = 0;
        for (; it < niters; it++) {
            const int v = mx[it - it0], good = v >> 4;
            if (good > max(best, 4)) {
                best = good;
                ws = it;
                wkk = v & 15;
                niters = update_num_iters(job.prob, (double)(n - good) / n, 5, niters);
            }
        }
        sh_best = best;
        sh_it = it;
        sh_ws = ws;
        sh_wk = wkk;
    }
    __syncthreads();
    if (sh_ws >= 0 && tid < 32) {  // re-solve the winning sample (deterministic: the same models in the same order)
        double* models = fp_scratch + kFiveptScratchDoubles;
        five_point_warp(x1, x2, wk.subsets + (size_t)sh_ws * 5, fp_scratch, models);
        __syncwarp();
        if (tid < 9) bestE[tid] = models[sh_wk * 9 + tid];
    }
    __syncthreads();
    write_result(job, pair, bestE, sh_best, sh_it, n, x1, x2);
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

// slamcu_fivept_solve: x1, x2 = [n_samples][5][2] doubles; one warp per sample
__global__ void __launch_bounds__(128) fivept_probe_kernel(const double* x1, const double* x2, int n_samples, double* models, int* counts) {
    __shared__ __align__(16) double scratch[4 * kFiveptScratchDoubles];
    __shared__ int ident[5];
    if (threadIdx.x < 5) ident[threadIdx.x] = threadIdx.x;
    __syncthreads();
    const int s = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (s >= n_samples) return;
    const int cnt = five_point_warp(reinterpret_cast<const double2*>(x1) + (size_t)s * 5, reinterpret_cast<const double2*>(x2) + (size_t)s * 5,
                                    ident, scratch + (threadIdx.x >> 5) * kFiveptScratchDoubles, models + (size_t)s * MAXM * 9);
    if ((threadIdx.x & 31) == 0) counts[s] = cnt;
}

}  // namespace

void init_essential_attributes() {
    cudaFuncSetAttribute(essential_ransac_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, NW * kFiveptScratchDoubles * 8);
    cudaFuncSetAttribute(essential_spec_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, NW * (kFiveptScratchDoubles + MAXM * 9) * 8);
    cudaFuncSetAttribute(essential_finish_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         (kFiveptScratchDoubles + MAXM * 9) * 8 + kEssentialMaxIters * 4);
}
size_t essential_work_bytes_per_pair(int max_iters) { return pair_work_bytes(max_iters); }

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
    SLAM_KERNEL("essential_ransac", st, essential_ransac_kernel<<<n_pairs, RT, NW * kFiveptScratchDoubles * 8, st>>>(job));
    if (job.max_iters <= P1_ITERS) return 1;
    const int spec_blocks = (job.max_iters - P1_ITERS + NW - 1) / NW;  // an unfinished pair has done P1_ITERS iterations
    SLAM_KERNEL("essential_spec", st,
                essential_spec_kernel<<<dim3(spec_blocks, n_pairs), RT, NW * (kFiveptScratchDoubles + MAXM * 9) * 8, st>>>(job));
    const size_t fin_smem = (size_t)(kFiveptScratchDoubles + MAXM * 9) * 8 + (size_t)job.max_iters * 4;
    SLAM_KERNEL("essential_finish", st, essential_finish_kernel<<<n_pairs, RT, fin_smem, st>>>(job));
    return 3;
}

int launch_fivept_probe(const double* x1, const double* x2, int n_samples, double* models, int* counts, cudaStream_t st) {
    fivept_probe_kernel<<<(n_samples + 3) / 4, 128, 0, st>>>(x1, x2, n_samples, models, counts);
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


// ---- slam::triangulate (common.hpp:201-221) as PoseEstimator::triangulatePoints uses it (pose_estimator.cpp:69-104) ----------
// One thread per correspondence: the 4x4 DLT system from the two projection matrices and the pixel coordinates, its null
// vector by the same one-sided Jacobi SVD as the cheirality vote above (the reference: cv::SVD, vt.row(3)).
// The reference then copies the CV_64F solution into a column view of a CV_32F matrix with Mat::copyTo, which re-allocates
// the temporary header instead of writing the column (common.hpp:217) -- its output is indeterminate; this kernel returns
// what the code evidently means: the homogeneous solution, and x / x[3] (pose_estimator.cpp:95-100).
__global__ void __launch_bounds__(128) triangulate_kernel(const double* __restrict__ Pm, const float* __restrict__ p1, const float* __restrict__ p2,
                                                          int n, double* __restrict__ out4, double* __restrict__ out3) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const double ax = (double)p1[2 * j], ay = (double)p1[2 * j + 1], bx = (double)p2[2 * j], by = (double)p2[2 * j + 1];
    double A[4][4], V[4][4], w[4];
    for (int k = 0; k < 4; k++) {
        A[0][k] = ax * Pm[8 + k] - Pm[k];
        A[1][k] = ay * Pm[8 + k] - Pm[4 + k];
        A[2][k] = bx * Pm[12 + 8 + k] - Pm[12 + k];
        A[3][k] = by * Pm[12 + 8 + k] - Pm[12 + 4 + k];
    }
    jacobi_svd<4>(A, V, w);
    int m = 0;
    for (int k = 1; k < 4; k++)
        if (w[k] < w[m]) m = k;
    const double sgn = V[3][m] < 0 ? -1.0 : 1.0;  // unit-norm null vector, last component non-negative
    if (out4)
        for (int k = 0; k < 4; k++) out4[4 * (size_t)j + k] = sgn * V[k][m];
    if (out3)
        for (int k = 0; k < 3; k++) out3[3 * (size_t)j + k] = V[k][m] / V[3][m];
}

// ---- LoopClosure::verifyGeometricConsistency's RANSAC (loop_closure.cpp:177-222) ---------------------------------------------
// One block per hypothesis.  Thread 0 runs LoopClosure::solvePnP (:238-274) on the six sampled correspondences LITERALLY:
// the 12x12 DLT system (:246-253), its null vector p, P = Eigen::Map<Matrix<double,3,4>>(p.data()) -- a COLUMN-major view
// of a vector that was laid out row-major (:258), so R = [p0 p3 p6; p1 p4 p7; p2 p5 p8], t = (p9, p10, p11) --, then
// rotation = U diag(1, 1, det(U V')) V' of that R and translation = t / |R|_F (:263-270); K is never removed.
// The sign of a null vector is an implementation detail of the SVD (Eigen::JacobiSVD in the reference), and it matters
// here (it flips R and t, and with them det and the z > 0 test), so BOTH signs are solved and scored: counts[2 h + s].
// All threads then score every correspondence (:201-215): X' = R X + t, skipped when z <= 0, projected = K (X' / z),
// inlier iff |x - projected| < threshold; the inlier count is a ballot / popc reduction.
__global__ void __launch_bounds__(128) pnp_ransac_kernel(const double* __restrict__ X3, const double* __restrict__ x2, int n,
                                                         const int* __restrict__ samples6, const double* __restrict__ K9, double thr,
                                                         int* __restrict__ counts, double* __restrict__ Rt) {
    __shared__ double sR[2][9], st[2][3];
    __shared__ int cnt[2];
    const int h = blockIdx.x;
    if (threadIdx.x < 2) cnt[threadIdx.x] = 0;
    if (threadIdx.x == 0) {
        double A[12][12], V[12][12], w[12];
        for (int i = 0; i < 6; i++) {
            const int idx = samples6[6 * h + i];
            const double X = X3[3 * idx], Y = X3[3 * idx + 1], Z = X3[3 * idx + 2], u = x2[2 * idx], v = x2[2 * idx + 1];
            const double r0[12] = {X, Y, Z, 1, 0, 0, 0, 0, -u * X, -u * Y, -u * Z, -u};
            const double r1[12] = {0, 0, 0, 0, X, Y, Z, 1, -v * X, -v * Y, -v * Z, -v};
            for (int k = 0; k < 12; k++) { A[2 * i][k] = r0[k]; A[2 * i + 1][k] = r1[k]; }
        }
        jacobi_svd<12>(A, V, w);
        int m = 0;
        for (int k = 1; k < 12; k++)
            if (w[k] < w[m]) m = k;
        double p[12];
        for (int k = 0; k < 12; k++) p[k] = V[k][m];
        int big = 0;
        for (int k = 1; k < 12; k++)
            if (fabs(p[k]) > fabs(p[big])) big = k;
        const double s0 = p[big] < 0 ? -1.0 : 1.0;  // sign 0: the component of largest magnitude is positive
        for (int sgn = 0; sgn < 2; sgn++) {
            const double f = sgn == 0 ? s0 : -s0;
            double R[3][3], U[3][3], W[3][3], sw[3];
            double fro = 0;
            for (int r = 0; r < 3; r++)
                for (int c = 0; c < 3; c++) {
                    R[r][c] = f * p[3 * c + r];  // column-major Map of the row-major vector
                    fro += R[r][c] * R[r][c];
                }
            fro = sqrt(fro);
            for (int r = 0; r < 3; r++)
                for (int c = 0; c < 3; c++) U[r][c] = R[r][c];
            jacobi_svd<3>(U, W, sw);  // U <- U diag(sw) (columns), W = V
            int ord[3] = {0, 1, 2};
            for (int i = 0; i < 2; i++)
                for (int j = i + 1; j < 3; j++)
                    if (sw[ord[j]] > sw[ord[i]]) { const int t = ord[i]; ord[i] = ord[j]; ord[j] = t; }
            double Uo[3][3], Vo[3][3];
            for (int c = 0; c < 3; c++)
                for (int k = 0; k < 3; k++) {
                    Uo[k][c] = sw[ord[c]] > 0 ? U[k][ord[c]] / sw[ord[c]] : 0.0;
                    Vo[k][c] = W[k][ord[c]];
                }
            if (!(sw[ord[2]] > 1e-300)) {  // rank-deficient block: complete U with a cross product
                Uo[0][2] = Uo[1][0] * Uo[2][1] - Uo[2][0] * Uo[1][1];
                Uo[1][2] = Uo[2][0] * Uo[0][1] - Uo[0][0] * Uo[2][1];
                Uo[2][2] = Uo[0][0] * Uo[1][1] - Uo[1][0] * Uo[0][1];
            }
            double UVt[9];
            for (int r = 0; r < 3; r++)
                for (int c = 0; c < 3; c++) UVt[3 * r + c] = (Uo[r][0] * Vo[c][0] + Uo[r][1] * Vo[c][1]) + Uo[r][2] * Vo[c][2];
            const double det = det3(UVt);
            for (int r = 0; r < 3; r++)
                for (int c = 0; c < 3; c++) sR[sgn][3 * r + c] = (Uo[r][0] * Vo[c][0] + Uo[r][1] * Vo[c][1]) + det * Uo[r][2] * Vo[c][2];
            for (int k = 0; k < 3; k++) st[sgn][k] = f * p[9 + k] / fro;
        }
    }
    __syncthreads();
    const unsigned lane = lane_id();
    for (int sgn = 0; sgn < 2; sgn++) {
        const double* R = sR[sgn];
        const double* t = st[sgn];
        int good = 0;
        for (int base = 0; base < n; base += blockDim.x) {
            const int j = base + threadIdx.x;
            bool in = false;
            if (j < n) {
                const double X = X3[3 * j], Y = X3[3 * j + 1], Z = X3[3 * j + 2];
                const double tx = ((R[0] * X + R[1] * Y) + R[2] * Z) + t[0];
                const double ty = ((R[3] * X + R[4] * Y) + R[5] * Z) + t[1];
                const double tz = ((R[6] * X + R[7] * Y) + R[8] * Z) + t[2];
                if (tz > 0) {
                    const double nx = tx / tz, ny = ty / tz, nz = tz / tz;
                    const double px = (K9[0] * nx + K9[1] * ny) + K9[2] * nz;
                    const double py = (K9[3] * nx + K9[4] * ny) + K9[5] * nz;
                    const double dx = x2[2 * j] - px, dy = x2[2 * j + 1] - py;
                    in = sqrt(dx * dx + dy * dy) < thr;
                }
            }
            good += __popc(__ballot_sync(0xffffffffu, in));
        }
        if (lane == 0 && good) atomicAdd(&cnt[sgn], good);
    }
    __syncthreads();
    if (threadIdx.x < 2) counts[2 * h + threadIdx.x] = cnt[threadIdx.x];
    if (threadIdx.x < 24) {
        const int sgn = threadIdx.x / 12, k = threadIdx.x % 12;
        Rt[(size_t)(2 * h + sgn) * 12 + k] = k < 9 ? sR[sgn][k] : st[sgn][k - 9];
    }
}

}  // namespace

int launch_recover_pose(const EssentialJob& job, int n_pairs, const double* K4, double* R, double* t, int* front, cudaStream_t st) {
    SLAM_KERNEL("recover_pose", st, recover_pose_kernel<<<n_pairs, 128, 0, st>>>(job, K4[0], K4[1], K4[2], K4[3], R, t, front));
    return 1;
}

int launch_triangulate(const double* P24, const float* p1, const float* p2, int n, double* out4, double* out3, cudaStream_t st) {
    if (n <= 0) return 0;
    SLAM_KERNEL("triangulate", st, triangulate_kernel<<<(n + 127) / 128, 128, 0, st>>>(P24, p1, p2, n, out4, out3));
    return 1;
}

int launch_pnp_ransac(const double* X3, const double* x2, int n, const int* samples6, int n_hyp, const double* K9, double thr, int* counts,
                      double* Rt, cudaStream_t st) {
    if (n_hyp <= 0) return 0;
    SLAM_KERNEL("pnp_ransac", st, pnp_ransac_kernel<<<n_hyp, 128, 0, st>>>(X3, x2, n, samples6, K9, thr, counts, Rt));
    return 1;
}

}  // namespace slamcu
