// describe_ref.cu -- reference-mode descriptor stage (rows A4, A5, A6 of SURVEY.md section 8a).
//
//   blur5_kernel    : FeatureDetector::gaussianBlur(image, 5, 1.0) (feature_detector.cpp:315-364).
//                     FP64, 25 products accumulated ky-outer / kx-inner exactly like the reference
//                     (this TU is compiled with -fmad=false, so no DFMA), std::round half-away,
//                     2-pixel frame copied from the input.  The 25 weights come from the host's
//                     std::exp (slamcu_detector_config::blur_weights).
//   describe_kernel : computeOrientation + computeBRIEFDescriptor (feature_detector.cpp:205-284),
//                     one warp per keypoint.  Moments are integer sums (exact, order-free), the angle
//                     goes through the glibc atan2f port, the rotation through the glibc sinf/cosf
//                     ports, samples are truncated toward zero, and out-of-image pairs are skipped
//                     *without consuming a bit index* (ballot + popc prefix).
#include "common.cuh"
#include "exact.cuh"

namespace slamcu {

namespace {

constexpr int BW = 64, BH = 16;           // output tile
constexpr int BSW = BW + 8;               // smem stride (x0-4 .. x0+68), word aligned
constexpr int BSH = BH + 4;

struct BlurW {
    double w[25];
    uint32_t q[25];  // weights in 8.24 fixed point (round to nearest); all zero = filter disabled
};

// exact value of one blurred pixel: the reference's double accumulation, ky outer / kx inner, no FMA, std::round
__device__ __forceinline__ uint32_t blur5_exact(const uint8_t* c, const BlurW& bw) {
    double acc = 0.0;
#pragma unroll
    for (int ky = -2; ky <= 2; ky++)
#pragma unroll
        for (int kx = -2; kx <= 2; kx++)
            acc = __dadd_rn(acc, __dmul_rn((double)c[ky * BSW + kx], bw.w[(ky + 2) * 5 + (kx + 2)]));
    return (uint32_t)(int)round(acc) & 0xffu;  // std::round: half away from zero (:351)
}

// Exactness filter.  The result is round(S) with S the reference's double sum; all that matters is on which side of
// k + 0.5 it lies.  An integer sum with the weights rounded to 2^-24 differs from the real-arithmetic sum by at most
// 255 * 25 * 2^-25 (3187.5 units of 2^-24), and the reference's own rounding errors are below 1e-11, so whenever the
// fixed-point fraction is farther than 4096 units from one half the rounded byte is already decided; the (0.05 % of)
// pixels inside that band take the exact FP64 path.  25 IMADs instead of 25 I2F + 25 DMUL + 25 DADD for the rest.
__global__ void __launch_bounds__(256) blur5_kernel(SeqView s, int first, BlurW bw) {
    __shared__ __align__(16) uint8_t tile[BSH * BSW];
    const int f = first + blockIdx.z;
    const int x0 = blockIdx.x * BW, y0 = blockIdx.y * BH;
    const uint8_t* img = s.img + (size_t)f * s.frame_bytes;
    uint8_t* out = s.blur + (size_t)f * s.frame_bytes;
    constexpr int WPR = BSW / 4;
    {
        uint32_t val[2];
#pragma unroll
        for (int k = 0; k < 2; k++) {
            const int v = threadIdx.x + 256 * k;
            const int r = v / WPR, cw = v - r * WPR;
            const int gy = y0 - 2 + r, gx = x0 - 4 + cw * 4;
            val[k] = 0;
            if (v < BSH * WPR && gy >= 0 && gy < s.rows && gx >= 0 && gx < s.pitch)
                val[k] = __ldg(reinterpret_cast<const uint32_t*>(img + (size_t)gy * s.pitch + gx));
        }
#pragma unroll
        for (int k = 0; k < 2; k++) {
            const int v = threadIdx.x + 256 * k;
            if (v < BSH * WPR) reinterpret_cast<uint32_t*>(tile)[v] = val[k];
        }
    }
    __syncthreads();
    // each thread produces 4 horizontally adjacent pixels -> one 32-bit store
    const int tx = (threadIdx.x & 15) * 4, ty = threadIdx.x >> 4;
    const int gy = y0 + ty;
    if (gy >= s.rows) return;
    const uint8_t* c0 = tile + (ty + 2) * BSW + 4 + tx;
    uint32_t sum[4] = {0, 0, 0, 0};
    const bool filter = bw.q[12] != 0;
    if (filter) {
#pragma unroll
        for (int ky = -2; ky <= 2; ky++) {
            uint32_t px[8];  // columns tx-2 .. tx+5 of this row
#pragma unroll
            for (int j = 0; j < 8; j++) px[j] = c0[ky * BSW + j - 2];
#pragma unroll
            for (int k = 0; k < 4; k++)
#pragma unroll
                for (int kx = 0; kx < 5; kx++) sum[k] += px[k + kx] * bw.q[(ky + 2) * 5 + kx];
        }
    }
    uint32_t packed = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int gx = x0 + tx + k;
        uint32_t v;
        if (gx >= 2 && gx < s.cols - 2 && gy >= 2 && gy < s.rows - 2) {
            const uint32_t frac = sum[k] & 0xffffffu;
            const uint32_t off = frac > 0x800000u ? frac - 0x800000u : 0x800000u - frac;  // distance from one half
            if (filter && off > 4096u) v = (sum[k] + 0x800000u) >> 24;
            else v = blur5_exact(c0 + k, bw);
        } else {
            v = c0[k];  // border frame copied from the input (:356-361)
        }
        packed |= v << (8 * k);
    }
    if (x0 + tx < s.pitch) *reinterpret_cast<uint32_t*>(out + (size_t)gy * s.pitch + x0 + tx) = packed;
}

__global__ void __launch_bounds__(128) describe_kernel(SeqView s, int first, int patch, int n_pattern,
                                                       const int* __restrict__ pattern) {
    __shared__ uint32_t sdesc[4][64];
    const int f = first + blockIdx.y;
    const int n = s.n_kp[f];
    const int warp = threadIdx.x >> 5;
    const unsigned lane = lane_id();
    const int kpi = blockIdx.x * 4 + warp;
    if (kpi >= n) return;
    slamcu_keypoint* kp = s.kps + (size_t)f * s.cap_kp + kpi;
    const uint8_t* img = s.blur + (size_t)f * s.frame_bytes;
    uint32_t* dout = s.desc + ((size_t)f * s.cap_kp + kpi) * s.desc_words;
    const int x = (int)kp->x, y = (int)kp->y;  // static_cast<int>(keypoint.x) (:207-208)
    const int r = patch / 2;
    uint32_t* sd = sdesc[warp];
    for (int w = lane; w < s.desc_words; w += 32) sd[w] = 0u;
    __syncwarp();
    const bool inside = !(x - r < 0 || x + r >= s.cols || y - r < 0 || y + r >= s.rows);
    float angle = 0.0f;
    if (inside) {
        // intensity-centroid moments over the disc u^2+v^2 <= r^2 (:216-227)
        float m01, m10;
        if (r <= 36) {
            // |partial sums| <= 255 * sum|u| < 2^24: float accumulation is exact => integer sums are too
            int sm01 = 0, sm10 = 0;
            for (int u = -r + (int)lane; u <= r; u += 32) {
                int col = 0, colv = 0;
                for (int v = -r; v <= r; v++) {
                    if (u * u + v * v <= r * r) {
                        const int p = img[(size_t)(y + v) * s.pitch + (x + u)];
                        col += p;
                        colv += v * p;
                    }
                }
                sm10 += u * col;
                sm01 += colv;
            }
#pragma unroll
            for (int o = 16; o; o >>= 1) {
                sm01 += __shfl_xor_sync(0xffffffffu, sm01, o);
                sm10 += __shfl_xor_sync(0xffffffffu, sm10, o);
            }
            m01 = (float)sm01;
            m10 = (float)sm10;
        } else {
            // large patches: partial sums may round; replay the reference's sequential float adds
            m01 = 0.0f;
            m10 = 0.0f;
            for (int v = -r; v <= r; v++)
                for (int u = -r; u <= r; u++)
                    if (u * u + v * v <= r * r) {
                        const float p = (float)img[(size_t)(y + v) * s.pitch + (x + u)];
                        m01 = m01 + (float)v * p;
                        m10 = m10 + (float)u * p;
                    }
        }
        const float kRad2Deg = (float)(180.0F / 3.14159265358979323846);
        angle = glibc_atan2f(m01, m10) * kRad2Deg;  // (:229)
    }
    if (lane == 0) kp->angle = angle;
    if (inside) {
        const float kDeg2Rad = (float)(3.14159265358979323846 / 180.0F);
        const float a = angle * kDeg2Rad;  // (:247)
        const float ca = glibc_cosf(a), sa = glibc_sinf(a);
        const int nbits = s.desc_bytes * 8;
        int bit_index = 0;
        for (int base = 0; base < n_pattern && bit_index < nbits; base += 32) {
            const int i = base + (int)lane;
            bool valid = false, less = false;
            if (i < n_pattern) {
                const int4 q = *reinterpret_cast<const int4*>(pattern + 4 * i);
                const float p1x = (float)q.x, p1y = (float)q.y, p2x = (float)q.z, p2y = (float)q.w;
                const int x1 = (int)((p1x * ca) - (p1y * sa)) + x;  // truncation toward zero (:263-266)
                const int y1 = (int)((p1x * sa) + (p1y * ca)) + y;
                const int x2 = (int)((p2x * ca) - (p2y * sa)) + x;
                const int y2 = (int)((p2x * sa) + (p2y * ca)) + y;
                valid = x1 >= 0 && x1 < s.cols && y1 >= 0 && y1 < s.rows && x2 >= 0 && x2 < s.cols && y2 >= 0 &&
                        y2 < s.rows;
                if (valid) less = img[(size_t)y1 * s.pitch + x1] < img[(size_t)y2 * s.pitch + x2];
            }
            const unsigned vm = __ballot_sync(0xffffffffu, valid);
            const int pos = bit_index + __popc(vm & lanemask_lt());
            if (valid && less && pos < nbits) atomicOr(&sd[pos >> 5], 1u << (pos & 31));
            bit_index += __popc(vm);
        }
    }
    __syncwarp();
    for (int w = lane; w < s.desc_words; w += 32) dout[w] = sd[w];
}

// OR of all descriptors of a frame (lets the matcher skip words that are zero everywhere).
__global__ void __launch_bounds__(256) desc_or_kernel(const uint32_t* __restrict__ desc, size_t set_stride,
                                                      const int* __restrict__ n_dev, int desc_words,
                                                      uint32_t* __restrict__ out) {
    const int f = blockIdx.x;
    const int n = n_dev[f];
    const uint32_t* d = desc + (size_t)f * set_stride;
    __shared__ uint32_t acc[64];
    if (threadIdx.x < 64) acc[threadIdx.x] = 0u;
    __syncthreads();
    const size_t total = (size_t)n * desc_words;
    for (size_t i = threadIdx.x; i < total; i += blockDim.x) {
        const uint32_t v = d[i];
        if (v) atomicOr(&acc[i % desc_words], v);
    }
    __syncthreads();
    if (threadIdx.x < desc_words) out[(size_t)f * desc_words + threadIdx.x] = acc[threadIdx.x];
}

}  // namespace

int launch_blur(const SeqView& s, int first, int n, const DetParams& p, cudaStream_t st) {
    BlurW bw;
    // fixed-point copy for the exactness filter: only for non-negative weights whose integer sum cannot overflow
    bool ok = true;
    unsigned long long total = 0;
    for (int i = 0; i < 25; i++) {
        bw.w[i] = p.blur_w[i];
        if (!(p.blur_w[i] >= 0.0 && p.blur_w[i] < 1.0)) ok = false;
        bw.q[i] = ok ? (uint32_t)llround(p.blur_w[i] * 16777216.0) : 0u;
        total += bw.q[i];
    }
    if (!ok || total * 255ull >= (1ull << 32) || bw.q[12] == 0)
        for (int i = 0; i < 25; i++) bw.q[i] = 0u;
    dim3 grid((s.cols + BW - 1) / BW, (s.rows + BH - 1) / BH, n);
    SLAM_KERNEL("blur5", st, blur5_kernel<<<grid, 256, 0, st>>>(s, first, bw));
    return 1;
}

int launch_describe(const SeqView& s, int first, int n, const DetParams& p, const int* d_pattern, cudaStream_t st) {
    dim3 grid((s.cap_kp + 3) / 4, n);
    SLAM_KERNEL("describe", st, describe_kernel<<<grid, 128, 0, st>>>(s, first, p.patch, p.n_pattern, d_pattern));
    SLAM_KERNEL("desc_or", st,
                desc_or_kernel<<<n, 256, 0, st>>>(s.desc + (size_t)first * s.cap_kp * s.desc_words,
                                                  (size_t)s.cap_kp * s.desc_words, s.n_kp + first, s.desc_words,
                                                  s.desc_or + (size_t)first * s.desc_words));
    return 2;
}

int launch_desc_or(const uint32_t* desc, const int* n_dev, int desc_words, uint32_t* out, cudaStream_t st, int sets,
                   size_t set_stride) {
    SLAM_KERNEL("desc_or", st, desc_or_kernel<<<sets, 256, 0, st>>>(desc, set_stride, n_dev, desc_words, out));
    return 1;
}

}  // namespace slamcu
