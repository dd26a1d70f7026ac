// fast_ref.cu -- reference-mode corner detection (rows A1, A2 of SURVEY.md section 8a).
//
//   fast_mask_kernel   : FeatureDetector::isFASTCorner over the whole frame (feature_detector.cpp:56-145)
//                        -> 1 bit / pixel corner mask, one 32-bit word per warp via __ballot_sync.
//   corner_list_kernel : turns the mask into the raster-ordered corner list the reference's
//                        detectFASTKeypoints() emits (feature_detector.cpp:61-67) and evaluates
//                        computeFASTScore (feature_detector.cpp:190-203) for every corner.
//
// Raster order matters: the reference feeds this list to an *unstable* std::sort, whose permutation
// is a function of the input order (sortnms.cu reproduces it).
#include "common.cuh"

namespace slamcu {

namespace {

constexpr int TW = 128;          // tile width  (pixels) = 4 mask words
constexpr int TH = 32;           // tile height (rows)
constexpr int HALO_X = 16;       // keeps 128-bit loads aligned
constexpr int SW = TW + 2 * HALO_X;  // smem row stride in bytes (160)
constexpr int SH = TH + 6;

// ring offsets (dx, dy): feature_detector.hpp:138-153, index 0 = (0,-3), clockwise in image coords
__constant__ int8_t c_ring_dx[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
__constant__ int8_t c_ring_dy[16] = {-3, -3, -2, -1, 0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3};

__device__ __forceinline__ bool segment_test(const uint8_t* t /* -> centre pixel in smem */, int thr, int arc) {
    const int c = t[0];
    const int up = c + thr, dn = c - thr;
    // cardinal pre-test, pixels 0 and 8 (feature_detector.cpp:81-96)
    const int p0 = t[-3 * SW], p8 = t[3 * SW];
    int hi = (p0 > up) + (p8 > up);
    int lo = (p0 < dn) + (p8 < dn);
    if ((hi | lo) == 0) return false;
    // pixels 4 and 12 (feature_detector.cpp:98-113)
    const int p4 = t[3], p12 = t[-3];
    hi += (p4 > up) + (p12 > up);
    lo += (p4 < dn) + (p12 < dn);
    if (hi < 3 && lo < 3) return false;
    if (arc <= 0) return true;  // `brighterPixels >= 0` holds at the first step of the scan (:138-141)
    // full segment test: 32 steps over the ring == circular run of >= arc equal-signed pixels (:121-142)
    unsigned mh = 0, ml = 0;
#pragma unroll
    for (int k = 0; k < 16; k++) {
        constexpr int dxs[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
        constexpr int dys[16] = {-3, -3, -2, -1, 0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3};
        const int v = t[dys[k] * SW + dxs[k]];
        mh |= (unsigned)(v > up) << k;
        ml |= (unsigned)(v < dn) << k;
    }
    mh |= mh << 16;
    ml |= ml << 16;
    unsigned rh = mh, rl = ml;
    for (int k = 1; k < arc; k++) {
        rh &= mh >> k;
        rl &= ml >> k;
    }
    return (rh | rl) != 0;
}

// bytes of |a - b| that exceed thr -> bit 7 of the byte (SWAR; thr in [0, 255]); k7 = (thr < 128 ? 127 - thr : 255 - thr) * 0x01010101
__device__ __forceinline__ unsigned swar_absdiff_gt(unsigned a, unsigned b, unsigned k7, bool big) {
    const unsigned d = __vabsdiffu4(a, b);
    const unsigned s = (d & 0x7f7f7f7fu) + k7;  // no carry between bytes: both addends are <= 127
    return big ? (s & d) : (s | d);
}

// Compacting phases, so that the (divergent) literal test runs on full warps:
//   1  SWAR filter on packed bytes, 4 pixels per thread: the reference's cardinal pre-test (feature_detector.cpp:81-113)
//      needs three of the four compass pixels beyond the threshold on one side, hence three of four with |diff| > thr;
//      this is implied by the literal pre-test for every ContiguousPixelsThreshold, so nothing is lost
//   2  the literal isFASTCorner on the survivors -> bits OR-ed into the tile's mask words in shared memory
__global__ void __launch_bounds__(256) fast_mask_kernel(SeqView s, int first, int thr, int arc) {
    __shared__ __align__(16) uint8_t tile[SH * SW];
    __shared__ uint16_t list[TH * TW];
    __shared__ unsigned mw[TH * (TW / 32)];
    __shared__ int n_list;
    const int f = first + blockIdx.z;
    const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH;
    const uint8_t* img = s.img + (size_t)f * s.frame_bytes;

    // stage the tile (+3 rows, +16 cols of halo) with aligned 128-bit loads, all of a thread's loads in flight at once
    constexpr int VPR = SW / 16;  // vectors per smem row
    {
        uint4 val[2];
#pragma unroll
        for (int k = 0; k < 2; k++) {
            const int v = threadIdx.x + 256 * k;
            const int r = v / VPR, cv = v - r * VPR;
            const int gy = y0 - 3 + r, gx = x0 - HALO_X + cv * 16;
            val[k] = make_uint4(0, 0, 0, 0);
            if (v < SH * VPR && gy >= 0 && gy < s.rows && gx >= 0 && gx < s.pitch)
                val[k] = __ldg(reinterpret_cast<const uint4*>(img + (size_t)gy * s.pitch + gx));
        }
#pragma unroll
        for (int k = 0; k < 2; k++) {
            const int v = threadIdx.x + 256 * k;
            if (v < SH * VPR) reinterpret_cast<uint4*>(tile)[v] = val[k];
        }
    }
    if (threadIdx.x < TH * (TW / 32)) mw[threadIdx.x] = 0;
    if (threadIdx.x == 0) n_list = 0;
    __syncthreads();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    {
        const bool big = thr >= 128;
        const unsigned k7 = (unsigned)(big ? 255 - thr : 127 - thr) * 0x01010101u;
        const int lo = 3, hi = s.cols - 3;  // isFASTCorner is evaluated for x in [3, cols - 3) (feature_detector.cpp:59-66)
        for (int r = warp; r < TH; r += 8) {
            const int gy = y0 + r;
            if (gy < 3 || gy >= s.rows - 3) continue;
            const uint32_t* row = reinterpret_cast<const uint32_t*>(tile + (r + 3) * SW) + HALO_X / 4 + lane;
            const unsigned w0 = row[0], wm = row[-1], wp = row[1];
            const unsigned wn = row[-3 * (SW / 4)], ws = row[3 * (SW / 4)];
            const unsigned we = __funnelshift_r(w0, wp, 24);  // pixels x+3 .. x+6
            const unsigned ww = __funnelshift_r(wm, w0, 8);   // pixels x-3 .. x
            const unsigned an = swar_absdiff_gt(wn, w0, k7, big), as = swar_absdiff_gt(ws, w0, k7, big);
            const unsigned ae = swar_absdiff_gt(we, w0, k7, big), aw = swar_absdiff_gt(ww, w0, k7, big);
            unsigned m = ((an & as & (ae | aw)) | (ae & aw & (an | as))) & 0x80808080u;
            const int gxb = x0 + 4 * lane;
            if (m != 0 && (gxb < lo || gxb + 4 > hi)) {
#pragma unroll
                for (int k = 0; k < 4; k++)
                    if ((unsigned)(gxb + k - lo) >= (unsigned)(hi - lo)) m &= ~(0x80u << (8 * k));
            }
            if (m != 0) {
                int pos = atomicAdd(&n_list, __popc(m));
                const int base = r * TW + 4 * lane;
#pragma unroll
                for (int k = 0; k < 4; k++)
                    if (m & (0x80u << (8 * k))) {
                        SLAMCU_BOUND(pos, TH * TW);
                        list[pos++] = (uint16_t)(base + k);
                    }
            }
        }
    }
    __syncthreads();
    const int cnt = n_list;
    for (int e = threadIdx.x; e < cnt; e += 256) {
        const int idx = list[e];
        const int r = idx / TW, lx = idx - r * TW;
        if (segment_test(tile + (r + 3) * SW + HALO_X + lx, thr, arc)) atomicOr(&mw[r * (TW / 32) + (lx >> 5)], 1u << (lx & 31));
    }
    __syncthreads();
    if (threadIdx.x < TH * (TW / 32)) {
        const int gy = y0 + (threadIdx.x >> 2), wi = (x0 >> 5) + (threadIdx.x & 3);
        if (gy < s.rows && wi < s.mwords) (s.mask + (size_t)f * s.rows * s.mwords)[(size_t)gy * s.mwords + wi] = mw[threadIdx.x];
    }
}

// One block per frame: mask -> raster-ordered list + SAD scores + sort keys.
__global__ void __launch_bounds__(256) corner_list_kernel(SeqView s, int first) {
    extern __shared__ int row_off[];  // [rows + 1]
    __shared__ int warp_tot[8];
    __shared__ int carry;
    const int f = first + blockIdx.x;
    const uint32_t* mask = s.mask + (size_t)f * s.rows * s.mwords;
    const uint8_t* img = s.img + (size_t)f * s.frame_bytes;
    uint32_t* xy = s.raw_xy + (size_t)f * s.cap_raw;
    uint32_t* keys = s.keys + (size_t)f * s.cap_raw;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // 1. per-row corner counts
    for (int r = warp; r < s.rows; r += 8) {
        int c = 0;
        for (int w = lane; w < s.mwords; w += 32) c += __popc(mask[(size_t)r * s.mwords + w]);
#pragma unroll
        for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
        if (lane == 0) row_off[r] = c;
    }
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    // 2. exclusive scan of the row counts (block-wide, 256 rows per round)
    for (int base = 0; base < s.rows; base += 256) {
        const int r = base + threadIdx.x;
        const int v = (r < s.rows) ? row_off[r] : 0;
        int inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) warp_tot[warp] = inc;
        __syncthreads();
        int woff = carry;
        for (int w = 0; w < warp; w++) woff += warp_tot[w];
        if (r < s.rows) row_off[r] = woff + inc - v;
        __syncthreads();
        if (threadIdx.x == 255) carry = woff + inc;
        __syncthreads();
    }
    const int total = carry;
    const int n = min(total, s.cap_raw);
    if (threadIdx.x == 0) {
        s.n_raw[f] = n;
        s.status[f] = (total > s.cap_raw) ? kStRawOverflow : 0;
    }
    // 3. expand bits in raster order
    for (int r = warp; r < s.rows; r += 8) {
        int base = row_off[r];
        for (int w0 = 0; w0 < s.mwords; w0 += 32) {
            const int w = w0 + lane;
            unsigned word = (w < s.mwords) ? mask[(size_t)r * s.mwords + w] : 0u;
            const int c = __popc(word);
            int inc = c;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += t;
            }
            int pos = base + inc - c;
            while (word) {
                const int b = __ffs(word) - 1;
                word &= word - 1;
                if (pos < s.cap_raw) xy[pos] = ((uint32_t)r << 16) | (uint32_t)(w * 32 + b);
                pos++;
            }
            base += __shfl_sync(0xffffffffu, inc, 31);
        }
    }
    __syncthreads();
    // 4. SAD score per corner (computeFASTScore), packed with the raster index into the sort key
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const uint32_t p = xy[i];
        const int x = p & 0xffff, y = p >> 16;
        const uint8_t* c = img + (size_t)y * s.pitch + x;
        const int cv = c[0];
        int score = 0;
#pragma unroll
        for (int k = 0; k < 16; k++) score += abs((int)c[c_ring_dy[k] * s.pitch + c_ring_dx[k]] - cv);
        keys[i] = ((uint32_t)score << kKeyIdxBits) | (uint32_t)i;
    }
}

// NMS disabled (or probe): keypoints straight from the raster list. response = score when `scored`
// (probe for the parity test) else 0 (feature_detector.cpp:13-15: response stays 0 without NMS).
__global__ void raster_keypoints_kernel(SeqView s, int first, int scored) {
    const int f = first + blockIdx.y;
    const int n = s.n_raw[f];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) {
        s.n_kp[f] = min(n, s.cap_kp);
        if (n > s.cap_kp) atomicOr(&s.status[f], kStKpOverflow);
    }
    if (i >= n || i >= s.cap_kp) return;
    const uint32_t p = s.raw_xy[(size_t)f * s.cap_raw + i];
    const uint32_t k = s.keys[(size_t)f * s.cap_raw + i];
    slamcu_keypoint kp;
    kp.x = (float)(p & 0xffff);
    kp.y = (float)(p >> 16);
    kp.size = 6.0f;
    kp.angle = 0.0f;
    kp.response = scored ? (float)(k >> kKeyIdxBits) : 0.0f;
    s.kps[(size_t)f * s.cap_kp + i] = kp;
}

}  // namespace

int launch_fast_corners(const SeqView& s, int first, int n, const DetParams& p, cudaStream_t st) {
    dim3 grid((s.cols + TW - 1) / TW, (s.rows + TH - 1) / TH, n);
    SLAM_KERNEL("fast_mask", st, fast_mask_kernel<<<grid, 256, 0, st>>>(s, first, p.thr, p.arc));
    SLAM_KERNEL("corner_list", st, corner_list_kernel<<<n, 256, (s.rows + 1) * sizeof(int), st>>>(s, first));
    return 2;
}

int launch_raster_keypoints(const SeqView& s, int first, int n, bool scored, cudaStream_t st) {
    dim3 grid((s.cap_kp + 255) / 256, n);
    SLAM_KERNEL("raster_keypoints", st, raster_keypoints_kernel<<<grid, 256, 0, st>>>(s, first, scored ? 1 : 0));
    return 1;
}

}  // namespace slamcu
