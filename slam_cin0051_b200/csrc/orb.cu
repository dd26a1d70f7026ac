// orb.cu -- OpenCV-ORB-compatible extractor ("mode B", rows B1-B8 of SURVEY.md section 8a).
//
// This is the algorithm BASELINE.json's headline config names (8-level pyramid, FAST-9 + Harris, 2000
// keypoint budget, rBRIEF-256).  It lives in OpenCV (features2d orb.cpp / fast.cpp, imgproc resize.cpp,
// filter.simd.hpp) -- an un-vendored dependency of the reference (conanfile.txt:2) -- so every kernel
// restates OpenCV's published arithmetic, and parity is pinned against cv2 itself
// (tests/test_orb_oracle.py, tests/test_gpu_orb_parity.py).
//
//   pyr_down_kernel     B1  INTER_LINEAR_EXACT: 8.8 fixed-point taps from host-built tables, level l from l-1
//   fast9_mask_kernel   B2  FAST-9/16 segment test + corner score + 3x3 strict NMS + 31-px border filter,
//                           smem-tiled, one mask word per warp via ballot
//   orb_select_kernel   B3  raster list, score histogram, retainBest(2*quota) threshold, ordered compaction
//   harris_kernel       B4  7x7 Harris response, warp per candidate, integer sums, float formula w/o FMA
//   orb_retain_kernel   B5  retainBest(quota) by rank counting (ties kept), ordered compaction
//   orb_assemble_kernel     level-major concatenation, pt = level coords * scale
//   blur7_kernel        B7  7x7 sigma=2 float separable blur with OpenCV's FMA placement, cvRound
//   orb_describe_kernel B6+B8  intensity-centroid angle (fastAtan2 polynomial, no FMA) + rBRIEF-256,
//                           warp per keypoint, one descriptor byte per lane
#include <cstdlib>

#include "orb.cuh"
#include <type_traits>

namespace slamcu {
namespace {

__device__ __forceinline__ const uint8_t* level_ptr(const SeqView& s, const OrbView& o, int f, int l) {
    return l == 0 ? s.img + (size_t)f * s.frame_bytes : o.pyr + (size_t)f * o.pyr_bytes + o.lv[l].off;
}
__device__ __forceinline__ uint8_t* level_ptr_w(const SeqView& s, const OrbView& o, int f, int l) {
    return l == 0 ? s.img + (size_t)f * s.frame_bytes : o.pyr + (size_t)f * o.pyr_bytes + o.lv[l].off;
}
__device__ __forceinline__ uint8_t* blur_ptr(const SeqView& s, const OrbView& o, int f, int l) {
    return l == 0 ? s.blur + (size_t)f * s.frame_bytes : o.pyrb + (size_t)f * o.pyr_bytes + o.lv[l].off;
}

// ---- B1 ------------------------------------------------------------------------------------------
// INTER_LINEAR_EXACT: out = ((256-ay) * ((256-ax) s00 + ax s01) + ay * ((256-ax) s10 + ax s11) + 32768) >> 16,
// written as a + w (b - a) on 8.8 / 16.16 integers (identical values: everything is exact in int32).
__global__ void __launch_bounds__(256) pyr_down_gather_kernel(SeqView s, OrbView o, int first, int l) {
    const int f = first + blockIdx.z;
    const OrbLevel& d = o.lv[l];
    const OrbLevel& p = o.lv[l - 1];
    const int x = blockIdx.x * 128 + (threadIdx.x & 31) * 4;  // 4 pixels per thread -> one 32-bit store
    const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (y >= d.rows || x >= d.pitch) return;
    const uint8_t* src = level_ptr(s, o, f, l - 1);
    uint8_t* dst = level_ptr_w(s, o, f, l);
    const uint32_t ty = __ldg(d.yt + y);
    const int sy0 = ty >> 8, ay = ty & 255, sy1 = min(sy0 + 1, p.rows - 1);
    const uint8_t* r0 = src + (size_t)sy0 * p.pitch;
    const uint8_t* r1 = src + (size_t)sy1 * p.pitch;
    uint32_t packed = 0;
    if (x + 3 < d.cols) {
        const uint4 tx = __ldg(reinterpret_cast<const uint4*>(d.xt + x));
        const uint32_t t4[4] = {tx.x, tx.y, tx.z, tx.w};
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int sx0 = t4[k] >> 8, ax = t4[k] & 255, sx1 = min(sx0 + 1, p.cols - 1);
            const int a0 = r0[sx0], b0 = r0[sx1], a1 = r1[sx0], b1 = r1[sx1];
            const int h0 = (a0 << 8) + ax * (b0 - a0);
            const int h1 = (a1 << 8) + ax * (b1 - a1);
            packed |= (uint32_t)(((h0 << 8) + ay * (h1 - h0) + 32768) >> 16) << (8 * k);
        }
    } else {
#pragma unroll
        for (int k = 0; k < 4; k++) {
            if (x + k >= d.cols) break;
            const uint32_t t1 = __ldg(d.xt + x + k);
            const int sx0 = t1 >> 8, ax = t1 & 255, sx1 = min(sx0 + 1, p.cols - 1);
            const int a0 = r0[sx0], b0 = r0[sx1], a1 = r1[sx0], b1 = r1[sx1];
            const int h0 = (a0 << 8) + ax * (b0 - a0);
            const int h1 = (a1 << 8) + ax * (b1 - a1);
            packed |= (uint32_t)(((h0 << 8) + ay * (h1 - h0) + 32768) >> 16) << (8 * k);
        }
    }
    *reinterpret_cast<uint32_t*>(dst + (size_t)y * d.pitch + x) = packed;
}


// The same arithmetic with the source staged through shared memory: one block produces a 128 x 64 tile of level l from
// the (128 s + 2) x (64 s + 2) source window of level l-1, loaded once with aligned 128-bit loads.  A thread walks
// 4 columns down PT_RPT output rows and keeps the horizontally interpolated source row it shares with the next output row
// (consecutive output rows are 1.2 source rows apart, so most rows cost one new source row, not two).
constexpr int PT_W = 128, PT_H = 64;        // output tile
constexpr int PT_RPT = 16;                   // output rows per thread (the per-thread set-up -- column selectors, table loads -- is paid once)
constexpr int PT_THREADS = 32 * (PT_H / PT_RPT);
constexpr int PS_PITCH = 192, PS_ROWS = 84;  // source window held in shared memory (bytes per row, rows)
__global__ void __launch_bounds__(PT_THREADS) pyr_down_kernel(SeqView s, OrbView o, int first, int l) {
    __shared__ __align__(16) uint8_t tile[PS_ROWS * PS_PITCH];
    const int f = first + blockIdx.z;
    const OrbLevel& d = o.lv[l];
    const OrbLevel& p = o.lv[l - 1];
    const int x0 = blockIdx.x * PT_W, y0 = blockIdx.y * PT_H;
    const int rows_here = min(PT_H, d.rows - y0);
    const uint8_t* src = level_ptr(s, o, f, l - 1);
    uint8_t* dst = level_ptr_w(s, o, f, l);
    const int x = x0 + (threadIdx.x & 31) * 4;
    const int yw = y0 + (threadIdx.x >> 5) * PT_RPT;  // first of this thread's PT_RPT output rows
    if (x0 >= d.cols) {  // a block in the row padding: zeros, like the gather kernel writes them
        if (x < d.pitch)
            for (int r = 0; r < PT_RPT && yw + r < d.rows; r++) *reinterpret_cast<uint32_t*>(dst + (size_t)(yw + r) * d.pitch + x) = 0u;
        return;
    }
    const int sy_lo = (int)(__ldg(d.yt + y0) >> 8);
    const int sy_hi = min((int)(__ldg(d.yt + y0 + rows_here - 1) >> 8) + 1, p.rows - 1);
    const int sx_lo = (int)(__ldg(d.xt + x0) >> 8) & ~15;
    const int sx_hi = min((int)(__ldg(d.xt + min(x0 + PT_W, d.cols) - 1) >> 8) + 1, p.cols - 1);
    const int nrows = sy_hi - sy_lo + 1, nvec = (sx_hi - sx_lo) / 16 + 1;  // host guarantees nrows <= PS_ROWS, nvec * 16 <= PS_PITCH
    {  // 16 threads per staged row (nvec <= 12 of them load a 128-bit vector), rows strided over the thread groups
        const int c = threadIdx.x & 15;
        if (c < nvec) {
            const bool inside = sx_lo + c * 16 < p.pitch;
            const uint8_t* g = src + (size_t)sy_lo * p.pitch + sx_lo + c * 16;
            for (int r = threadIdx.x >> 4; r < nrows; r += PT_THREADS / 16) {
                uint4 v = make_uint4(0, 0, 0, 0);
                if (inside) v = __ldg(reinterpret_cast<const uint4*>(g + (size_t)r * p.pitch));
                *reinterpret_cast<uint4*>(tile + r * PS_PITCH + c * 16) = v;
            }
        }
    }
    __syncthreads();
    if (x >= d.pitch || yw >= d.rows) return;
    // Per thread (4 output columns): the source bytes all four outputs touch lie within 8 bytes of the first one (3 * 2.0 + 1
    // at the steepest supported step), so a source row costs 3 aligned LDS.32 from the thread's own window, 2 funnel shifts to
    // byte-align it, and per output one PRMT (picks the pair a, b) + one IDP.2A ((256 - ax) a + ax b, exact in 16 bits).
    unsigned wx[4], sel[4];
    int nvalid = 0, o0 = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const bool ok = x + k < d.cols;
        const uint32_t t = ok ? __ldg(d.xt + x + k) : 0u;
        const int sx0 = ok ? (int)(t >> 8) : sx_lo + o0;
        const int a_off = sx0 - sx_lo, b_off = min(sx0 + 1, p.cols - 1) - sx_lo;
        if (k == 0) o0 = a_off;
        const unsigned axk = t & 255u;
        wx[k] = (256u - axk) | (axk << 16);                          // IDP.2A weights: (256 - ax, ax)
        sel[k] = (unsigned)(a_off - o0) | ((unsigned)(b_off - o0) << 4);           // PRMT selector: byte 0 <- a, byte 1 <- b
        nvalid += ok;
    }
    const int wbase = o0 & ~3, sh = 8 * (o0 & 3);
    uint32_t tys[PT_RPT];  // the row-table entries up front: independent loads instead of one dependent load per row
#pragma unroll
    for (int r = 0; r < PT_RPT; r++) tys[r] = __ldg(d.yt + min(yw + r, d.rows - 1));
    const int prows1 = p.rows - 1, dpitch = d.pitch, nrow = min(PT_RPT, d.rows - yw);
    uint8_t* out = dst + (size_t)yw * dpitch + x;
    auto hrow4 = [&](int sy, unsigned (&h)[4]) {  // the horizontally interpolated source row sy at the 4 output columns
        const uint32_t* w = reinterpret_cast<const uint32_t*>(tile + (sy - sy_lo) * PS_PITCH + wbase);
        const unsigned w0 = w[0], w1 = w[1], w2 = w[2];
        const unsigned lo = __funnelshift_r(w0, w1, sh), hi = __funnelshift_r(w1, w2, sh);
#pragma unroll
        for (int k = 0; k < 4; k++) h[k] = __dp2a_lo(wx[k], __byte_perm(lo, hi, sel[k]), 0u);  // IDP.2A.LO reads bytes 0, 1 of the pair only
    };
    int hrow = -1;  // source row whose horizontal interpolation is cached in hc
    unsigned hc[4] = {0, 0, 0, 0};
#pragma unroll
    for (int r = 0; r < PT_RPT; r++) {
        if (r >= nrow) break;
        const int sy0 = (int)(tys[r] >> 8), sy1 = min(sy0 + 1, prows1);
        const unsigned ay = tys[r] & 255u;
        if (sy0 != hrow) hrow4(sy0, hc);
        unsigned h1[4], v[4];
        hrow4(sy1, h1);
#pragma unroll
        for (int k = 0; k < 4; k++) {
            // ((h0 << 8) + ay (h1 - h0) + 32768) >> 16, < 2^24: the result is byte 2 of the sum
            v[k] = (hc[k] << 8) + 32768u + ay * (h1[k] - hc[k]);
            hc[k] = h1[k];
        }
        hrow = sy1;
        unsigned packed = __byte_perm(__byte_perm(v[0], v[1], 0x4462u), __byte_perm(v[2], v[3], 0x4462u), 0x5410u);
        if (nvalid < 4) packed &= nvalid == 0 ? 0u : (0xffffffffu >> (8 * (4 - nvalid)));
        *reinterpret_cast<uint32_t*>(out) = packed;
        out += dpitch;
    }
}

// ---- B2 ------------------------------------------------------------------------------------------
// FAST-9/16 + cornerScore + 3x3 NMS + border filter, one 128x48 output tile per block, in compacting phases so
// that every phase runs with full warps (the one-thread-per-pixel form spent half its issue slots diverged):
//   0  stage the tile (+16 px / 4 row halo) in shared memory with aligned 128-bit loads; clear the score grid
//   1  SWAR pre-test, 4 pixels per thread on packed bytes: a 9-arc contains two ring pixels 90 degrees apart
//      (one of N/S and one of E/W) that differ from the centre by more than t -> survivors to list 1
//   2  segment test + cornerScore on list 1 in one pass over the ring (fast9_test_and_score) -> dense score grid
//      (1 px ring around the tile for the NMS), corners to list 2
//   4  3x3 strict NMS + edgeThreshold border filter for the list-2 entries inside the tile -> bit mask
constexpr int FTW = 128, FTH = 48, FHX = 16, FSW = FTW + 2 * FHX, FSH = FTH + 8;  // pixel tile with halo
constexpr int SCW = FTW + 8, SCH = FTH + 2;  // score grid: x0-4 .. x0+131 (4-px groups), y0-1 .. y0+32

// cornerScore of cv::FAST (9/16) = max over the 16 arcs of 9 ring pixels of min |v - p| (one sign), minus 1.
// With d = v - p:  min over an arc of d = v - max(p),  min of -d = min(p) - v, so the score needs the sliding-window
// (length 9, circular) min or max of the raw ring pixels: windows of 3, then of 9 (fast9_test_and_score below).
// The subtraction from v is applied AFTER the min/max network on purpose: nvcc 12.9 folds max(a, -b) chains into
// VIMNMX3 for sm_100a and loses the negation ("toolchain finding").

// ring of radius 3, OpenCV order; only the set of arcs matters
__device__ __forceinline__ void fast9_load_ring(const uint8_t* t, int stride, int (&p)[16]) {
    constexpr int dxs[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
    constexpr int dys[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};
#pragma unroll
    for (int k = 0; k < 16; k++) p[k] = t[dys[k] * stride + dxs[k]];
}

// Segment test AND corner score in one pass over the ring: with m9 = max over the 16 arcs of (min of the arc's 9 pixels) and
// M9 = min over the arcs of (max of the arc's 9 pixels), the pixel is a bright corner iff m9 > v + thr, a dark one iff
// M9 < v - thr (both cannot hold: two disjoint 9-arcs do not fit in 16), and cornerScore = m9 - v - 1 resp. v - M9 - 1.
// Both polarities share ONE network: every ring pixel is widened to the halfword pair (p, 255 - p) by a single IMAD, so the
// sliding minimum of the low halves is min(p) and that of the high halves is 255 - max(p); 16 + 16 + 8 VIMNMX3.U16x2.
// Returns the score, or -1 when the pixel is no corner.
__device__ __forceinline__ int fast9_test_and_score(int v, const int (&p)[16], int thr) {
    unsigned q[16], lo3[16];
#pragma unroll
    for (int i = 0; i < 16; i++) q[i] = (unsigned)p[i] * 0xFFFF0001u + 0x00FF0000u;  // p | (255 - p) << 16
#pragma unroll
    for (int i = 0; i < 16; i++) lo3[i] = __vimin3_u16x2(q[i], q[(i + 1) & 15], q[(i + 2) & 15]);
    unsigned m = 0;
#pragma unroll
    for (int i = 0; i < 16; i += 2) {
        const unsigned a0 = __vimin3_u16x2(lo3[i], lo3[(i + 3) & 15], lo3[(i + 6) & 15]);
        const unsigned a1 = __vimin3_u16x2(lo3[i + 1], lo3[(i + 4) & 15], lo3[(i + 7) & 15]);
        m = __vimax3_u16x2(m, a0, a1);
    }
    const int bright = (int)(m & 0xffffu) - v;      // m9 - v
    const int dark = (int)(m >> 16) - (255 - v);    // (255 - M9) - (255 - v) = v - M9
    const int best = max(bright, dark);             // > thr <=> corner of that polarity
    return best > thr ? best - 1 : -1;
}

// bytes of |a - b| that exceed thr -> bit 7 of the byte (SWAR; thr in [0, 255]).  thr < 128: bit 7 of d + (127 - thr), or of d
// itself.  The addition runs over the whole word: a byte with d >= 129 + thr carries into its neighbour, which can only turn
// that neighbour's "d == thr" into a hit (a wrap of the neighbour needs d >= 128, which the OR catches).  The pre-test may
// pass extra pixels, never drop one: the exact test follows.  thr >= 128 (never the default): exact form, bit 7 of both.
// `one` is 1 in a register the compiler cannot see through: d * one + k7 stays an IMAD (FMA pipe) instead of an ALU-pipe add,
// and the ALU pipe is this kernel's limiter.
__device__ __forceinline__ unsigned swar_absdiff_gt(unsigned a, unsigned b, unsigned k7, bool big, unsigned one) {
    const unsigned d = __vabsdiffu4(a, b);
    unsigned sum;
    if (!big) {
        asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(sum) : "r"(d), "r"(one), "r"(k7));
        return sum | d;
    }
    return ((d & 0x7f7f7f7fu) + k7) & d;
}

// ---- TMA helpers (sm_90+ PTX): one thread arms an mbarrier with the box size and issues cp.async.bulk.tensor; the
// hardware zero-fills the part of the box that lies outside the tensor, so the staging loop and its bounds tests go away
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, int x, int y, int z, uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(smem_u32(dst)),
        "l"(map), "r"(x), "r"(y), "r"(z), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(phase)
        : "memory");
}

template <bool kTma>
__global__ void __launch_bounds__(256) fast9_mask_kernel(SeqView s, OrbView o, int first, int l, const __grid_constant__ CUtensorMap tmap) {
    __shared__ __align__(128) uint8_t tile[FSH * FSW];
    __shared__ __align__(8) uint64_t tma_bar;
    __shared__ __align__(16) uint8_t sc[SCH * SCW];
    __shared__ uint16_t list1[SCH * SCW];
    __shared__ uint16_t list2[SCH * SCW];
    __shared__ unsigned mw[FTH * (FTW / 32)];
    __shared__ int n1, n2;
    const int f = first + blockIdx.z;
    const OrbLevel& L = o.lv[l];
    const int x0 = blockIdx.x * FTW, y0 = blockIdx.y * FTH;
    const int thr = o.fast_threshold;
    const uint8_t* img = level_ptr(s, o, f, l);
    // ---- phase 0
    if (kTma) {
        if (threadIdx.x == 0) mbar_init(&tma_bar, 1);
        __syncthreads();
        if (threadIdx.x == 0) tma_load_3d(tile, &tmap, x0 - FHX, y0 - 4, f, &tma_bar, FSH * FSW);
        for (int v = threadIdx.x; v < SCH * SCW / 16; v += 256) reinterpret_cast<uint4*>(sc)[v] = make_uint4(0, 0, 0, 0);
    } else {
        constexpr int VPR = FSW / 16;
        uint4 val[2];  // both 128-bit loads of a thread are in flight before the first store
#pragma unroll
        for (int k = 0; k < 2; k++) {
            const int v = threadIdx.x + 256 * k;
            const int r = v / VPR, cv = v - r * VPR;
            const int gy = y0 - 4 + r, gx = x0 - FHX + cv * 16;
            val[k] = make_uint4(0, 0, 0, 0);
            if (v < FSH * VPR && gy >= 0 && gy < L.rows && gx >= 0 && gx < L.pitch)
                val[k] = __ldg(reinterpret_cast<const uint4*>(img + (size_t)gy * L.pitch + gx));
        }
        for (int v = threadIdx.x; v < SCH * SCW / 16; v += 256) reinterpret_cast<uint4*>(sc)[v] = make_uint4(0, 0, 0, 0);
#pragma unroll
        for (int k = 0; k < 2; k++) {
            const int v = threadIdx.x + 256 * k;
            if (v < FSH * VPR) reinterpret_cast<uint4*>(tile)[v] = val[k];
        }
    }
    if (threadIdx.x < FTH * (FTW / 32)) mw[threadIdx.x] = 0;
    if (threadIdx.x == 0) { n1 = 0; n2 = 0; }
    if (kTma) mbar_wait(&tma_bar, 0);
    __syncthreads();
    // ---- phase 1: a warp per score-grid row, a lane per 4-pixel group of the tile's 128 columns; the two ring
    // columns (x0-1, x0+128) skip the pre-test and go straight to list 1
    // valid score-grid columns: the FAST domain [3, cols-3) intersected with [x0-1, x0+129)
    const int lo = max(3, x0 - 1), hi = min(L.cols - 3, x0 + FTW + 1);
    auto pretest = [&](auto bigc) {  // the threshold class is block-uniform: two loops, not predicated twins in one
        constexpr bool big = decltype(bigc)::value;
        const unsigned k7 = (unsigned)(big ? 255 - thr : 127 - thr) * 0x01010101u;
        const unsigned one = (unsigned)min(o.nlevels, 1);
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        for (int ry = warp; ry < SCH; ry += 8) {
            const int gy = y0 - 1 + ry;
            if ((unsigned)(gy - 3) >= (unsigned)(L.rows - 6)) continue;
            const uint32_t* row = reinterpret_cast<const uint32_t*>(tile + (ry + 3) * FSW) + 4 + lane;
            const unsigned w0 = row[0], wm = row[-1], wp = row[1];
            const unsigned wn = row[-3 * (FSW / 4)], ws = row[3 * (FSW / 4)];
            const unsigned we = __funnelshift_r(w0, wp, 24);  // pixels x+3 .. x+6
            const unsigned ww = __funnelshift_r(wm, w0, 8);   // pixels x-3 .. x
            unsigned m = (swar_absdiff_gt(wn, w0, k7, big, one) | swar_absdiff_gt(ws, w0, k7, big, one)) &
                         (swar_absdiff_gt(we, w0, k7, big, one) | swar_absdiff_gt(ww, w0, k7, big, one)) & 0x80808080u;
            const int gxb = x0 + 4 * lane;
            if (m != 0 && (gxb < lo || gxb + 4 > hi)) {
#pragma unroll
                for (int k = 0; k < 4; k++)
                    if ((unsigned)(gxb + k - lo) >= (unsigned)(hi - lo)) m &= ~(0x80u << (8 * k));
            }
            if (m != 0) {
                int pos = atomicAdd(&n1, __popc(m));
                const int base = ry * SCW + 4 + 4 * lane;
#pragma unroll
                for (int k = 0; k < 4; k++)
                    if (m & (0x80u << (8 * k))) {
                        SLAMCU_BOUND(pos, SCH * SCW);
                        list1[pos++] = (uint16_t)(base + k);
                    }
            }
        }
    };
    {
        if (thr >= 128) pretest(std::true_type{});
        else pretest(std::false_type{});
        if (threadIdx.x < 2 * SCH) {  // the two ring columns of every score-grid row
            const int ry = threadIdx.x >> 1, gy = y0 - 1 + ry;
            const int cx = (threadIdx.x & 1) ? FTW + 4 : 3;  // score-grid columns of x0+128 and x0-1
            const int gx = x0 - 4 + cx;
            if ((unsigned)(gy - 3) < (unsigned)(L.rows - 6) && gx >= lo && gx < hi)
            {
                const int pos = atomicAdd(&n1, 1);
                SLAMCU_BOUND(pos, SCH * SCW);
                list1[pos] = (uint16_t)(ry * SCW + cx);
            }
        }
    }
    __syncthreads();
    // ---- phase 2 + 3: segment test and corner score in one pass over the ring (fast9_test_and_score) -> score grid, list 2
    const int c1 = n1;
    for (int e = threadIdx.x; e < c1; e += 256) {
        const int idx = list1[e];
        const int ry = idx / SCW, cx = idx - ry * SCW;
        const uint8_t* t = tile + (ry + 3) * FSW + (FHX - 4) + cx;
        SLAMCU_BOUND((ry + 3 - 3) * FSW + (FHX - 4) + cx - 3, FSH * FSW);
        SLAMCU_BOUND((ry + 3 + 3) * FSW + (FHX - 4) + cx + 3, FSH * FSW);
        int p[16];
        fast9_load_ring(t, FSW, p);
        const int score = fast9_test_and_score(t[0], p, thr);
        if (score >= 0) {
            sc[idx] = (uint8_t)score;
            const int pos = atomicAdd(&n2, 1);
            SLAMCU_BOUND(pos, SCH * SCW);
            SLAMCU_BOUND(idx, SCH * SCW);
            list2[pos] = (uint16_t)idx;
        }
    }
    __syncthreads();
    const int c2 = n2;
    // ---- phase 4: NMS + border filter
    for (int e = threadIdx.x; e < c2; e += 256) {
        const int idx = list2[e];
        const int ry = idx / SCW, cx = idx - ry * SCW;
        const int r = ry - 1, lx = cx - 4;
        if ((unsigned)r >= (unsigned)FTH || (unsigned)lx >= (unsigned)FTW) continue;
        const int gx = x0 + lx, gy = y0 + r;
        if (gx < kOrbEdge || gx >= L.cols - kOrbEdge || gy < kOrbEdge || gy >= L.rows - kOrbEdge) continue;
        const uint8_t* c = sc + idx;
        const int v = c[0];
        if (v > c[-1] && v > c[1] && v > c[-SCW - 1] && v > c[-SCW] && v > c[-SCW + 1] && v > c[SCW - 1] && v > c[SCW] &&
            v > c[SCW + 1]) {
            atomicOr(&mw[r * (FTW / 32) + (lx >> 5)], 1u << (lx & 31));
            (o.fscore + (size_t)f * o.score_bytes + L.soff)[(size_t)gy * L.pitch + gx] = (uint8_t)v;  // read back by orb_select
        }
    }
    __syncthreads();
    if (threadIdx.x < FTH * (FTW / 32)) {
        const int gy = y0 + (threadIdx.x >> 2), wi = (x0 >> 5) + (threadIdx.x & 3);
        if (gy < L.rows && wi < L.mwords)
            (o.mask + (size_t)f * o.mask_words + L.moff)[(size_t)gy * L.mwords + wi] = mw[threadIdx.x];
    }
}

// ---- B3: one block per (level, frame) ---------------------------------------------------------------
__global__ void __launch_bounds__(256) orb_select_kernel(SeqView s, OrbView o, int first) {
    extern __shared__ int row_off[];  // [rows + 1]
    __shared__ int warp_tot[8];
    __shared__ int carry;
    __shared__ int hist[256];
    __shared__ int thr_sh;
    const int l = blockIdx.x, f = first + blockIdx.y;
    const OrbLevel& L = o.lv[l];
    const uint32_t* mask = o.mask + (size_t)f * o.mask_words + L.moff;
    const uint8_t* fsc = o.fscore + (size_t)f * o.score_bytes + L.soff;
    uint32_t* cxy = o.cxy + (size_t)f * o.cand_total + L.coff;
    int* csc = o.cscore + (size_t)f * o.cand_total + L.coff;
    uint32_t* sxy = o.sxy + (size_t)f * o.cand_total + L.coff;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    hist[threadIdx.x] = 0;
    for (int r = warp; r < L.rows; r += 8) {
        int c = 0;
        for (int w = lane; w < L.mwords; w += 32) c += __popc(mask[(size_t)r * L.mwords + w]);
#pragma unroll
        for (int q = 16; q; q >>= 1) c += __shfl_xor_sync(0xffffffffu, c, q);
        if (lane == 0) row_off[r] = c;
    }
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < L.rows; base += 256) {
        const int r = base + threadIdx.x;
        const int v = (r < L.rows) ? row_off[r] : 0;
        int inc = v;
#pragma unroll
        for (int q = 1; q < 32; q <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, inc, q);
            if (lane >= q) inc += t;
        }
        if (lane == 31) warp_tot[warp] = inc;
        __syncthreads();
        int woff = carry;
        for (int w = 0; w < warp; w++) woff += warp_tot[w];
        if (r < L.rows) row_off[r] = woff + inc - v;
        __syncthreads();
        if (threadIdx.x == 255) carry = woff + inc;
        __syncthreads();
    }
    const int total = carry;
    const int n = min(total, L.capc);
    if (threadIdx.x == 0) {
        o.n_cand[f * kMaxLevels + l] = n;
        if (total > L.capc) atomicOr(&s.status[f], kStRawOverflow);
    }
    for (int r = warp; r < L.rows; r += 8) {
        int base = row_off[r];
        for (int w0 = 0; w0 < L.mwords; w0 += 32) {
            const int w = w0 + lane;
            unsigned word = (w < L.mwords) ? mask[(size_t)r * L.mwords + w] : 0u;
            const int c = __popc(word);
            int inc = c;
#pragma unroll
            for (int q = 1; q < 32; q <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, inc, q);
                if (lane >= q) inc += t;
            }
            int pos = base + inc - c;
            while (word) {
                const int b = __ffs(word) - 1;
                word &= word - 1;
                if (pos < L.capc) cxy[pos] = ((uint32_t)r << 16) | (uint32_t)(w * 32 + b);
                pos++;
            }
            base += __shfl_sync(0xffffffffu, inc, 31);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const uint32_t p = cxy[i];
        const int sc = fsc[(size_t)(p >> 16) * L.pitch + (p & 0xffff)];  // written by fast9_mask_kernel for every survivor
        csc[i] = sc;
        atomicAdd(&hist[min(max(sc, 0), 255)], 1);
    }
    __syncthreads();
    // KeyPointsFilter::retainBest(2 * quota): keep everything >= the (2q)-th largest score
    const int want = 2 * L.quota;
    if (threadIdx.x == 0) {
        int thr = 0;
        if (want <= 0) thr = 1 << 30;  // n_points == 0 -> clear
        else if (n > want) {
            int acc = 0;
            for (int b = 255; b >= 0; b--) {
                acc += hist[b];
                if (acc >= want) { thr = b; break; }
            }
        }
        thr_sh = thr;
        carry = 0;
    }
    __syncthreads();
    const int thr = thr_sh;
    for (int base = 0; base < n; base += 256) {
        const int i = base + threadIdx.x;
        const bool keep = i < n && csc[i] >= thr;
        const unsigned b = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) warp_tot[warp] = __popc(b);
        __syncthreads();
        int off = carry;
        for (int w = 0; w < warp; w++) off += warp_tot[w];
        if (keep) {
            SLAMCU_BOUND(off + __popc(b & lanemask_lt()), L.capc);
            sxy[off + __popc(b & lanemask_lt())] = cxy[i];
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            int t = 0;
            for (int w = 0; w < 8; w++) t += warp_tot[w];
            carry += t;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) o.n_sel[f * kMaxLevels + l] = carry;
}

// ---- B4: thread per candidate ---------------------------------------------------------------------------
// HarrisResponses (orb.cpp): over the 7x7 block, Ix / Iy are 3x3 Sobel sums; a = sum Ix^2, b = sum Iy^2, c = sum IxIy
// in int, then the float formula without FMA.  One thread walks its own 9x9 patch with a three-row sliding
// window held in registers: per row the horizontal differences dx[c] = p[c+1] - p[c-1] and the smoothed values
// sx[c] = p[c-1] + 2 p[c] + p[c+1], so Ix = dx(r-1) + 2 dx(r) + dx(r+1) and Iy = sx(r+1) - sx(r-1).
__global__ void __launch_bounds__(128) harris_kernel(SeqView s, OrbView o, int first) {
    const int l = blockIdx.y, f = first + blockIdx.z;
    const OrbLevel& L = o.lv[l];
    const int n = o.n_sel[f * kMaxLevels + l];
    const uint8_t* img = level_ptr(s, o, f, l);
    const float scale = 1.f / ((1 << 2) * 7 * 255.f);
    const float scale4 = scale * scale * scale * scale;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t p = o.sxy[(size_t)f * o.cand_total + L.coff + i];
        const int x = p & 0xffff, y = p >> 16;
        const uint8_t* q = img + (size_t)(y - 4) * L.pitch + (x - 4);
        int dx[3][7], sx[3][7];
        int a = 0, b = 0, c = 0;
#pragma unroll
        for (int r = 0; r < 9; r++) {
            int px[9];
#pragma unroll
            for (int k = 0; k < 9; k++) px[k] = q[k];
            q += L.pitch;
#pragma unroll
            for (int k = 0; k < 7; k++) {
                dx[r % 3][k] = px[k + 2] - px[k];
                sx[r % 3][k] = px[k] + 2 * px[k + 1] + px[k + 2];
            }
            if (r >= 2) {  // rows r-2, r-1, r are complete: emit the block row centred on r-1
#pragma unroll
                for (int k = 0; k < 7; k++) {
                    const int Ix = dx[(r - 2) % 3][k] + 2 * dx[(r - 1) % 3][k] + dx[r % 3][k];
                    const int Iy = sx[r % 3][k] - sx[(r - 2) % 3][k];
                    a += Ix * Ix;
                    b += Iy * Iy;
                    c += Ix * Iy;
                }
            }
        }
        const float fa = (float)a, fb = (float)b, fc = (float)c;
        // ((float)a * b - (float)c * c - harris_k * ((float)a + b) * ((float)a + b)) * scale_sq_sq   (no FMA)
        o.sresp[(size_t)f * o.cand_total + L.coff + i] = ((fa * fb - fc * fc) - (0.04f * (fa + fb)) * (fa + fb)) * scale4;
    }
}

// ---- B5: one block per (level, frame) --------------------------------------------------------------------
// KeyPointsFilter::retainBest(quota): keep everything >= the quota-th largest response (ties kept).  The threshold is
// found by an MSB-first radix select over the order-preserving integer image of the float responses (four 8-bit
// passes with a shared-memory histogram) instead of counting, for every element, how many are larger.
__device__ __forceinline__ uint32_t float_order_key(float v) {
    uint32_t b = __float_as_uint(v);
    if (b == 0x80000000u) b = 0u;  // -0.0 compares equal to +0.0
    return b ^ ((b >> 31) ? 0xffffffffu : 0x80000000u);
}

__global__ void __launch_bounds__(256) orb_retain_kernel(SeqView s, OrbView o, int first, int smem_cap) {
    __shared__ int hist[256];
    __shared__ int warp_tot[8];
    __shared__ int carry;
    __shared__ uint32_t sh_prefix;
    __shared__ int sh_want;
    const int l = blockIdx.x, f = first + blockIdx.y;
    const OrbLevel& L = o.lv[l];
    const int n = o.n_sel[f * kMaxLevels + l];
    const size_t base_off = (size_t)f * o.cand_total + L.coff;
    const float* r = o.sresp + base_off;
    const uint32_t* sxy = o.sxy + base_off;
    uint32_t* fxy = o.fxy + base_off;
    float* fr = o.fresp + base_off;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int want = L.quota;
    (void)smem_cap;
    uint32_t thr_key = 0u;  // keep everything
    if (want <= 0) thr_key = 0xffffffffu;  // n_points == 0 -> nothing survives (keys never reach all ones: no NaNs)
    else if (n > want) {
        if (threadIdx.x == 0) { sh_prefix = 0u; sh_want = want; }
        for (int pass = 0; pass < 4; pass++) {
            const int shift = 24 - 8 * pass;
            hist[threadIdx.x] = 0;
            __syncthreads();
            const uint32_t prefix = sh_prefix;
            const uint32_t himask = pass == 0 ? 0u : (0xffffffffu << (shift + 8));
            for (int i = threadIdx.x; i < n; i += 256) {
                const uint32_t k = float_order_key(r[i]);
                if ((k & himask) == prefix) atomicAdd(&hist[(k >> shift) & 255], 1);
            }
            __syncthreads();
            if (threadIdx.x == 0) {  // the digit in which the want-th largest key falls
                int acc = 0, d = 255;
                const int w = sh_want;
                for (; d > 0; d--) {
                    if (acc + hist[d] >= w) break;
                    acc += hist[d];
                }
                sh_want = w - acc;
                sh_prefix = prefix | ((uint32_t)d << shift);
            }
            __syncthreads();
        }
        thr_key = sh_prefix;
    }
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < n; base += 256) {
        const int i = base + threadIdx.x;
        const float v = i < n ? r[i] : 0.f;
        const bool keep = i < n && thr_key != 0xffffffffu && float_order_key(v) >= thr_key;
        const unsigned b = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) warp_tot[warp] = __popc(b);
        __syncthreads();
        int off = carry;
        for (int w = 0; w < warp; w++) off += warp_tot[w];
        if (keep) {
            const int pos = off + __popc(b & lanemask_lt());
            SLAMCU_BOUND(pos, L.capc);
            fxy[pos] = sxy[i];
            fr[pos] = v;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            int t = 0;
            for (int w = 0; w < 8; w++) t += warp_tot[w];
            carry += t;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) o.n_fin[f * kMaxLevels + l] = carry;
}

// ---- concatenate levels: one block per frame -----------------------------------------------------------------
__global__ void __launch_bounds__(256) orb_assemble_kernel(SeqView s, OrbView o, int first) {
    __shared__ int off[kMaxLevels + 1];
    const int f = first + blockIdx.x;
    if (threadIdx.x == 0) {
        int acc = 0;
        for (int l = 0; l < o.nlevels; l++) {
            off[l] = acc;
            acc += o.n_fin[f * kMaxLevels + l];
        }
        off[o.nlevels] = acc;
        s.n_kp[f] = min(acc, s.cap_kp);
        s.n_raw[f] = 0;
        if (acc > s.cap_kp) atomicOr(&s.status[f], kStKpOverflow);
    }
    __syncthreads();
    slamcu_keypoint* kps = s.kps + (size_t)f * s.cap_kp;
    for (int l = 0; l < o.nlevels; l++) {
        const OrbLevel& L = o.lv[l];
        const int n = off[l + 1] - off[l];
        const size_t base_off = (size_t)f * o.cand_total + L.coff;
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const int pos = off[l] + i;
            if (pos >= s.cap_kp) break;
            const uint32_t p = o.fxy[base_off + i];
            slamcu_keypoint k;
            k.x = (float)(p & 0xffff) * L.scale;   // keypoint.pt *= scale (float)
            k.y = (float)(p >> 16) * L.scale;
            k.size = (float)kOrbPatch * L.scale;
            k.angle = 0.f;
            k.response = o.fresp[base_off + i];
            SLAMCU_BOUND(pos, s.cap_kp);
            kps[pos] = k;
            o.octave[(size_t)f * s.cap_kp + pos] = l;
            o.lxy[(size_t)f * s.cap_kp + pos] = p;
        }
    }
    if (threadIdx.x == 0) {
        int raw = 0;
        for (int l = 0; l < o.nlevels; l++) raw += o.n_cand[f * kMaxLevels + l];
        s.n_raw[f] = raw;
    }
}

// ---- B7 ----------------------------------------------------------------------------------------------------
// 7x7 sigma=2 separable float blur with OpenCV's rounding: row pass s = k0*S0, s = fma(k_j, S_j, s) (j = 1..6),
// column pass s = k3*R0, s = fma(k_{3+j}, R_{+j} + R_{-j}, s) (j = 1..3), cvRound (half to even), saturate.
//
// 128 x 64 output tile per block of 4 warps.  The tile (+4 px / 3 row halo, reflected at the image border) is staged ONCE as
// floats in shared memory (PRMT + FADD: 0x4B000000 | byte is the float 2^23 + byte).  Then every lane walks a 4-pixel-wide
// column strip down 16 output rows with the last seven row-pass results in registers (the loop is fully unrolled, so the
// window rotates by renaming): per input row 3 LDS.128 + 28 FMA, per output row 28 column-pass operations, 4 conversions
// and one 32-bit store.  No second shared-memory buffer, no barrier after staging, and almost no address arithmetic -- the
// previous two-pass version (row results through shared memory) executed 50 instructions per pixel for 20 of arithmetic.
constexpr int BW = 128, BRS = 15, BWARPS = 4, BTH = BRS * BWARPS;  // tile width, rows per warp strip, warps, tile height (60)
constexpr int BIW = BW + 8, BIH = BTH + 6;                          // staged floats per row (x0-4 .. x0+BW+3), staged rows (66)
static_assert((BRS + 6) % 7 == 0, "the strip loop is unrolled by the window depth");
__device__ __forceinline__ int reflect101(int i, int n) {
    if (i < 0) i = -i;
    if (i >= n) i = 2 * n - 2 - i;
    return min(max(i, 0), n - 1);
}
__device__ __forceinline__ float byte_to_float(unsigned w, unsigned sel) {
    return __uint_as_float(__byte_perm(w, 0x4B000000u, sel)) - 8388608.0f;
}

// cvRound (round half to even) + saturate_cast<uchar> of a blurred pixel.  The value is a convex combination of bytes (the float
// weights sum to 1 + a few ulp), so it lies in [0, 255.0001]: adding 2^23 rounds it to an integer exactly like cvt.rni and leaves
// that integer in the low byte -- an FADD on the FMA pipe instead of an F2I on the XU pipe; no saturation can trigger.
__device__ __forceinline__ unsigned sat_u8_rn(float v) { return __float_as_uint(v + 8388608.0f); }

constexpr int BBW = kBlurBoxW, BBH = kBlurBoxH, BBX = 16;  // byte tile: pixels x0-16 .. x0+143 (160 B rows), rows y0-3 .. y0+62
static_assert(BBW == BW + 2 * BBX && BBH == BIH, "blur tile geometry");  // (the TMA start coordinate stays a multiple of 16 bytes)

constexpr int kBlurTileBytes = (BBH * BBW + 127) / 128 * 128;  // every TMA destination stays 128-byte aligned
constexpr size_t kBlurSmemBytes = 2 * kBlurTileBytes + BIH * BIW * sizeof(float) + 16;
static_assert((BIH * BIW * sizeof(float)) % 8 == 0, "blur shared-memory carve-up");
// One block walks `tiles_per_block` vertically adjacent tiles of one tile column: while it converts and filters tile t, the
// TMA engine is already writing tile t+1 into the other byte buffer (two mbarriers, one per buffer), so the global-memory
// latency of the staging is off the critical path and the block's set-up cost is paid once per column.
template <bool kTma>
__global__ void __launch_bounds__(BWARPS * 32) blur7_kernel(SeqView s, OrbView o, int first, int l, int tiles_per_block,
                                                            const __grid_constant__ CUtensorMap tmap) {
    extern __shared__ __align__(128) uint8_t blur_smem[];  // kBlurSmemBytes: two byte tiles, the float tile, two mbarriers
    uint8_t(*tb)[kBlurTileBytes] = reinterpret_cast<uint8_t(*)[kBlurTileBytes]>(blur_smem);
    float* tin = reinterpret_cast<float*>(blur_smem + 2 * kBlurTileBytes);
    uint64_t* tma_bar = reinterpret_cast<uint64_t*>(blur_smem + 2 * kBlurTileBytes + BIH * BIW * sizeof(float));
    const int f = first + blockIdx.z;
    const OrbLevel& L = o.lv[l];
    const uint8_t* img = level_ptr(s, o, f, l);
    uint8_t* out = blur_ptr(s, o, f, l);
    const int x0 = blockIdx.x * BW;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int tiles_y = (L.rows + BTH - 1) / BTH;
    const int t_begin = blockIdx.y * tiles_per_block, t_end = min(tiles_y, t_begin + tiles_per_block);
    if (t_begin >= t_end) return;
    // getGaussianKernel(7, 2, CV_32F) (bit patterns of the cv2 result; OpenCV computes exp(-x^2/8) normalised)
    const float k0 = __uint_as_float(0x3d8fafb1u), k1 = __uint_as_float(0x3e06387eu), k2 = __uint_as_float(0x3e434a39u),
                k3 = __uint_as_float(0x3e5d4ae0u);
    if (kTma) {
        if (threadIdx.x == 0) {
            mbar_init(&tma_bar[0], 1);
            mbar_init(&tma_bar[1], 1);
        }
        __syncthreads();
        if (threadIdx.x == 0) tma_load_3d(tb[0], &tmap, x0 - BBX, t_begin * BTH - 3, f, &tma_bar[0], BBH * BBW);
    }
    constexpr int WPR = BIW / 4;  // 34 float4 per staged row
    const int x = x0 + 4 * lane;
    const int pitch = L.pitch;
    for (int t = t_begin; t < t_end; t++) {
        const int y0 = t * BTH, buf = (t - t_begin) & 1;
        // ---- stage 1: the byte tile, zero outside the level (TMA: issued one tile ahead; LDG fallback: filled here)
        if (kTma) {
            if (threadIdx.x == 0 && t + 1 < t_end) {  // buffer buf^1 was last read (and, on border tiles, patched) before the barrier that ended tile t-1
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                tma_load_3d(tb[buf ^ 1], &tmap, x0 - BBX, (t + 1) * BTH - 3, f, &tma_bar[buf ^ 1], BBH * BBW);
            }
            mbar_wait(&tma_bar[buf], ((t - t_begin) >> 1) & 1);
        } else {
            for (int i = threadIdx.x; i < BBH * (BBW / 4); i += BWARPS * 32) {
                const int r = i / (BBW / 4), c = i - r * (BBW / 4);
                const int gy = y0 - 3 + r, gx = x0 - BBX + 4 * c;
                unsigned w = 0;
                if (gy >= 0 && gy < L.rows && gx >= 0 && gx < L.pitch) {
                    w = __ldg(reinterpret_cast<const uint32_t*>(img + (size_t)gy * L.pitch + gx));
                    if (gx + 3 >= L.cols) w &= gx >= L.cols ? 0u : (0xffffffffu >> (8 * (gx + 4 - L.cols)));
                }
                reinterpret_cast<uint32_t*>(tb[buf])[i] = w;
            }
            __syncthreads();
        }
        // ---- stage 2: BORDER_REFLECT_101 is applied by mirroring the halo INSIDE the byte tile (a reflected pixel is never more
        // than 4 px / 3 rows away from the border; only tiles on the level's border do this: rows first, then columns, so the
        // corners come out doubly reflected), then every tile converts bytes -> floats once per pixel with one uniform loop
        const int rows_needed = min(BTH, L.rows - y0) + 6;
        uint32_t* tb32 = reinterpret_cast<uint32_t*>(tb[buf]);
        const bool edge_y = y0 - 3 < 0 || y0 + BTH + 3 > L.rows, edge_x = x0 - 4 < 0 || x0 + BW + 4 > L.cols;
        if (edge_y) {
            for (int i = threadIdx.x; i < rows_needed * (BBW / 4); i += BWARPS * 32) {
                const int r = i / (BBW / 4), c = i - r * (BBW / 4);
                const int vy = y0 - 3 + r;
                if (vy < 0 || vy >= L.rows) tb32[i] = tb32[max(reflect101(vy, L.rows) - (y0 - 3), 0) * (BBW / 4) + c];
            }
            __syncthreads();
        }
        if (edge_x) {
            for (int i = threadIdx.x; i < rows_needed * 8; i += BWARPS * 32) {
                const int r = i >> 3, k = i & 7;
                const int gx = k < 4 ? k - 4 : L.cols + (k - 4);  // the four pixels left of the level, the four right of it
                const int tc = gx - (x0 - BBX);
                if ((k < 4 ? x0 - 4 < 0 : x0 + BW + 4 > L.cols) && tc >= 0 && tc < BBW)
                    tb[buf][r * BBW + tc] = tb[buf][r * BBW + min(max(reflect101(gx, L.cols) - (x0 - BBX), 0), BBW - 1)];
            }
            __syncthreads();
        }
        for (int i = threadIdx.x; i < rows_needed * WPR; i += BWARPS * 32) {
            const int r = i / WPR, c = i - r * WPR;
            const unsigned w = tb32[r * (BBW / 4) + c + (BBX - 4) / 4];
            float4 v;
            v.x = byte_to_float(w, 0x7440u);
            v.y = byte_to_float(w, 0x7441u);
            v.z = byte_to_float(w, 0x7442u);
            v.w = byte_to_float(w, 0x7443u);
            reinterpret_cast<float4*>(tin)[i] = v;
        }
        __syncthreads();
        // ---- filter: this lane's 4-pixel-wide strip of 15 output rows, seven row-pass results in registers
        const int ys = y0 + warp * BRS;
        if (ys < L.rows && x < pitch) {
            const int nout = min(BRS, L.rows - ys);
            const float4* src = reinterpret_cast<const float4*>(tin + (warp * BRS) * BIW) + lane;
            uint8_t* dst = out + (size_t)ys * pitch + x;
            float win[7][4];
            for (int base = 0; base < BRS + 6; base += 7) {  // unrolled by the window depth: the slots are compile-time registers
#pragma unroll
                for (int jj = 0; jj < 7; jj++) {
                    const int i = base + jj;
                    if (i < nout + 6) {
                        const float4 a = src[0], b = src[1], d = src[2];
                        src += BIW / 4;
                        const float p[10] = {a.y, a.z, a.w, b.x, b.y, b.z, b.w, d.x, d.y, d.z};  // pixels x-3 .. x+6
#pragma unroll
                        for (int k = 0; k < 4; k++) {
                            float v = k0 * p[k];
                            v = fmaf(k1, p[k + 1], v);
                            v = fmaf(k2, p[k + 2], v);
                            v = fmaf(k3, p[k + 3], v);
                            v = fmaf(k2, p[k + 4], v);
                            v = fmaf(k1, p[k + 5], v);
                            v = fmaf(k0, p[k + 6], v);
                            win[jj][k] = v;
                        }
                        if (i >= 6) {  // staged rows i-6 .. i are in the window: emit output row i-6, centred on staged row i-3
                            unsigned v[4];
#pragma unroll
                            for (int k = 0; k < 4; k++) {
                                float acc = k3 * win[(jj + 4) % 7][k];
                                acc = fmaf(k2, win[(jj + 5) % 7][k] + win[(jj + 3) % 7][k], acc);
                                acc = fmaf(k1, win[(jj + 6) % 7][k] + win[(jj + 2) % 7][k], acc);
                                acc = fmaf(k0, win[jj][k] + win[(jj + 1) % 7][k], acc);
                                v[k] = sat_u8_rn(acc);
                            }
                            *reinterpret_cast<uint32_t*>(dst) = __byte_perm(__byte_perm(v[0], v[1], 0x0040), __byte_perm(v[2], v[3], 0x0040), 0x5410);
                            dst += pitch;
                        }
                    }
                }
            }
        }
        __syncthreads();  // tin (and, one tile later, this byte buffer) may be overwritten
    }
}

// ---- B6 + B8: warp per keypoint ----------------------------------------------------------------------------
__device__ __forceinline__ float cv_fast_atan2(float y, float x) {
    const float scale = (float)(180.0 / 3.14159265358979323846);
    const float p1 = 0.9997878412794807f * scale, p3 = -0.3258083974640975f * scale, p5 = 0.1555786518463281f * scale,
                p7 = -0.04432655554792128f * scale;
    const float ax = fabsf(x), ay = fabsf(y);
    float a, c, c2;
    if (ax >= ay) {
        c = ay / (ax + 2.2204460492503131e-16f);
        c2 = c * c;
        a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    } else {
        c = ax / (ay + 2.2204460492503131e-16f);
        c2 = c * c;
        a = 90.f - (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    }
    if (x < 0) a = 180.f - a;
    if (y < 0) a = 360.f - a;
    return a;
}

// One warp owns 32 consecutive keypoints of a frame.  Phase A: per keypoint, the 31 lanes of a disc row set sum
// their column of the unblurred level (integer moments, any order is exact) and lane j keeps keypoint j's angle.
// Phase B: every lane does the double-precision cos / sin of ITS keypoint (once per keypoint instead of 32
// redundant copies).  Phase C: per keypoint, lane k builds descriptor byte k from 8 rotated pattern pairs.
__global__ void __launch_bounds__(128) orb_describe_kernel(SeqView s, OrbView o, int first, const float2* __restrict__ patf) {
    const int f = first + blockIdx.y;
    const int n = s.n_kp[f];
    const int base = (blockIdx.x * 4 + (threadIdx.x >> 5)) * 32;
    if (base >= n) return;
    const int lane = lane_id();
    const int mine = base + lane;
    const bool have = mine < n;
    const size_t kbase = (size_t)f * s.cap_kp;
    const int my_l = have ? o.octave[kbase + mine] : 0;
    const uint32_t my_p = have ? o.lxy[kbase + mine] : 0u;
    const int cnt = min(32, n - base);
    // IC_Angle on the unblurred level: umax-limited disc of radius 15
    constexpr int umax[16] = {15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3};
    float my_angle = 0.f;
    const int u = lane - kOrbHalfPatch;  // lanes 0..30 -> u = -15..15
    const int au = abs(u);
    for (int j = 0; j < cnt; j++) {
        const int l = __shfl_sync(0xffffffffu, my_l, j);
        const uint32_t p = __shfl_sync(0xffffffffu, my_p, j);
        const int x = p & 0xffff, y = p >> 16;
        const int pitch = o.lv[l].pitch;
        const uint8_t* img = level_ptr(s, o, f, l);
        int m01 = 0, m10 = 0;
        if (u <= kOrbHalfPatch) {
            const uint8_t* c = img + (size_t)y * pitch + x + u;
            int col = c[0];  // v = 0 row
            int vsum = 0;
#pragma unroll
            for (int v = 1; v <= kOrbHalfPatch; v++) {
                if (au <= umax[v]) {
                    SLAMCU_BOUND((y + v) * pitch + x + u, o.lv[l].rows * pitch);
                    SLAMCU_BOUND((y - v) * pitch + x + u, o.lv[l].rows * pitch);
                    const int below = c[v * pitch], above = c[-v * pitch];
                    col += below + above;
                    vsum += v * (below - above);
                }
            }
            m10 = u * col;
            m01 = vsum;
        }
#pragma unroll
        for (int q = 16; q; q >>= 1) {
            m01 += __shfl_xor_sync(0xffffffffu, m01, q);
            m10 += __shfl_xor_sync(0xffffffffu, m10, q);
        }
        if (lane == j) my_angle = cv_fast_atan2((float)m01, (float)m10);
    }
    // rBRIEF: a = (float)cos(angle_rad), b = (float)sin(angle_rad) in double, narrowed
    float my_a = 1.f, my_b = 0.f;
    if (have) {
        s.kps[kbase + mine].angle = my_angle;
        const float ar = my_angle * (float)(3.14159265358979323846 / 180.f);
        my_a = (float)cos((double)ar);
        my_b = (float)sin((double)ar);
    }
    float2 pp[16];  // this lane's 16 pattern points (8 pairs -> descriptor byte `lane`), the same for every keypoint
#pragma unroll
    for (int k = 0; k < 16; k++) pp[k] = __ldg(patf + lane * 16 + k);
    for (int j = 0; j < cnt; j++) {
        const int l = __shfl_sync(0xffffffffu, my_l, j);
        const uint32_t p = __shfl_sync(0xffffffffu, my_p, j);
        const float a = __shfl_sync(0xffffffffu, my_a, j), b = __shfl_sync(0xffffffffu, my_b, j);
        const int x = p & 0xffff, y = p >> 16;
        const int pitch = o.lv[l].pitch;
        const uint8_t* cb = blur_ptr(s, o, f, l) + (size_t)y * pitch + x;
        unsigned byte = 0;
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const float2 p0 = pp[2 * k], p1 = pp[2 * k + 1];
            const int ix0 = __float2int_rn(p0.x * a - p0.y * b), iy0 = __float2int_rn(p0.x * b + p0.y * a);
            const int ix1 = __float2int_rn(p1.x * a - p1.y * b), iy1 = __float2int_rn(p1.x * b + p1.y * a);
            SLAMCU_BOUND((y + iy0) * pitch + x + ix0, o.lv[l].rows * pitch);
            SLAMCU_BOUND((y + iy1) * pitch + x + ix1, o.lv[l].rows * pitch);
            const int t0 = cb[iy0 * pitch + ix0], t1 = cb[iy1 * pitch + ix1];
            byte |= (unsigned)(t0 < t1) << k;
        }
        // pack 4 lanes' bytes into one 32-bit word
        unsigned w = byte << (8 * (lane & 3));
        w |= __shfl_xor_sync(0xffffffffu, w, 1);
        w |= __shfl_xor_sync(0xffffffffu, w, 2);
        if ((lane & 3) == 0) s.desc[(kbase + base + j) * s.desc_words + (lane >> 2)] = w;
    }
}

}  // namespace

void init_orb_attributes(int smem_optin) {
    cudaFuncSetAttribute(orb_retain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    // 56 KB of dynamic shared memory per 4-warp block (above the 48 KB default limit), four blocks per SM
    cudaFuncSetAttribute(blur7_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBlurSmemBytes);
    cudaFuncSetAttribute(blur7_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBlurSmemBytes);
    cudaFuncSetAttribute(blur7_kernel<true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(blur7_kernel<false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    (void)smem_optin;
}

int launch_orb_extract(const SeqView& s, const OrbView& o, int first, int n, cudaStream_t st, cudaStream_t aux, cudaEvent_t ev_fork,
                       cudaEvent_t ev_join, const OrbTmaps* tmaps) {
    int launches = 0;
    int max_rows = 0, max_capc = 0;
    for (int l = 0; l < o.nlevels; l++) {
        max_rows = max(max_rows, o.lv[l].rows);
        max_capc = max(max_capc, o.lv[l].capc);
    }
    for (int l = 1; l < o.nlevels; l++) {
        // source window of one output tile: (tile * step + 2) pixels, + 15 for the 16-byte alignment of its left edge
        const double sx = (double)o.lv[l - 1].cols / o.lv[l].cols, sy = (double)o.lv[l - 1].rows / o.lv[l].rows;
        const bool fits = PT_W * sx + 3 + 15 + 16 <= PS_PITCH && PT_H * sy + 4 <= PS_ROWS;
        if (fits) {
            dim3 grid((o.lv[l].pitch + PT_W - 1) / PT_W, (o.lv[l].rows + PT_H - 1) / PT_H, n);
            SLAM_KERNEL("pyr_down", st, pyr_down_kernel<<<grid, PT_THREADS, 0, st>>>(s, o, first, l));
        } else {  // steep scale factors: per-pixel gather
            dim3 grid((o.lv[l].pitch + 127) / 128, (o.lv[l].rows + 7) / 8, n);
            SLAM_KERNEL("pyr_down", st, pyr_down_gather_kernel<<<grid, 256, 0, st>>>(s, o, first, l));
        }
        launches++;
    }
    // The blurred levels depend only on the pyramid: they run on the auxiliary stream next to the corner chain
    // (FAST -> select -> Harris -> retain -> assemble) and join before the descriptor kernel.
    // (not while per-kernel timing is on: overlapped kernels would each be charged the other's time)
    const bool forked = aux != nullptr && aux != st && slamcu::g_prof == nullptr;
    cudaStream_t bs = forked ? aux : st;
    if (forked) {
        cudaEventRecord(ev_fork, st);
        cudaStreamWaitEvent(aux, ev_fork, 0);
    }
    for (int l = 0; l < o.nlevels; l++) {
        // a block takes a whole column of tiles when there are frames enough to fill the device, fewer for small batches
        const int tx = (o.lv[l].cols + BW - 1) / BW, ty = (o.lv[l].rows + BTH - 1) / BTH;
        const int tpb = max(1, min(ty, (int)((long long)tx * ty * n / 1200)));
        dim3 grid(tx, (ty + tpb - 1) / tpb, n);
        static const bool no_tma = getenv("SLAMCU_NO_TMA") != nullptr;  // debugging aid: stage every tile with plain loads
        if (tmaps && tmaps->valid && !no_tma)
            SLAM_KERNEL("blur7", bs, blur7_kernel<true><<<grid, BWARPS * 32, kBlurSmemBytes, bs>>>(s, o, first, l, tpb, tmaps->blur[l]));
        else
            SLAM_KERNEL("blur7", bs, blur7_kernel<false><<<grid, BWARPS * 32, kBlurSmemBytes, bs>>>(s, o, first, l, tpb, CUtensorMap{}));
        launches++;
    }
    if (forked) cudaEventRecord(ev_join, aux);
    for (int l = 0; l < o.nlevels; l++) {
        dim3 grid((o.lv[l].cols + FTW - 1) / FTW, (o.lv[l].rows + FTH - 1) / FTH, n);
        if (tmaps && tmaps->valid)
            SLAM_KERNEL("fast9_mask", st, fast9_mask_kernel<true><<<grid, 256, 0, st>>>(s, o, first, l, tmaps->fast[l]));
        else
            SLAM_KERNEL("fast9_mask", st, fast9_mask_kernel<false><<<grid, 256, 0, st>>>(s, o, first, l, CUtensorMap{}));
        launches++;
    }
    SLAM_KERNEL("orb_select", st,
                orb_select_kernel<<<dim3(o.nlevels, n), 256, (max_rows + 1) * sizeof(int), st>>>(s, o, first));
    int max_quota = 1;
    for (int l = 0; l < o.nlevels; l++) max_quota = max(max_quota, o.lv[l].quota);
    SLAM_KERNEL("harris", st, harris_kernel<<<dim3((2 * max_quota + 127) / 128, o.nlevels, n), 128, 0, st>>>(s, o, first));
    SLAM_KERNEL("orb_retain", st, orb_retain_kernel<<<dim3(o.nlevels, n), 256, 0, st>>>(s, o, first, 0));
    SLAM_KERNEL("orb_assemble", st, orb_assemble_kernel<<<n, 256, 0, st>>>(s, o, first));
    launches += 4;
    if (forked) cudaStreamWaitEvent(st, ev_join, 0);
    SLAM_KERNEL("orb_describe", st,
                orb_describe_kernel<<<dim3((s.cap_kp + 127) / 128, n), 128, 0, st>>>(s, o, first, o.patf));
    launches++;
    return launches;
}

}  // namespace slamcu
