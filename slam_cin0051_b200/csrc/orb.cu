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
#include "orb.cuh"

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
__global__ void __launch_bounds__(256) pyr_down_kernel(SeqView s, OrbView o, int first, int l) {
    const int f = first + blockIdx.z;
    const OrbLevel& d = o.lv[l];
    const OrbLevel& p = o.lv[l - 1];
    const int x = blockIdx.x * 64 + (threadIdx.x & 15) * 4;  // 4 pixels per thread -> one 32-bit store
    const int y = blockIdx.y * 16 + (threadIdx.x >> 4);
    if (y >= d.rows || x >= d.pitch) return;
    const uint8_t* src = level_ptr(s, o, f, l - 1);
    uint8_t* dst = level_ptr_w(s, o, f, l);
    const int y0 = d.y0[y], y1 = d.y1[y], ay = d.ay[y];
    const uint8_t* r0 = src + (size_t)y0 * p.pitch;
    const uint8_t* r1 = src + (size_t)y1 * p.pitch;
    uint32_t packed = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int xx = x + k;
        uint32_t v = 0;
        if (xx < d.cols) {
            const int x0 = d.x0[xx], x1 = d.x1[xx], ax = d.ax[xx];
            const int h0 = (256 - ax) * r0[x0] + ax * r0[x1];
            const int h1 = (256 - ax) * r1[x0] + ax * r1[x1];
            v = (uint32_t)(((256 - ay) * h0 + ay * h1 + 32768) >> 16);
        }
        packed |= v << (8 * k);
    }
    *reinterpret_cast<uint32_t*>(dst + (size_t)y * d.pitch + x) = packed;
}

// ---- B2 ------------------------------------------------------------------------------------------
constexpr int FTW = 128, FTH = 32, FHX = 16, FSW = FTW + 2 * FHX, FSH = FTH + 8;  // pixel tile with 4-row halo
constexpr int SCW = FTW + 2, SCH = FTH + 2;                                      // score tile with 1-px halo

// cornerScore of cv::FAST (9/16): max over the 16 arcs of 9 ring pixels of min |v - p| (one sign), minus 1.
// With d = v - p:  min over an arc of d = v - max(p),  min of -d = min(p) - v, so the score needs the
// sliding-window (length 9, circular) min and max of the raw ring pixels: doubling steps 2, 4, 8, then +1.
// The subtraction from v is applied AFTER the min/max network on purpose: nvcc 12.9 folds
// max(a, -b) chains into VIMNMX3 for sm_100a and loses the negation ("toolchain findings").
__device__ __forceinline__ int fast9_ring_score(int v, const int (&p)[16], int thr) {
    int n2[16], x2[16], n4[16], x4[16];
#pragma unroll
    for (int i = 0; i < 16; i++) {
        n2[i] = min(p[i], p[(i + 1) & 15]);
        x2[i] = max(p[i], p[(i + 1) & 15]);
    }
#pragma unroll
    for (int i = 0; i < 16; i++) {
        n4[i] = min(n2[i], n2[(i + 2) & 15]);
        x4[i] = max(x2[i], x2[(i + 2) & 15]);
    }
    int brightest_min = 0, darkest_max = 255;
#pragma unroll
    for (int i = 0; i < 16; i++) {
        const int n9 = min(min(n4[i], n4[(i + 4) & 15]), p[(i + 8) & 15]);
        const int x9 = max(max(x4[i], x4[(i + 4) & 15]), p[(i + 8) & 15]);
        brightest_min = max(brightest_min, n9);  // best arc whose pixels are all brighter than v
        darkest_max = min(darkest_max, x9);      // best arc whose pixels are all darker than v
    }
    const int bright = brightest_min - v, dark = v - darkest_max;
    return max(thr, max(bright, dark)) - 1;
}

// ring of radius 3, OpenCV order; only the set of arcs matters
__device__ __forceinline__ int fast9_score(const uint8_t* t, int thr) {
    constexpr int dxs[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
    constexpr int dys[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};
    const int v = t[0];
    const int up = v + thr, dn = v - thr;
    // opposite-pixel rejection: a 9-arc contains one pixel of every antipodal pair
    {
        const int a = t[3 * FSW], b = t[-3 * FSW];
        if (!((a > up) | (b > up) | (a < dn) | (b < dn))) return 0;
        const int c = t[3], e = t[-3];
        if (!((c > up) | (e > up) | (c < dn) | (e < dn))) return 0;
    }
    int p[16];
    unsigned mh = 0, ml = 0;
#pragma unroll
    for (int k = 0; k < 16; k++) {
        p[k] = t[dys[k] * FSW + dxs[k]];
        mh |= (unsigned)(p[k] > up) << k;
        ml |= (unsigned)(p[k] < dn) << k;
    }
    mh |= mh << 16;
    ml |= ml << 16;
    unsigned rh = mh & (mh >> 1), rl = ml & (ml >> 1);
    rh &= rh >> 2;
    rl &= rl >> 2;
    rh &= rh >> 4;
    rl &= rl >> 4;            // runs of 8
    rh &= mh >> 8;
    rl &= ml >> 8;            // runs of 9
    if ((rh | rl) == 0) return 0;
    return fast9_ring_score(v, p, thr);
}

__global__ void __launch_bounds__(256) fast9_mask_kernel(SeqView s, OrbView o, int first, int l) {
    __shared__ __align__(16) uint8_t tile[FSH * FSW];
    __shared__ uint8_t sc[SCH * SCW];
    const int f = first + blockIdx.z;
    const OrbLevel& L = o.lv[l];
    const int x0 = blockIdx.x * FTW, y0 = blockIdx.y * FTH;
    const uint8_t* img = level_ptr(s, o, f, l);
    constexpr int VPR = FSW / 16;
    for (int v = threadIdx.x; v < FSH * VPR; v += blockDim.x) {
        const int r = v / VPR, cv = v - r * VPR;
        const int gy = y0 - 4 + r, gx = x0 - FHX + cv * 16;
        uint4 val = make_uint4(0, 0, 0, 0);
        if (gy >= 0 && gy < L.rows && gx >= 0 && gx < L.pitch)
            val = __ldg(reinterpret_cast<const uint4*>(img + (size_t)gy * L.pitch + gx));
        *reinterpret_cast<uint4*>(tile + r * FSW + cv * 16) = val;
    }
    __syncthreads();
    // scores on the output tile plus a 1-pixel ring (needed by the 3x3 NMS)
    for (int i = threadIdx.x; i < SCH * SCW; i += blockDim.x) {
        const int ry = i / SCW, rx = i - ry * SCW;
        const int gx = x0 - 1 + rx, gy = y0 - 1 + ry;
        int v = 0;
        if (gx >= 3 && gx < L.cols - 3 && gy >= 3 && gy < L.rows - 3)
            v = fast9_score(tile + (ry + 3) * FSW + (FHX - 1) + rx, o.fast_threshold);
        sc[i] = (uint8_t)v;
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t* mask = o.mask + (size_t)f * o.mask_words + L.moff;
    for (int r = warp; r < FTH; r += 8) {
        const int gy = y0 + r;
        if (gy >= L.rows) break;
#pragma unroll
        for (int wx = 0; wx < FTW / 32; wx++) {
            const int lx = wx * 32 + lane;
            const int gx = x0 + lx;
            const uint8_t* c = sc + (r + 1) * SCW + lx + 1;
            const int v = c[0];
            bool keep = v > 0 && gx >= kOrbEdge && gx < L.cols - kOrbEdge && gy >= kOrbEdge && gy < L.rows - kOrbEdge;
            if (keep)
                keep = v > c[-1] && v > c[1] && v > c[-SCW - 1] && v > c[-SCW] && v > c[-SCW + 1] && v > c[SCW - 1] &&
                       v > c[SCW] && v > c[SCW + 1];
            const unsigned word = __ballot_sync(0xffffffffu, keep);
            const int wi = (x0 >> 5) + wx;
            if (lane == 0 && wi < L.mwords) mask[(size_t)gy * L.mwords + wi] = word;
        }
    }
}

// global-memory version of the score for the list kernel (same arithmetic, pitch-strided)
__device__ int fast9_score_global(const uint8_t* c, int pitch, int thr) {
    constexpr int dxs[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
    constexpr int dys[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};
    int p[16];
#pragma unroll
    for (int k = 0; k < 16; k++) p[k] = c[dys[k] * pitch + dxs[k]];
    return fast9_ring_score(c[0], p, thr);
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
    const uint8_t* img = level_ptr(s, o, f, l);
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
        const int sc = fast9_score_global(img + (size_t)(p >> 16) * L.pitch + (p & 0xffff), L.pitch, o.fast_threshold);
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
        if (keep) sxy[off + __popc(b & lanemask_lt())] = cxy[i];
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

// ---- B4: warp per candidate ---------------------------------------------------------------------------
__global__ void __launch_bounds__(128) harris_kernel(SeqView s, OrbView o, int first) {
    const int l = blockIdx.y, f = first + blockIdx.z;
    const OrbLevel& L = o.lv[l];
    const int n = o.n_sel[f * kMaxLevels + l];
    const unsigned lane = lane_id();
    const uint8_t* img = level_ptr(s, o, f, l);
    for (int i = blockIdx.x * 4 + (threadIdx.x >> 5); i < n; i += gridDim.x * 4) {
    const uint32_t p = o.sxy[(size_t)f * o.cand_total + L.coff + i];
    const int x = p & 0xffff, y = p >> 16;
    int a = 0, b = 0, c = 0;
    for (int k = lane; k < 49; k += 32) {
        const int dy = k / 7 - 3, dx = k - (k / 7) * 7 - 3;
        const uint8_t* q = img + (size_t)(y + dy) * L.pitch + (x + dx);
        const int st = L.pitch;
        const int Ix = ((int)q[1] - (int)q[-1]) * 2 + ((int)q[-st + 1] - (int)q[-st - 1]) + ((int)q[st + 1] - (int)q[st - 1]);
        const int Iy = ((int)q[st] - (int)q[-st]) * 2 + ((int)q[st - 1] - (int)q[-st - 1]) + ((int)q[st + 1] - (int)q[-st + 1]);
        a += Ix * Ix;
        b += Iy * Iy;
        c += Ix * Iy;
    }
#pragma unroll
    for (int q = 16; q; q >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, q);
        b += __shfl_xor_sync(0xffffffffu, b, q);
        c += __shfl_xor_sync(0xffffffffu, c, q);
    }
    if (lane == 0) {
        const float scale = 1.f / ((1 << 2) * 7 * 255.f);
        const float scale4 = scale * scale * scale * scale;
        const float fa = (float)a, fb = (float)b, fc = (float)c;
        // ((float)a * b - (float)c * c - harris_k * ((float)a + b) * ((float)a + b)) * scale_sq_sq   (no FMA)
        const float r = ((fa * fb - fc * fc) - (0.04f * (fa + fb)) * (fa + fb)) * scale4;
        o.sresp[(size_t)f * o.cand_total + L.coff + i] = r;
    }
    }
}

// ---- B5: one block per (level, frame) --------------------------------------------------------------------
__global__ void __launch_bounds__(256) orb_retain_kernel(SeqView s, OrbView o, int first, int smem_cap) {
    extern __shared__ float sr[];
    __shared__ int warp_tot[8];
    __shared__ int carry;
    const int l = blockIdx.x, f = first + blockIdx.y;
    const OrbLevel& L = o.lv[l];
    const int n = o.n_sel[f * kMaxLevels + l];
    const size_t base_off = (size_t)f * o.cand_total + L.coff;
    const float* gr = o.sresp + base_off;
    const uint32_t* sxy = o.sxy + base_off;
    uint32_t* fxy = o.fxy + base_off;
    float* fr = o.fresp + base_off;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool in_smem = n <= smem_cap;
    if (in_smem)
        for (int i = threadIdx.x; i < n; i += blockDim.x) sr[i] = gr[i];
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const float* r = in_smem ? sr : gr;
    const int want = L.quota;
    for (int base = 0; base < n; base += 256) {
        const int i = base + threadIdx.x;
        bool keep = false;
        if (i < n) {
            if (n <= want) keep = true;
            else if (want > 0) {
                const float v = r[i];
                int greater = 0;
                for (int j = 0; j < n; j++) greater += (r[j] > v) ? 1 : 0;
                keep = greater < want;  // v >= the want-th largest (ties kept)
            }
        }
        const unsigned b = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) warp_tot[warp] = __popc(b);
        __syncthreads();
        int off = carry;
        for (int w = 0; w < warp; w++) off += warp_tot[w];
        if (keep) {
            const int pos = off + __popc(b & lanemask_lt());
            fxy[pos] = sxy[i];
            fr[pos] = r[i];
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
constexpr int GW = 64, GH = 16;
__device__ __forceinline__ int reflect101(int i, int n) {
    if (i < 0) i = -i;
    if (i >= n) i = 2 * n - 2 - i;
    return min(max(i, 0), n - 1);
}

__global__ void __launch_bounds__(256) blur7_kernel(SeqView s, OrbView o, int first, int l) {
    __shared__ uint8_t tin[(GH + 6) * (GW + 8)];
    __shared__ float trow[(GH + 6) * GW];
    const int f = first + blockIdx.z;
    const OrbLevel& L = o.lv[l];
    const uint8_t* img = level_ptr(s, o, f, l);
    uint8_t* out = blur_ptr(s, o, f, l);
    const int x0 = blockIdx.x * GW, y0 = blockIdx.y * GH;
    // getGaussianKernel(7, 2, CV_32F) (bit patterns of the cv2 result; OpenCV computes exp(-x^2/8) normalised)
    const float k0 = __uint_as_float(0x3d8fafb1u), k1 = __uint_as_float(0x3e06387eu), k2 = __uint_as_float(0x3e434a39u),
                k3 = __uint_as_float(0x3e5d4ae0u);
    for (int i = threadIdx.x; i < (GH + 6) * (GW + 6); i += blockDim.x) {
        const int r = i / (GW + 6), c = i - r * (GW + 6);
        const int gy = reflect101(y0 - 3 + r, L.rows), gx = reflect101(x0 - 3 + c, L.cols);
        tin[r * (GW + 8) + c] = img[(size_t)gy * L.pitch + gx];
    }
    __syncthreads();
    // row pass: s = k0*S0; s = fma(k_j, S_j, s), j = 1..6
    for (int i = threadIdx.x; i < (GH + 6) * GW; i += blockDim.x) {
        const int r = i / GW, c = i - r * GW;
        const uint8_t* p = tin + r * (GW + 8) + c;
        float acc = k0 * (float)p[0];
        acc = fmaf(k1, (float)p[1], acc);
        acc = fmaf(k2, (float)p[2], acc);
        acc = fmaf(k3, (float)p[3], acc);
        acc = fmaf(k2, (float)p[4], acc);
        acc = fmaf(k1, (float)p[5], acc);
        acc = fmaf(k0, (float)p[6], acc);
        trow[i] = acc;
    }
    __syncthreads();
    // column pass: s = k3*R0; s = fma(k_{3+j}, R_{+j} + R_{-j}, s), j = 1..3; cvRound, saturate
    const int tx = (threadIdx.x & 15) * 4, ty = threadIdx.x >> 4;
    const int gy = y0 + ty;
    if (gy >= L.rows) return;
    uint32_t packed = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const float* q = trow + (ty + 3) * GW + tx + k;
        float acc = k3 * q[0];
        acc = fmaf(k2, q[GW] + q[-GW], acc);
        acc = fmaf(k1, q[2 * GW] + q[-2 * GW], acc);
        acc = fmaf(k0, q[3 * GW] + q[-3 * GW], acc);
        const int v = min(max(__float2int_rn(acc), 0), 255);
        packed |= (uint32_t)v << (8 * k);
    }
    if (x0 + tx < L.pitch) *reinterpret_cast<uint32_t*>(out + (size_t)gy * L.pitch + x0 + tx) = packed;
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

__global__ void __launch_bounds__(128) orb_describe_kernel(SeqView s, OrbView o, int first) {
    const int f = first + blockIdx.y;
    const int n = s.n_kp[f];
    const int kpi = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (kpi >= n) return;
    const unsigned lane = lane_id();
    const int l = o.octave[(size_t)f * s.cap_kp + kpi];
    const uint32_t p = o.lxy[(size_t)f * s.cap_kp + kpi];
    const int x = p & 0xffff, y = p >> 16;
    const OrbLevel& L = o.lv[l];
    const uint8_t* img = level_ptr(s, o, f, l);
    const uint8_t* bl = blur_ptr(s, o, f, l);
    // IC_Angle on the unblurred level: umax-limited disc of radius 15
    constexpr int umax[16] = {15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3};
    int m01 = 0, m10 = 0;
    {
        const int u = (int)lane - kOrbHalfPatch;  // lanes 0..30 -> u = -15..15
        if (u <= kOrbHalfPatch) {
            const uint8_t* c = img + (size_t)y * L.pitch + x + u;
            int col = c[0];  // v = 0 row
            int vsum = 0;
#pragma unroll
            for (int v = 1; v <= kOrbHalfPatch; v++) {
                if (abs(u) <= umax[v]) {
                    const int below = c[v * L.pitch], above = c[-v * L.pitch];
                    col += below + above;
                    vsum += v * (below - above);
                }
            }
            m10 = u * col;
            m01 = vsum;
        }
    }
#pragma unroll
    for (int q = 16; q; q >>= 1) {
        m01 += __shfl_xor_sync(0xffffffffu, m01, q);
        m10 += __shfl_xor_sync(0xffffffffu, m10, q);
    }
    const float angle = cv_fast_atan2((float)m01, (float)m10);
    if (lane == 0) s.kps[(size_t)f * s.cap_kp + kpi].angle = angle;
    // rBRIEF: a = (float)cos(angle_rad), b = (float)sin(angle_rad) in double, narrowed
    const float ar = angle * (float)(3.14159265358979323846 / 180.f);
    const float a = (float)cos((double)ar), b = (float)sin((double)ar);
    const uint8_t* cb = bl + (size_t)y * L.pitch + x;
    const int8_t* pat = o.pattern + lane * 32;  // 16 points (x, y) per lane
    unsigned byte = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const float x0 = (float)pat[4 * k + 0], y0 = (float)pat[4 * k + 1];
        const float x1 = (float)pat[4 * k + 2], y1 = (float)pat[4 * k + 3];
        const int ix0 = __float2int_rn(x0 * a - y0 * b), iy0 = __float2int_rn(x0 * b + y0 * a);
        const int ix1 = __float2int_rn(x1 * a - y1 * b), iy1 = __float2int_rn(x1 * b + y1 * a);
        const int t0 = cb[iy0 * L.pitch + ix0], t1 = cb[iy1 * L.pitch + ix1];
        byte |= (unsigned)(t0 < t1) << k;
    }
    // pack 4 lanes' bytes into one 32-bit word
    unsigned w = byte << (8 * (lane & 3));
    w |= __shfl_xor_sync(0xffffffffu, w, 1);
    w |= __shfl_xor_sync(0xffffffffu, w, 2);
    if ((lane & 3) == 0) s.desc[((size_t)f * s.cap_kp + kpi) * s.desc_words + (lane >> 2)] = w;
}

}  // namespace

void init_orb_attributes(int smem_optin) {
    cudaFuncSetAttribute(orb_retain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    (void)smem_optin;
}

int launch_orb_extract(const SeqView& s, const OrbView& o, int first, int n, cudaStream_t st) {
    int launches = 0;
    int max_rows = 0, max_capc = 0;
    for (int l = 0; l < o.nlevels; l++) {
        max_rows = max(max_rows, o.lv[l].rows);
        max_capc = max(max_capc, o.lv[l].capc);
    }
    for (int l = 1; l < o.nlevels; l++) {
        dim3 grid((o.lv[l].pitch + 63) / 64, (o.lv[l].rows + 15) / 16, n);
        SLAM_KERNEL("pyr_down", st, pyr_down_kernel<<<grid, 256, 0, st>>>(s, o, first, l));
        launches++;
    }
    for (int l = 0; l < o.nlevels; l++) {
        dim3 grid((o.lv[l].cols + FTW - 1) / FTW, (o.lv[l].rows + FTH - 1) / FTH, n);
        SLAM_KERNEL("fast9_mask", st, fast9_mask_kernel<<<grid, 256, 0, st>>>(s, o, first, l));
        launches++;
    }
    SLAM_KERNEL("orb_select", st,
                orb_select_kernel<<<dim3(o.nlevels, n), 256, (max_rows + 1) * sizeof(int), st>>>(s, o, first));
    int max_quota = 1;
    for (int l = 0; l < o.nlevels; l++) max_quota = max(max_quota, o.lv[l].quota);
    SLAM_KERNEL("harris", st, harris_kernel<<<dim3((2 * max_quota + 3) / 4, o.nlevels, n), 128, 0, st>>>(s, o, first));
    const int retain_cap = 12 * 1024;
    SLAM_KERNEL("orb_retain", st,
                orb_retain_kernel<<<dim3(o.nlevels, n), 256, retain_cap * sizeof(float), st>>>(s, o, first, retain_cap));
    SLAM_KERNEL("orb_assemble", st, orb_assemble_kernel<<<n, 256, 0, st>>>(s, o, first));
    launches += 4;
    for (int l = 0; l < o.nlevels; l++) {
        dim3 grid((o.lv[l].cols + GW - 1) / GW, (o.lv[l].rows + GH - 1) / GH, n);
        SLAM_KERNEL("blur7", st, blur7_kernel<<<grid, 256, 0, st>>>(s, o, first, l));
        launches++;
    }
    SLAM_KERNEL("orb_describe", st, orb_describe_kernel<<<dim3((s.cap_kp + 3) / 4, n), 128, 0, st>>>(s, o, first));
    launches++;
    return launches;
}

}  // namespace slamcu
