// sortnms.cu -- FeatureDetector::applyNonMaxSuppression (feature_detector.cpp:147-188), row A3.
//
// The reference sorts the raw corners by response with an *unstable* std::sort (:160-161) and then
// walks the sorted list greedily (:165-185).  Both the kept set and -- more visibly -- the output
// order (which match indices refer to) depend on the exact permutation libstdc++'s introsort yields
// on ties, and responses are small integers so ties are everywhere.
//
// sort_kernel reproduces that permutation *in parallel*.  Facts used (bits/stl_algo.h, libstdc++ 13):
//  * __introsort_loop partitions independent sub-ranges, so they can be processed level by level
//    (depth limit = 2*lg n decrements once per level; on 0 the range is heap-sorted).
//  * __unguarded_partition is a Hoare partition.  With L_1<L_2<... the positions whose element does
//    not precede the pivot (left-scan stops) and R_1>R_2>... the positions the pivot does not precede
//    (right-scan stops), swap k happens iff L_k < R_k, both scans only ever see untouched elements,
//    and the returned cut is L_{K+1} if it lies left of R_K, else R_K (K = number of swaps).
//    So a warp finds both stop lists with ballots, K with one comparison per pair, and swaps in parallel.
//  * __final_insertion_sort never moves an element across a cut (left >= pivot >= right), so it is
//    equivalent to a stable insertion sort of every final sub-range (<= 16 elements, or heap-sorted).
// tests/native/host_exact.cpp holds a scalar model of this formulation checked against std::sort.
//
// nms_kernel is the greedy walk: one warp per frame, a 1-bit/pixel "suppressed" bitmap (shared memory
// when it fits), 32 sorted corners per step with an in-warp resolution of intra-chunk conflicts, and
// disc stamping for the survivors.  The disc uses the reference's float predicate
// sqrt(dx*dx+dy*dy) < float(window) evaluated per offset, so it is exact for any window.
#include "common.cuh"
#include "exact.cuh"

namespace slamcu {

namespace {

struct KeyDesc {
    __device__ __forceinline__ bool operator()(uint32_t a, uint32_t b) const {
        return (a >> kKeyIdxBits) > (b >> kKeyIdxBits);
    }
};

// Warp-cooperative exact Hoare partition of keys[first,last) (last-first > 16). Returns the cut.
__device__ int warp_partition(uint32_t* keys, int first, int last, uint32_t* Lpos, uint32_t* Rasc) {
    const unsigned lane = lane_id();
    const KeyDesc less;
    if (lane == 0) {  // __move_median_to_first(first, first+1, mid, last-1)
        const int ia = first + 1, ib = first + (last - first) / 2, ic = last - 1;
        const uint32_t a = keys[ia], b = keys[ib], c = keys[ic];
        int pick;
        if (less(a, b)) {
            if (less(b, c)) pick = ib;
            else if (less(a, c)) pick = ic;
            else pick = ia;
        } else if (less(a, c)) pick = ia;
        else if (less(b, c)) pick = ic;
        else pick = ib;
        const uint32_t t = keys[first];
        keys[first] = keys[pick];
        keys[pick] = t;
    }
    __syncwarp();
    const uint32_t pivot = keys[first];
    int nL = 0, nR = 0;
    for (int base = first + 1; base < last; base += 32) {
        const int i = base + (int)lane;
        bool isL = false, isR = false;
        if (i < last) {
            const uint32_t k = keys[i];
            isL = !less(k, pivot);
            isR = !less(pivot, k);
        }
        const unsigned bL = __ballot_sync(0xffffffffu, isL), bR = __ballot_sync(0xffffffffu, isR);
        if (isL) Lpos[first + nL + __popc(bL & lanemask_lt())] = (uint32_t)i;
        if (isR) Rasc[first + nR + __popc(bR & lanemask_lt())] = (uint32_t)i;
        nL += __popc(bL);
        nR += __popc(bR);
    }
    __syncwarp();
    // K = #{k : L_k < R_k}; the predicate is monotone in k
    const int m = min(nL, nR);
    int K = 0;
    for (int base = 0; base < m; base += 32) {
        const int k = base + (int)lane;
        const bool ok = k < m && Lpos[first + k] < Rasc[first + nR - 1 - k];
        const unsigned b = __ballot_sync(0xffffffffu, ok);
        K += __popc(b);
        if (b != 0xffffffffu) break;
    }
    for (int k = (int)lane; k < K; k += 32) {
        const uint32_t i = Lpos[first + k], j = Rasc[first + nR - 1 - k];
        const uint32_t t = keys[i];
        keys[i] = keys[j];
        keys[j] = t;
    }
    const int rk = (K > 0) ? (int)Rasc[first + nR - K] : last;
    int cut = rk;
    if (K < nL) {
        const int l = (int)Lpos[first + K];
        if (l < rk) cut = l;
    }
    __syncwarp();
    return cut;
}

// Frames with n_lo < n_raw <= n_hi are sorted by this launch (two launches cover all frames: see launch_sort_nms).
__global__ void __launch_bounds__(256) sort_kernel(SeqView s, int first_frame, int smem_keys, int n_lo, int n_hi) {
    extern __shared__ uint32_t sk[];
    __shared__ int qn[2];
    const int f = first_frame + blockIdx.x;
    const int n = s.n_raw[f];
    if (n <= n_lo || n > n_hi) return;
    uint32_t* gkeys = s.keys + (size_t)f * s.cap_raw;
    uint32_t* scratch = s.sort_scratch + (size_t)f * s.scratch_words;
    uint32_t* Lpos = scratch;
    uint32_t* Rasc = scratch + s.cap_raw;
    uint2* q[2] = {reinterpret_cast<uint2*>(scratch + 2 * (size_t)s.cap_raw),
                   reinterpret_cast<uint2*>(scratch + 2 * (size_t)s.cap_raw) + s.qcap};
    uint32_t* bnd = scratch + 2 * (size_t)s.cap_raw + 4 * (size_t)s.qcap;
    const int bwords = (n >> 5) + 1;
    const bool in_smem = n <= smem_keys;
    uint32_t* keys = in_smem ? sk : gkeys;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (n <= 1) return;
    if (in_smem)
        for (int i = tid; i < n; i += blockDim.x) sk[i] = gkeys[i];
    for (int i = tid; i < bwords; i += blockDim.x) bnd[i] = (i == 0) ? 1u : 0u;
    if (tid == 0) {
        qn[0] = 0;
        qn[1] = 0;
        if (n > 16) {
            q[0][0] = make_uint2(0u, (uint32_t)n);
            qn[0] = 1;
        }
    }
    int depth = 2 * std_lg(n);
    int cur = 0;
    while (true) {
        __syncthreads();
        const int nseg = qn[cur];
        if (nseg == 0) break;
        __syncthreads();
        if (tid == 0) qn[cur ^ 1] = 0;
        __syncthreads();
        if (depth == 0) {
            // depth limit hit: std::__partial_sort(first, last, last) == heap sort of the range
            for (int sgi = tid; sgi < nseg; sgi += blockDim.x) {
                const uint2 sg = q[cur][sgi];
                std_partial_sort(keys + sg.x, (int)(sg.y - sg.x), (int)(sg.y - sg.x), KeyDesc());
            }
        } else {
            for (int sgi = warp; sgi < nseg; sgi += (blockDim.x >> 5)) {
                const uint2 sg = q[cur][sgi];
                const int cut = warp_partition(keys, (int)sg.x, (int)sg.y, Lpos, Rasc);
                if (lane == 0) {
                    atomicOr(&bnd[cut >> 5], 1u << (cut & 31));
                    if (cut - (int)sg.x > 16) q[cur ^ 1][atomicAdd(&qn[cur ^ 1], 1)] = make_uint2(sg.x, (uint32_t)cut);
                    if ((int)sg.y - cut > 16) q[cur ^ 1][atomicAdd(&qn[cur ^ 1], 1)] = make_uint2((uint32_t)cut, sg.y);
                }
            }
        }
        depth--;
        cur ^= 1;
    }
    __syncthreads();
    // final insertion sort, one thread per boundary-delimited sub-range
    for (int w = tid; w < bwords; w += blockDim.x) {
        uint32_t bits = bnd[w];
        while (bits) {
            const int b = __ffs(bits) - 1;
            bits &= bits - 1;
            const int st = w * 32 + b;
            if (st >= n) break;
            // end = next boundary after st
            int en = n;
            {
                uint32_t rest = bits;  // remaining higher bits of this word
                int ww = w;
                while (true) {
                    if (rest) {
                        en = min(n, ww * 32 + __ffs(rest) - 1);
                        break;
                    }
                    ww++;
                    if (ww >= bwords) break;
                    rest = bnd[ww];
                }
            }
            for (int i = st + 1; i < en; i++) {
                const uint32_t v = keys[i];
                int j = i;
                while (j > st && (v >> kKeyIdxBits) > (keys[j - 1] >> kKeyIdxBits)) {
                    keys[j] = keys[j - 1];
                    j--;
                }
                keys[j] = v;
            }
        }
    }
    __syncthreads();
    if (in_smem)
        for (int i = tid; i < n; i += blockDim.x) gkeys[i] = sk[i];
}

constexpr int kHwTab = 512;
// largest |dx| with sqrt(dx^2+dy^2) < window under the reference's float arithmetic (-1: none)
__device__ int disc_half_width(int dyi, int window, float fw) {
    const float fdy = (float)dyi;
    int hw = window - 1;
    while (hw >= 0) {
        const float fdx = (float)hw;
        if (sqrtf((fdx * fdx) + (fdy * fdy)) < fw) break;
        hw--;
    }
    return hw;
}

// Greedy radius suppression over the sorted list; one warp per frame.
__global__ void __launch_bounds__(32) nms_kernel(SeqView s, int first_frame, int window, int use_smem, int only_flagged) {
    extern __shared__ uint32_t sbits[];
    const int f = first_frame + blockIdx.x;
    if (only_flagged && s.n_kp[f] != -1) return;  // already done by nms_parallel_kernel
    const int n = s.n_raw[f];
    const unsigned lane = lane_id();
    const uint32_t* keys = s.keys + (size_t)f * s.cap_raw;
    const uint32_t* xy = s.raw_xy + (size_t)f * s.cap_raw;
    slamcu_keypoint* out = s.kps + (size_t)f * s.cap_kp;
    const int bm_words = s.rows * s.mwords;
    uint32_t* bm = use_smem ? sbits : s.nms_bitmap + (size_t)f * bm_words;
    for (int i = lane; i < bm_words; i += 32) bm[i] = 0u;
    __syncwarp();
    const float fw = (float)window;
    const int nrows = 2 * window - 1;  // dy in (-window, window)
    // half-width of the suppression disc per row offset, by the reference's float predicate
    __shared__ short hwtab[kHwTab];
    for (int r = lane; r < min(nrows, kHwTab); r += 32) hwtab[r] = (short)disc_half_width(r - (window - 1), window, fw);
    __syncwarp();
    int kept = 0;
    for (int base = 0; base < n; base += 32) {
        const int i = base + (int)lane;
        int x = 0, y = 0, score = 0;
        bool alive = false;
        if (i < n) {
            const uint32_t k = keys[i];
            const uint32_t p = xy[k & kKeyIdxMask];
            x = p & 0xffff;
            y = p >> 16;
            score = k >> kKeyIdxBits;
            alive = !((bm[y * s.mwords + (x >> 5)] >> (x & 31)) & 1u);
        }
        unsigned am = __ballot_sync(0xffffffffu, alive);
        unsigned keepm = 0;
        while (am) {
            const int leader = __ffs(am) - 1;
            const int lx = __shfl_sync(0xffffffffu, x, leader), ly = __shfl_sync(0xffffffffu, y, leader);
            keepm |= 1u << leader;
            // feature_detector.cpp:177-183 (coordinates are integral floats)
            const float dx = (float)(lx - x), dy = (float)(ly - y);
            const bool sup = sqrtf((dx * dx) + (dy * dy)) < fw;
            am &= ~(1u << leader);
            am &= ~__ballot_sync(0xffffffffu, sup);
        }
        if ((keepm >> lane) & 1u) {
            const int pos = kept + __popc(keepm & lanemask_lt());
            if (pos < s.cap_kp) {
                slamcu_keypoint kp;
                kp.x = (float)x;
                kp.y = (float)y;
                kp.size = 6.0f;
                kp.angle = 0.0f;
                kp.response = (float)score;
                out[pos] = kp;
            }
        }
        const int nk = __popc(keepm);
        kept += nk;
        // stamp the suppression discs of the survivors (only later chunks read them)
        if (base + 32 < n) {
            const int total = nk * nrows;
            for (int t0 = 0; t0 < total; t0 += 32) {
                const int t = t0 + (int)lane;
                const int ki = (t < total) ? t / nrows : 0;
                const int src = __fns(keepm, 0, ki + 1);
                const int cx = __shfl_sync(0xffffffffu, x, src & 31), cy = __shfl_sync(0xffffffffu, y, src & 31);
                if (t < total) {
                    const int dyi = (t - ki * nrows) - (window - 1);
                    const int yy = cy + dyi;
                    if (yy >= 0 && yy < s.rows) {
                        const int ri = dyi + window - 1;
                        const int hw = (ri < kHwTab) ? (int)hwtab[ri] : disc_half_width(dyi, window, fw);
                        if (hw >= 0) {
                            const int xa = max(cx - hw, 0), xb = min(cx + hw, s.cols - 1);
                            for (int w = xa >> 5; w <= (xb >> 5); w++) {
                                const int lo = max(xa, w * 32) - w * 32, hi = min(xb, w * 32 + 31) - w * 32;
                                const uint32_t m = (hi == 31 ? 0xffffffffu : ((1u << (hi + 1)) - 1u)) & ~((1u << lo) - 1u);
                                atomicOr(&bm[yy * s.mwords + w], m);
                            }
                        }
                    }
                }
            }
        }
        __syncwarp();
    }
    if (lane == 0) {
        s.n_kp[f] = min(kept, s.cap_kp);
        if (kept > s.cap_kp) atomicOr(&s.status[f], kStKpOverflow);
    }
}

// ---- exact greedy NMS, parallel form ------------------------------------------------------------------------------
// The greedy walk of feature_detector.cpp:165-185 is the lexicographically-first maximal independent set of the
// "closer than window" graph in sorted order.  It is computed here as a fixed point: a corner is KEPT once every
// earlier-ranked neighbour is known to be suppressed, and SUPPRESSED once some earlier-ranked neighbour is known to be
// kept; decided states are final, so by induction on the rank the result equals the sequential walk.
//   * kept corners stamp their suppression disc into a 1-bit/pixel bitmap (the reference's float predicate decides the
//     half width of every disc row), so "some earlier neighbour is kept" is one bit test.  (A kept corner that covers a
//     still-undecided corner r is necessarily earlier: a later one could only be kept after r was suppressed.)
//   * "every earlier neighbour is decided" scans a uniform grid (cell >= window, 3x3 cells cover the disc) whose
//     per-cell lists are sorted by rank, and stops at the first earlier neighbour that is still undecided.
// Everything lives in shared memory, one block per frame.  Frames that do not fit (n_kp = -1) are left to the
// one-warp walker above.
constexpr int kNmsThreads = 1024;
enum : uint8_t { kUnknown = 0, kKept = 1, kSuppressed = 2 };

__global__ void __launch_bounds__(kNmsThreads) nms_parallel_kernel(SeqView s, int first_frame, int window, int smem_bytes) {
    extern __shared__ __align__(16) uint8_t nsm[];
    __shared__ int sh_scan[kNmsThreads / 32];
    __shared__ int sh_carry;
    __shared__ short hwtab[kHwTab];
    const int f = first_frame + blockIdx.x;
    const int n = s.n_raw[f];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (n <= 0) {
        if (tid == 0) s.n_kp[f] = 0;
        return;
    }
    const int bm_words = s.rows * s.mwords;
    const int nrows = 2 * window - 1;
    // grid geometry: the smallest cell >= window whose table fits next to the bitmap and the per-corner arrays
    int cell = window;
    int gw = (s.cols + cell - 1) / cell, gh = (s.rows + cell - 1) / cell;
    const long long per_corner = 4 + 2 + 1;  // position, cell-list entry (u16), state
    const long long avail_cells = ((long long)smem_bytes - 4LL * bm_words - per_corner * n - 64) / 8 - 2;
    while ((long long)gw * gh > avail_cells && cell < 4096) {
        cell += max(cell / 4, 1);
        gw = (s.cols + cell - 1) / cell;
        gh = (s.rows + cell - 1) / cell;
    }
    const int ncell = gw * gh;
    if (n >= 65536 || (long long)ncell > avail_cells || nrows > kHwTab) {  // does not fit: the sequential walker takes it
        if (tid == 0) s.n_kp[f] = -1;
        return;
    }
    uint32_t* bm = reinterpret_cast<uint32_t*>(nsm);                        // [rows][mwords] suppressed pixels
    uint32_t* pxy = bm + bm_words;                                          // [n]  (y << 16) | x by rank
    int* cstart = reinterpret_cast<int*>(pxy + n);                          // [ncell + 1]
    int* cfill = cstart + ncell + 1;                                         // [ncell]
    uint16_t* clist = reinterpret_cast<uint16_t*>(cfill + ncell);           // [n] ranks grouped by cell, ascending
    uint8_t* state = reinterpret_cast<uint8_t*>(clist + n + (n & 1));       // [n]
    const uint32_t* keys = s.keys + (size_t)f * s.cap_raw;
    const uint32_t* xy = s.raw_xy + (size_t)f * s.cap_raw;
    const float fw = (float)window;
    for (int i = tid; i < bm_words; i += kNmsThreads) bm[i] = 0u;
    for (int i = tid; i <= ncell; i += kNmsThreads) cstart[i] = 0;
    for (int r = tid; r < nrows; r += kNmsThreads) hwtab[r] = (short)disc_half_width(r - (window - 1), window, fw);
    __syncthreads();
    for (int r = tid; r < n; r += kNmsThreads) {
        const uint32_t p = xy[keys[r] & kKeyIdxMask];
        pxy[r] = p;
        state[r] = kUnknown;
        atomicAdd(&cstart[((p >> 16) / cell) * gw + (p & 0xffff) / cell], 1);
    }
    __syncthreads();
    // exclusive scan of the cell counts
    if (tid == 0) sh_carry = 0;
    __syncthreads();
    for (int base = 0; base < ncell; base += kNmsThreads) {
        const int i = base + tid;
        const int v = i < ncell ? cstart[i] : 0;
        int inc = v;
#pragma unroll
        for (int q = 1; q < 32; q <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, inc, q);
            if (lane >= q) inc += t;
        }
        if (lane == 31) sh_scan[warp] = inc;
        __syncthreads();
        int off = sh_carry;
        for (int w = 0; w < warp; w++) off += sh_scan[w];
        if (i < ncell) {
            cstart[i] = off + inc - v;
            cfill[i] = off + inc - v;
        }
        __syncthreads();
        if (tid == kNmsThreads - 1) sh_carry = off + inc;
        __syncthreads();
    }
    if (tid == 0) cstart[ncell] = n;
    __syncthreads();
    for (int r = tid; r < n; r += kNmsThreads) {
        const uint32_t p = pxy[r];
        clist[atomicAdd(&cfill[((p >> 16) / cell) * gw + (p & 0xffff) / cell], 1)] = (uint16_t)r;
    }
    __syncthreads();
    for (int c = tid; c < ncell; c += kNmsThreads) {  // ascending rank inside every cell (insertion sort, short lists)
        const int a = cstart[c], b = cstart[c + 1];
        for (int i = a + 1; i < b; i++) {
            const uint16_t v = clist[i];
            int j = i - 1;
            while (j >= a && clist[j] > v) { clist[j + 1] = clist[j]; j--; }
            clist[j + 1] = v;
        }
    }
    __syncthreads();
    // fixed-point rounds
    for (;;) {
        int pending = 0;
        for (int r = tid; r < n; r += kNmsThreads) {
            if (state[r] != kUnknown) continue;
            const uint32_t p = pxy[r];
            const int x = p & 0xffff, y = p >> 16;
            if ((bm[y * s.mwords + (x >> 5)] >> (x & 31)) & 1u) {
                state[r] = kSuppressed;
                continue;
            }
            const int cx = x / cell, cy = y / cell;
            bool wait = false, dead = false;
            for (int ny = max(cy - 1, 0); ny <= min(cy + 1, gh - 1) && !(dead | wait); ny++)
                for (int nx = max(cx - 1, 0); nx <= min(cx + 1, gw - 1) && !(dead | wait); nx++) {
                    const int c = ny * gw + nx;
                    for (int k = cstart[c]; k < cstart[c + 1]; k++) {
                        const int j = clist[k];
                        if (j >= r) break;  // ascending: only earlier ranks matter
                        const uint8_t sj = state[j];
                        if (sj == kSuppressed) continue;
                        const uint32_t q = pxy[j];
                        // feature_detector.cpp:177-183 (coordinates are integral floats)
                        const float dx = (float)((int)(q & 0xffff) - x), dy = (float)((int)(q >> 16) - y);
                        if (!(sqrtf((dx * dx) + (dy * dy)) < fw)) continue;
                        if (sj == kKept) dead = true;  // kept this round, its disc is not stamped yet
                        else wait = true;
                        break;
                    }
                }
            if (dead) {
                state[r] = kSuppressed;
            } else if (wait) {
                pending = 1;
            } else {
                state[r] = kKept;
                for (int ri = 0; ri < nrows; ri++) {  // stamp the suppression disc
                    const int yy = y + ri - (window - 1);
                    const int hw = hwtab[ri];
                    if (yy < 0 || yy >= s.rows || hw < 0) continue;
                    const int xa = max(x - hw, 0), xb = min(x + hw, s.cols - 1);
                    for (int w = xa >> 5; w <= (xb >> 5); w++) {
                        const int lo = max(xa, w * 32) - w * 32, hi = min(xb, w * 32 + 31) - w * 32;
                        const uint32_t m = (hi == 31 ? 0xffffffffu : ((1u << (hi + 1)) - 1u)) & ~((1u << lo) - 1u);
                        atomicOr(&bm[yy * s.mwords + w], m);
                    }
                }
            }
        }
        if (!__syncthreads_or(pending)) break;
    }
    // ordered emission of the kept corners
    slamcu_keypoint* out = s.kps + (size_t)f * s.cap_kp;
    if (tid == 0) sh_carry = 0;
    __syncthreads();
    for (int base = 0; base < n; base += kNmsThreads) {
        const int r = base + tid;
        const bool keep = r < n && state[r] == kKept;
        const unsigned b = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) sh_scan[warp] = __popc(b);
        __syncthreads();
        int off = sh_carry;
        for (int w = 0; w < warp; w++) off += sh_scan[w];
        const int pos = off + __popc(b & lanemask_lt());
        if (keep && pos < s.cap_kp) {
            const uint32_t p = pxy[r];
            slamcu_keypoint kp;
            kp.x = (float)(p & 0xffff);
            kp.y = (float)(p >> 16);
            kp.size = 6.0f;
            kp.angle = 0.0f;
            kp.response = (float)(keys[r] >> kKeyIdxBits);
            out[pos] = kp;
        }
        __syncthreads();
        if (tid == 0) {
            int tot = 0;
            for (int w = 0; w < kNmsThreads / 32; w++) tot += sh_scan[w];
            sh_carry += tot;
        }
        __syncthreads();
    }
    if (tid == 0) {
        const int kept = sh_carry;
        s.n_kp[f] = min(kept, s.cap_kp);
        if (kept > s.cap_kp) atomicOr(&s.status[f], kStKpOverflow);
    }
}

}  // namespace

// per-device opt-in to large dynamic shared memory (called from slamcu_create)
void init_sortnms_attributes(int smem_optin) {
    cudaFuncSetAttribute(sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_optin - 1024);
    cudaFuncSetAttribute(nms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_optin - 1024);
    cudaFuncSetAttribute(nms_parallel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_optin - 4096);
}

int launch_sort_nms(const SeqView& s, int first, int n, const DetParams& p, int smem_optin, cudaStream_t st) {
    // keys in shared memory when the frame's raw-corner count fits; else the same code runs on HBM.  The sort is one
    // latency-bound block per frame, so the shared-memory request decides how many frames an SM sorts at once: a first
    // launch takes the frames that fit a third of the SM's shared memory (3 blocks / SM), a second one the rest
    const int smem_keys = min(s.cap_raw, (smem_optin - 1024) / 4);
    const int small_keys = min(smem_keys, (smem_optin / 3 - 2048) / 4);
    int launches = 0;
    if (small_keys > 0 && small_keys < smem_keys) {
        SLAM_KERNEL("sort", st, sort_kernel<<<n, 256, (size_t)small_keys * 4, st>>>(s, first, small_keys, 0, small_keys));
        SLAM_KERNEL("sort_large", st, sort_kernel<<<n, 256, (size_t)smem_keys * 4, st>>>(s, first, smem_keys, small_keys, INT_MAX));
        launches = 2;
    } else {
        SLAM_KERNEL("sort", st, sort_kernel<<<n, 256, (size_t)smem_keys * 4, st>>>(s, first, smem_keys, 0, INT_MAX));
        launches = 1;
    }
    const size_t bm_bytes = (size_t)s.rows * s.mwords * 4;
    const int use_smem = bm_bytes <= (size_t)(smem_optin - 1024);
    const int par_smem = smem_optin - 4096;  // leaves room for the kernel's static shared memory
    SLAM_KERNEL("nms", st, nms_parallel_kernel<<<n, kNmsThreads, par_smem, st>>>(s, first, p.window, par_smem));
    SLAM_KERNEL("nms_fallback", st, nms_kernel<<<n, 32, use_smem ? bm_bytes : 0, st>>>(s, first, p.window, use_smem, 1));
    return launches + 2;
}

}  // namespace slamcu
