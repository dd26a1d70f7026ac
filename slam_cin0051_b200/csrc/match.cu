// match.cu -- FeatureMatcher (rows A10-A13 of SURVEY.md section 8a; B9 reuses the k=2 search).
//
//   match_kernel    : findBestMatchesHamming's double loop (feature_matcher.cpp:153-173) as a popc-tiled
//                     all-pairs kernel on the integer pipe: one query per thread held in registers,
//                     train descriptors (+ keypoint xy) staged through shared memory with 128-bit
//                     loads and read back as warp-wide broadcasts, running top-2 per thread with the
//                     reference's strict-< tie rule (lowest train index wins).  calculateHammingDistance
//                     (common.hpp:40-50, byte table) == sum of __popc over 32-bit words.
//                     The image-distance penalty (:161-170) uses the same IEEE float ops, no FMA.
//   finalize_kernel : ratio test (:175-187), emission in query order, then filterAndSortMatches
//                     (:191-204) with libstdc++'s std::partial_sort / std::sort permutation (exact.cuh).
#include <cstdlib>

#include "common.cuh"
#include "exact.cuh"
#include "match_tc.cuh"

namespace slamcu {

namespace {

constexpr int QT = 128;  // queries per block (one per thread)
constexpr int TT = 128;  // train descriptors per shared-memory tile

struct Top2 {
    int best, second, bidx, sidx;
};

__device__ __forceinline__ void top2_update(Top2& t, int d, int j) {  // feature_matcher.cpp:132-141
    if (d < t.best) {
        t.second = t.best;
        t.sidx = t.bidx;
        t.best = d;
        t.bidx = j;
    } else if (d < t.second) {
        t.second = d;
        t.sidx = j;
    }
}

// Packed running top-2 for the penalty-free paths: key = (distance << 20) | train index, so integer order on the
// key is the reference's (distance, lowest index first) order; a tie with the best lands in second, exactly
// like the strict-< updateBestMatches (feature_matcher.cpp:132-141).  Needs nt <= 2^20 and distance < 2^11.
constexpr int kKeyShift = 20;
__device__ __forceinline__ void top2_update_key(uint32_t& best, uint32_t& second, uint32_t key) {
    second = min(second, max(best, key));
    best = min(best, key);
}

// 256-bit Hamming distance with a carry-save adder tree (Harley-Seal): seven of the eight XOR words are
// compressed to ones / twos / fours bit-planes by four full adders (two LOP3 each), so a comparison costs
// 4 POPC instead of 8.  POPC issues at a quarter of the LOP3 rate on sm_100, which makes the popc pipe the
// limiter of the plain form (ncu: sm__inst_executed_pipe_xu 89 %).
__device__ __forceinline__ int hamming256_csa(const uint32_t (&q)[8], const uint4& a, const uint4& b) {
    const uint32_t x0 = q[0] ^ a.x, x1 = q[1] ^ a.y, x2 = q[2] ^ a.z, x3 = q[3] ^ a.w;
    const uint32_t x4 = q[4] ^ b.x, x5 = q[5] ^ b.y, x6 = q[6] ^ b.z, x7 = q[7] ^ b.w;
    const uint32_t s1 = x0 ^ x1 ^ x2, c1 = (x0 & x1) | (x2 & (x0 | x1));
    const uint32_t s2 = x3 ^ x4 ^ x5, c2 = (x3 & x4) | (x5 & (x3 | x4));
    const uint32_t ones = s1 ^ s2 ^ x6, c3 = (s1 & s2) | (x6 & (s1 | s2));
    const uint32_t twos = c1 ^ c2 ^ c3, fours = (c1 & c2) | (c3 & (c1 | c2));
    return __popc(ones) + __popc(x7) + 2 * __popc(twos) + 4 * __popc(fours);
}

__device__ __forceinline__ int penalise(int dist, float qx, float qy, float tx, float ty) {
    // feature_matcher.cpp:162-169; MAX_JUMP_RADIUS = 500 (feature_matcher.hpp:12)
    const float dx = qx - tx, dy = qy - ty;
    const float id = sqrtf(dx * dx + dy * dy);
    if (id > 500.0f) {
        const float pen = 1.0f + (id / 500.0f);
        dist = (int)((float)dist * pen);
    }
    return dist;
}

// W = descriptor words handled by the unrolled path (8 = 256 bits); W = 0 -> generic loop.
// n_seg > 1: blockIdx.z selects a slice of the train set; each slice writes its own top-2 (cand + seg * seg_stride)
// and merge_segments_kernel combines them -- the global top-2 by (distance, index) is the top-2 of the slices' top-2s.
template <int W>
__global__ void __launch_bounds__(QT) match_kernel(MatchJob job, int with_kp, int n_seg, int seg_len, size_t seg_stride) {
    extern __shared__ __align__(16) uint32_t smem[];
    const int pair = blockIdx.y;
    const int nq = job.nq[(size_t)pair * job.count_stride];
    const int nt = job.nt[(size_t)pair * job.count_stride];
    const int q0 = blockIdx.x * QT;
    if (q0 >= nq || nt <= 0) return;
    const int dw = job.desc_words;
    const uint32_t* dq = job.dq + (size_t)pair * job.desc_pair_stride;
    const uint32_t* dt = job.dt + (size_t)pair * job.desc_pair_stride;
    const slamcu_keypoint* kq = with_kp ? job.kq + (size_t)pair * job.kp_pair_stride : nullptr;
    const slamcu_keypoint* kt = with_kp ? job.kt + (size_t)pair * job.kp_pair_stride : nullptr;
    uint32_t* tile = smem;                                     // [TT][dw]
    float2* txy = reinterpret_cast<float2*>(smem + TT * dw);   // [TT]

    // words that can be non-zero in either set (XOR of all-zero words contributes nothing)
    int wmax = dw;
    if (job.orq && job.ort) {
        const uint32_t* oq = job.orq + (size_t)pair * job.or_stride;
        const uint32_t* ot = job.ort + (size_t)pair * job.or_stride;
        while (wmax > 1 && (oq[wmax - 1] | ot[wmax - 1]) == 0u) wmax--;
    }

    const int q = q0 + threadIdx.x;
    const bool qok = q < nq;
    uint32_t qd[W > 0 ? W : 1];
    if (W > 0) {
#pragma unroll
        for (int w = 0; w < W; w++) qd[w] = qok ? dq[(size_t)q * dw + w] : 0u;
    }
    float qx = 0.f, qy = 0.f;
    if (with_kp && qok) {
        qx = kq[q].x;
        qy = kq[q].y;
    }
    Top2 t{INT_MAX, INT_MAX, -1, -1};
    // penalty-free, full-width 256-bit case: packed keys + carry-save popcount
    const bool packed = W == 8 && !with_kp && wmax > 2 && nt <= (1 << kKeyShift);
    uint32_t kbest = 0xffffffffu, ksecond = 0xffffffffu;

    const int t_begin = n_seg > 1 ? (int)blockIdx.z * seg_len : 0;
    const int t_end = n_seg > 1 ? min(nt, t_begin + seg_len) : nt;
    for (int t0 = t_begin; t0 < t_end; t0 += TT) {
        const int cnt = min(TT, t_end - t0);
        __syncthreads();
        if ((dw & 3) == 0) {
            const uint4* src = reinterpret_cast<const uint4*>(dt + (size_t)t0 * dw);
            uint4* dst = reinterpret_cast<uint4*>(tile);
            for (int v = threadIdx.x; v < cnt * (dw >> 2); v += QT) dst[v] = __ldg(src + v);
        } else {
            for (int v = threadIdx.x; v < cnt * dw; v += QT) tile[v] = __ldg(dt + (size_t)t0 * dw + v);
        }
        if (with_kp)
            for (int v = threadIdx.x; v < cnt; v += QT) txy[v] = make_float2(kt[t0 + v].x, kt[t0 + v].y);
        __syncthreads();
        if (!qok) continue;
        if (W == 8 && packed) {
            uint32_t qq[8];
#pragma unroll
            for (int w = 0; w < 8; w++) qq[w] = qd[W > 0 ? (w < W ? w : 0) : 0];
#pragma unroll 4
            for (int j = 0; j < cnt; j++) {
                const uint4 a = *reinterpret_cast<const uint4*>(tile + j * 8);
                const uint4 b = *reinterpret_cast<const uint4*>(tile + j * 8 + 4);
                const int d = hamming256_csa(qq, a, b);
                top2_update_key(kbest, ksecond, ((uint32_t)d << kKeyShift) + (uint32_t)(t0 + j));
            }
        } else if (W == 8 && wmax <= 2) {
            for (int j = 0; j < cnt; j++) {
                const uint2 a = *reinterpret_cast<const uint2*>(tile + j * 8);
                int d = __popc(qd[0] ^ a.x) + __popc(qd[W > 1 ? 1 : 0] ^ a.y);
                if (with_kp && d < t.second) d = penalise(d, qx, qy, txy[j].x, txy[j].y);  // the penalty only grows d: skipped when d cannot enter the top 2
                top2_update(t, d, t0 + j);
            }
        } else if (W == 8) {
#pragma unroll 2
            for (int j = 0; j < cnt; j++) {
                const uint4 a = *reinterpret_cast<const uint4*>(tile + j * 8);
                const uint4 b = *reinterpret_cast<const uint4*>(tile + j * 8 + 4);
                int d = __popc(qd[0] ^ a.x) + __popc(qd[W > 1 ? 1 : 0] ^ a.y) + __popc(qd[W > 2 ? 2 : 0] ^ a.z) +
                        __popc(qd[W > 3 ? 3 : 0] ^ a.w) + __popc(qd[W > 4 ? 4 : 0] ^ b.x) +
                        __popc(qd[W > 5 ? 5 : 0] ^ b.y) + __popc(qd[W > 6 ? 6 : 0] ^ b.z) +
                        __popc(qd[W > 7 ? 7 : 0] ^ b.w);
                if (with_kp && d < t.second) d = penalise(d, qx, qy, txy[j].x, txy[j].y);  // the penalty only grows d: skipped when d cannot enter the top 2
                top2_update(t, d, t0 + j);
            }
        } else {
            for (int j = 0; j < cnt; j++) {
                int d = 0;
                for (int w = 0; w < wmax; w++) d += __popc(dq[(size_t)q * dw + w] ^ tile[j * dw + w]);
                if (with_kp && d < t.second) d = penalise(d, qx, qy, txy[j].x, txy[j].y);  // the penalty only grows d: skipped when d cannot enter the top 2
                top2_update(t, d, t0 + j);
            }
        }
    }
    if (packed) {
        if (kbest != 0xffffffffu) { t.best = (int)(kbest >> kKeyShift); t.bidx = (int)(kbest & ((1u << kKeyShift) - 1)); }
        if (ksecond != 0xffffffffu) { t.second = (int)(ksecond >> kKeyShift); t.sidx = (int)(ksecond & ((1u << kKeyShift) - 1)); }
    }
    if (qok) job.cand[(size_t)blockIdx.z * seg_stride + (size_t)pair * job.cand_pair_stride + q] = make_int4(t.bidx, t.best, t.second, t.sidx);
}

// ---- the penalty-free 256-bit path (cv::BFMatcher::knnMatch(k = 2) / FeatureMatcher::match without keypoints) -------
// One block = 128 threads x 2 queries held in registers; the train set is staged through shared memory in tiles of 512
// descriptors (16 KB; up to 2048 through SLAMCU_MATCH_TILE), one barrier pair per 1024 comparisons of a thread (the
// 128-descriptor tiles of match_kernel paid one per 128: 15 % of its stall samples).
// Per train descriptor two broadcast LDS.128 serve both queries; per comparison: 8 XOR + 8 LOP3 (carry-save tree) on the ALU
// pipe, 4 POPC on the XU pipe, the weighted sum and the (distance << 20 | index) key as integer multiply-adds (FMA pipe),
// and the running top-2 as min / max on packed keys.  Instruction budget per comparison ~28 (was 32).
constexpr int MTHR = 128;
constexpr int kMaxTile256 = 2048;   // largest tile the kernel may be given (64 KB)
constexpr int kDefTile256 = 512;    // default: 16 KB per block -> nine 4-warp blocks per SM (measured: 2048 -> 780, 1024 -> 841, 512 -> 848, 256 -> 847 Gcmp/s)

__device__ __forceinline__ uint32_t mad_u32(uint32_t a, uint32_t b, uint32_t c) {  // IMAD: runs on the FMA pipe, not the ALU pipe
    uint32_t r;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}

// full adder on bit planes with explicit LOP3 truth tables (sum = a ^ b ^ c: 0x96, carry = majority: 0xE8): one instruction
// each.  Written as inline PTX so that the comparison costs exactly 8 XOR + 8 LOP3: left to itself ptxas re-associates the
// query / train XORs into the adders' inputs and ends up with 22 LOP3 per comparison when two queries share the train words.
__device__ __forceinline__ uint32_t lop3_sum(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t r;
    asm("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}
__device__ __forceinline__ uint32_t lop3_maj(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t r;
    asm("lop3.b32 %0, %1, %2, %3, 0xE8;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}

__device__ __forceinline__ uint32_t hamming256_key(const uint32_t (&q)[8], const uint4& a, const uint4& b, uint32_t idx, uint32_t w1, uint32_t w2,
                                                   uint32_t w4) {
    const uint32_t x0 = q[0] ^ a.x, x1 = q[1] ^ a.y, x2 = q[2] ^ a.z, x3 = q[3] ^ a.w;
    const uint32_t x4 = q[4] ^ b.x, x5 = q[5] ^ b.y, x6 = q[6] ^ b.z, x7 = q[7] ^ b.w;
    const uint32_t s1 = lop3_sum(x0, x1, x2), c1 = lop3_maj(x0, x1, x2);
    const uint32_t s2 = lop3_sum(x3, x4, x5), c2 = lop3_maj(x3, x4, x5);
    const uint32_t ones = lop3_sum(s1, s2, x6), c3 = lop3_maj(s1, s2, x6);
    const uint32_t twos = lop3_sum(c1, c2, c3), fours = lop3_maj(c1, c2, c3);
    // key = (popc(ones) + popc(x7) + 2 popc(twos) + 4 popc(fours)) << 20 | idx, as four multiply-adds; the weights come in
    // registers (w1 = 1 << 20 is a kernel argument) so that ptxas keeps IMADs instead of turning them into ALU-pipe LEAs
    uint32_t k = mad_u32((uint32_t)__popc(ones), w1, idx);
    k = mad_u32((uint32_t)__popc(x7), w1, k);
    k = mad_u32((uint32_t)__popc(twos), w2, k);
    return mad_u32((uint32_t)__popc(fours), w4, k);
}

template <int MQ>  // queries per thread: 2 for batches and large problems, 1 for small single problems (twice the blocks)
__global__ void __launch_bounds__(MTHR) match256_kernel(MatchJob job, int n_seg, int seg_len, size_t seg_stride, int tile_cap, uint32_t w1) {
    constexpr int MQB = MQ * MTHR;
    extern __shared__ __align__(16) uint32_t smem[];
    uint4* tile = reinterpret_cast<uint4*>(smem);  // [tile_cap][2]
    const int pair = blockIdx.y;
    const int nq = job.nq[(size_t)pair * job.count_stride];
    const int nt = job.nt[(size_t)pair * job.count_stride];
    const int q0 = blockIdx.x * MQB;
    if (q0 >= nq || nt <= 0) return;
    const uint32_t* dq = job.dq + (size_t)pair * job.desc_pair_stride;
    const uint4* dt = reinterpret_cast<const uint4*>(job.dt + (size_t)pair * job.desc_pair_stride);
    // words that can be non-zero in either set: the reference's own 46-bit descriptors populate two words
    int wmax = 8;
    if (job.orq && job.ort) {
        const uint32_t* oq = job.orq + (size_t)pair * job.or_stride;
        const uint32_t* ot = job.ort + (size_t)pair * job.or_stride;
        while (wmax > 1 && (oq[wmax - 1] | ot[wmax - 1]) == 0u) wmax--;
    }
    uint32_t qd[MQ][8];
    bool qok[MQ];
#pragma unroll
    for (int m = 0; m < MQ; m++) {
        const int q = q0 + m * MTHR + threadIdx.x;
        qok[m] = q < nq;
        const uint4* src = reinterpret_cast<const uint4*>(dq + (size_t)(qok[m] ? q : q0) * 8);
        const uint4 lo = __ldg(src), hi = __ldg(src + 1);
        qd[m][0] = lo.x; qd[m][1] = lo.y; qd[m][2] = lo.z; qd[m][3] = lo.w;
        qd[m][4] = hi.x; qd[m][5] = hi.y; qd[m][6] = hi.z; qd[m][7] = hi.w;
    }
    uint32_t kbest[MQ], ksecond[MQ];
#pragma unroll
    for (int m = 0; m < MQ; m++) kbest[m] = ksecond[m] = 0xffffffffu;
    const uint32_t w2 = w1 * 2u, w4 = w1 * 4u;

    const int t_begin = n_seg > 1 ? (int)blockIdx.z * seg_len : 0;
    const int t_end = n_seg > 1 ? min(nt, t_begin + seg_len) : nt;
    for (int t0 = t_begin; t0 < t_end; t0 += tile_cap) {
        const int cnt = min(tile_cap, t_end - t0);
        if (t0 != t_begin) __syncthreads();  // everyone is done with the previous tile
        {
            const uint4* src = dt + (size_t)t0 * 2;
            const int nv = cnt * 2;
            int v = threadIdx.x;
            for (; v + 7 * MTHR < nv; v += 8 * MTHR) {  // eight 128-bit loads in flight per thread
                uint4 r[8];
#pragma unroll
                for (int k = 0; k < 8; k++) r[k] = __ldg(src + v + k * MTHR);
#pragma unroll
                for (int k = 0; k < 8; k++) tile[v + k * MTHR] = r[k];
            }
            for (; v < nv; v += MTHR) tile[v] = __ldg(src + v);
        }
        __syncthreads();
        if (wmax > 2) {
            // two train descriptors per step: their keys are ordered first, then merged into the running top-2 with a 3-input
            // minimum -- five min / max per two comparisons instead of six
            int j = 0;
#pragma unroll 2
            for (; j + 1 < cnt; j += 2) {
                const uint4 a0 = tile[2 * j], b0 = tile[2 * j + 1], a1 = tile[2 * j + 2], b1 = tile[2 * j + 3];
#pragma unroll
                for (int m = 0; m < MQ; m++) {
                    const uint32_t k0 = hamming256_key(qd[m], a0, b0, (uint32_t)(t0 + j), w1, w2, w4);
                    const uint32_t k1 = hamming256_key(qd[m], a1, b1, (uint32_t)(t0 + j + 1), w1, w2, w4);
                    const uint32_t lo = min(k0, k1), hi = max(k0, k1);
                    ksecond[m] = __vimin3_u32(ksecond[m], max(kbest[m], lo), hi);
                    kbest[m] = min(kbest[m], lo);
                }
            }
            if (j < cnt) {
                const uint4 a = tile[2 * j], b = tile[2 * j + 1];
#pragma unroll
                for (int m = 0; m < MQ; m++) top2_update_key(kbest[m], ksecond[m], hamming256_key(qd[m], a, b, (uint32_t)(t0 + j), w1, w2, w4));
            }
        } else {
#pragma unroll 4
            for (int j = 0; j < cnt; j++) {
                const uint4 a = tile[2 * j];
#pragma unroll
                for (int m = 0; m < MQ; m++) {
                    const uint32_t d = (uint32_t)(__popc(qd[m][0] ^ a.x) + __popc(qd[m][1] ^ a.y));
                    top2_update_key(kbest[m], ksecond[m], mad_u32(d, w1, (uint32_t)(t0 + j)));
                }
            }
        }
    }
#pragma unroll
    for (int m = 0; m < MQ; m++) {
        if (!qok[m]) continue;
        int4 c = make_int4(-1, INT_MAX, INT_MAX, -1);
        if (kbest[m] != 0xffffffffu) { c.y = (int)(kbest[m] >> kKeyShift); c.x = (int)(kbest[m] & ((1u << kKeyShift) - 1)); }
        if (ksecond[m] != 0xffffffffu) { c.z = (int)(ksecond[m] >> kKeyShift); c.w = (int)(ksecond[m] & ((1u << kKeyShift) - 1)); }
        SLAMCU_BOUND(q0 + m * MTHR + threadIdx.x, job.max_q);
        SLAMCU_BOUND(c.x + 1, nt + 1);  // -1 = no candidate (an empty train slice)
        job.cand[(size_t)blockIdx.z * seg_stride + (size_t)pair * job.cand_pair_stride + q0 + m * MTHR + threadIdx.x] = c;
    }
}

__global__ void __launch_bounds__(256) merge_segments_kernel(MatchJob job, int n_seg, size_t seg_stride) {
    const int pair = blockIdx.y;
    const int nq = job.nq[(size_t)pair * job.count_stride];
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    Top2 t{INT_MAX, INT_MAX, -1, -1};
    for (int s = 0; s < n_seg; s++) {  // slices are in ascending index order, so strict < keeps the lowest index
        const int4 c = job.cand[(size_t)s * seg_stride + (size_t)pair * job.cand_pair_stride + q];
        if (c.x >= 0) top2_update(t, c.y, c.x);
        if (c.w >= 0) top2_update(t, c.z, c.w);
    }
    job.cand[(size_t)pair * job.cand_pair_stride + q] = make_int4(t.bidx, t.best, t.second, t.sidx);
}

struct MatchLess {  // sortPredicate: a.distance < b.distance (feature_matcher.cpp:192-194); key = dist<<32 | slot
    __device__ __forceinline__ bool operator()(unsigned long long a, unsigned long long b) const {
        return (uint32_t)(a >> 32) < (uint32_t)(b >> 32);
    }
};

// One block per pair: ratio test, ordered emission, optional top-K / sort with libstdc++ tie order.
__global__ void __launch_bounds__(256) finalize_kernel(MatchJob job, MatchParams p, unsigned long long* gkeys,
                                                       int smem_cap) {
    extern __shared__ __align__(16) unsigned long long skeys[];
    __shared__ int warp_tot[8];
    __shared__ int carry;
    const int pair = blockIdx.x;
    const int nq = job.nq[(size_t)pair * job.count_stride];
    const int nt = job.nt[(size_t)pair * job.count_stride];
    const int4* cand = job.cand + (size_t)pair * job.cand_pair_stride;
    slamcu_dmatch* out = job.matches + (size_t)pair * job.cand_pair_stride;
    unsigned long long* keys_g = gkeys + (size_t)pair * job.cand_pair_stride;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) carry = 0;
    __syncthreads();
    if (nq <= 0 || nt <= 0) {  // the reference throws std::invalid_argument here; the adapters do too
        if (tid == 0) job.n_match[(size_t)pair * job.count_stride] = 0;
        return;
    }
    const bool staged = p.filter != 0;  // with filtering, emit into a staging order first
    for (int base = 0; base < nq; base += 256) {
        const int q = base + tid;
        bool good = false;
        int4 c = make_int4(-1, 0, 0, -1);
        if (q < nq) {
            c = cand[q];
            good = c.x != -1;
            // Lowe ratio (feature_matcher.cpp:176-182); second == INT_MAX when nt == 1
            if (p.use_ratio && (float)c.y >= p.ratio * (float)c.z) good = false;
        }
        const unsigned b = __ballot_sync(0xffffffffu, good);
        const int wcount = __popc(b);
        if (lane == 0) warp_tot[warp] = wcount;
        __syncthreads();
        int off = carry;
        for (int w = 0; w < warp; w++) off += warp_tot[w];
        const int pos = off + __popc(b & lanemask_lt());
        if (good && pos < job.cap_out) {
            SLAMCU_BOUND(pos, job.cap_out);
            if (staged) keys_g[pos] = ((unsigned long long)(uint32_t)c.y << 32) | (uint32_t)q;
            else {
                slamcu_dmatch m;
                m.queryIdx = q;
                m.trainIdx = c.x;
                m.distance = (float)c.y;
                out[pos] = m;
            }
        }
        __syncthreads();
        if (tid == 0) {
            int tot = 0;
            for (int w = 0; w < 8; w++) tot += warp_tot[w];
            carry += tot;
        }
        __syncthreads();
    }
    int n = carry;
    if (n > job.cap_out) {
        if (tid == 0) atomicOr(&job.status[(size_t)pair * job.count_stride], kStMatchOverflow);
        n = job.cap_out;
    }
    if (!staged) {
        if (tid == 0) job.n_match[(size_t)pair * job.count_stride] = n;
        return;
    }
    // filterAndSortMatches: keys carry (distance, query index); the permutation is what matters.
    const bool in_smem = n <= smem_cap;
    unsigned long long* keys = in_smem ? skeys : keys_g;
    if (in_smem) {
        for (int i = tid; i < n; i += 256) skeys[i] = keys_g[i];
    }
    __syncthreads();
    int n_out = n;
    if (tid == 0) {
        if (n > p.good) {
            std_partial_sort(keys, p.good, n, MatchLess());
        } else {
            std_sort(keys, n, MatchLess());
        }
    }
    if (n > p.good) n_out = p.good;
    __syncthreads();
    for (int i = tid; i < n_out; i += 256) {
        const unsigned long long k = keys[i];
        const int q = (int)(uint32_t)k;
        slamcu_dmatch m;
        m.queryIdx = q;
        m.trainIdx = cand[q].x;
        m.distance = (float)(uint32_t)(k >> 32);
        out[i] = m;
    }
    if (tid == 0) job.n_match[(size_t)pair * job.count_stride] = n_out;
}

}  // namespace

void init_match_attributes() {
    cudaFuncSetAttribute(match256_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxTile256 * 32);
    cudaFuncSetAttribute(match256_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxTile256 * 32);
}

int launch_match(const MatchJob& job, int n_pairs, const MatchParams& p, bool emit_matches, int with_kp,
                 unsigned long long* sort_keys, cudaStream_t st, int n_seg, int seg_len, size_t seg_stride, const MatchTc* tc) {
    if (n_pairs <= 0) return 0;
    if (n_seg < 1) n_seg = 1;
    if (tc && job.desc_words == 8 && !with_kp && job.max_t <= (1 << kKeyShift) && (n_seg == 1 || seg_len % 128 == 0)) {
        launch_match_tc(job, *tc, n_pairs, st, n_seg, seg_len, seg_stride);
    } else if (job.desc_words == 8 && !with_kp && job.max_t <= (1 << kKeyShift)) {
        // 256-bit descriptors, no image-distance penalty: whole train slice in shared memory, two queries per thread
        const int span = n_seg > 1 ? seg_len : job.max_t;
        static const int tile_env = [] { const char* e = getenv("SLAMCU_MATCH_TILE"); return e ? atoi(e) : 0; }();  // tuning knob (descriptors per smem tile)
        const int tile_max = tile_env >= 64 && tile_env <= kMaxTile256 ? tile_env / 64 * 64 : kDefTile256;
        const int tile_cap = max(64, min(tile_max, (span + 63) / 64 * 64));
        // two queries per thread when that still leaves a few blocks per SM, else one (small single problems)
        const bool two = (long long)n_pairs * ((job.max_q + 2 * MTHR - 1) / (2 * MTHR)) * n_seg >= 4 * 148;
        const int qpb = (two ? 2 : 1) * MTHR;
        dim3 grid((job.max_q + qpb - 1) / qpb, n_pairs, n_seg);
        if (two)
            SLAM_KERNEL("match", st, match256_kernel<2><<<grid, MTHR, (size_t)tile_cap * 32, st>>>(job, n_seg, seg_len, seg_stride, tile_cap, 1u << kKeyShift));
        else
            SLAM_KERNEL("match", st, match256_kernel<1><<<grid, MTHR, (size_t)tile_cap * 32, st>>>(job, n_seg, seg_len, seg_stride, tile_cap, 1u << kKeyShift));
    } else {
        dim3 grid((job.max_q + QT - 1) / QT, n_pairs, n_seg);
        const size_t smem = (size_t)TT * job.desc_words * 4 + TT * sizeof(float2);
        if (job.desc_words == 8)
            SLAM_KERNEL("match", st, match_kernel<8><<<grid, QT, smem, st>>>(job, with_kp, n_seg, seg_len, seg_stride));
        else
            SLAM_KERNEL("match", st, match_kernel<0><<<grid, QT, smem, st>>>(job, with_kp, n_seg, seg_len, seg_stride));
    }
    int launches = 1;
    if (n_seg > 1) {
        SLAM_KERNEL("match_merge", st,
                    merge_segments_kernel<<<dim3((job.max_q + 255) / 256, n_pairs), 256, 0, st>>>(job, n_seg, seg_stride));
        launches++;
    }
    if (emit_matches) {
        const int smem_cap = 4096;
        SLAM_KERNEL("match_finalize", st, finalize_kernel<<<n_pairs, 256, smem_cap * sizeof(unsigned long long), st>>>(job, p, sort_keys, smem_cap));
        launches++;
    }
    return launches;
}

}  // namespace slamcu
