// pack.cu -- result compaction for the end-to-end path.
//
// The per-frame result arrays live in HBM as [F][cap_kp] rows (keypoints, descriptors, matches).  Copying them to the host
// as they are moves cap_kp rows per frame although only n_kp (n_match) are defined: 40 % of the D2H bytes of the headline
// configuration were padding.  dense_scan_kernel turns the per-frame device counts into running offsets, dense_copy_kernel
// writes the defined rows back to back into dense DEVICE buffers; the host fetches them with exact-size copy-engine
// transfers once the totals have reached it (slamcu_sequence_wait).  Round 2 first wrote the rows straight into mapped host
// memory from the SMs: the posted writes queued behind the link and cost the concurrent compute kernels ~10 %.
//   pack_counts_kernel: the four per-frame count arrays -> one int32[n][4] device array (the only data the multi-GPU path
//   exchanges: one NCCL all-gather per step).
#include <cstdlib>

#include "common.cuh"

namespace slamcu {
namespace {

// off[first + i + 1] = off[first] + counts[first] + ... + counts[first + i]; off[0] must have been zeroed by the caller
__global__ void __launch_bounds__(256) dense_scan_kernel(const int* __restrict__ counts, int first, int n, int* __restrict__ off) {
    __shared__ int warp_tot[8];
    __shared__ int carry;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry = off[first];
    __syncthreads();
    for (int base = 0; base < n; base += 256) {
        const int i = base + threadIdx.x;
        const int v = i < n ? counts[first + i] : 0;
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
        if (i < n) off[first + i + 1] = woff + inc;
        __syncthreads();
        if (threadIdx.x == 255) carry = woff + inc;
        __syncthreads();
    }
}

// rows [off[f], off[f] + counts[f]) of the dense destination <- the first counts[f] rows of frame f's block
template <int kWordsPerRow>
__device__ __forceinline__ void copy_rows(const uint32_t* __restrict__ src, uint32_t* __restrict__ dst, int n_rows, int part, int parts) {
    const long long words = (long long)n_rows * kWordsPerRow;
    if ((kWordsPerRow & 3) == 0) {  // 16-byte rows: 128-bit loads and stores (both sides are 16-byte aligned)
        const uint4* s4 = reinterpret_cast<const uint4*>(src);
        uint4* d4 = reinterpret_cast<uint4*>(dst);
        for (long long i = (long long)part * blockDim.x + threadIdx.x; i < words / 4; i += (long long)parts * blockDim.x) d4[i] = s4[i];
    } else {
        for (long long i = (long long)part * blockDim.x + threadIdx.x; i < words; i += (long long)parts * blockDim.x) dst[i] = src[i];
    }
}

// A persistent grid of kDenseBlocks blocks.  Work items: (frame, keypoints + descriptors) and (pair, matches), strided over
// the blocks.
constexpr int kDenseBlocks = 4 * 148;
__global__ void __launch_bounds__(256) dense_copy_kernel(SeqView s, int first, int n_frames, int pair_first, int n_pairs, const int* __restrict__ kp_off,
                                                         const int* __restrict__ m_off, uint32_t* __restrict__ h_kps, uint32_t* __restrict__ h_desc,
                                                         uint32_t* __restrict__ h_matches, int kp_cap, int m_cap, int* __restrict__ overflow) {
    for (int item = blockIdx.x; item < n_frames + n_pairs; item += gridDim.x) {
        if (item < n_frames) {
            const int f = first + item;
            const int n = s.n_kp[f], base = kp_off[f];
            if (base + n > kp_cap) {
                if (threadIdx.x == 0) atomicOr(overflow, 1);
                continue;
            }
            SLAMCU_BOUND(base, kp_cap + 1);
            SLAMCU_BOUND(n, s.cap_kp + 1);
            if (h_kps) copy_rows<5>(reinterpret_cast<const uint32_t*>(s.kps + (size_t)f * s.cap_kp), h_kps + (size_t)base * 5, n, 0, 1);
            if (h_desc) {
                const uint32_t* src = s.desc + (size_t)f * s.cap_kp * s.desc_words;
                uint32_t* dst = h_desc + (size_t)base * s.desc_words;
                const long long words = (long long)n * s.desc_words;
                if ((s.desc_words & 3) == 0) {
                    const uint4* s4 = reinterpret_cast<const uint4*>(src);
                    uint4* d4 = reinterpret_cast<uint4*>(dst);
                    for (long long i = threadIdx.x; i < words / 4; i += blockDim.x) d4[i] = s4[i];
                } else {
                    for (long long i = threadIdx.x; i < words; i += blockDim.x) dst[i] = src[i];
                }
            }
        } else if (h_matches) {
            const int p = pair_first + item - n_frames;
            const int n = s.n_match[p], base = m_off[p];
            if (base + n > m_cap) {
                if (threadIdx.x == 0) atomicOr(overflow, 2);
                continue;
            }
            copy_rows<3>(reinterpret_cast<const uint32_t*>(s.matches + (size_t)p * s.cap_kp), h_matches + (size_t)base * 3, n, 0, 1);
        }
    }
}

__global__ void pack_counts_kernel(SeqView s, int first, int n, int* __restrict__ out4) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int f = first + i;
    reinterpret_cast<int4*>(out4)[i] = make_int4(s.n_kp[f], s.n_match[f], s.n_raw[f], s.status[f]);
}

}  // namespace

int launch_dense_scan(const int* counts, int first, int n, int* off, cudaStream_t st) {
    if (n <= 0) return 0;
    SLAM_KERNEL("dense_scan", st, dense_scan_kernel<<<1, 256, 0, st>>>(counts, first, n, off));
    return 1;
}

int launch_dense_copy(const SeqView& s, int first, int n_frames, int pair_first, int n_pairs, const int* kp_off, const int* m_off, void* h_kps,
                      void* h_desc, void* h_matches, int kp_cap, int m_cap, int* overflow, cudaStream_t st) {
    if (n_frames + n_pairs <= 0) return 0;
    static const int blocks_env = [] { const char* e = getenv("SLAMCU_DENSE_BLOCKS"); return e ? atoi(e) : 0; }();  // tuning knob
    const int cap = blocks_env > 0 ? blocks_env : kDenseBlocks;
    const int grid = n_frames + n_pairs < cap ? n_frames + n_pairs : cap;
    SLAM_KERNEL("dense_copy", st,
                dense_copy_kernel<<<grid, 256, 0, st>>>(s, first, n_frames, pair_first, n_pairs, kp_off, m_off, static_cast<uint32_t*>(h_kps),
                                                        static_cast<uint32_t*>(h_desc), static_cast<uint32_t*>(h_matches), kp_cap, m_cap, overflow));
    return 1;
}

int launch_pack_counts(const SeqView& s, int first, int n, int* out4, cudaStream_t st) {
    if (n <= 0) return 0;
    SLAM_KERNEL("pack_counts", st, pack_counts_kernel<<<(n + 255) / 256, 256, 0, st>>>(s, first, n, out4));
    return 1;
}

}  // namespace slamcu
