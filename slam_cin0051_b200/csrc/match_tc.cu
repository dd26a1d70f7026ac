// match_tc.cu -- the penalty-free 256-bit all-pairs Hamming search (cv::BFMatcher::knnMatch(k = 2), FeatureMatcher::match
// without keypoints: feature_matcher.cpp:153-173) on the 5th-generation tensor cores.
//
//   hamming(q, t) = popc(q) + popc(t) - 2 <q, t>      with <q, t> the dot product of the two descriptors as 0/1 vectors,
//
// so the n_q x n_t distance table of a frame pair is one integer GEMM with K = 256: the descriptors are widened to one byte
// per bit (expand_bits_kernel; 256 B per descriptor, K-major), a CTA owns 128 queries (UMMA M = 128 = the TMEM lanes) and
// walks the train set in tiles of 128 descriptors (UMMA N = 128):
//
//   warp 4, one lane   TMA producer: the query tile once, the train tiles through a two-stage ring (cp.async.bulk.tensor,
//                      128-byte swizzle, mbarrier complete_tx), each with its 128 packed key constants (cp.async.bulk)
//   warp 5, one lane   MMA issuer: 8 x tcgen05.mma.kind::i8 (128 x 128 x 32, u8 x u8 -> s32) per train tile into one of two
//                      128-column TMEM accumulator stages; tcgen05.commit releases the smem stage and publishes the accumulator
//   warps 0-3          epilogue: thread = query = TMEM lane.  tcgen05.ld 32 columns at a time; per element ONE integer
//                      multiply-add turns the dot product into the reference's packed key
//                          key = ((popc(t) + 512 - 2 <q,t>) << 20) | t_index     (popc(q) is constant per thread: added last)
//                      and the running top-2 of a thread is 2.5 min / max per element (pairs ordered first, 3-input merge).
//                      Integer order on the key is (distance, lowest index first): the strict-< rule of updateBestMatches
//                      (feature_matcher.cpp:132-141), ties included, exactly as match256_kernel.
//
// Everything is exact integer arithmetic (0/1 operands, s32 accumulators).  The integer-pipe kernels of match.cu remain for
// the image-distance-penalty path, descriptor widths other than 256 bits and tiny problems.
#include <cuda.h>

#include "common.cuh"
#include "match_tc.cuh"

namespace slamcu {

namespace {

constexpr int TCM = 128;                       // queries per CTA
constexpr int TCN = 128;                       // train descriptors per accumulator stage
constexpr int TCSTAGES = 2;                    // train tiles in flight in shared memory
constexpr uint32_t kKBlockBytes = 128;         // bytes of K per swizzled row (one 128-byte swizzle span)
constexpr uint32_t kTileBytes = TCM * kKBlockBytes;  // one (128 rows x 128 K-bytes) operand block: 16 KB
constexpr uint32_t kOperandBytes = 2 * kTileBytes;   // K = 256 bytes = two blocks
constexpr int kTcThreads = 192;
constexpr uint32_t kTmemCols = 2 * TCN;        // two accumulator stages
static_assert(TCN == TCM, "one tile geometry for both operands (one TMA box shape)");

// shared-memory carve-up (offsets from a 1024-byte aligned base: the 128-byte swizzle atoms are 1024 bytes)
constexpr uint32_t kOffA = 0;
constexpr uint32_t kOffB = kOffA + kOperandBytes;
constexpr uint32_t kOffKeys = kOffB + TCSTAGES * kOperandBytes;
constexpr int kKeySlots = 8;  // key constants of tile j live in slot j % 8.  Slot reuse: tile j + 8 is loaded after MMA j + 6 completed,
                             // which was issued after all four epilogue warps had drained tile j + 4 -- and a warp that has reached
                             // tile j + 4 is done with every read of tile j's keys.  (An epilogue warp releases the accumulator stage
                             // BEFORE its last fold of the tile, so four slots left that fold racing the refill: seen as wrong
                             // matches in the last wave of a launch when kernels of the other compute lane shared its SMs.)
constexpr uint32_t kOffBars = kOffKeys + kKeySlots * TCN * 4;
constexpr uint32_t kTcSmemBytes = kOffBars + 128 + 1024;  // barriers + TMEM base + alignment slack

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// bounded wait: a protocol error traps (the launch fails loudly) instead of hanging the device
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    const long long t0 = clock64();
    for (;;) {
        uint32_t ok;
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
        if (ok) return;
        if (clock64() - t0 > 4000000000LL) __trap();
    }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int x, int y, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
                 "l"(map), "r"(x), "r"(y), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes),
                 "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// shared-memory matrix descriptor, K-major, 128-byte swizzle (cute::UMMA::SmemDescriptor): start address >> 4 in bits
// [0,14), leading byte offset (unused for swizzled K-major: 1) in [16,30), stride byte offset = 8 rows x 128 B = 1024 >> 4
// in [32,46), version 1 in [46,48), layout type SWIZZLE_128B = 2 in [61,64)
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
           ((uint64_t)2 << 61);
}
// instruction descriptor (cute::UMMA::InstrDescriptor), kind::i8: D = s32 (2 << 4), A = B = unsigned 8 bit (0), both K-major,
// N >> 3 in [17,23), M >> 4 in [24,29)
constexpr uint32_t kIdesc = (2u << 4) | ((uint32_t)(TCN >> 3) << 17) | ((uint32_t)(TCM >> 4) << 24);

__device__ __forceinline__ void umma_i8(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(kIdesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// 32 lanes x 32 columns of 32 bits: thread = lane of the warp's TMEM quadrant, v[i] = column i
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, "
        "%19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
          "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
          "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
// wait for the outstanding tcgen05.ld, and pin the registers behind it: the empty volatile asm per register keeps the compiler
// from scheduling a consumer of v[] above the wait (the load is asynchronous; the data dependence alone does not say so)
__device__ __forceinline__ void tmem_ld_wait(uint32_t (&v)[32]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; i++) asm volatile("" : "+r"(v[i]));
}

__device__ __forceinline__ uint32_t mad_u32(uint32_t a, uint32_t b, uint32_t c) {  // IMAD: FMA pipe, not the ALU pipe
    uint32_t r;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}

// 32 dot products -> 32 keys -> running top-2 (pairs ordered first, then a 3-input merge: 5 min / max per two elements)
__device__ __forceinline__ void fold32(const uint32_t (&v)[32], const uint4* __restrict__ ck, uint32_t mul, uint32_t& best, uint32_t& second) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const uint4 c = ck[i];  // the same address for every thread: one broadcast LDS.128 per four columns
        const uint32_t k0 = mad_u32(v[4 * i], mul, c.x), k1 = mad_u32(v[4 * i + 1], mul, c.y);
        const uint32_t k2 = mad_u32(v[4 * i + 2], mul, c.z), k3 = mad_u32(v[4 * i + 3], mul, c.w);
        const uint32_t lo0 = min(k0, k1), hi0 = max(k0, k1);
        second = __vimin3_u32(second, max(best, lo0), hi0);
        best = min(best, lo0);
        const uint32_t lo1 = min(k2, k3), hi1 = max(k2, k3);
        second = __vimin3_u32(second, max(best, lo1), hi1);
        best = min(best, lo1);
    }
}

// ---- descriptors -> one byte per bit (K-major GEMM operand rows) + the per-row key constant ------------------------------
// set s: rows [0, n_s) of desc + s * set_stride_words; x8 / ck rows s * rows_pad + r.  Rows [n_s, round_up(n_s, 128)) are
// zeroed with ck = 0xffffffff (a key that never enters a top-2); rows beyond are never read.
// 16 threads per row, each 16 bits -> 16 bytes (one 128-bit store; a row's 256 bytes are contiguous).
__global__ void __launch_bounds__(256) expand_bits_kernel(const uint32_t* __restrict__ desc, size_t set_stride_words, const int* __restrict__ counts,
                                                           int count_stride, int rows_pad, uint8_t* __restrict__ x8, uint32_t* __restrict__ ck) {
    const int s = blockIdx.y;
    const int n = min(counts[(size_t)s * count_stride], rows_pad);
    const int n_up = min((n + 127) / 128 * 128, rows_pad);
    const int r = blockIdx.x * 16 + (threadIdx.x >> 4), part = threadIdx.x & 15;
    if (r >= n_up) return;
    uint32_t bits = 0;
    if (r < n) {
        const uint32_t w = __ldg(desc + (size_t)s * set_stride_words + (size_t)r * 8 + (part >> 1));
        bits = (part & 1) ? (w >> 16) : (w & 0xffffu);
    }
    uint4 o;  // bit b of a nibble -> byte b of a word: the partial products land on 16 distinct bit positions, no carries
    o.x = ((bits & 0xfu) * 0x00204081u) & 0x01010101u;
    o.y = (((bits >> 4) & 0xfu) * 0x00204081u) & 0x01010101u;
    o.z = (((bits >> 8) & 0xfu) * 0x00204081u) & 0x01010101u;
    o.w = (((bits >> 12) & 0xfu) * 0x00204081u) & 0x01010101u;
    const size_t row = (size_t)s * rows_pad + r;
    reinterpret_cast<uint4*>(x8 + row * 256)[part] = o;
    // popcount of the row: the 16 parts of a row sit in one half-warp
    int pc = __popc(bits);
#pragma unroll
    for (int d = 8; d >= 1; d >>= 1) pc += __shfl_xor_sync(0xffffffffu, pc, d);
    if (part == 0) ck[row] = r < n ? (((uint32_t)(pc + 512) << 20) | (uint32_t)r) : 0xffffffffu;
}

__global__ void __launch_bounds__(kTcThreads) match_tc_kernel(MatchJob job, MatchTcView tc, const __grid_constant__ CUtensorMap map_q,
                                                               const __grid_constant__ CUtensorMap map_t, int n_seg, int seg_len, size_t seg_stride,
                                                               uint32_t mul) {
    extern __shared__ uint8_t smem_raw[];
    const int pair = blockIdx.y;
    const int nq = job.nq[(size_t)pair * job.count_stride];
    const int nt = job.nt[(size_t)pair * job.count_stride];
    const int q0 = blockIdx.x * TCM;
    if (q0 >= nq || nt <= 0) return;  // block-uniform: nothing allocated yet
    const int t_begin = n_seg > 1 ? (int)blockIdx.z * seg_len : 0;
    const int t_end = n_seg > 1 ? min(nt, t_begin + seg_len) : nt;
    const int n_tiles = t_end > t_begin ? (t_end - t_begin + TCN - 1) / TCN : 0;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t sA = base + kOffA, sB = base + kOffB, sKeys = base + kOffKeys, sBars = base + kOffBars;
    const uint32_t bar_a = sBars, bar_bfull = sBars + 8, bar_bempty = sBars + 8 + 8 * TCSTAGES;
    const uint32_t bar_accfull = sBars + 8 + 16 * TCSTAGES, bar_accempty = bar_accfull + 16, s_tmem = bar_accempty + 16;
    uint8_t* gen_base = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t* keys_s = reinterpret_cast<const uint32_t*>(gen_base + kOffKeys);
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(gen_base + (s_tmem - base));

    if (n_tiles > 0) {
        if (threadIdx.x == 0) {
            mbar_init(bar_a, 1);
            for (int s = 0; s < TCSTAGES; s++) {
                mbar_init(bar_bfull + 8 * s, 1);
                mbar_init(bar_bempty + 8 * s, 1);
            }
            for (int a = 0; a < 2; a++) {
                mbar_init(bar_accfull + 8 * a, 1);
                mbar_init(bar_accempty + 8 * a, 4);  // one arrival per epilogue warp
            }
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        if (warp == 5) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_tmem), "r"(kTmemCols) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
        const uint32_t tmem_base = *tmem_slot;

        const int row_q0 = tc.q_row0 + pair * tc.q_rows_per_pair + q0;
        const int row_t0 = tc.t_row0 + pair * tc.t_rows_per_pair + t_begin;
        if (warp == 4) {
            if (lane == 0) {  // ---- TMA producer
                mbar_expect_tx(bar_a, kOperandBytes);
                tma_load_2d(sA, &map_q, 0, row_q0, bar_a);
                tma_load_2d(sA + kTileBytes, &map_q, (int)kKBlockBytes, row_q0, bar_a);
                for (int j = 0; j < n_tiles; j++) {
                    const int s = j % TCSTAGES;
                    if (j >= TCSTAGES) mbar_wait(bar_bempty + 8 * s, (uint32_t)((j / TCSTAGES) - 1) & 1u);
                    mbar_expect_tx(bar_bfull + 8 * s, kOperandBytes + TCN * 4);
                    tma_load_2d(sB + s * kOperandBytes, &map_t, 0, row_t0 + j * TCN, bar_bfull + 8 * s);
                    tma_load_2d(sB + s * kOperandBytes + kTileBytes, &map_t, (int)kKBlockBytes, row_t0 + j * TCN, bar_bfull + 8 * s);
                    bulk_load(sKeys + (uint32_t)(j % kKeySlots) * TCN * 4, tc.ckt + row_t0 + j * TCN, TCN * 4, bar_bfull + 8 * s);
                }
            }
        } else if (warp == 5) {
            if (lane == 0) {  // ---- MMA issuer
                mbar_wait(bar_a, 0);
                for (int j = 0; j < n_tiles; j++) {
                    const int s = j % TCSTAGES, a = j & 1;
                    mbar_wait(bar_bfull + 8 * s, (uint32_t)(j / TCSTAGES) & 1u);
                    if (j >= 2) mbar_wait(bar_accempty + 8 * a, (uint32_t)((j >> 1) - 1) & 1u);
                    tc_fence_after();
#pragma unroll
                    for (int kb = 0; kb < 2; kb++) {
#pragma unroll
                        for (int k = 0; k < 4; k++) {  // UMMA K = 32 bytes: advance the start address inside the swizzle span
                            const uint64_t ad = umma_desc(sA + kb * kTileBytes + k * 32);
                            const uint64_t bd = umma_desc(sB + s * kOperandBytes + kb * kTileBytes + k * 32);
                            umma_i8(tmem_base + a * TCN, ad, bd, (kb | k) ? 1u : 0u);
                        }
                    }
                    umma_commit(bar_bempty + 8 * s);    // the smem stage may be refilled once these MMAs have read it
                    umma_commit(bar_accfull + 8 * a);   // ... and the accumulator stage is complete
                }
            }
        } else {  // ---- epilogue: warps 0-3 own TMEM lanes 32 * warp .. + 31 = queries q0 + 32 * warp + lane
            uint32_t best = 0xffffffffu, second = 0xffffffffu;
            for (int j = 0; j < n_tiles; j++) {
                const int a = j & 1;
                mbar_wait(bar_accfull + 8 * a, (uint32_t)(j >> 1) & 1u);
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(a * TCN);
                const uint4* ck = reinterpret_cast<const uint4*>(keys_s + (j % kKeySlots) * TCN);  // landed with the train tile
                uint32_t v0[32], v1[32];
                tmem_ld32(taddr, v0);
                tmem_ld_wait(v0);
                tmem_ld32(taddr + 32, v1);
                fold32(v0, ck, mul, best, second);
                tmem_ld_wait(v1);
                tmem_ld32(taddr + 64, v0);
                fold32(v1, ck + 8, mul, best, second);
                tmem_ld_wait(v0);
                tmem_ld32(taddr + 96, v1);
                fold32(v0, ck + 16, mul, best, second);
                tmem_ld_wait(v1);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_accempty + 8 * a);  // this warp has drained its quadrant of the stage
                fold32(v1, ck + 24, mul, best, second);
            }
            const int q = q0 + warp * 32 + lane;
            if (q < nq) {
                // popc(q) from the query's own key constant; key >> 20 = popc(t) + 512 - 2 <q, t>
                const int pq = (int)(__ldg(tc.ckq + row_q0 + warp * 32 + lane) >> 20) - 512;
                int4 c = make_int4(-1, INT_MAX, INT_MAX, -1);
                if (best != 0xffffffffu) { c.y = (int)(best >> 20) - 512 + pq; c.x = (int)(best & 0xfffffu); }
                if (second != 0xffffffffu) { c.z = (int)(second >> 20) - 512 + pq; c.w = (int)(second & 0xfffffu); }
                SLAMCU_BOUND(q, job.max_q);
                SLAMCU_BOUND(c.x + 1, nt + 1);
                job.cand[(size_t)blockIdx.z * seg_stride + (size_t)pair * job.cand_pair_stride + q] = c;
            }
        }
        tc_fence_before();
        __syncthreads();
        if (warp == 5) {
            tc_fence_after();
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
        }
    } else if (threadIdx.x < TCM) {  // an empty train slice
        const int q = q0 + threadIdx.x;
        if (q < nq) job.cand[(size_t)blockIdx.z * seg_stride + (size_t)pair * job.cand_pair_stride + q] = make_int4(-1, INT_MAX, INT_MAX, -1);
    }
}

}  // namespace

void init_match_tc_attributes() { cudaFuncSetAttribute(match_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmemBytes); }

int launch_expand_bits(const uint32_t* desc, size_t set_stride_words, const int* counts, int count_stride, int n_sets, int rows_pad, uint8_t* x8,
                       uint32_t* ck, cudaStream_t st) {
    if (n_sets <= 0) return 0;
    SLAM_KERNEL("match_expand", st,
                expand_bits_kernel<<<dim3((rows_pad + 15) / 16, n_sets), 256, 0, st>>>(desc, set_stride_words, counts, count_stride, rows_pad, x8, ck));
    return 1;
}

int launch_match_tc(const MatchJob& job, const MatchTc& tc, int n_pairs, cudaStream_t st, int n_seg, int seg_len, size_t seg_stride) {
    dim3 grid((job.max_q + TCM - 1) / TCM, n_pairs, n_seg);
    SLAM_KERNEL("match", st,
                match_tc_kernel<<<grid, kTcThreads, kTcSmemBytes, st>>>(job, tc.view, tc.map_q, tc.map_t, n_seg, seg_len, seg_stride,
                                                                      (uint32_t)(-(2 << 20))));
    return 1;
}

}  // namespace slamcu
