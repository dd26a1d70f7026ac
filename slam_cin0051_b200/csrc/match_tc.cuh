// match_tc.cuh -- host-visible pieces of the tensor-core Hamming matcher (match_tc.cu).
#pragma once
#include <cuda.h>  // CUtensorMap (types only)

#include "common.cuh"

namespace slamcu {

// where the widened operands of a launch live: rows of 256 bytes (one byte per descriptor bit), key constants per row
struct MatchTcView {
    const uint32_t* ckq;  // [rows]  (popcount + 512) << 20 | index in its set; 0xffffffff for the zero rows that pad a set to 128
    const uint32_t* ckt;
    int q_row0, t_row0;                 // first row of pair 0's query / train set
    int q_rows_per_pair, t_rows_per_pair;  // rows between consecutive pairs' sets (multiples of 128)
};
struct MatchTc {
    MatchTcView view;
    CUtensorMap map_q, map_t;  // rank-2 u8 tensors {256, rows}, box {128, 128}, 128-byte swizzle, zero fill
};
constexpr int kTcRowBytes = 256;
constexpr int kTcRowAlign = 128;  // sets are padded to a multiple of 128 rows

void init_match_tc_attributes();
// n_sets descriptor sets (8 words per row) -> x8 / ck rows s * rows_pad + r
int launch_expand_bits(const uint32_t* desc, size_t set_stride_words, const int* counts, int count_stride, int n_sets, int rows_pad, uint8_t* x8,
                       uint32_t* ck, cudaStream_t st);
// fills job.cand exactly like match256_kernel (n_seg slices of seg_len train descriptors, seg_len a multiple of 128)
int launch_match_tc(const MatchJob& job, const MatchTc& tc, int n_pairs, cudaStream_t st, int n_seg, int seg_len, size_t seg_stride);

}  // namespace slamcu
