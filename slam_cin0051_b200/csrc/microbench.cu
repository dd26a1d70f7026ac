// microbench.cu -- measured integer-pipe ceiling for the Hamming matcher's roofline.
// The matcher is bound by the POPC issue rate, not HBM; MEASURED_PEAKS.json has no such figure, so
// bench.py measures it on the box: independent POPC chains (LOP3 + POPC + IADD, like the matcher's
// inner loop) on every SM, timed with CUDA events.
#include "common.cuh"

namespace slamcu {
namespace {
__global__ void __launch_bounds__(256) popc_peak_kernel(unsigned* out, int iters, unsigned seed) {
    unsigned a0 = seed + threadIdx.x, a1 = a0 * 3u, a2 = a0 * 5u, a3 = a0 * 7u, a4 = a0 * 11u, a5 = a0 * 13u,
             a6 = a0 * 17u, a7 = a0 * 19u;
    unsigned s0 = 0, s1 = 0, s2 = 0, s3 = 0;
    for (int i = 0; i < iters; i++) {
        const unsigned k = (unsigned)i * 0x9e3779b9u;
        s0 += __popc(a0 ^ k) + __popc(a4 ^ k);
        s1 += __popc(a1 ^ k) + __popc(a5 ^ k);
        s2 += __popc(a2 ^ k) + __popc(a6 ^ k);
        s3 += __popc(a3 ^ k) + __popc(a7 ^ k);
    }
    if ((s0 + s1 + s2 + s3) == 0xffffffffu) out[0] = s0;
}
__global__ void trip_bound_kernel(int index, int limit) { SLAMCU_BOUND(index, limit); }
}  // namespace

// test hook of the debug-bounds build: an out-of-range index through the very macro the kernels use (a no-op in release builds)
void launch_trip_bound(cudaStream_t st) { trip_bound_kernel<<<1, 1, 0, st>>>(1, 1); }

// returns the number of POPCs executed; caller times it
long long launch_popc_peak(unsigned* scratch, int blocks, int iters, cudaStream_t st) {
    SLAM_KERNEL("popc_peak", st, popc_peak_kernel<<<blocks, 256, 0, st>>>(scratch, iters, 12345u));
    return (long long)blocks * 256 * iters * 8;
}
}  // namespace slamcu
