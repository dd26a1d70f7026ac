// prep.cu -- image preparation (rows P1, P2 of SURVEY.md section 8a).
//
//   bgr2gray_kernel      : cv::cvtColor(BGR2GRAY) as called at preprocessor.cpp:136.  OpenCV's 8-bit path is
//                          fixed point: (3735*B + 19235*G + 9798*R + 16384) >> 15.
//   undistort_map_kernel : the per-camera part of Camera::undistortImage (common.hpp:146-162): forward
//                          distortion of every output pixel in FP64 (no FMA), std::round, bounds test
//                          -> int32 source index (-1 = outside).  Constant per camera, built once.
//   remap_kernel         : the per-frame part (:159-170): gather; emits the u8 image the detector
//                          consumes and/or the reference's value/255.0 double image.
#include <algorithm>

#include "common.cuh"

namespace slamcu {
namespace {

__global__ void bgr2gray_kernel(const uint8_t* __restrict__ bgr, int rows, int cols, int stride,
                                uint8_t* __restrict__ gray, int gstride) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= cols || y >= rows) return;
    const uint8_t* p = bgr + (size_t)y * stride + 3 * x;
    gray[(size_t)y * gstride + x] = (uint8_t)((3735u * p[0] + 19235u * p[1] + 9798u * p[2] + 16384u) >> 15);
}

// Everything but pow() is IEEE-exact on the device (no FMA, correctly rounded sqrt and division); CUDA's pow() may differ
// from the host libm's by an ulp or two, which moves ud / vd by < 1e-9 px.  Pixels whose ud or vd lies within 1e-6 of a
// rounding boundary (k + 0.5) are therefore appended to `fix` and recomputed by the host with the reference's own libm
// calls (undistort_index_host in api.cu); every other pixel rounds to the same source index either way.
__global__ void undistort_map_kernel(int rows, int cols, CamParams c, int* __restrict__ map, int* __restrict__ fix, int fix_cap) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
    if (j >= cols || i >= rows) return;
    const double x = ((double)j - c.cx) / c.fx;
    const double y = ((double)i - c.cy) / c.fy;
    const double r = sqrt(x * x + y * y);
    const double r2 = r * r;
    const double r4 = pow(r, 4.0);
    const double xd = x * (1 + c.k1 * r2 + c.k2 * r4) + 2 * c.p1 * x * y + c.p2 * (r2 + 2 * (x * x));
    const double yd = y * (1 + c.k1 * r2 + c.k2 * r4) + 2 * c.p2 * x * y + c.p1 * (r2 + 2 * (y * y));
    const double ud = c.fx * xd + c.cx;
    const double vd = c.fy * yd + c.cy;
    const int u = (int)round(ud), v = (int)round(vd);
    map[(size_t)i * cols + j] = (u >= 0 && v >= 0 && u < cols && v < rows) ? v * cols + u : -1;
    const bool near = fabs(fabs(ud - floor(ud)) - 0.5) < 1e-6 || fabs(fabs(vd - floor(vd)) - 0.5) < 1e-6 || !(fabs(ud) < 1e9) || !(fabs(vd) < 1e9);
    if (near) {
        const int pos = atomicAdd(fix, 1);
        if (pos < fix_cap) fix[1 + pos] = i * cols + j;
    }
}

__global__ void remap_kernel(const uint8_t* __restrict__ gray, int rows, int cols, int stride,
                             const int* __restrict__ map, uint8_t* __restrict__ out_u8, double* __restrict__ out_f64) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
    if (j >= cols || i >= rows) return;
    const int src = map[(size_t)i * cols + j];
    uint8_t v = 0;
    if (src >= 0) v = gray[(size_t)(src / cols) * stride + (src % cols)];
    if (out_u8) out_u8[(size_t)i * cols + j] = v;
    if (out_f64) out_f64[(size_t)i * cols + j] = (src >= 0) ? (double)v / 255.0 : 0.0;
}

// dense host-layout frames (stride bytes per row) -> the pitched frame store; 4 output bytes per thread.
// (A strided cudaMemcpy2D of 1241-byte rows runs at ~10 GB/s over PCIe 5; one linear copy + this kernel runs
// at the link rate, tools/xfer_probe.py.)
__global__ void __launch_bounds__(256) repitch_kernel(const uint8_t* __restrict__ src, int stride, uint8_t* __restrict__ dst,
                                                      int pitch, int cols, long long rows_total) {
    // grid-stride over the rows: grid.y is capped at 65535 blocks of 8 rows, sequences may hold more rows than that
    for (long long row = (long long)blockIdx.y * 8 + (threadIdx.x >> 5); row < rows_total; row += (long long)gridDim.y * 8) {
        const uint8_t* s = src + row * stride;
        uint8_t* d = dst + row * pitch;
        for (int x = (blockIdx.x * 32 + (threadIdx.x & 31)) * 4; x < pitch; x += gridDim.x * 128) {
            uint32_t w = 0;
#pragma unroll
            for (int k = 0; k < 4; k++)
                if (x + k < cols) w |= (uint32_t)s[x + k] << (8 * k);
            *reinterpret_cast<uint32_t*>(d + x) = w;
        }
    }
}

// Batched Preprocessor::yield arithmetic (preprocessor.cpp:136-137) straight into the frame store: optional BGR2GRAY at
// the gathered source pixel, optional undistortion gather through the per-camera map (-1 = outside -> 0).
__global__ void __launch_bounds__(256) prepare_kernel(const uint8_t* __restrict__ src, int channels, int src_stride,
                                                      size_t src_frame_bytes, const int* __restrict__ map, uint8_t* __restrict__ dst,
                                                      int pitch, size_t dst_frame_bytes, int rows, int cols) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, f = blockIdx.z;
    if (x >= pitch) return;
    uint8_t v = 0;
    if (x < cols) {
        int sy = y, sx = x;
        bool inside = true;
        if (map) {
            const int m = map[(size_t)y * cols + x];
            inside = m >= 0;
            sy = m / cols;
            sx = m - sy * cols;
        }
        if (inside) {
            const uint8_t* p = src + (size_t)f * src_frame_bytes + (size_t)sy * src_stride + (size_t)sx * channels;
            v = channels == 3 ? (uint8_t)((3735u * p[0] + 19235u * p[1] + 9798u * p[2] + 16384u) >> 15) : p[0];
        }
    }
    dst[(size_t)f * dst_frame_bytes + (size_t)y * pitch + x] = v;  // the pitch padding is written as zeros
}

}  // namespace

int launch_prepare(const uint8_t* src, int channels, int src_stride, size_t src_frame_bytes, const int* map, uint8_t* dst, int pitch,
                   size_t dst_frame_bytes, int rows, int cols, int n, cudaStream_t st) {
    dim3 grid((pitch + 255) / 256, rows, n);
    SLAM_KERNEL("prepare", st, prepare_kernel<<<grid, 256, 0, st>>>(src, channels, src_stride, src_frame_bytes, map, dst, pitch,
                                                                    dst_frame_bytes, rows, cols));
    return 1;
}

int launch_repitch(const uint8_t* src, int stride, uint8_t* dst, int pitch, int cols, long long rows_total, cudaStream_t st) {
    dim3 grid(min((pitch / 4 + 31) / 32, 4), (unsigned)std::min<long long>((rows_total + 7) / 8, 65535));
    SLAM_KERNEL("repitch", st, repitch_kernel<<<grid, 256, 0, st>>>(src, stride, dst, pitch, cols, rows_total));
    return 1;
}

int launch_bgr2gray(const uint8_t* bgr, int rows, int cols, int stride, uint8_t* gray, int gstride, cudaStream_t st) {
    dim3 grid((cols + 255) / 256, rows);
    bgr2gray_kernel<<<grid, 256, 0, st>>>(bgr, rows, cols, stride, gray, gstride);
    return 1;
}
int launch_undistort_map(int rows, int cols, const CamParams& cam, int* map, int* fix, int fix_cap, cudaStream_t st) {
    dim3 grid((cols + 255) / 256, rows);
    cudaMemsetAsync(fix, 0, sizeof(int), st);
    undistort_map_kernel<<<grid, 256, 0, st>>>(rows, cols, cam, map, fix, fix_cap);
    return 1;
}
int launch_remap(const uint8_t* gray, int rows, int cols, int stride, const int* map, uint8_t* out_u8, double* out_f64,
                 cudaStream_t st) {
    dim3 grid((cols + 255) / 256, rows);
    remap_kernel<<<grid, 256, 0, st>>>(gray, rows, cols, stride, map, out_u8, out_f64);
    return 1;
}

}  // namespace slamcu
