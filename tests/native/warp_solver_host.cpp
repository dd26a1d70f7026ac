// Host build of the WARP-COOPERATIVE five-point solver (slam_cin0051_b200/csrc/fivept_warp.cuh) under the SIMT
// emulation of simt_emu.hpp, so that the device source of the RANSAC kernels' solver runs in the CPU test suite.
#include <thread>
#include <vector>

#include "simt_emu.hpp"
#include "../../slam_cin0051_b200/csrc/fivept_warp.cuh"

// x1, x2: [n_samples][5][2]; models: [n_samples][10][9]; counts: [n_samples]
extern "C" void hw_five_point_warp(const double* x1, const double* x2, int n_samples, double* models, int* counts) {
    simt::Warp warp;
    simt::g_warp = &warp;
    std::vector<double> scratch(slamcu::kFiveptScratchDoubles);
    const int ident[5] = {0, 1, 2, 3, 4};
    std::vector<std::thread> lanes;
    for (int l = 0; l < 32; l++)
        lanes.emplace_back([&, l] {
            threadIdx.x = l;
            for (int s = 0; s < n_samples; s++) {
                const int c = slamcu::five_point_warp(reinterpret_cast<const double2*>(x1) + (size_t)s * 5,
                                                      reinterpret_cast<const double2*>(x2) + (size_t)s * 5, ident, scratch.data(),
                                                      models + (size_t)s * 90);
                if (l == 0) counts[s] = c;
                __syncwarp();
            }
        });
    for (auto& t : lanes) t.join();
    simt::g_warp = nullptr;
}
