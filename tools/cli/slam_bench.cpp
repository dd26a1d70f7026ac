// slam_bench -- C++ benchmark driver of the CUDA frontend (the reference's tools/cli/cli.cpp is a getopt stub that
// only constructs an empty SLAMModel; this is the driver north_star asks for in its place).  Host C++ only: it talks
// to the GPU through the C ABI (include/slam/cuda/slamcu.h) and links libslamcu.so.
//
//   slam_bench -c detector.yml -m matcher.yml [-W 1241 -H 376] [-f frames] [-s steps] [-w warmup] [-k] [-r raw.u8]
//
// A step = detectAndCompute on every frame + match(f, f+1) on every consecutive pair.  Prints one JSON line with
// frames/s for device-resident frames and for the pipelined host-buffer path (H2D / D2H inside the timed region).
// Frames: a deterministic synthetic translating scene (squares on texture), or raw 8-bit frames from -r.
#include <getopt.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include <slam/cuda/frontend.hpp>

namespace {

struct XorShift {
    uint64_t s;
    uint32_t next() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return static_cast<uint32_t>(s >> 16); }
    int range(int lo, int hi) { return lo + static_cast<int>(next() % static_cast<uint32_t>(hi - lo)); }
};

void make_frames(uint8_t* out, int n, int rows, int cols, int pitch_px, uint64_t seed) {
    const int H = rows + 128, W = cols + 128;
    std::vector<uint8_t> canvas(static_cast<size_t>(H) * W);
    XorShift r{seed * 0x9E3779B97F4A7C15ULL + 1};
    for (auto& v : canvas) v = static_cast<uint8_t>(110 + r.range(-6, 7));
    for (int y = pitch_px / 2; y < H - 8; y += pitch_px)
        for (int x = pitch_px / 2; x < W - 8; x += pitch_px) {
            const int yy = std::max(y + r.range(-1, 2), 0), xx = std::max(x + r.range(-1, 2), 0), side = r.range(3, 6);
            const uint8_t val = static_cast<uint8_t>(r.range(0, 2) ? r.range(0, 50) : r.range(190, 256));
            for (int dy = 0; dy < side; dy++) std::memset(&canvas[static_cast<size_t>(yy + dy) * W + xx], val, static_cast<size_t>(side));
        }
    for (int f = 0; f < n; f++) {
        const int oy = 64 + (f % 16), ox = 64 + ((2 * f) % 32);
        for (int y = 0; y < rows; y++)
            std::memcpy(out + (static_cast<size_t>(f) * rows + y) * cols, &canvas[static_cast<size_t>(oy + y) * W + ox], static_cast<size_t>(cols));
    }
}

double now() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

}  // namespace

int main(int argc, char** argv) {
    std::string det_cfg, mat_cfg, raw;
    int rows = 376, cols = 1241, frames = 256, steps = 5, warmup = 3, with_kp = 0, max_kp = 2560, chunk = 128, pitch_px = 14;
    bool ransac = false;
    double K4[4] = {525.0, 525.0, 319.5, 239.5};
    int opt;
    while ((opt = getopt(argc, argv, "hc:m:W:H:f:s:w:kr:K:C:e:p:")) != -1) {
        switch (opt) {
            case 'c': det_cfg = optarg; break;
            case 'm': mat_cfg = optarg; break;
            case 'W': cols = std::atoi(optarg); break;
            case 'H': rows = std::atoi(optarg); break;
            case 'f': frames = std::atoi(optarg); break;
            case 's': steps = std::atoi(optarg); break;
            case 'w': warmup = std::atoi(optarg); break;
            case 'k': with_kp = 1; break;
            case 'r': raw = optarg; break;
            case 'K': max_kp = std::atoi(optarg); break;
            case 'C': chunk = std::atoi(optarg); break;
            case 'p': pitch_px = std::atoi(optarg); break;
            case 'e':  // -e fx,fy,cx,cy : also run findEssentialMat(RANSAC) on every consecutive pair
                ransac = std::sscanf(optarg, "%lf,%lf,%lf,%lf", &K4[0], &K4[1], &K4[2], &K4[3]) == 4;
                if (!ransac) { std::fprintf(stderr, "slam_bench: -e expects fx,fy,cx,cy\n"); return 2; }
                break;
            default:
                std::printf("usage: %s -c detector.yml -m matcher.yml [-W cols -H rows -f frames -s steps -w warmup -k -r raw.u8 -K max_kp -C chunk -p scene_pitch -e fx,fy,cx,cy]\n", argv[0]);
                return opt == 'h' ? 0 : 2;
        }
    }
    if (det_cfg.empty() || mat_cfg.empty()) {
        std::fprintf(stderr, "slam_bench: -c and -m are required\n");
        return 2;
    }
    try {
        slam::cuda::Context& ctx = slam::cuda::Context::instance();
        slam::cuda::FeatureDetector det(det_cfg, ctx);
        slam::cuda::FeatureMatcher mat(mat_cfg, ctx);
        const size_t fb = static_cast<size_t>(rows) * cols;
        void* p = nullptr;
        ctx.check(slamcu_alloc_pinned(fb * frames, &p));
        uint8_t* host = static_cast<uint8_t*>(p);
        if (!raw.empty()) {
            FILE* f = std::fopen(raw.c_str(), "rb");
            if (!f || std::fread(host, 1, fb * frames, f) != fb * frames) throw std::runtime_error("could not read " + raw);
            std::fclose(f);
        } else {
            for (int f0 = 0; f0 < frames; f0 += 16) make_frames(host + fb * f0, std::min(16, frames - f0), rows, cols, pitch_px, 1000 + f0 / 16);
        }
        slamcu_sequence* seq = nullptr;
        ctx.check(slamcu_sequence_create(ctx.get(), rows, cols, frames, 0, max_kp, det.descriptorBytes(), &seq));
        void *hk = nullptr, *hd = nullptr, *hm = nullptr, *hc = nullptr;
        ctx.check(slamcu_alloc_pinned(static_cast<size_t>(frames) * max_kp * sizeof(slamcu_keypoint), &hk));
        ctx.check(slamcu_alloc_pinned(static_cast<size_t>(frames) * max_kp * det.descriptorBytes(), &hd));
        ctx.check(slamcu_alloc_pinned(static_cast<size_t>(frames) * max_kp * sizeof(slamcu_dmatch), &hm));
        ctx.check(slamcu_alloc_pinned(static_cast<size_t>(frames) * 16, &hc));
        auto resident = [&]() {
            ctx.check(slamcu_sequence_extract(seq, det.handle(), 0, frames));
            ctx.check(slamcu_sequence_match(seq, mat.handle(), 0, frames - 1, with_kp));
            if (ransac) ctx.check(slamcu_sequence_essential(seq, 0, frames - 1, K4, 0.999, 1.0, 1000));
        };
        auto e2e = [&]() {
            ctx.check(slamcu_sequence_process(seq, det.handle(), mat.handle(), host, cols, frames, chunk, with_kp,
                                              static_cast<slamcu_keypoint*>(hk), static_cast<uint8_t*>(hd),
                                              static_cast<slamcu_dmatch*>(hm), static_cast<int32_t*>(hc)));
            if (ransac) ctx.check(slamcu_sequence_essential(seq, 0, frames - 1, K4, 0.999, 1.0, 1000));
            ctx.check(slamcu_synchronize(ctx.get()));
        };
        ctx.check(slamcu_sequence_upload(seq, 0, frames, host, cols));
        for (int i = 0; i < std::max(warmup, 3); i++) resident();
        ctx.check(slamcu_synchronize(ctx.get()));
        const int64_t l0 = slamcu_launch_count(ctx.get());
        double t0 = now();
        for (int i = 0; i < steps; i++) resident();
        ctx.check(slamcu_synchronize(ctx.get()));
        const double t_res = now() - t0;
        const int64_t launches = slamcu_launch_count(ctx.get()) - l0;
        for (int i = 0; i < 2; i++) e2e();
        t0 = now();
        for (int i = 0; i < steps; i++) e2e();
        const double t_e2e = now() - t0;
        const int32_t* c = static_cast<const int32_t*>(hc);
        long kp = 0, nm = 0, bad = 0, inl = 0, its = 0;
        for (int f = 0; f < frames; f++) { kp += c[4 * f]; nm += c[4 * f + 1]; bad += c[4 * f + 3] != 0; }
        if (ransac)
            for (int f = 0; f + 1 < frames; f += std::max(1, (frames - 1) / 16)) {  // sample 16 pairs
                int ni = 0, nit = 0;
                ctx.check(slamcu_sequence_essential_read(seq, f, nullptr, &ni, &nit, nullptr, 0, nullptr));
                inl += ni; its += nit;
            }
        std::printf("{\"tool\": \"slam_bench\", \"rows\": %d, \"cols\": %d, \"frames_per_step\": %d, \"steps\": %d, "
                    "\"frames_per_s_resident\": %.1f, \"frames_per_s_e2e\": %.1f, \"ms_per_step_resident\": %.3f, "
                    "\"ms_per_step_e2e\": %.3f, \"keypoints_per_frame\": %.1f, \"matches_per_pair\": %.1f, "
                    "\"overflowed_frames\": %ld, \"gpu_launches\": %lld, \"ransac\": %s, \"ransac_inliers_sampled_sum\": %ld, "
                    "\"ransac_iterations_sampled_sum\": %ld}\n",
                    rows, cols, frames, steps, frames * steps / t_res, frames * steps / t_e2e, 1e3 * t_res / steps,
                    1e3 * t_e2e / steps, static_cast<double>(kp) / frames, static_cast<double>(nm) / std::max(frames - 1, 1), bad,
                    static_cast<long long>(launches), ransac ? "true" : "false", inl, its);
        slamcu_sequence_destroy(seq);
        slamcu_free_pinned(hk); slamcu_free_pinned(hd); slamcu_free_pinned(hm); slamcu_free_pinned(hc); slamcu_free_pinned(host);
    } catch (const std::exception& e) {
        std::fprintf(stderr, "slam_bench: %s\n", e.what());
        return 1;
    }
    return 0;
}
