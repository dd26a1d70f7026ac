// oracle/ref_frontend.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// CPU restatement ("port") of the reference monocular-SLAM frontend's hot path, used only as the
// parity checker by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs.  Nothing under slam_cin0051_b200/ may link, import or call this file.
//
// Every function cites the reference file:line it follows (paths relative to the reference repo).
// The reference itself cannot be compiled in this image (no Eigen / OpenCV C++ / spdlog headers), so
// the restatement keeps the *arithmetic and library calls* of the reference -- std::sort,
// std::partial_sort, std::default_random_engine, std::normal_distribution<float>, std::atan2,
// std::cos, std::sin, std::exp, std::round, std::sqrt -- so it inherits libstdc++ 13 / glibc 2.39
// behaviour exactly, while replacing Eigen containers with flat row-major arrays.
//
// PARITY PIN: the reference's own tests hold no golden vectors for this path (they assert only exit
// codes), so "parity unpinned" by the reference; the oracle is pinned instead against the
// survey-time known answers in tests/test_oracle_pins.py.
//
// Build: see oracle/Makefile (g++ -O2, no -march, no -ffast-math: the reference has neither, so no
// FMA contraction can occur on baseline x86-64).

#include <algorithm>
#include <atomic>
#include <chrono>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <random>
#include <thread>
#include <utility>
#include <vector>

namespace {

// include/slam/frontend/feature_detector.hpp:14-23
constexpr int kCircle = 16;
constexpr float kKeypointSize = 6.0F;
constexpr float kRad2Deg = 180.0F / M_PI;
constexpr float kDeg2Rad = M_PI / 180.0F;
// include/slam/frontend/feature_matcher.hpp:12
constexpr uint32_t kMaxJumpRadius = 500;

// include/slam/frontend/feature_detector.hpp:138-153 (dx, dy), index 0 = straight up, clockwise.
constexpr int kRing[16][2] = {{0, -3}, {1, -3}, {2, -2}, {3, -1}, {3, 0},  {3, 1},   {2, 2},   {1, 3},
                              {0, 3},  {-1, 3}, {-2, 2}, {-3, 1}, {-3, 0}, {-3, -1}, {-2, -2}, {-1, -3}};

struct Kp {  // include/slam/frontend/feature_detector.hpp:28-38
    float x, y, size, angle, response;
};
struct Mt {  // include/slam/frontend/feature_matcher.hpp:18-25
    int queryIdx, trainIdx;
    float distance;
};

struct Img {
    const uint8_t* p;
    int rows, cols;
    uint8_t at(int y, int x) const { return p[static_cast<size_t>(y) * cols + x]; }
};

// src/frontend/feature_detector.cpp:70-145
bool is_corner(const Img& im, int x, int y, int thr, int arc) {
    const int c = im.at(y, x);
    int hi = 0, lo = 0;
    auto vote = [&](int k) {
        const int n = im.at(y + kRing[k][1], x + kRing[k][0]);
        if (n > c + thr) {
            hi++;
        } else if (n < c - thr) {
            lo++;
        }
    };
    vote(0);
    vote(8);
    if (hi == 0 && lo == 0) return false;
    vote(4);
    vote(12);
    if (hi < 3 && lo < 3) return false;
    hi = lo = 0;
    for (int i = 0; i < 2 * kCircle; i++) {
        const int k = i % kCircle;
        const int n = im.at(y + kRing[k][1], x + kRing[k][0]);
        if (n > c + thr) {
            hi++;
            lo = 0;
        } else if (n < c - thr) {
            lo++;
            hi = 0;
        } else {
            hi = lo = 0;
        }
        if (hi >= arc || lo >= arc) return true;
    }
    return false;
}

// src/frontend/feature_detector.cpp:56-68
void fast_scan(const Img& im, int thr, int arc, std::vector<Kp>& out) {
    out.clear();
    for (int r = 3; r < im.rows - 3; r++)
        for (int c = 3; c < im.cols - 3; c++)
            if (is_corner(im, c, r, thr, arc))
                out.push_back(Kp{static_cast<float>(c), static_cast<float>(r), kKeypointSize, 0.0F, 0.0F});
}

// src/frontend/feature_detector.cpp:190-203
float sad_score(const Img& im, int x, int y) {
    const uint8_t c = im.at(y, x);
    float s = 0.0F;
    for (int i = 0; i < kCircle; i++) {
        const uint8_t n = im.at(y + kRing[i][1], x + kRing[i][0]);
        s += static_cast<float>(std::abs(n - c));
    }
    return s;
}

// src/frontend/feature_detector.cpp:147-188
void greedy_nms(const Img& im, int window, std::vector<Kp>& kps) {
    if (kps.empty()) return;
    for (auto& k : kps) k.response = sad_score(im, static_cast<int>(k.x), static_cast<int>(k.y));
    std::sort(kps.begin(), kps.end(), [](const Kp& a, const Kp& b) { return a.response > b.response; });
    std::vector<bool> dead(kps.size(), false);
    std::vector<Kp> keep;
    for (size_t i = 0; i < kps.size(); i++) {
        if (dead[i]) continue;
        keep.push_back(kps[i]);
        for (size_t j = i + 1; j < kps.size(); j++) {
            if (dead[j]) continue;
            float dx = kps[i].x - kps[j].x;
            float dy = kps[i].y - kps[j].y;
            float d = std::sqrt((dx * dx) + (dy * dy));
            if (d < static_cast<float>(window)) dead[j] = true;
        }
    }
    kps = keep;
}

// src/frontend/feature_detector.cpp:315-364 (kernel construction :321-335)
void blur_weights(int ksize, double sigma, std::vector<double>& w) {
    const int h = ksize / 2;
    w.assign(static_cast<size_t>(ksize) * ksize, 0.0);
    double sum = 0.0;
    for (int i = -h; i <= h; i++)
        for (int j = -h; j <= h; j++) {
            double v = std::exp(-((i * i) + (j * j)) / (2 * sigma * sigma));
            w[static_cast<size_t>(i + h) * ksize + (j + h)] = v;
            sum += v;
        }
    for (auto& v : w) v /= sum;  // Eigen 3.4 `kernel /= sum` is a per-coefficient division
}

// src/frontend/feature_detector.cpp:337-363
void gaussian_blur(const Img& im, int ksize, double sigma, uint8_t* out) {
    std::vector<double> w;
    blur_weights(ksize, sigma, w);
    const int h = ksize / 2;
    std::memset(out, 0, static_cast<size_t>(im.rows) * im.cols);
    for (int y = h; y < im.rows - h; y++)
        for (int x = h; x < im.cols - h; x++) {
            double acc = 0.0;
            for (int ky = -h; ky <= h; ky++)
                for (int kx = -h; kx <= h; kx++)
                    acc += static_cast<double>(im.at(y + ky, x + kx)) * w[static_cast<size_t>(ky + h) * ksize + (kx + h)];
            out[static_cast<size_t>(y) * im.cols + x] = static_cast<uint8_t>(std::round(acc));
        }
    // border frame copied from the input, rows first then columns (:356-361)
    for (int i = 0; i < h; i++) {
        if (i >= im.rows || i >= im.cols) break;
        std::memcpy(out + static_cast<size_t>(i) * im.cols, im.p + static_cast<size_t>(i) * im.cols, im.cols);
        std::memcpy(out + static_cast<size_t>(im.rows - 1 - i) * im.cols,
                    im.p + static_cast<size_t>(im.rows - 1 - i) * im.cols, im.cols);
        for (int r = 0; r < im.rows; r++) {
            out[static_cast<size_t>(r) * im.cols + i] = im.at(r, i);
            out[static_cast<size_t>(r) * im.cols + im.cols - 1 - i] = im.at(r, im.cols - 1 - i);
        }
    }
}

// src/frontend/feature_detector.cpp:286-313
void brief_pattern(int patch, int pairs, std::vector<int>& pat) {
    pat.clear();
    const float scale = static_cast<float>(patch) / 2.0F;
    std::default_random_engine gen;
    std::normal_distribution<float> dist(0.0F, 1.0F);
    for (int i = 0; i < pairs; i++) {
        float x1 = dist(gen) * scale;
        float y1 = dist(gen) * scale;
        float x2 = dist(gen) * scale;
        float y2 = dist(gen) * scale;
        if (std::abs(x1) < scale && std::abs(y1) < scale && std::abs(x2) < scale && std::abs(y2) < scale) {
            pat.push_back(static_cast<int>(x1));
            pat.push_back(static_cast<int>(y1));
            pat.push_back(static_cast<int>(x2));
            pat.push_back(static_cast<int>(y2));
        }
    }
}

// src/frontend/feature_detector.cpp:205-231
float orientation(const Img& im, const Kp& k, int patch) {
    int x = static_cast<int>(k.x);
    int y = static_cast<int>(k.y);
    const int r = patch / 2;
    if (x - r < 0 || x + r >= im.cols || y - r < 0 || y + r >= im.rows) return 0.0F;
    float m01 = 0.0F, m10 = 0.0F;
    for (int v = -r; v <= r; v++)
        for (int u = -r; u <= r; u++)
            if (u * u + v * v <= r * r) {
                uint8_t p = im.at(y + v, x + u);
                m01 += static_cast<float>(v) * static_cast<float>(p);
                m10 += static_cast<float>(u) * static_cast<float>(p);
            }
    return static_cast<float>(std::atan2(m01, m10) * kRad2Deg);
}

// src/frontend/feature_detector.cpp:233-284
void brief(const Img& im, const Kp& k, int patch, int pairs, const std::vector<int>& pat, uint8_t* d) {
    const int nbytes = pairs / 8;
    std::memset(d, 0, nbytes);
    int x = static_cast<int>(k.x);
    int y = static_cast<int>(k.y);
    if (x - patch / 2 < 0 || x + patch / 2 >= im.cols || y - patch / 2 < 0 || y + patch / 2 >= im.rows) return;
    float a = k.angle * kDeg2Rad;
    float ca = std::cos(a);
    float sa = std::sin(a);
    int bit = 0;
    for (size_t i = 0; i < pat.size() / 4 && bit < nbytes * 8; i++) {
        const auto p1x = static_cast<float>(pat[4 * i]);
        const auto p1y = static_cast<float>(pat[4 * i + 1]);
        const auto p2x = static_cast<float>(pat[4 * i + 2]);
        const auto p2y = static_cast<float>(pat[4 * i + 3]);
        auto x1 = static_cast<int>((p1x * ca) - (p1y * sa)) + x;
        auto y1 = static_cast<int>((p1x * sa) + (p1y * ca)) + y;
        auto x2 = static_cast<int>((p2x * ca) - (p2y * sa)) + x;
        auto y2 = static_cast<int>((p2x * sa) + (p2y * ca)) + y;
        if (x1 >= 0 && x1 < im.cols && y1 >= 0 && y1 < im.rows && x2 >= 0 && x2 < im.cols && y2 >= 0 &&
            y2 < im.rows) {
            if (im.at(y1, x1) < im.at(y2, x2)) d[bit / 8] |= static_cast<uint8_t>(1 << (bit % 8));
            bit++;
        }
    }
}

// include/slam/common/common.hpp:18-50
int hamming(const uint8_t* a, const uint8_t* b, int n) {
    static const auto table = [] {
        std::vector<uint8_t> t(256);
        for (int i = 0; i < 256; i++) {
            int v = i;
            uint8_t c = 0;
            while (v > 0) {
                v &= (v - 1);
                c++;
            }
            t[i] = c;
        }
        return t;
    }();
    int d = 0;
    for (int k = 0; k < n; k++) d += table[a[k] ^ b[k]];
    return d;
}

struct DetCfg {
    int thr, arc, nms, window, patch, pairs;
};
struct MatCfg {
    int filter, good, use_ratio;
    float ratio;
};

// src/frontend/feature_matcher.cpp:143-189 (+ :132-141)
void best_matches(const uint8_t* d1, int n1, const uint8_t* d2, int n2, int width, const Kp* k1, int nk1,
                  const Kp* k2, int nk2, const MatCfg& cfg, std::vector<Mt>& out, long long* penalised) {
    const bool use_kp = nk1 > 0 && nk2 > 0;
    long long pen = 0;
    for (int i = 0; i < n1; i++) {
        int best = INT_MAX, second = INT_MAX, bidx = -1;
        for (int j = 0; j < n2; j++) {
            int dist = hamming(d1 + static_cast<size_t>(i) * width, d2 + static_cast<size_t>(j) * width, width);
            if (use_kp) {
                const float dx = k1[i].x - k2[j].x;
                const float dy = k1[i].y - k2[j].y;
                const float idist = std::sqrt(dx * dx + dy * dy);
                if (idist > kMaxJumpRadius) {
                    float penalty = 1.0f + (idist / static_cast<float>(kMaxJumpRadius));
                    dist = static_cast<int>(static_cast<float>(dist) * penalty);
                    pen++;
                }
            }
            if (dist < best) {
                second = best;
                best = dist;
                bidx = j;
            } else if (dist < second) {
                second = dist;
            }
        }
        bool good = true;
        if (cfg.use_ratio && static_cast<float>(best) >= cfg.ratio * static_cast<float>(second)) good = false;
        if (good && bidx != -1) out.push_back(Mt{i, bidx, static_cast<float>(best)});
    }
    if (penalised) *penalised = pen;
}

// src/frontend/feature_matcher.cpp:191-204
void filter_sort(std::vector<Mt>& m, int good) {
    const auto pred = [](const Mt& a, const Mt& b) { return a.distance < b.distance; };
    if (m.size() > static_cast<size_t>(good)) {
        std::partial_sort(m.begin(), m.begin() + good, m.end(), pred);
        m.erase(m.begin() + good, m.end());
    } else {
        std::sort(m.begin(), m.end(), pred);
    }
}

void detect_impl(const Img& im, const DetCfg& c, std::vector<Kp>& kps) {  // feature_detector.cpp:8-18
    fast_scan(im, c.thr, c.arc, kps);
    if (c.nms) greedy_nms(im, c.window, kps);
}

void compute_impl(const Img& im, const DetCfg& c, const std::vector<int>& pat, std::vector<Kp>& kps,
                  std::vector<uint8_t>& desc) {  // feature_detector.cpp:20-47
    desc.clear();
    if (kps.empty()) return;
    const int nb = c.pairs / 8;
    desc.assign(kps.size() * nb, 0);
    std::vector<uint8_t> blurred(static_cast<size_t>(im.rows) * im.cols);
    gaussian_blur(im, 5, 1.0, blurred.data());
    Img b{blurred.data(), im.rows, im.cols};
    for (size_t i = 0; i < kps.size(); i++) {
        kps[i].angle = orientation(b, kps[i], c.patch);
        brief(b, kps[i], c.patch, c.pairs, pat, desc.data() + i * nb);
    }
}

}  // namespace

extern "C" {

// All entry points: caller-allocated outputs; return value = element count (or -1 on capacity overflow).

int orc_brief_pattern(int patch, int pairs, int* out4, int cap_pairs) {
    std::vector<int> pat;
    brief_pattern(patch, pairs, pat);
    int n = static_cast<int>(pat.size() / 4);
    if (n > cap_pairs) return -1;
    std::memcpy(out4, pat.data(), pat.size() * sizeof(int));
    return n;
}

int orc_blur_weights(int ksize, double sigma, double* out) {
    std::vector<double> w;
    blur_weights(ksize, sigma, w);
    std::memcpy(out, w.data(), w.size() * sizeof(double));
    return static_cast<int>(w.size());
}

int orc_fast_scan(const uint8_t* img, int rows, int cols, int thr, int arc, float* kps5, int cap) {
    std::vector<Kp> k;
    fast_scan(Img{img, rows, cols}, thr, arc, k);
    if (static_cast<int>(k.size()) > cap) return -1;
    std::memcpy(kps5, k.data(), k.size() * sizeof(Kp));
    return static_cast<int>(k.size());
}

// raw corners + SAD scores in raster order (before the sort) -- for the first-kernel parity test
int orc_fast_scan_scored(const uint8_t* img, int rows, int cols, int thr, int arc, float* kps5, int cap) {
    std::vector<Kp> k;
    Img im{img, rows, cols};
    fast_scan(im, thr, arc, k);
    if (static_cast<int>(k.size()) > cap) return -1;
    for (auto& e : k) e.response = sad_score(im, static_cast<int>(e.x), static_cast<int>(e.y));
    std::memcpy(kps5, k.data(), k.size() * sizeof(Kp));
    return static_cast<int>(k.size());
}

int orc_detect(const uint8_t* img, int rows, int cols, const int* cfg6, float* kps5, int cap) {
    DetCfg c{cfg6[0], cfg6[1], cfg6[2], cfg6[3], cfg6[4], cfg6[5]};
    std::vector<Kp> k;
    detect_impl(Img{img, rows, cols}, c, k);
    if (static_cast<int>(k.size()) > cap) return -1;
    std::memcpy(kps5, k.data(), k.size() * sizeof(Kp));
    return static_cast<int>(k.size());
}

void orc_gaussian_blur(const uint8_t* img, int rows, int cols, int ksize, double sigma, uint8_t* out) {
    gaussian_blur(Img{img, rows, cols}, ksize, sigma, out);
}

// keypoints in/out (angle written); desc = n * (pairs/8) bytes
int orc_compute(const uint8_t* img, int rows, int cols, const int* cfg6, float* kps5, int n, uint8_t* desc) {
    DetCfg c{cfg6[0], cfg6[1], cfg6[2], cfg6[3], cfg6[4], cfg6[5]};
    std::vector<int> pat;
    brief_pattern(c.patch, c.pairs, pat);
    std::vector<Kp> k(n);
    std::memcpy(k.data(), kps5, static_cast<size_t>(n) * sizeof(Kp));
    std::vector<uint8_t> d;
    compute_impl(Img{img, rows, cols}, c, pat, k, d);
    std::memcpy(kps5, k.data(), static_cast<size_t>(n) * sizeof(Kp));
    if (!d.empty()) std::memcpy(desc, d.data(), d.size());
    return n;
}

int orc_detect_and_compute(const uint8_t* img, int rows, int cols, const int* cfg6, float* kps5, uint8_t* desc,
                           int cap) {
    DetCfg c{cfg6[0], cfg6[1], cfg6[2], cfg6[3], cfg6[4], cfg6[5]};
    std::vector<int> pat;
    brief_pattern(c.patch, c.pairs, pat);
    std::vector<Kp> k;
    std::vector<uint8_t> d;
    Img im{img, rows, cols};
    detect_impl(im, c, k);
    if (static_cast<int>(k.size()) > cap) return -1;
    compute_impl(im, c, pat, k, d);
    std::memcpy(kps5, k.data(), k.size() * sizeof(Kp));
    if (!d.empty()) std::memcpy(desc, d.data(), d.size());
    return static_cast<int>(k.size());
}

// cfg: filter, good, use_ratio ; ratio separately.  kp1/kp2 may be null (nk = 0).
// stage: 0 = after ratio test (query order), 1 = after filterAndSortMatches when cfg.filter.
int orc_match(const uint8_t* d1, int n1, const uint8_t* d2, int n2, int width, const float* kp1, int nk1,
              const float* kp2, int nk2, const int* cfg3, float ratio, int stage, int* out_q, int* out_t,
              float* out_d, int cap, long long* penalised) {
    MatCfg c{cfg3[0], cfg3[1], cfg3[2], ratio};
    std::vector<Mt> m;
    best_matches(d1, n1, d2, n2, width, reinterpret_cast<const Kp*>(kp1), nk1, reinterpret_cast<const Kp*>(kp2),
                 nk2, c, m, penalised);
    if (stage >= 1 && c.filter) filter_sort(m, c.good);
    if (static_cast<int>(m.size()) > cap) return -1;
    for (size_t i = 0; i < m.size(); i++) {
        out_q[i] = m[i].queryIdx;
        out_t[i] = m[i].trainIdx;
        out_d[i] = m[i].distance;
    }
    return static_cast<int>(m.size());
}

// std::sort / std::partial_sort permutation probes, for checking the device emulation on CPU.
// keys: response (desc) -> writes the permutation (original index at each sorted slot).
void orc_sort_perm_desc(const float* resp, int n, int* perm) {
    struct E {
        float r;
        int i;
    };
    std::vector<E> v(n);
    for (int i = 0; i < n; i++) v[i] = E{resp[i], i};
    std::sort(v.begin(), v.end(), [](const E& a, const E& b) { return a.r > b.r; });
    for (int i = 0; i < n; i++) perm[i] = v[i].i;
}
// distance asc; n <= k -> std::sort of all; else partial_sort top-k.  Returns output count.
int orc_topk_perm_asc(const float* dist, int n, int k, int* perm) {
    struct E {
        float d;
        int i;
    };
    std::vector<E> v(n);
    for (int i = 0; i < n; i++) v[i] = E{dist[i], i};
    const auto pred = [](const E& a, const E& b) { return a.d < b.d; };
    int m = n;
    if (n > k) {
        std::partial_sort(v.begin(), v.begin() + k, v.end(), pred);
        m = k;
    } else {
        std::sort(v.begin(), v.end(), pred);
    }
    for (int i = 0; i < m; i++) perm[i] = v[i].i;
    return m;
}

// libm probes (glibc float routines the reference calls through std::atan2/cos/sin on floats)
void orc_atan2f(const float* y, const float* x, float* out, long long n) {
    for (long long i = 0; i < n; i++) out[i] = std::atan2(y[i], x[i]);
}
void orc_sincosf(const float* a, float* s, float* c, long long n) {
    for (long long i = 0; i < n; i++) {
        s[i] = std::sin(a[i]);
        c[i] = std::cos(a[i]);
    }
}

// include/slam/common/common.hpp:127-173 (Camera::undistortImage); K = fx,fy,cx,cy ; D = k1,k2,p1,p2.
// Output: row-major double image in [0,1] (the reference returns a column-major MatrixXd; values per
// (row, col) are what is compared) and, when map_out != null, the int32 source index (-1 = outside).
void orc_undistort(const uint8_t* img, int rows, int cols, const double* K4, const double* D4, double* out,
                   int* map_out) {
    const double fx = K4[0], fy = K4[1], cx = K4[2], cy = K4[3];
    const double k1 = D4[0], k2 = D4[1], p1 = D4[2], p2 = D4[3];
    for (int i = 0; i < rows; i++)
        for (int j = 0; j < cols; j++) {
            // LinSpaced(n, 0, n-1) yields exact integers for these sizes
            double x = (static_cast<double>(j) - cx) / fx;
            double y = (static_cast<double>(i) - cy) / fy;
            double r = std::sqrt(x * x + y * y);
            double r2 = r * r;
            double r4 = std::pow(r, 4);
            double xd = x * (1 + k1 * r2 + k2 * r4) + 2 * p1 * x * y + p2 * (r2 + 2 * (x * x));
            double yd = y * (1 + k1 * r2 + k2 * r4) + 2 * p2 * x * y + p1 * (r2 + 2 * (y * y));
            double ud = fx * xd + cx;
            double vd = fy * yd + cy;
            int u = static_cast<int>(std::round(ud));
            int v = static_cast<int>(std::round(vd));
            double val = 0.0;
            int src = -1;
            if (u >= 0 && v >= 0 && u < cols && v < rows) {
                src = v * cols + u;
                val = static_cast<double>(img[src]) / 255.0;
            }
            out[static_cast<size_t>(i) * cols + j] = val;
            if (map_out) map_out[static_cast<size_t>(i) * cols + j] = src;
        }
}

// CPU baseline driver: runs detectAndCompute on every frame and match(f, f+1) with keypoints on
// `threads` host threads (frame-parallel; the reference itself is single-threaded).  Returns seconds.
// counts3[f] = {n_kp, n_matches(after filter), 0}
double orc_frontend_run(const uint8_t* frames, int nframes, int rows, int cols, const int* det6, const int* mat3,
                        float ratio, int with_kp, int threads, int* counts3) {
    DetCfg dc{det6[0], det6[1], det6[2], det6[3], det6[4], det6[5]};
    MatCfg mc{mat3[0], mat3[1], mat3[2], ratio};
    std::vector<int> pat;
    brief_pattern(dc.patch, dc.pairs, pat);
    std::vector<std::vector<Kp>> kps(nframes);
    std::vector<std::vector<uint8_t>> desc(nframes);
    const size_t fsz = static_cast<size_t>(rows) * cols;
    const int nb = dc.pairs / 8;
    auto t0 = std::chrono::steady_clock::now();
    {
        std::atomic<int> next{0};
        auto work = [&] {
            for (int f; (f = next.fetch_add(1)) < nframes;) {
                Img im{frames + fsz * f, rows, cols};
                detect_impl(im, dc, kps[f]);
                compute_impl(im, dc, pat, kps[f], desc[f]);
                counts3[3 * f] = static_cast<int>(kps[f].size());
            }
        };
        std::vector<std::thread> pool;
        for (int t = 1; t < threads; t++) pool.emplace_back(work);
        work();
        for (auto& t : pool) t.join();
    }
    {
        std::atomic<int> next{0};
        auto work = [&] {
            for (int f; (f = next.fetch_add(1)) < nframes - 1;) {
                std::vector<Mt> m;
                counts3[3 * f + 1] = 0;
                counts3[3 * f + 2] = 0;
                if (kps[f].empty() || kps[f + 1].empty()) continue;  // reference would throw invalid_argument
                best_matches(desc[f].data(), static_cast<int>(kps[f].size()), desc[f + 1].data(),
                             static_cast<int>(kps[f + 1].size()), nb, kps[f].data(),
                             with_kp ? static_cast<int>(kps[f].size()) : 0, kps[f + 1].data(),
                             with_kp ? static_cast<int>(kps[f + 1].size()) : 0, mc, m, nullptr);
                if (mc.filter) filter_sort(m, mc.good);
                counts3[3 * f + 1] = static_cast<int>(m.size());
            }
        };
        std::vector<std::thread> pool;
        for (int t = 1; t < threads; t++) pool.emplace_back(work);
        work();
        for (auto& t : pool) t.join();
    }
    if (nframes > 0) {
        counts3[3 * (nframes - 1) + 1] = 0;
        counts3[3 * (nframes - 1) + 2] = 0;
    }
    auto t1 = std::chrono::steady_clock::now();
    return std::chrono::duration<double>(t1 - t0).count();
}

}  // extern "C"
