// oracle/ref_api.cpp -- TEST INFRASTRUCTURE ONLY.
//
// C entry points over the reference's OWN, UNMODIFIED classes: slam::FeatureDetector (src/frontend/feature_detector.cpp),
// slam::FeatureMatcher (src/frontend/feature_matcher.cpp) and slam::Camera::undistortImage (include/slam/common/common.hpp:127-173),
// compiled from the sources where they lie under /root/reference by oracle/Makefile (target _ref/libslam_ref.so) against the
// header stand-ins in oracle/shim/.  tests/test_ref_build.py checks the C++ restatement (oracle/ref_frontend.cpp) against this
// library: that is what pins the restatement -- and through it the CUDA path -- to the reference's code.
//
// `#define private public` gives the tests the reference's private stages (isFASTCorner scan, gaussianBlur, BRIEF pattern);
// it changes neither the sources nor the object layout.
#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <filesystem>
#include <map>
#include <memory>
#include <random>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include <Eigen/Eigen>
#include <opencv2/core.hpp>

#define private public
#include <slam/frontend/feature_detector.hpp>
#include <slam/frontend/feature_matcher.hpp>
#undef private

namespace {
thread_local std::string g_err;
slam::EigenGrayMatrix to_eigen(const uint8_t* img, int rows, int cols) {
    slam::EigenGrayMatrix m(rows, cols);
    std::memcpy(m.data(), img, (size_t)rows * cols);
    return m;
}
int put_keypoints(const std::vector<slam::Keypoint>& k, float* kps5, int cap) {
    if ((int)k.size() > cap) return -1;
    static_assert(sizeof(slam::Keypoint) == 20, "Keypoint is 5 floats");
    if (!k.empty()) std::memcpy(kps5, k.data(), k.size() * sizeof(slam::Keypoint));
    return (int)k.size();
}
std::vector<slam::Keypoint> get_keypoints(const float* kps5, int n) {
    std::vector<slam::Keypoint> k;
    k.reserve((size_t)n);
    for (int i = 0; i < n; i++) {
        slam::Keypoint p(kps5[5 * i], kps5[5 * i + 1], kps5[5 * i + 2]);
        p.angle = kps5[5 * i + 3];
        p.response = kps5[5 * i + 4];
        k.push_back(p);
    }
    return k;
}
template <class F>
int guarded(F f) {
    try {
        return f();
    } catch (const std::invalid_argument& e) {
        g_err = std::string("invalid_argument: ") + e.what();
        return -2;
    } catch (const std::runtime_error& e) {
        g_err = std::string("runtime_error: ") + e.what();
        return -3;
    } catch (const std::exception& e) {
        g_err = std::string("exception: ") + e.what();
        return -4;
    }
}
}  // namespace

extern "C" {

const char* ref_last_error() { return g_err.c_str(); }

// constructor only: returns 0, or a negative code with ref_last_error() holding the exception type and message
int ref_detector_check(const char* yml) { return guarded([&] { slam::FeatureDetector d{std::filesystem::path(yml)}; return 0; }); }
int ref_matcher_check(const char* yml) { return guarded([&] { slam::FeatureMatcher m{std::filesystem::path(yml)}; return 0; }); }

int ref_brief_pattern(const char* yml, int* out4, int cap_pairs) {
    return guarded([&] {
        slam::FeatureDetector d{std::filesystem::path(yml)};
        const auto& p = d.m_briefPattern;
        if ((int)p.size() > cap_pairs) return -1;
        for (size_t i = 0; i < p.size(); i++) {
            out4[4 * i + 0] = p[i].first.x();
            out4[4 * i + 1] = p[i].first.y();
            out4[4 * i + 2] = p[i].second.x();
            out4[4 * i + 3] = p[i].second.y();
        }
        return (int)p.size();
    });
}

int ref_fast_scan(const char* yml, const uint8_t* img, int rows, int cols, float* kps5, int cap) {
    return guarded([&] {
        slam::FeatureDetector d{std::filesystem::path(yml)};
        std::vector<slam::Keypoint> k;
        d.detectFASTKeypoints(to_eigen(img, rows, cols), k);
        return put_keypoints(k, kps5, cap);
    });
}

int ref_gaussian_blur(const uint8_t* img, int rows, int cols, int ksize, double sigma, uint8_t* out) {
    return guarded([&] {
        const slam::EigenGrayMatrix b = slam::FeatureDetector::gaussianBlur(to_eigen(img, rows, cols), ksize, sigma);
        std::memcpy(out, b.data(), (size_t)rows * cols);
        return 0;
    });
}

int ref_detect(const char* yml, const uint8_t* img, int rows, int cols, float* kps5, int cap) {
    return guarded([&] {
        slam::FeatureDetector d{std::filesystem::path(yml)};
        std::vector<slam::Keypoint> k;
        d.detect(to_eigen(img, rows, cols), k);
        return put_keypoints(k, kps5, cap);
    });
}

int ref_compute(const char* yml, const uint8_t* img, int rows, int cols, float* kps5, int n, uint8_t* desc) {
    return guarded([&] {
        slam::FeatureDetector d{std::filesystem::path(yml)};
        std::vector<slam::Keypoint> k = get_keypoints(kps5, n);
        slam::DescriptorMatrix dm;
        d.compute(to_eigen(img, rows, cols), k, dm);
        put_keypoints(k, kps5, n);
        if (dm.rows() > 0) std::memcpy(desc, dm.data(), (size_t)(dm.rows() * dm.cols()));
        return (int)dm.rows();
    });
}

int ref_detect_and_compute(const char* yml, const uint8_t* img, int rows, int cols, float* kps5, uint8_t* desc, int cap) {
    return guarded([&] {
        slam::FeatureDetector d{std::filesystem::path(yml)};
        std::vector<slam::Keypoint> k;
        slam::DescriptorMatrix dm;
        d.detectAndCompute(to_eigen(img, rows, cols), k, dm);
        if ((int)k.size() > cap) return -1;
        put_keypoints(k, kps5, cap);
        if (dm.rows() > 0) std::memcpy(desc, dm.data(), (size_t)(dm.rows() * dm.cols()));
        return (int)k.size();
    });
}

int ref_match(const char* yml, const uint8_t* d1, int n1, int w1, const uint8_t* d2, int n2, int w2, const float* kp1, int nk1,
              const float* kp2, int nk2, int* q, int* t, float* dist, int cap) {
    return guarded([&] {
        slam::FeatureMatcher m{std::filesystem::path(yml)};
        slam::DescriptorMatrix a(n1, w1), b(n2, w2);
        if (n1 > 0 && w1 > 0) std::memcpy(a.data(), d1, (size_t)n1 * w1);
        if (n2 > 0 && w2 > 0) std::memcpy(b.data(), d2, (size_t)n2 * w2);
        std::vector<slam::Match> out;
        m.match(a, b, out, get_keypoints(kp1, nk1), get_keypoints(kp2, nk2));
        if ((int)out.size() > cap) return -1;
        for (size_t i = 0; i < out.size(); i++) {
            q[i] = out[i].queryIdx;
            t[i] = out[i].trainIdx;
            dist[i] = out[i].distance;
        }
        return (int)out.size();
    });
}

int ref_hamming(const uint8_t* a, const uint8_t* b, int width) {
    slam::DescriptorMatrix m(2, width);
    std::memcpy(m.data(), a, (size_t)width);
    std::memcpy(m.data() + width, b, (size_t)width);
    return slam::calculateHammingDistance(m.row(0), m.row(1));
}

// Camera(configPath, index).undistortImage(cv::Mat&&): out = rows x cols doubles, written row-major
int ref_undistort(const char* camera_yml, int camera_index, const uint8_t* img, int rows, int cols, double* out) {
    return guarded([&] {
        slam::Camera cam{std::filesystem::path(camera_yml), camera_index};
        std::vector<uint8_t> copy(img, img + (size_t)rows * cols);
        cv::Mat raw = rows > 0 && cols > 0 ? cv::Mat(rows, cols, CV_8UC1, copy.data()) : cv::Mat();
        const Eigen::MatrixXd u = cam.undistortImage(std::move(raw));
        for (Eigen::Index r = 0; r < u.rows(); r++)
            for (Eigen::Index c = 0; c < u.cols(); c++) out[r * u.cols() + c] = u(r, c);
        return 0;
    });
}

}  // extern "C"
