// frontend.hpp -- host C++ adapters that re-seat the reference's src/frontend classes on the CUDA C ABI (slamcu.h).
//
//   slam::cuda::FeatureDetector  <->  slam::FeatureDetector  (include/slam/frontend/feature_detector.hpp:47-192)
//   slam::cuda::FeatureMatcher   <->  slam::FeatureMatcher   (include/slam/frontend/feature_matcher.hpp:38-87)
//   slam::cuda::PoseEstimator    <->  slam::PoseEstimator    (include/slam/frontend/pose_estimator.hpp:13-36)
//   slam::cuda::Preprocessor     <->  slam::Preprocessor     (include/slam/preprocessing/preprocessor.hpp:22-54)
//   slam::cuda::EssentialSolver  <->  the cv::findEssentialMat call of PoseEstimator::estimate (pose_estimator.cpp:42)
//   slam::cuda::Camera           <->  slam::Camera (common.hpp:67-190): calibration YAML, undistortImage
//   slam::cuda::bgrToGray        <->  cv::cvtColor(BGR2GRAY) of Preprocessor::yield (preprocessor.cpp:136)
//
// Same constructors (a YAML path), same method names and argument order, same exception types and messages.  The
// methods are templates over the container types so that the header works both inside the reference tree, with
//   EigenGrayMatrix / DescriptorMatrix (row-major Eigen matrices: .data() .rows() .cols() .resize(r, c)),
//   slam::Keypoint {x, y, size, angle, response; Keypoint(x, y, size)} and slam::Match(q, t, dist),
// and without Eigen / OpenCV (tools/cli, test/frontend in this repository) with the look-alike types below.
// There is no CPU fallback: construction throws if no CUDA device is available.
#pragma once
#include <algorithm>
#include <chrono>
#include <cstdint>
#include <cstring>
#include <ctime>
#include <filesystem>
#include <fstream>
#include <iomanip>
#include <sstream>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

#include "slamcu.h"
#include "yaml_lite.hpp"

namespace slam::cuda {

// ---- Eigen-free look-alikes (layout-identical to the reference's types) --------------------------------------------
struct Keypoint {  // slam::Keypoint, feature_detector.hpp:28-38
    float x{}, y{}, size{}, angle{}, response{};
    Keypoint() = default;
    Keypoint(float X, float Y, float s = 6.0F) : x(X), y(Y), size(s) {}
};
struct Match {  // slam::Match, feature_matcher.hpp:18-25
    int queryIdx;
    int trainIdx;
    float distance;
    Match(int q, int t, float d) : queryIdx(q), trainIdx(t), distance(d) {}
};
template <class T>
class RowMajorMatrix {  // the subset of Eigen::Matrix<T, Dynamic, Dynamic, RowMajor> the adapters touch
public:
    RowMajorMatrix() = default;
    RowMajorMatrix(long r, long c) : m_rows(r), m_cols(c), m_data(static_cast<size_t>(r * c)) {}
    void resize(long r, long c) { m_rows = r; m_cols = c; m_data.assign(static_cast<size_t>(r * c), T{}); }
    long rows() const { return m_rows; }
    long cols() const { return m_cols; }
    T* data() { return m_data.data(); }
    const T* data() const { return m_data.data(); }
    T& operator()(long r, long c) { return m_data[static_cast<size_t>(r * m_cols + c)]; }
    const T& operator()(long r, long c) const { return m_data[static_cast<size_t>(r * m_cols + c)]; }
private:
    long m_rows = 0, m_cols = 0;
    std::vector<T> m_data;
};
using GrayMatrix = RowMajorMatrix<uint8_t>;
using DescriptorMatrix = RowMajorMatrix<uint8_t>;

// ---- context: one per device, shared by the adapters ----------------------------------------------------------------
class Context {
public:
    explicit Context(int device = 0) {
        const int st = slamcu_create(device, &m_ctx);
        if (st != SLAMCU_OK) throw std::runtime_error("slamcu_create failed: no usable CUDA device (there is no CPU fallback)");
    }
    ~Context() { slamcu_destroy(m_ctx); }
    Context(const Context&) = delete;
    Context& operator=(const Context&) = delete;
    slamcu_context* get() const { return m_ctx; }
    static Context& instance() {
        static Context ctx(0);
        return ctx;
    }
    // status -> the exception the reference throws for the same condition
    void check(int status) const {
        if (status == SLAMCU_OK) return;
        const std::string msg = slamcu_last_error(m_ctx);
        if (status == SLAMCU_EMPTY_INPUT) throw std::invalid_argument(msg.empty() ? "Empty descriptors provided." : msg);
        throw std::runtime_error(msg.empty() ? slamcu_status_string(status) : msg);
    }
private:
    slamcu_context* m_ctx = nullptr;
};

// ---- FeatureDetector -----------------------------------------------------------------------------------------------
class FeatureDetector {
public:
    explicit FeatureDetector(const std::filesystem::path& configPath, Context& ctx = Context::instance()) : m_ctx(ctx) {
        YamlLite fs(configPath.string());
        if (!fs.isOpened()) throw std::runtime_error("Could not open feature detector file: " + configPath.string());
        slamcu_detector_config c{};
        c.intensity_threshold = fs.getInt("IntensityThreshold");
        if (c.intensity_threshold < 0 || c.intensity_threshold > 255)
            throw std::runtime_error("Intensity threshold must be in the range [0, 255].");
        c.contiguous_pixels_threshold = fs.getInt("ContiguousPixelsThreshold");
        if (c.contiguous_pixels_threshold < 0 || c.contiguous_pixels_threshold > 16)
            throw std::runtime_error("Contiguous pixels threshold must be in the range [0, 16].");
        c.non_max_suppression = fs.getInt("NonMaxSuppression");
        if (c.non_max_suppression != 0 && c.non_max_suppression != 1)
            throw std::runtime_error("Non-max suppression must be either 0 (false) or 1 (true).");
        c.suppression_window_size = fs.getInt("SuppressionWindowSize");
        if (c.suppression_window_size <= 0) throw std::runtime_error("Suppression window size must be a positive integer.");
        c.patch_size = fs.getInt("PatchSize");
        if (c.patch_size <= 0 || c.patch_size % 2 == 0) throw std::runtime_error("Patch size must be a positive odd integer.");
        c.num_brief_pairs = fs.getInt("NumBRIEFPairs");
        if (c.num_brief_pairs <= 0 || c.num_brief_pairs % 8 != 0)
            throw std::runtime_error("Number of BRIEF pairs must be a positive multiple of 8.");
        // generateBRIEFPattern() and the blur kernel come from the host C++ library, like in the reference
        m_pattern.resize(static_cast<size_t>(c.num_brief_pairs) * 4);
        int n = 0;
        m_ctx.check(slamcu_default_brief_pattern(c.patch_size, c.num_brief_pairs, m_pattern.data(), c.num_brief_pairs, &n));
        m_ctx.check(slamcu_default_blur_weights(m_blur));
        c.n_pattern = n;
        c.pattern = m_pattern.data();
        c.blur_weights = m_blur;
        // opt-in OpenCV-ORB-compatible mode: keys the reference's YAML does not have
        if (fs.has("NumLevels") || fs.has("MaxFeatures") || fs.has("ScaleFactor")) {
            c.mode = SLAMCU_MODE_ORB;
            c.n_levels = fs.has("NumLevels") ? fs.getInt("NumLevels") : 8;
            c.scale_factor = fs.has("ScaleFactor") ? fs.getFloat("ScaleFactor") : 1.2F;
            c.max_features = fs.has("MaxFeatures") ? fs.getInt("MaxFeatures") : 2000;
            c.fast_threshold = fs.has("FastThreshold") ? fs.getInt("FastThreshold") : c.intensity_threshold;
            c.orb_pattern = nullptr;  // OpenCV's bit_pattern_31_, built into the library
        }
        m_descBytes = c.num_brief_pairs / 8;
        m_ctx.check(slamcu_detector_create(m_ctx.get(), &c, &m_det));
    }
    ~FeatureDetector() { slamcu_detector_destroy(m_det); }
    FeatureDetector(const FeatureDetector&) = delete;
    FeatureDetector& operator=(const FeatureDetector&) = delete;

    template <class Gray, class KP>
    void detect(const Gray& image, std::vector<KP>& keypoints) {
        keypoints.clear();
        run(image, keypoints, static_cast<std::vector<uint8_t>*>(nullptr), false);
    }

    template <class Gray, class KP, class Desc>
    void compute(const Gray& image, std::vector<KP>& keypoints, Desc& descriptors) {
        if (keypoints.empty()) {  // feature_detector.cpp:22-25
            descriptors = Desc(0, 0);
            return;
        }
        std::vector<slamcu_keypoint> k(keypoints.size());
        for (size_t i = 0; i < k.size(); i++) k[i] = {keypoints[i].x, keypoints[i].y, keypoints[i].size, keypoints[i].angle, keypoints[i].response};
        descriptors.resize(static_cast<long>(k.size()), m_descBytes);
        m_ctx.check(slamcu_compute(m_det, image.data(), static_cast<int>(image.rows()), static_cast<int>(image.cols()),
                                   static_cast<int>(image.cols()), k.data(), static_cast<int>(k.size()), descriptors.data(), m_descBytes));
        for (size_t i = 0; i < k.size(); i++) keypoints[i].angle = k[i].angle;
    }

    template <class Gray, class KP, class Desc>
    void detectAndCompute(const Gray& image, std::vector<KP>& keypoints, Desc& descriptors) {
        keypoints.clear();
        std::vector<uint8_t> d;
        run(image, keypoints, &d, true);
        if (keypoints.empty()) {
            descriptors = Desc(0, 0);
            return;
        }
        descriptors.resize(static_cast<long>(keypoints.size()), m_descBytes);
        std::copy(d.begin(), d.begin() + static_cast<long>(keypoints.size()) * m_descBytes, descriptors.data());
    }

    slamcu_detector* handle() const { return m_det; }
    int descriptorBytes() const { return m_descBytes; }

private:
    template <class Gray, class KP>
    void run(const Gray& image, std::vector<KP>& keypoints, std::vector<uint8_t>* desc, bool withDesc) {
        const int rows = static_cast<int>(image.rows()), cols = static_cast<int>(image.cols());
        int cap = std::max(4096, rows * cols / 16), n = 0;
        std::vector<slamcu_keypoint> k;
        for (;;) {
            k.resize(static_cast<size_t>(cap));
            int st;
            if (withDesc) {
                desc->resize(static_cast<size_t>(cap) * m_descBytes);
                st = slamcu_detect_and_compute(m_det, image.data(), rows, cols, cols, k.data(), desc->data(), m_descBytes, cap, &n);
            } else {
                st = slamcu_detect(m_det, image.data(), rows, cols, cols, k.data(), cap, &n);
            }
            if (st == SLAMCU_CAPACITY && n > cap) { cap = n; continue; }
            m_ctx.check(st);
            break;
        }
        keypoints.reserve(static_cast<size_t>(n));
        for (int i = 0; i < n; i++) {
            KP kp(k[i].x, k[i].y, k[i].size);
            kp.angle = k[i].angle;
            kp.response = k[i].response;
            keypoints.push_back(kp);
        }
    }
    Context& m_ctx;
    slamcu_detector* m_det = nullptr;
    std::vector<int32_t> m_pattern;
    double m_blur[25]{};
    int m_descBytes = 32;
};

// ---- FeatureMatcher --------------------------------------------------------------------------------------------------
class FeatureMatcher {
public:
    explicit FeatureMatcher(const std::filesystem::path& configPath, Context& ctx = Context::instance()) : m_ctx(ctx) {
        YamlLite fs(configPath.string());
        if (!fs.isOpened()) throw std::runtime_error("Could not open feature matcher config file: " + configPath.string());
        slamcu_matcher_config c{};
        const std::string dt = fs.getString("DistanceType");
        if (dt == "HAMMING") c.distance_type = SLAMCU_DISTANCE_HAMMING;
        else if (dt == "L2") c.distance_type = SLAMCU_DISTANCE_L2;
        else throw std::runtime_error("Invalid distance type. Must be 'HAMMING' or 'L2'.");
        c.filter_matches = fs.getInt("FilterMatches");
        if (c.filter_matches != 0 && c.filter_matches != 1) throw std::runtime_error("FilterMatches must be either 0 (false) or 1 (true).");
        c.good_matches_count = fs.getInt("GoodMatchesCount");
        if (c.filter_matches && c.good_matches_count <= 0)
            throw std::runtime_error("GoodMatchesCount must be positive when filtering is enabled.");
        c.use_ratio_test = fs.getInt("UseRatioTest");
        if (c.use_ratio_test != 0 && c.use_ratio_test != 1) throw std::runtime_error("UseRatioTest must be either 0 (false) or 1 (true).");
        c.ratio_test_threshold = fs.getFloat("RatioTestThreshold");
        if (c.ratio_test_threshold < 0.0F || c.ratio_test_threshold > 1.0F)
            throw std::runtime_error("RatioTestThreshold must be in the range [0, 1].");
        m_ctx.check(slamcu_matcher_create(m_ctx.get(), &c, &m_matcher));
    }
    ~FeatureMatcher() { slamcu_matcher_destroy(m_matcher); }
    FeatureMatcher(const FeatureMatcher&) = delete;
    FeatureMatcher& operator=(const FeatureMatcher&) = delete;

    template <class Desc, class M, class KP = Keypoint>
    void match(const Desc& descriptors1, const Desc& descriptors2, std::vector<M>& matches,
               const std::vector<KP>& keypoints1 = {}, const std::vector<KP>& keypoints2 = {}) const {
        matches.clear();  // feature_matcher.cpp:76
        auto pack = [](const std::vector<KP>& in) {
            std::vector<slamcu_keypoint> out(in.size());
            for (size_t i = 0; i < in.size(); i++) out[i] = {in[i].x, in[i].y, in[i].size, in[i].angle, in[i].response};
            return out;
        };
        const std::vector<slamcu_keypoint> k1 = pack(keypoints1), k2 = pack(keypoints2);
        const int n1 = static_cast<int>(descriptors1.rows()), n2 = static_cast<int>(descriptors2.rows());
        std::vector<slamcu_dmatch> out(static_cast<size_t>(std::max(n1, 1)));
        int n = 0;
        m_ctx.check(slamcu_match(m_matcher, n1 ? descriptors1.data() : nullptr, n1, static_cast<int>(descriptors1.cols()),
                                 n2 ? descriptors2.data() : nullptr, n2, static_cast<int>(descriptors2.cols()),
                                 k1.empty() ? nullptr : k1.data(), static_cast<int>(k1.size()), k2.empty() ? nullptr : k2.data(),
                                 static_cast<int>(k2.size()), out.data(), static_cast<int>(out.size()), &n));
        matches.reserve(static_cast<size_t>(n));
        for (int i = 0; i < n; i++) matches.emplace_back(out[i].queryIdx, out[i].trainIdx, out[i].distance);
    }
    slamcu_matcher* handle() const { return m_matcher; }

private:
    Context& m_ctx;
    slamcu_matcher* m_matcher = nullptr;
};

// ---- slam::Camera (common.hpp:67-190) and the grayscale step of Preprocessor::yield (preprocessor.cpp:136) -------------
class Camera {
public:
    // same argument meaning and error messages as the reference constructor (common.hpp:76-122)
    explicit Camera(const std::filesystem::path& configPath, int cameraIndex = 0, Context& ctx = Context::instance())
        : m_ctx(ctx), m_cameraIndex(cameraIndex) {
        YamlLite fs(configPath.string());
        if (!fs.isOpened()) throw std::runtime_error("Could not open calibration file: " + configPath.string());
        const std::string kKey = "K" + std::to_string(cameraIndex), dKey = "D" + std::to_string(cameraIndex);
        const std::vector<double> K = fs.getDoubles(kKey), D = fs.getDoubles(dKey), size = fs.getDoubles("ImageSize");
        if (K.size() < 9 || D.empty()) throw std::runtime_error("Could not find keys " + kKey + " or " + dKey + " in file.");
        for (int i = 0; i < 9; i++) m_K[i] = K[static_cast<size_t>(i)];
        m_D = D;
        m_width = size.size() > 0 ? static_cast<int>(size[0]) : 0;
        m_height = size.size() > 1 ? static_cast<int>(size[1]) : 0;
    }
    const double* getIntrinsicMatrix() const { return m_K; }                 // row-major 3x3
    const std::vector<double>& getDistortionCoefficients() const { return m_D; }
    int width() const { return m_width; }
    int height() const { return m_height; }
    void intrinsics4(double K4[4]) const { K4[0] = m_K[0]; K4[1] = m_K[4]; K4[2] = m_K[2]; K4[3] = m_K[5]; }

    // Camera::undistortImage (common.hpp:127-173): raw = rows x cols 8-bit grayscale (row-major, .data() / .rows() /
    // .cols()); out(i, j) receives the reference's value / 255.0 image (any matrix type with resize(rows, cols) and
    // operator()(i, j): Eigen::MatrixXd in the reference tree).  Same exceptions as the reference (:130-135).
    template <class Gray, class Out>
    void undistortImage(const Gray& raw, Out& out) const {
        std::vector<double> f64;
        run(raw, nullptr, &f64);
        const int rows = static_cast<int>(raw.rows()), cols = static_cast<int>(raw.cols());
        out.resize(rows, cols);
        for (int i = 0; i < rows; i++)
            for (int j = 0; j < cols; j++) out(i, j) = f64[static_cast<size_t>(i) * cols + j];
    }
    // the same gather emitted as the 8-bit image the detector consumes (the /255.0 double image is a dead end in the
    // reference: nothing downstream reads it)
    template <class Gray>
    void undistortImageU8(const Gray& raw, Gray& out) const {
        out.resize(raw.rows(), raw.cols());
        run(raw, out.data(), nullptr);
    }

private:
    template <class Gray>
    void run(const Gray& raw, uint8_t* u8, std::vector<double>* f64) const {
        const int rows = static_cast<int>(raw.rows()), cols = static_cast<int>(raw.cols());
        if (rows == 0 || cols == 0) throw std::runtime_error("Input image is empty.");
        if (cols != m_width || rows != m_height) throw std::runtime_error("Input image size does not match camera image size.");
        double K4[4];
        intrinsics4(K4);
        double D4[4] = {0, 0, 0, 0};  // k1, k2, p1, p2; k3 is loaded but unused by the reference (:113, :151-154)
        for (size_t i = 0; i < 4 && i < m_D.size(); i++) D4[i] = m_D[i];
        if (f64) f64->assign(static_cast<size_t>(rows) * cols, 0.0);
        m_ctx.check(slamcu_undistort(m_ctx.get(), raw.data(), rows, cols, cols, K4, D4, u8, f64 ? f64->data() : nullptr));
    }
    Context& m_ctx;
    int m_cameraIndex = 0, m_width = 0, m_height = 0;
    double m_K[9]{};
    std::vector<double> m_D;
};

// cv::cvtColor(image, image, cv::COLOR_BGR2GRAY) of Preprocessor::yield (preprocessor.cpp:136): bgr = rows x cols x 3
// interleaved bytes; gray receives rows x cols bytes (bit-exact with OpenCV's fixed-point formula)
template <class Gray>
void bgrToGray(const uint8_t* bgr, int rows, int cols, Gray& gray, Context& ctx = Context::instance()) {
    gray.resize(rows, cols);
    ctx.check(slamcu_bgr_to_gray(ctx.get(), bgr, rows, cols, cols * 3, gray.data(), cols));
}

// ---- the cv::findEssentialMat call of PoseEstimator::estimate -------------------------------------------------------
struct EssentialResult {
    bool valid = false;       // false: fewer than 8 matches (pose_estimator.cpp:22-26) or no hypothesis accepted (:44-47)
    double E[9]{};            // row-major, |E|_F = 1
    std::vector<uint8_t> mask;
    int inliers = 0;
};
class EssentialSolver {
public:
    // K4 = fx, fy, cx, cy of slam::Camera::getIntrinsicMatrix()
    explicit EssentialSolver(const double K4[4], Context& ctx = Context::instance()) : m_ctx(ctx) {
        for (int i = 0; i < 4; i++) m_K[i] = K4[i];
    }
    // pts1 / pts2: matched pixel coordinates (x, y) as PoseEstimator::estimate gathers them (pose_estimator.cpp:30-35)
    EssentialResult solve(const std::vector<float>& pts1, const std::vector<float>& pts2, double prob = 0.999,
                          double threshold = 1.0, int maxIters = 1000) const {
        EssentialResult r;
        const int n = static_cast<int>(pts1.size() / 2);
        if (n < 8) return r;
        r.mask.resize(static_cast<size_t>(n));
        m_ctx.check(slamcu_find_essential(m_ctx.get(), pts1.data(), pts2.data(), n, m_K, prob, threshold, maxIters, r.E,
                                          r.mask.data(), &r.inliers));
        r.valid = r.inliers > 0;
        return r;
    }
    // PoseEstimator::estimate end to end: findEssentialMat + simpleRecoverPose (simple_pose_recover.cpp:35-97).
    // valid == false where the reference logs a warning and leaves R, t untouched.
    struct Pose {
        bool valid = false;
        double R[9]{}, t[3]{}, E[9]{};
        int inliers = 0, front[4]{};
        std::vector<uint8_t> mask;
    };
    Pose estimate(const std::vector<float>& pts1, const std::vector<float>& pts2) const {
        Pose p;
        const int n = static_cast<int>(pts1.size() / 2);
        if (n < 8) return p;
        p.mask.resize(static_cast<size_t>(n));
        const int st = slamcu_estimate_pose(m_ctx.get(), pts1.data(), pts2.data(), n, m_K, p.E, p.mask.data(), &p.inliers, p.R, p.t, p.front);
        if (st == SLAMCU_EMPTY_INPUT) return p;
        m_ctx.check(st);
        p.valid = true;
        return p;
    }

private:
    Context& m_ctx;
    double m_K[4]{};
};

// ---- slam::PoseEstimator (pose_estimator.hpp:13-36) ----------------------------------------------------------------------
namespace detail {
// fx, fy, cx, cy out of Camera::getIntrinsicMatrix(): an Eigen::Matrix3d in the reference (operator()(r, c)), a row-major
// double[9] in slam::cuda::Camera
template <class CameraT>
void intrinsics_of(const CameraT& camera, double K4[4]) {
    const auto& K = camera.getIntrinsicMatrix();
    if constexpr (std::is_pointer_v<std::decay_t<decltype(K)>>) {
        K4[0] = K[0]; K4[1] = K[4]; K4[2] = K[2]; K4[3] = K[5];
    } else {
        K4[0] = K(0, 0); K4[1] = K(1, 1); K4[2] = K(0, 2); K4[3] = K(1, 2);
    }
}
// cv::Mat-shaped outputs: create(rows, cols, CV_64F) + contiguous `data` (cv::Mat in the reference tree)
template <class MatT>
void assign_f64(MatT& m, int rows, int cols, const double* v) {
    m.create(rows, cols, 6 /* CV_64F */);
    std::memcpy(m.data, v, sizeof(double) * static_cast<size_t>(rows) * cols);
}
}  // namespace detail

class PoseEstimator {
public:
    // pose_estimator.hpp:15.  The reference keeps a reference to the camera; only its intrinsic matrix is ever read
    // (pose_estimator.cpp:40-41, :72-73), so the four intrinsics are copied here.
    template <class CameraT>
    explicit PoseEstimator(const CameraT& camera, Context& ctx = Context::instance()) : m_ctx(ctx) {
        detail::intrinsics_of(camera, m_K);
    }
    // PoseEstimator::estimate (pose_estimator.cpp:18-67).  pairs: std::vector<KeyDescriptorPair> (only .first.pt is read,
    // :33-34); R, t: cv::Mat.  Like the reference, returns WITHOUT touching R, t below 8 matches (:22-26) and when no essential
    // matrix was accepted (:44-47).
    template <class Pair, class MatT>
    void estimate(const std::vector<Pair>& pairs1, const std::vector<Pair>& pairs2, const std::vector<std::pair<int, int>>& matches, MatT& R,
                  MatT& t) {
        if (matches.size() < 8) return;
        std::vector<float> p1, p2;
        p1.reserve(matches.size() * 2);
        p2.reserve(matches.size() * 2);
        for (const auto& m : matches) {
            p1.push_back(pairs1[static_cast<size_t>(m.first)].first.pt.x);
            p1.push_back(pairs1[static_cast<size_t>(m.first)].first.pt.y);
            p2.push_back(pairs2[static_cast<size_t>(m.second)].first.pt.x);
            p2.push_back(pairs2[static_cast<size_t>(m.second)].first.pt.y);
        }
        double E[9], Rv[9], tv[3];
        int inliers = 0, front[4];
        const int st = slamcu_estimate_pose(m_ctx.get(), p1.data(), p2.data(), static_cast<int>(matches.size()), m_K, E, nullptr, &inliers, Rv, tv, front);
        if (st == SLAMCU_EMPTY_INPUT) return;  // "Essential Matrix could not be computed."
        m_ctx.check(st);
        detail::assign_f64(R, 3, 3, Rv);
        detail::assign_f64(t, 3, 1, tv);
    }
    // PoseEstimator::triangulatePoints (pose_estimator.cpp:69-104): P1 = K [I | 0], P2 = K [R | t], slam::triangulate per match,
    // x / x[3].  P3: cv::Point3d; KP: cv::KeyPoint; DM: cv::DMatch; R, t: CV_64F cv::Mat (read through .data like R.at<double>).
    // (The reference's own result is indeterminate -- Mat::copyTo into a differently typed column, common.hpp:217 -- this is the
    // value that code evidently means; slamcu.h.)
    template <class P3, class KP, class DM, class MatT>
    std::vector<P3> triangulatePoints(const std::vector<KP>& keypoints1, const std::vector<KP>& keypoints2, const std::vector<DM>& matches,
                                      const MatT& R, const MatT& t) {
        const double* r = reinterpret_cast<const double*>(R.data);
        const double* tv = reinterpret_cast<const double*>(t.data);
        const double fx = m_K[0], fy = m_K[1], cx = m_K[2], cy = m_K[3];
        const double P1[12] = {fx, 0, cx, 0, 0, fy, cy, 0, 0, 0, 1, 0};
        double P2[12];
        for (int j = 0; j < 4; j++) {
            const double c0 = j < 3 ? r[j] : tv[0], c1 = j < 3 ? r[3 + j] : tv[1], c2 = j < 3 ? r[6 + j] : tv[2];
            P2[j] = fx * c0 + cx * c2;
            P2[4 + j] = fy * c1 + cy * c2;
            P2[8 + j] = c2;
        }
        std::vector<float> p1, p2;
        for (const auto& m : matches) {
            p1.push_back(keypoints1[static_cast<size_t>(m.queryIdx)].pt.x);
            p1.push_back(keypoints1[static_cast<size_t>(m.queryIdx)].pt.y);
            p2.push_back(keypoints2[static_cast<size_t>(m.trainIdx)].pt.x);
            p2.push_back(keypoints2[static_cast<size_t>(m.trainIdx)].pt.y);
        }
        std::vector<double> x3(matches.size() * 3);
        m_ctx.check(slamcu_triangulate(m_ctx.get(), P1, P2, p1.data(), p2.data(), static_cast<int>(matches.size()), nullptr, x3.data()));
        std::vector<P3> out;
        out.reserve(matches.size());
        for (size_t i = 0; i < matches.size(); i++) out.emplace_back(x3[3 * i], x3[3 * i + 1], x3[3 * i + 2]);
        return out;
    }

private:
    Context& m_ctx;
    double m_K[4]{};
};

// ---- slam::Preprocessor (preprocessor.hpp:22-54, preprocessor.cpp) -------------------------------------------------------
// Directory streams: the same file selection, lexical order, timestamps.txt parsing and error messages as the reference
// (preprocessor.cpp:24-82).  Decoding a file into BGR bytes is the one step that stays a host library call (cv::imread there):
// the adapter takes it as a callable  bool decode(const std::filesystem::path&, int& rows, int& cols, std::vector<uint8_t>& bgr).
// yield() does cv::cvtColor(BGR2GRAY) + Camera::undistortImage on the device and returns the reference's double image;
// yieldInto() does the same for a batch straight into a device-resident sequence (slamcu_sequence_prepare) -- the 8-bit image
// the detector consumes, with no double image in between.  Video files need cv::VideoCapture and are not supported here.
template <class Decoder>
class Preprocessor {
public:
    using TimePoint = std::chrono::system_clock::time_point;
    template <class CameraT>
    Preprocessor(std::filesystem::path streamPath, const CameraT& camera, Decoder decode, int frameSkip = 0, Context& ctx = Context::instance())
        : m_ctx(ctx), m_decode(std::move(decode)), m_frameSkip(frameSkip), m_streamPath(std::move(streamPath)) {
        detail::intrinsics_of(camera, m_K);
        const auto& D = camera.getDistortionCoefficients();
        for (int i = 0; i < 4; i++) m_D[i] = static_cast<long>(D.size()) > i ? static_cast<double>(D[static_cast<size_t>(i)]) : 0.0;
        if (std::filesystem::is_directory(m_streamPath)) prepareDirectory();
        else if (std::filesystem::is_regular_file(m_streamPath)) throw std::runtime_error("Could not open video file: " + m_streamPath.string());
        else throw std::runtime_error("Unsupported stream type: " + m_streamPath.string());
    }
    int totalFrames() const { return m_totalFrames; }

    // Preprocessor::yield (preprocessor.cpp:95-141).  Out: Eigen::MatrixXd (resize(rows, cols), operator()(i, j)).  An empty
    // (0 x 0) image and a default time point at the end of the stream, like the reference's default-constructed pair.
    template <class Out>
    std::pair<Out, TimePoint> yield() {
        std::pair<Out, TimePoint> pair;
        if (m_frameNumber >= m_totalFrames) return pair;
        int rows = 0, cols = 0;
        std::vector<uint8_t> bgr;
        if (!m_decode(m_files[static_cast<size_t>(m_frameNumber)], rows, cols, bgr) || rows <= 0 || cols <= 0)
            throw std::runtime_error("Failed to read image from file: " + m_files[static_cast<size_t>(m_frameNumber)].string());
        std::vector<uint8_t> gray(static_cast<size_t>(rows) * cols);
        m_ctx.check(slamcu_bgr_to_gray(m_ctx.get(), bgr.data(), rows, cols, cols * 3, gray.data(), cols));
        std::vector<double> f64(gray.size());
        m_ctx.check(slamcu_undistort(m_ctx.get(), gray.data(), rows, cols, cols, m_K, m_D, nullptr, f64.data()));
        pair.first.resize(rows, cols);
        for (int i = 0; i < rows; i++)
            for (int j = 0; j < cols; j++) pair.first(i, j) = f64[static_cast<size_t>(i) * cols + j];
        pair.second = stampOf(m_frameNumber);
        m_frameNumber += 1 + m_frameSkip;
        return pair;
    }

    // Batched yield: up to n frames are decoded, uploaded as BGR with one copy and converted + undistorted on the device into
    // slots [first, first + count) of `seq`.  Returns count (0 at the end of the stream).
    int yieldInto(slamcu_sequence* seq, int first, int n, std::vector<TimePoint>* stamps = nullptr) {
        std::vector<uint8_t> batch;
        int count = 0, rows0 = 0, cols0 = 0;
        while (count < n && m_frameNumber < m_totalFrames) {
            int rows = 0, cols = 0;
            std::vector<uint8_t> bgr;
            if (!m_decode(m_files[static_cast<size_t>(m_frameNumber)], rows, cols, bgr) || rows <= 0 || cols <= 0)
                throw std::runtime_error("Failed to read image from file: " + m_files[static_cast<size_t>(m_frameNumber)].string());
            if (count == 0) { rows0 = rows; cols0 = cols; }
            if (rows != rows0 || cols != cols0) throw std::runtime_error("Input image size does not match camera image size.");
            batch.insert(batch.end(), bgr.begin(), bgr.end());
            if (stamps) stamps->push_back(stampOf(m_frameNumber));
            m_frameNumber += 1 + m_frameSkip;
            count++;
        }
        if (count > 0) {
            m_ctx.check(slamcu_sequence_prepare(seq, first, count, batch.data(), 3, cols0 * 3, m_K, m_D));
            m_ctx.check(slamcu_synchronize(m_ctx.get()));  // `batch` goes out of scope
        }
        return count;
    }

private:
    void prepareDirectory() {  // preprocessor.cpp:24-82
        for (const auto& entry : std::filesystem::directory_iterator(m_streamPath)) {
            if ((entry.is_regular_file() && entry.path().extension() == ".jpg") || entry.path().extension() == ".png") {  // (sic, :34-35)
                m_totalFrames++;
                m_files.push_back(entry.path());
            }
        }
        std::sort(m_files.begin(), m_files.end());
        std::ifstream file(m_streamPath / "timestamps.txt");
        if (!file) throw std::runtime_error("Could not open timestamps.txt in directory: " + m_streamPath.string());
        std::string line;
        while (std::getline(file, line)) {
            const auto decimalPos = line.find('.');
            if (decimalPos == std::string::npos) continue;
            std::tm timeStruct = {};
            std::stringstream ss(line.substr(0, decimalPos));
            ss >> std::get_time(&timeStruct, "%Y-%m-%d %H:%M:%S");
            if (ss.fail()) continue;
            auto timePoint = std::chrono::system_clock::from_time_t(std::mktime(&timeStruct));
            timePoint += std::chrono::duration_cast<std::chrono::system_clock::duration>(std::chrono::nanoseconds(std::stoll(line.substr(decimalPos + 1))));
            m_timestamps.push_back(timePoint);
        }
        if (static_cast<int>(m_timestamps.size()) != m_totalFrames) throw std::runtime_error("Number of timestamps does not match number of frames.");
    }
    TimePoint stampOf(int frame) const {  // preprocessor.cpp:114-119: the stamp truncated to milliseconds
        const double ms = static_cast<double>(m_timestamps[static_cast<size_t>(frame)].time_since_epoch().count()) / 1.0e6;
        return TimePoint(std::chrono::milliseconds(static_cast<int64_t>(ms)));
    }
    Context& m_ctx;
    Decoder m_decode;
    int m_frameNumber = 0, m_totalFrames = 0, m_frameSkip = 0;
    std::vector<TimePoint> m_timestamps;
    std::vector<std::filesystem::path> m_files;
    std::filesystem::path m_streamPath;
    double m_K[4]{}, m_D[4]{};
};

}  // namespace slam::cuda
