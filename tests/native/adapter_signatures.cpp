// tests/native/adapter_signatures.cpp -- compile-time proof that the slam::cuda adapters (include/slam/cuda/frontend.hpp) accept
// the REFERENCE'S OWN TYPES with the reference's own call shapes: this TU includes the reference's real headers
// (include/slam/{common,frontend}/*.hpp from /root/reference, through the header stand-ins in oracle/shim for Eigen / OpenCV /
// spdlog) next to the adapters and instantiates every adapter method the way the reference's callers and tests call the
// originals (test/frontend/test_feature_detector.cpp, test_feature_matcher.cpp, test_pose_estimator.cpp,
// src/backend/loop_closure.cpp:157-158).  A signature drift on either side fails this build.  TEST INFRASTRUCTURE: compiled by
// tests/test_adapter_signatures.py with -fsyntax-only semantics (it is never run).
#include <filesystem>
#include <utility>
#include <vector>

#include <slam/common/common.hpp>
#include <slam/frontend/feature_detector.hpp>
#include <slam/frontend/feature_matcher.hpp>

#include <slam/cuda/frontend.hpp>

namespace {
struct Decode {  // the role cv::imread plays in Preprocessor::yield
    bool operator()(const std::filesystem::path&, int& rows, int& cols, std::vector<uint8_t>& bgr) const {
        rows = cols = 0;
        bgr.clear();
        return false;
    }
};
}  // namespace

int adapter_signatures(const std::filesystem::path& cfg) {
    // slam::FeatureDetector: ctor from a path; detect / compute / detectAndCompute on EigenGrayMatrix, std::vector<slam::Keypoint>,
    // slam::DescriptorMatrix (feature_detector.hpp:53, :114-135)
    slam::cuda::FeatureDetector det(cfg);
    slam::EigenGrayMatrix image(8, 8);
    std::vector<slam::Keypoint> kps;
    slam::DescriptorMatrix desc;
    det.detect(image, kps);
    det.compute(image, kps, desc);
    det.detectAndCompute(image, kps, desc);
    // slam::FeatureMatcher: match(d1, d2, out, kp1 = {}, kp2 = {}) const (feature_matcher.hpp:64-66)
    const slam::cuda::FeatureMatcher matcher(cfg);
    std::vector<slam::Match> matches;
    matcher.match(desc, desc, matches);
    matcher.match(desc, desc, matches, kps, kps);
    // slam::PoseEstimator: ctor from slam::Camera; estimate(pairs1, pairs2, matches, R, t) with KeyDescriptorPair and cv::Mat
    // (pose_estimator.hpp:15-18); triangulatePoints with cv::KeyPoint / cv::DMatch / cv::Mat -> std::vector<cv::Point3d> (:30-33)
    const slam::Camera camera(cfg, 0);
    slam::cuda::PoseEstimator pose(camera);
    std::vector<slam::KeyDescriptorPair> pairs1, pairs2;
    std::vector<std::pair<int, int>> idx;
    cv::Mat R, t;
    pose.estimate(pairs1, pairs2, idx, R, t);
    std::vector<cv::KeyPoint> k1, k2;
    std::vector<cv::DMatch> dm;
    std::vector<cv::Point3d> pts = pose.triangulatePoints<cv::Point3d>(k1, k2, dm, R, t);
    // slam::Camera::undistortImage -> Eigen::MatrixXd (common.hpp:127); slam::cuda::Camera mirrors the constructor
    slam::cuda::Camera cam2(cfg, 0);
    Eigen::MatrixXd und;
    cam2.undistortImage(image, und);
    slam::cuda::PoseEstimator pose2(cam2);
    // slam::Preprocessor: ctor (streamPath, camera, frameSkip), yield() -> pair<Eigen::MatrixXd, time_point> (preprocessor.hpp:30-36)
    slam::cuda::Preprocessor<Decode> pre(cfg, camera, Decode{}, 0);
    std::pair<Eigen::MatrixXd, std::chrono::system_clock::time_point> frame = pre.yield<Eigen::MatrixXd>();
    return static_cast<int>(pts.size() + matches.size() + static_cast<size_t>(frame.first.rows()));
}
