// test_frontend_cuda -- the reference's three frontend smoke tests (test/frontend/test_feature_detector.cpp,
// test_feature_matcher.cpp, test_pose_estimator.cpp) re-seated on the CUDA adapters, with VALUE output: it prints
// counts and FNV-1a-64 digests of every result so that tests/test_gpu_cpp_host.py can compare the C++ host path with
// the oracle.  Images are binary PGM (the reference reads PNG through OpenCV, which this image does not have in C++).
//
//   test_frontend_cuda image0.pgm image1.pgm detector.yml matcher.yml [fx fy cx cy]
//   test_frontend_cuda --camera camera.yml [index] image.pgm      (test/preprocessing/test_preprocessor.cpp's image path:
//                                                                  BGR2GRAY + Camera::undistortImage, digests of both)
//   test_frontend_cuda --preprocess dir camera.yml                 (slam::Preprocessor over a directory stream: yield() per frame
//                                                                  and the batched yieldInto(); frames are binary PPM payloads)
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <string>
#include <vector>

#include <slam/cuda/frontend.hpp>

using namespace slam::cuda;

static GrayMatrix read_pgm(const std::string& path) {
    std::ifstream in(path, std::ios::binary);
    std::string magic;
    int w = 0, h = 0, maxv = 0;
    in >> magic >> w >> h >> maxv;
    in.get();
    if (!in || magic != "P5" || maxv != 255) throw std::runtime_error("Could not read image: " + path);
    GrayMatrix m(h, w);
    in.read(reinterpret_cast<char*>(m.data()), static_cast<std::streamsize>(h) * w);
    if (in.gcount() != static_cast<std::streamsize>(h) * w) throw std::runtime_error("Truncated image: " + path);
    return m;
}

static unsigned long long fnv(const void* p, size_t n, unsigned long long h = 1469598103934665603ULL) {
    const unsigned char* b = static_cast<const unsigned char*>(p);
    for (size_t i = 0; i < n; i++) { h ^= b[i]; h *= 1099511628211ULL; }
    return h;
}

static int camera_main(int argc, char** argv) {
    try {
        const int index = argc >= 5 ? std::atoi(argv[3]) : 0;
        const GrayMatrix img = read_pgm(argv[argc - 1]);
        Camera cam(argv[2], index);
        std::printf("camera %dx%d K", cam.width(), cam.height());
        for (int i = 0; i < 9; i++) std::printf(" %.17g", cam.getIntrinsicMatrix()[i]);
        std::printf(" D");
        for (double d : cam.getDistortionCoefficients()) std::printf(" %.17g", d);
        std::printf("\n");
        RowMajorMatrix<double> und;
        cam.undistortImage(img, und);
        GrayMatrix und8;
        cam.undistortImageU8(img, und8);
        std::printf("undistort f64 %016llx u8 %016llx\n", fnv(und.data(), static_cast<size_t>(und.rows() * und.cols()) * sizeof(double)),
                    fnv(und8.data(), static_cast<size_t>(und8.rows() * und8.cols())));
        // a deterministic colour image from the gray one: B = v, G = 255 - v, R = v / 2
        std::vector<uint8_t> bgr(static_cast<size_t>(img.rows() * img.cols()) * 3);
        for (long i = 0; i < img.rows() * img.cols(); i++) {
            const uint8_t v = img.data()[i];
            bgr[3 * i] = v; bgr[3 * i + 1] = static_cast<uint8_t>(255 - v); bgr[3 * i + 2] = static_cast<uint8_t>(v >> 1);
        }
        GrayMatrix gray;
        bgrToGray(bgr.data(), static_cast<int>(img.rows()), static_cast<int>(img.cols()), gray);
        std::printf("gray %016llx\n", fnv(gray.data(), static_cast<size_t>(gray.rows() * gray.cols())));
        try {
            GrayMatrix small(10, 10);
            RowMajorMatrix<double> o;
            cam.undistortImage(small, o);
            std::fprintf(stderr, "expected std::runtime_error\n");
            return -1;
        } catch (const std::runtime_error& e) {
            std::printf("mismatch: %s\n", e.what());
        }
        try {
            Camera missing(argv[2], 7);
            return -1;
        } catch (const std::runtime_error& e) {
            std::printf("missing: %s\n", e.what());
        }
    } catch (const std::exception& e) {
        std::fprintf(stderr, "Exception: %s\n", e.what());
        return -1;
    }
    return 0;
}

// the role cv::imread(path, IMREAD_COLOR) plays in Preprocessor::yield: here the files hold binary PPM (P6) payloads whose three
// bytes per pixel are taken as B, G, R
struct PpmDecoder {
    bool operator()(const std::filesystem::path& path, int& rows, int& cols, std::vector<uint8_t>& bgr) const {
        std::ifstream in(path, std::ios::binary);
        std::string magic;
        int maxv = 0;
        in >> magic >> cols >> rows >> maxv;
        in.get();
        if (!in || magic != "P6" || maxv != 255) return false;
        bgr.resize(static_cast<size_t>(rows) * cols * 3);
        in.read(reinterpret_cast<char*>(bgr.data()), static_cast<std::streamsize>(bgr.size()));
        return in.gcount() == static_cast<std::streamsize>(bgr.size());
    }
};

static int preprocess_main(char** argv) {
    try {
        Camera cam(argv[3], 0);
        Preprocessor<PpmDecoder> pre(argv[2], cam, PpmDecoder{}, 0);
        std::printf("frames %d\n", pre.totalFrames());
        for (;;) {  // test_preprocessing.cpp: yield until the stream ends
            auto frame = pre.yield<RowMajorMatrix<double>>();
            if (frame.first.rows() == 0) break;
            const long long ms = std::chrono::duration_cast<std::chrono::milliseconds>(frame.second.time_since_epoch()).count();
            std::printf("yield %ldx%ld %016llx ms_mod %lld\n", frame.first.rows(), frame.first.cols(),
                        fnv(frame.first.data(), static_cast<size_t>(frame.first.rows() * frame.first.cols()) * sizeof(double)), ms % 1000);
        }
        // the batched path: BGR2GRAY + undistortion straight into a device-resident sequence, every second frame (frameSkip = 1)
        Preprocessor<PpmDecoder> pre2(argv[2], cam, PpmDecoder{}, 1);
        slamcu_sequence* seq = nullptr;
        Context& ctx = Context::instance();
        ctx.check(slamcu_sequence_create(ctx.get(), cam.height(), cam.width(), 8, 0, 0, 32, &seq));
        const int got = pre2.yieldInto(seq, 0, 8);
        std::printf("batched %d\n", got);
        for (int f = 0; f < got; f++) {
            GrayMatrix img(cam.height(), cam.width());
            ctx.check(slamcu_sequence_image(seq, f, img.data(), cam.width()));
            std::printf("slot %d %016llx\n", f, fnv(img.data(), static_cast<size_t>(img.rows() * img.cols())));
        }
        slamcu_sequence_destroy(seq);
        try {
            Preprocessor<PpmDecoder> bad(std::string(argv[2]) + "/does-not-exist", cam, PpmDecoder{});
            return -1;
        } catch (const std::runtime_error& e) {
            std::printf("bad: %s\n", e.what());
        }
    } catch (const std::exception& e) {
        std::fprintf(stderr, "Exception: %s\n", e.what());
        return -1;
    }
    return 0;
}

// cv::Mat / cv::KeyPoint / cv::DMatch / cv::Point3d / KeyDescriptorPair look-alikes for slam::cuda::PoseEstimator (the shapes of
// the reference's types: tests/native/adapter_signatures.cpp compiles the same calls against the reference's own headers)
struct MockMat {
    int rows = 0, cols = 0;
    unsigned char* data = nullptr;
    std::vector<unsigned char> buf;
    void create(int r, int c, int /*type: CV_64F*/) { rows = r; cols = c; buf.assign(static_cast<size_t>(r) * c * 8, 0); data = buf.data(); }
    bool empty() const { return data == nullptr; }
};
struct MockPoint2f { float x, y; };
struct MockKeyPoint { MockPoint2f pt; };
struct MockDMatch { int queryIdx, trainIdx; };
struct MockPoint3d { double x, y, z; MockPoint3d(double X, double Y, double Z) : x(X), y(Y), z(Z) {} };
struct CameraK {  // only getIntrinsicMatrix() is read
    double K[9];
    const double* getIntrinsicMatrix() const { return K; }
};

int main(int argc, char** argv) {
    if (argc >= 4 && std::string(argv[1]) == "--camera") return camera_main(argc, argv);
    if (argc >= 4 && std::string(argv[1]) == "--preprocess") return preprocess_main(argv);
    if (argc < 5) {
        std::fprintf(stderr, "usage: %s image0.pgm image1.pgm detector.yml matcher.yml [fx fy cx cy]\n", argv[0]);
        return -1;
    }
    try {
        const GrayMatrix img0 = read_pgm(argv[1]), img1 = read_pgm(argv[2]);
        FeatureDetector detector(argv[3]);
        FeatureMatcher matcher(argv[4]);
        // test_feature_detector.cpp: detect, compute, detectAndCompute
        std::vector<Keypoint> kps;
        detector.detect(img0, kps);
        std::vector<Keypoint> kps0, kps1;
        DescriptorMatrix desc0, desc1;
        detector.detectAndCompute(img0, kps0, desc0);
        detector.detectAndCompute(img1, kps1, desc1);
        bool same = kps.size() == kps0.size();
        try {
            DescriptorMatrix desc;
            detector.compute(img0, kps, desc);
            same = same && fnv(desc.data(), desc.rows() * desc.cols()) == fnv(desc0.data(), desc0.rows() * desc0.cols());
        } catch (const std::runtime_error& e) {  // ORB mode: compute() on supplied keypoints is not offered
            std::fprintf(stderr, "compute: %s\n", e.what());
        }
        if (!same) {
            std::fprintf(stderr, "detect+compute differs from detectAndCompute\n");
            return -1;
        }
        std::printf("kp0 %zu %016llx desc0 %016llx\n", kps0.size(), fnv(kps0.data(), kps0.size() * sizeof(Keypoint)),
                    fnv(desc0.data(), static_cast<size_t>(desc0.rows() * desc0.cols())));
        std::printf("kp1 %zu %016llx desc1 %016llx\n", kps1.size(), fnv(kps1.data(), kps1.size() * sizeof(Keypoint)),
                    fnv(desc1.data(), static_cast<size_t>(desc1.rows() * desc1.cols())));
        // test_feature_matcher.cpp: match WITH keypoints; test_pose_estimator.cpp: match WITHOUT keypoints
        std::vector<Match> mk, mn;
        matcher.match(desc0, desc1, mk, kps0, kps1);
        matcher.match(desc0, desc1, mn);
        std::printf("match_kp %zu %016llx match_nokp %zu %016llx\n", mk.size(), fnv(mk.data(), mk.size() * sizeof(Match)), mn.size(),
                    fnv(mn.data(), mn.size() * sizeof(Match)));
        // empty descriptors -> std::invalid_argument("Empty descriptors provided.") like feature_matcher.cpp:99-102
        try {
            std::vector<Match> none;
            matcher.match(DescriptorMatrix(0, 32), desc1, none);
            std::fprintf(stderr, "expected std::invalid_argument\n");
            return -1;
        } catch (const std::invalid_argument& e) {
            std::printf("empty: %s\n", e.what());
        }
        if (argc >= 9) {  // PoseEstimator::estimate up to E (pose_estimator.cpp:18-47)
            const double K4[4] = {std::atof(argv[5]), std::atof(argv[6]), std::atof(argv[7]), std::atof(argv[8])};
            std::vector<float> p1, p2;
            for (const Match& m : mn) {
                p1.push_back(kps0[m.queryIdx].x); p1.push_back(kps0[m.queryIdx].y);
                p2.push_back(kps1[m.trainIdx].x); p2.push_back(kps1[m.trainIdx].y);
            }
            const EssentialResult r = EssentialSolver(K4).solve(p1, p2);
            std::printf("essential valid %d inliers %d of %zu mask %016llx E", r.valid ? 1 : 0, r.inliers, mn.size(),
                        fnv(r.mask.data(), r.mask.size()));
            for (double v : r.E) std::printf(" %.17g", v);
            std::printf("\n");
            const EssentialSolver::Pose pose = EssentialSolver(K4).estimate(p1, p2);
            std::printf("pose valid %d R", pose.valid ? 1 : 0);
            for (double v : pose.R) std::printf(" %.17g", v);
            std::printf(" t %.17g %.17g %.17g\n", pose.t[0], pose.t[1], pose.t[2]);
            // slam::PoseEstimator's own signature: estimate(pairs1, pairs2, matches, R, t) + triangulatePoints
            const CameraK camK{{K4[0], 0, K4[2], 0, K4[1], K4[3], 0, 0, 1}};
            PoseEstimator estimator(camK);
            std::vector<std::pair<MockKeyPoint, MockMat>> pairs0, pairs1;
            std::vector<MockKeyPoint> ck0, ck1;
            for (const Keypoint& k : kps0) { pairs0.push_back({MockKeyPoint{{k.x, k.y}}, MockMat{}}); ck0.push_back(MockKeyPoint{{k.x, k.y}}); }
            for (const Keypoint& k : kps1) { pairs1.push_back({MockKeyPoint{{k.x, k.y}}, MockMat{}}); ck1.push_back(MockKeyPoint{{k.x, k.y}}); }
            std::vector<std::pair<int, int>> idx;
            std::vector<MockDMatch> dm;
            for (const Match& m : mn) { idx.emplace_back(m.queryIdx, m.trainIdx); dm.push_back(MockDMatch{m.queryIdx, m.trainIdx}); }
            MockMat R, t;
            estimator.estimate(pairs0, pairs1, idx, R, t);
            std::printf("estimator touched %d", R.empty() ? 0 : 1);
            if (!R.empty()) {
                const double* r = reinterpret_cast<const double*>(R.data);
                const double* tv = reinterpret_cast<const double*>(t.data);
                std::printf(" R");
                for (int i = 0; i < 9; i++) std::printf(" %.17g", r[i]);
                std::printf(" t %.17g %.17g %.17g", tv[0], tv[1], tv[2]);
            }
            std::printf("\n");
            MockMat R2, t2;  // fewer than 8 matches: R, t stay untouched (pose_estimator.cpp:22-26)
            estimator.estimate(pairs0, pairs1, std::vector<std::pair<int, int>>(idx.begin(), idx.begin() + std::min<size_t>(idx.size(), 7)), R2, t2);
            std::printf("estimator_few touched %d\n", R2.empty() ? 0 : 1);
            if (!R.empty()) {
                const std::vector<MockPoint3d> pts = estimator.triangulatePoints<MockPoint3d>(ck0, ck1, dm, R, t);
                std::printf("triangulated %zu", pts.size());
                for (size_t i = 0; i < pts.size() && i < 3; i++) std::printf(" %.17g %.17g %.17g", pts[i].x, pts[i].y, pts[i].z);
                std::printf("\n");
            }
        }
    } catch (const std::exception& e) {
        std::fprintf(stderr, "Exception: %s\n", e.what());
        return -1;
    }
    return 0;
}
