// oracle/shim/opencv2/mini_cv.hpp -- TEST INFRASTRUCTURE ONLY (see ../README.md).
//
// The slice of OpenCV's core API that include/slam/common/common.hpp and the two frontend sources of the reference
// need in order to compile: cv::FileStorage (reader for the %YAML:1.0 files under test/data), a small cv::Mat,
// cv::Size / Point2f / KeyPoint, cv2eigen.  cv::SVD::compute throws (slam::triangulate is inline in common.hpp and has
// to compile; the oracle never runs it).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <fstream>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include <Eigen/Eigen>

#define CV_8U 0
#define CV_8UC1 0
#define CV_32F 5
#define CV_64F 6

namespace cv {

using uchar = unsigned char;

struct Size { int width = 0, height = 0; Size() = default; Size(int w, int h) : width(w), height(h) {} };
struct Point2f { float x = 0, y = 0; Point2f() = default; Point2f(float X, float Y) : x(X), y(Y) {} };
struct Point3d { double x = 0, y = 0, z = 0; Point3d() = default; Point3d(double X, double Y, double Z) : x(X), y(Y), z(Z) {} };
struct KeyPoint { Point2f pt; float size = 0, angle = -1, response = 0; int octave = 0, class_id = -1; };
struct DMatch { int queryIdx = -1, trainIdx = -1, imgIdx = -1; float distance = 0; };

class Mat;
struct MatExpr { std::shared_ptr<Mat> value; };

class Mat {
public:
    int rows = 0, cols = 0;
    uchar* data = nullptr;
    Mat() = default;
    Mat(int r, int c, int type) { create(r, c, type); }
    Mat(int r, int c, int type, void* external) : rows(r), cols(c), data(static_cast<uchar*>(external)), type_(type), step_((size_t)c * esz(type)) {}
    Mat(const MatExpr& e);
    Mat(const Mat&) = default;
    Mat& operator=(const Mat&) = default;
    Mat(Mat&& o) noexcept { *this = std::move(o); }
    Mat& operator=(Mat&& o) noexcept {
        rows = o.rows; cols = o.cols; data = o.data; type_ = o.type_; step_ = o.step_; buf_ = std::move(o.buf_);
        o.rows = o.cols = 0; o.data = nullptr;
        return *this;
    }
    Mat& operator=(const MatExpr& e);  // assigns INTO the existing elements when the shapes agree (row/col headers)
    void create(int r, int c, int type) {
        rows = r; cols = c; type_ = type; step_ = (size_t)c * esz(type);
        buf_ = std::make_shared<std::vector<uchar>>((size_t)r * step_);
        data = buf_->data();
    }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    int type() const { return type_; }
    double get(int r, int c) const {
        const uchar* p = data + (size_t)r * step_ + (size_t)c * esz(type_);
        if (type_ == CV_64F) { double v; std::memcpy(&v, p, 8); return v; }
        if (type_ == CV_32F) { float v; std::memcpy(&v, p, 4); return v; }
        return *p;
    }
    void set(int r, int c, double v) {
        uchar* p = data + (size_t)r * step_ + (size_t)c * esz(type_);
        if (type_ == CV_64F) std::memcpy(p, &v, 8);
        else if (type_ == CV_32F) { float f = (float)v; std::memcpy(p, &f, 4); }
        else *p = (uchar)v;
    }
    Mat row(int r) const { Mat h = *this; h.rows = 1; h.data = data + (size_t)r * step_; return h; }
    Mat col(int c) const { Mat h = *this; h.cols = 1; h.data = data + (size_t)c * esz(type_); return h; }
    Mat t() const {
        Mat o(cols, rows, type_);
        for (int r = 0; r < rows; r++) for (int c = 0; c < cols; c++) o.set(c, r, get(r, c));
        return o;
    }
    void copyTo(Mat dst) const {  // same-type, same-shape destination headers only
        for (int r = 0; r < rows && r < dst.rows; r++) for (int c = 0; c < cols && c < dst.cols; c++) dst.set(r, c, get(r, c));
    }
    static size_t esz(int type) { return type == CV_64F ? 8 : type == CV_32F ? 4 : 1; }

private:
    int type_ = CV_8U;
    size_t step_ = 0;
    std::shared_ptr<std::vector<uchar>> buf_;
};

inline Mat::Mat(const MatExpr& e) { *this = *e.value; }
inline Mat& Mat::operator=(const MatExpr& e) {
    const Mat& s = *e.value;
    if (!empty() && rows == s.rows && cols == s.cols) { for (int r = 0; r < rows; r++) for (int c = 0; c < cols; c++) set(r, c, s.get(r, c)); }
    else *this = s;
    return *this;
}
inline MatExpr make_expr(const Mat& a, const Mat& b, double sa, double sb) {
    auto o = std::make_shared<Mat>(a.rows, a.cols, CV_64F);
    for (int r = 0; r < a.rows; r++) for (int c = 0; c < a.cols; c++) o->set(r, c, sa * a.get(r, c) + (b.empty() ? 0.0 : sb * b.get(r, c)));
    return MatExpr{o};
}
inline MatExpr operator*(double s, const Mat& a) { return make_expr(a, Mat(), s, 0.0); }
inline MatExpr operator-(const MatExpr& a, const Mat& b) { return make_expr(*a.value, b, 1.0, -1.0); }
inline MatExpr operator-(const Mat& a, const Mat& b) { return make_expr(a, b, 1.0, -1.0); }

struct SVD {
    enum Flags { MODIFY_A = 1, NO_UV = 2, FULL_UV = 4 };
    static void compute(const Mat&, Mat&, Mat&, Mat&, int = 0) { throw std::logic_error("oracle shim: cv::SVD is not provided"); }
};

// ---- cv::FileStorage: %YAML:1.0 reader (scalars, quoted strings, flow sequences, !!opencv-matrix) ----
struct FileNode {
    enum Kind { NONE, SCALAR, SEQ, MATRIX } kind = NONE;
    std::string text;
    std::vector<double> seq;
    int rows = 0, cols = 0;
    std::string dt;
    bool empty() const { return kind == NONE; }
    bool is_real() const { return text.find_first_of(".eE") != std::string::npos; }
};

inline int cvRound(double v) { return (int)std::nearbyint(v); }
inline void operator>>(const FileNode& n, int& v) { v = n.kind == FileNode::SCALAR ? (n.is_real() ? cvRound(std::stod(n.text)) : std::stoi(n.text)) : 0; }
inline void operator>>(const FileNode& n, float& v) { v = n.kind == FileNode::SCALAR ? (float)std::stod(n.text) : 0.f; }
inline void operator>>(const FileNode& n, double& v) { v = n.kind == FileNode::SCALAR ? std::stod(n.text) : 0.0; }
inline void operator>>(const FileNode& n, std::string& v) { v = n.kind == FileNode::SCALAR ? n.text : std::string(); }
inline void operator>>(const FileNode& n, Size& v) { v = n.seq.size() >= 2 ? Size((int)n.seq[0], (int)n.seq[1]) : Size(); }
inline void operator>>(const FileNode& n, Mat& m) {
    if (n.kind != FileNode::MATRIX) { m = Mat(); return; }
    const int type = n.dt == "d" ? CV_64F : n.dt == "f" ? CV_32F : CV_8U;
    m.create(n.rows, n.cols, type);
    for (int r = 0; r < n.rows; r++) for (int c = 0; c < n.cols; c++) m.set(r, c, n.seq[(size_t)r * n.cols + c]);
}

class FileStorage {
public:
    enum Mode { READ = 0 };
    FileStorage(const std::string& path, int) {
        std::ifstream f(path);
        if (!f) return;
        open_ = true;
        std::string line, key;
        FileNode* cur = nullptr;
        bool in_data = false;
        auto trim = [](std::string s) {
            const size_t a = s.find_first_not_of(" \t\r\n"), b = s.find_last_not_of(" \t\r\n");
            return a == std::string::npos ? std::string() : s.substr(a, b - a + 1);
        };
        auto numbers = [](const std::string& s, std::vector<double>& out) {
            std::string t = s;
            for (char& ch : t) if (ch == '[' || ch == ']' || ch == ',') ch = ' ';
            std::istringstream is(t);
            double v;
            while (is >> v) out.push_back(v);
        };
        while (std::getline(f, line)) {
            const size_t hash = line.find('#');
            if (hash != std::string::npos && line.find('"') == std::string::npos) line = line.substr(0, hash);
            const std::string t = trim(line);
            if (t.empty() || t[0] == '%' || t == "---") continue;
            if (in_data && cur) {  // continuation of a multi-line data: [ ... ]
                numbers(t, cur->seq);
                if (t.find(']') != std::string::npos) in_data = false;
                continue;
            }
            const size_t colon = t.find(':');
            if (colon == std::string::npos) continue;
            const std::string k = trim(t.substr(0, colon)), v = trim(t.substr(colon + 1));
            const bool nested = line[0] == ' ' || line[0] == '\t';
            if (nested && cur && cur->kind == FileNode::MATRIX) {
                if (k == "rows") cur->rows = std::stoi(v);
                else if (k == "cols") cur->cols = std::stoi(v);
                else if (k == "dt") cur->dt = v;
                else if (k == "data") { numbers(v, cur->seq); in_data = v.find(']') == std::string::npos; }
                continue;
            }
            cur = &nodes_[k];
            if (v.rfind("!!opencv-matrix", 0) == 0) cur->kind = FileNode::MATRIX;
            else if (!v.empty() && v[0] == '[') { cur->kind = FileNode::SEQ; numbers(v, cur->seq); }
            else {
                cur->kind = FileNode::SCALAR;
                cur->text = (v.size() >= 2 && v.front() == '"' && v.back() == '"') ? v.substr(1, v.size() - 2) : v;
            }
        }
    }
    bool isOpened() const { return open_; }
    FileNode operator[](const std::string& k) const { auto it = nodes_.find(k); return it == nodes_.end() ? FileNode() : it->second; }
    FileNode operator[](const char* k) const { return (*this)[std::string(k)]; }
    void release() { open_ = false; }

private:
    bool open_ = false;
    std::map<std::string, FileNode> nodes_;
};

template <class T, int R, int C, int O>
void cv2eigen(const Mat& src, Eigen::Matrix<T, R, C, O>& dst) {
    if (R == Eigen::Dynamic || C == Eigen::Dynamic) dst.resize(src.rows, src.cols);
    for (int r = 0; r < src.rows; r++) for (int c = 0; c < src.cols; c++) dst(r, c) = (T)src.get(r, c);
}

}  // namespace cv
