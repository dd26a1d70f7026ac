// oracle/shim/Eigen/mini_eigen.hpp -- TEST INFRASTRUCTURE ONLY (see ../README.md).
//
// The subset of Eigen 3.4.0's dense API that the reference's frontend sources touch
// (include/slam/common/common.hpp, include/slam/frontend/*.hpp, src/frontend/feature_detector.cpp,
// src/frontend/feature_matcher.cpp), with EAGER evaluation.  Each coefficient-wise operator applies the
// same IEEE operation as Eigen's functor for it (scalar_sum_op -> a + b, scalar_quotient_op -> a / b,
// scalar_sqrt_op -> std::sqrt, scalar_pow_op with a promoted exponent -> std::pow(x, (T)e), ...), and C++
// operator precedence builds the same expression tree Eigen would evaluate lazily, so the values agree
// bit for bit on a target without FMA contraction (baseline x86-64, the reference's build).
#pragma once
// the standard headers Eigen/Core itself pulls in (the reference relies on some of them transitively, e.g. <array>)
#include <algorithm>
#include <array>
#include <cassert>
#include <climits>
#include <cmath>
#include <complex>
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <iosfwd>
#include <limits>
#include <sstream>
#include <string>
#include <type_traits>
#include <vector>

namespace Eigen {

using Index = std::ptrdiff_t;
constexpr int Dynamic = -1;
enum StorageOptions { ColMajor = 0, RowMajor = 1 };

// ---- MatrixBase: what calculateHammingDistance (common.hpp:40-50) needs from any dense expression ----
template <class Derived>
struct MatrixBase {
    const Derived& derived() const { return static_cast<const Derived&>(*this); }
    Derived& derived() { return static_cast<Derived&>(*this); }
    Index size() const { return derived().rows() * derived().cols(); }
    auto operator()(Index k) const { return derived().coeff(k); }
};

template <class T, int R, int C, int O = ((R == 1 && C != 1) ? RowMajor : ColMajor)>
class Matrix;
template <class T, int R, int C>
class Array;

template <class M>
class RowBlock : public MatrixBase<RowBlock<M>> {
public:
    using Scalar = typename M::Scalar;
    RowBlock(M& m, Index r) : m_(m), r_(r) {}
    Index rows() const { return 1; }
    Index cols() const { return m_.cols(); }
    Scalar coeff(Index k) const { return m_(r_, k); }
    Scalar operator()(Index, Index c) const { return m_(r_, c); }
    template <class Other>
    RowBlock& operator=(const MatrixBase<Other>& o) {
        for (Index k = 0; k < cols(); k++) m_(r_, k) = o.derived().coeff(k);
        return *this;
    }
    RowBlock& operator=(const RowBlock& o) {
        for (Index k = 0; k < cols(); k++) m_(r_, k) = o.coeff(k);
        return *this;
    }
    Scalar minCoeff(Index* idx) const {  // first minimum, like Eigen's visitor
        Index best = 0;
        for (Index k = 1; k < cols(); k++)
            if (coeff(k) < coeff(best)) best = k;
        *idx = best;
        return coeff(best);
    }

private:
    M& m_;
    Index r_;
};

template <class M>
class ColBlock : public MatrixBase<ColBlock<M>> {
public:
    using Scalar = typename M::Scalar;
    ColBlock(M& m, Index c) : m_(m), c_(c) {}
    Index rows() const { return m_.rows(); }
    Index cols() const { return 1; }
    Scalar coeff(Index k) const { return m_(k, c_); }
    template <class Other>
    ColBlock& operator=(const MatrixBase<Other>& o) {
        for (Index k = 0; k < rows(); k++) m_(k, c_) = o.derived().coeff(k);
        return *this;
    }
    ColBlock& operator=(const ColBlock& o) {
        for (Index k = 0; k < rows(); k++) m_(k, c_) = o.coeff(k);
        return *this;
    }

private:
    M& m_;
    Index c_;
};

template <class M>
struct RowwiseOp {
    const M& m;
    auto squaredNorm() const {
        using T = typename M::Scalar;
        Matrix<T, Dynamic, 1> out(m.rows(), 1);
        for (Index r = 0; r < m.rows(); r++) {
            T s = T(0);
            for (Index c = 0; c < m.cols(); c++) s += m(r, c) * m(r, c);
            out(r) = s;
        }
        return out;
    }
};

// ---- Matrix -----------------------------------------------------------------------------------
template <class T, int R, int C, int O>
class Matrix : public MatrixBase<Matrix<T, R, C, O>> {
public:
    using Scalar = T;
    using Index = Eigen::Index;
    static constexpr bool kFixed = (R != Dynamic && C != Dynamic);

    Matrix() : r_(R == Dynamic ? 0 : R), c_(C == Dynamic ? 0 : C), d_((size_t)(r_ * c_)) {}
    // (rows, cols) for dynamic matrices; (x, y) for fixed two-vectors, as in Eigen
    template <class A, class B, class = std::enable_if_t<std::is_arithmetic<A>::value && std::is_arithmetic<B>::value>>
    Matrix(A a, B b) {
        if constexpr (kFixed && R * C == 2) {
            r_ = R; c_ = C; d_.resize(2);
            d_[0] = (T)a; d_[1] = (T)b;
        } else {
            r_ = (Index)a; c_ = (Index)b; d_.assign((size_t)(r_ * c_), T());
        }
    }
    template <class T2, int R2, int C2, int O2>
    Matrix(const Matrix<T2, R2, C2, O2>& o) : r_(o.rows()), c_(o.cols()), d_((size_t)(o.rows() * o.cols())) {
        for (Index r = 0; r < r_; r++)
            for (Index c = 0; c < c_; c++) (*this)(r, c) = (T)o(r, c);
    }
    template <class M>
    Matrix(const RowBlock<M>& b) : r_(1), c_(b.cols()), d_((size_t)b.cols()) {
        for (Index k = 0; k < c_; k++) d_[(size_t)k] = b.coeff(k);
    }
    Matrix(const Array<T, Dynamic, Dynamic>& a);

    static Matrix Zero(Index r, Index c) { Matrix m; m.resize(r, c); std::fill(m.d_.begin(), m.d_.end(), T(0)); return m; }
    // linspaced_op_impl<Scalar, /*IsInteger*/ false> of Eigen 3.4.0 (NullaryFunctors.h)
    static Matrix LinSpaced(Index n, T low, T high) {
        Matrix m;
        if (R == 1) m.resize(1, n); else m.resize(n, 1);
        const Index size1 = n == 1 ? 1 : n - 1;
        const T step = n == 1 ? T(1) : (high - low) / T(n - 1);
        const bool flip = std::abs(high) < std::abs(low);
        for (Index i = 0; i < n; i++)
            m.d_[(size_t)i] = flip ? (i == 0 ? low : T(high - T(size1 - i) * step)) : (i == size1 ? high : T(low + T(i) * step));
        return m;
    }
    template <class F>
    static Matrix NullaryExpr(Index r, Index c, F f) {
        Matrix m; m.resize(r, c);
        for (Index j = 0; j < c; j++)
            for (Index i = 0; i < r; i++) m(i, j) = f(i, j);
        return m;
    }

    void resize(Index r, Index c) { r_ = r; c_ = c; d_.assign((size_t)(r * c), T()); }
    Index rows() const { return r_; }
    Index cols() const { return c_; }
    Index size() const { return r_ * c_; }
    T* data() { return d_.data(); }
    const T* data() const { return d_.data(); }

    T& operator()(Index r, Index c) { return d_[(size_t)(O == RowMajor ? r * c_ + c : c * r_ + r)]; }
    const T& operator()(Index r, Index c) const { return d_[(size_t)(O == RowMajor ? r * c_ + c : c * r_ + r)]; }
    T& operator()(Index k) { return d_[(size_t)k]; }             // vectors
    const T& operator()(Index k) const { return d_[(size_t)k]; }
    T& operator[](Index k) { return d_[(size_t)k]; }             // vectors, like Eigen's DenseCoeffsBase
    const T& operator[](Index k) const { return d_[(size_t)k]; }
    T coeff(Index k) const { return d_[(size_t)k]; }
    T x() const { return d_[0]; }
    T y() const { return d_[1]; }

    RowBlock<Matrix> row(Index r) { return RowBlock<Matrix>(*this, r); }
    RowBlock<const Matrix> row(Index r) const { return RowBlock<const Matrix>(*this, r); }
    ColBlock<Matrix> col(Index c) { return ColBlock<Matrix>(*this, c); }
    ColBlock<const Matrix> col(Index c) const { return ColBlock<const Matrix>(*this, c); }
    RowwiseOp<Matrix> rowwise() const { return RowwiseOp<Matrix>{*this}; }

    template <class S, class = std::enable_if_t<std::is_arithmetic<S>::value>>
    Matrix& operator/=(S s) {  // scalar_quotient_op: a true division per coefficient
        for (auto& v : d_) v = v / (T)s;
        return *this;
    }
    template <class S, class = std::enable_if_t<std::is_arithmetic<S>::value>>
    Matrix<T, Dynamic, Dynamic, O> operator/(S s) const {
        Matrix<T, Dynamic, Dynamic, O> out(r_, c_);
        for (Index r = 0; r < r_; r++)
            for (Index c = 0; c < c_; c++) out(r, c) = (*this)(r, c) / (T)s;
        return out;
    }
    Matrix<T, Dynamic, Dynamic> transpose() const {
        Matrix<T, Dynamic, Dynamic> out(c_, r_);
        for (Index r = 0; r < r_; r++)
            for (Index c = 0; c < c_; c++) out(c, r) = (*this)(r, c);
        return out;
    }
    Matrix<T, Dynamic, Dynamic> replicate(Index rf, Index cf) const {
        Matrix<T, Dynamic, Dynamic> out(r_ * rf, c_ * cf);
        for (Index r = 0; r < r_ * rf; r++)
            for (Index c = 0; c < c_ * cf; c++) out(r, c) = (*this)(r % r_, c % c_);
        return out;
    }
    template <class T2>
    Matrix<T2, Dynamic, Dynamic, O> cast() const {
        Matrix<T2, Dynamic, Dynamic, O> out(r_, c_);
        for (Index r = 0; r < r_; r++)
            for (Index c = 0; c < c_; c++) out(r, c) = (T2)(*this)(r, c);
        return out;
    }
    Array<T, Dynamic, Dynamic> array() const;

private:
    Index r_, c_;
    std::vector<T> d_;
};

template <class T, int R1, int C1, int O1, int R2, int C2, int O2>
Matrix<T, Dynamic, Dynamic> operator+(const Matrix<T, R1, C1, O1>& a, const Matrix<T, R2, C2, O2>& b) {
    Matrix<T, Dynamic, Dynamic> out(a.rows(), a.cols());
    for (Index r = 0; r < a.rows(); r++)
        for (Index c = 0; c < a.cols(); c++) out(r, c) = a(r, c) + b(r, c);
    return out;
}
template <class T, int R1, int C1, int O1, int R2, int C2, int O2>
Matrix<T, Dynamic, Dynamic> operator-(const Matrix<T, R1, C1, O1>& a, const Matrix<T, R2, C2, O2>& b) {
    Matrix<T, Dynamic, Dynamic> out(a.rows(), a.cols());
    for (Index r = 0; r < a.rows(); r++)
        for (Index c = 0; c < a.cols(); c++) out(r, c) = a(r, c) - b(r, c);
    return out;
}
template <class S, class T, int R, int C, int O, class = std::enable_if_t<std::is_arithmetic<S>::value>>
Matrix<T, Dynamic, Dynamic> operator*(S s, const Matrix<T, R, C, O>& a) {
    Matrix<T, Dynamic, Dynamic> out(a.rows(), a.cols());
    for (Index r = 0; r < a.rows(); r++)
        for (Index c = 0; c < a.cols(); c++) out(r, c) = (T)s * a(r, c);
    return out;
}
template <class T, int R1, int C1, int O1, int R2, int C2, int O2>
Matrix<T, Dynamic, Dynamic> operator*(const Matrix<T, R1, C1, O1>& a, const Matrix<T, R2, C2, O2>& b) {
    Matrix<T, Dynamic, Dynamic> out(a.rows(), b.cols());
    for (Index r = 0; r < a.rows(); r++)
        for (Index c = 0; c < b.cols(); c++) {
            T s = T(0);
            for (Index k = 0; k < a.cols(); k++) s += a(r, k) * b(k, c);
            out(r, c) = s;
        }
    return out;
}

// ---- Map: a view over caller memory (only Map<EigenGrayMatrix>::cast<double>() is used, common.hpp:137-138) ----
template <class M>
class Map;
template <class T, int R, int C, int O>
class Map<Matrix<T, R, C, O>> {
public:
    Map(T* p, Index r, Index c) : p_(p), r_(r), c_(c) {}
    Index rows() const { return r_; }
    Index cols() const { return c_; }
    const T& operator()(Index r, Index c) const { return p_[O == RowMajor ? r * c_ + c : c * r_ + r]; }
    template <class T2>
    Matrix<T2, Dynamic, Dynamic, O> cast() const {
        Matrix<T2, Dynamic, Dynamic, O> out(r_, c_);
        for (Index r = 0; r < r_; r++)
            for (Index c = 0; c < c_; c++) out(r, c) = (T2)(*this)(r, c);
        return out;
    }

private:
    T* p_;
    Index r_, c_;
};

// ---- Array: coefficient-wise arithmetic (Camera::undistortImage, common.hpp:143-157) ----------------
template <class T, int R, int C>
class Array {
public:
    using Scalar = T;
    Array() : r_(0), c_(0) {}
    Array(Index r, Index c) : r_(r), c_(c), d_((size_t)(r * c)) {}
    Index rows() const { return r_; }
    Index cols() const { return c_; }
    T& operator()(Index r, Index c) { return d_[(size_t)(c * r_ + r)]; }
    const T& operator()(Index r, Index c) const { return d_[(size_t)(c * r_ + r)]; }
    template <class F>
    Array map(F f) const {
        Array out(r_, c_);
        for (size_t i = 0; i < d_.size(); i++) out.d_[i] = f(d_[i]);
        return out;
    }
    template <class F>
    Array zip(const Array& o, F f) const {
        Array out(r_, c_);
        for (size_t i = 0; i < d_.size(); i++) out.d_[i] = f(d_[i], o.d_[i]);
        return out;
    }
    Array square() const { return map([](T v) { return v * v; }); }                      // scalar_square_op
    Array sqrt() const { return map([](T v) { return std::sqrt(v); }); }                 // scalar_sqrt_op
    template <class E>
    Array pow(E e) const { return map([e](T v) { return std::pow(v, (T)e); }); }         // scalar_pow_op<T, T>, exponent promoted

private:
    Index r_, c_;
    std::vector<T> d_;
};

template <class T, int R, int C, int O>
Array<T, Dynamic, Dynamic> Matrix<T, R, C, O>::array() const {
    Array<T, Dynamic, Dynamic> out(r_, c_);
    for (Index r = 0; r < r_; r++)
        for (Index c = 0; c < c_; c++) out(r, c) = (*this)(r, c);
    return out;
}
template <class T, int R, int C, int O>
Matrix<T, R, C, O>::Matrix(const Array<T, Dynamic, Dynamic>& a) : r_(a.rows()), c_(a.cols()), d_((size_t)(a.rows() * a.cols())) {
    for (Index r = 0; r < r_; r++)
        for (Index c = 0; c < c_; c++) (*this)(r, c) = a(r, c);
}

#define MINI_EIGEN_ARRAY_OP(OP)                                                                              \
    template <class T, int R, int C>                                                                         \
    Array<T, R, C> operator OP(const Array<T, R, C>& a, const Array<T, R, C>& b) {                          \
        return a.zip(b, [](T x, T y) { return x OP y; });                                                    \
    }                                                                                                        \
    template <class T, int R, int C, class S, class = std::enable_if_t<std::is_arithmetic<S>::value>>        \
    Array<T, R, C> operator OP(const Array<T, R, C>& a, S s) {                                               \
        const T t = (T)s;                                                                                    \
        return a.map([t](T x) { return x OP t; });                                                           \
    }                                                                                                        \
    template <class T, int R, int C, class S, class = std::enable_if_t<std::is_arithmetic<S>::value>>        \
    Array<T, R, C> operator OP(S s, const Array<T, R, C>& a) {                                               \
        const T t = (T)s;                                                                                    \
        return a.map([t](T x) { return t OP x; });                                                           \
    }
MINI_EIGEN_ARRAY_OP(+)
MINI_EIGEN_ARRAY_OP(-)
MINI_EIGEN_ARRAY_OP(*)
MINI_EIGEN_ARRAY_OP(/)
#undef MINI_EIGEN_ARRAY_OP

using MatrixXd = Matrix<double, Dynamic, Dynamic>;
using MatrixXf = Matrix<float, Dynamic, Dynamic>;
using VectorXd = Matrix<double, Dynamic, 1>;
using VectorXf = Matrix<float, Dynamic, 1>;
using RowVectorXd = Matrix<double, 1, Dynamic>;
using Matrix3d = Matrix<double, 3, 3>;
using Vector2i = Matrix<int, 2, 1>;
using Vector3d = Matrix<double, 3, 1>;
using ArrayXXd = Array<double, Dynamic, Dynamic>;

}  // namespace Eigen
