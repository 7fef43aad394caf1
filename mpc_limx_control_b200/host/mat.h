// mat.h -- minimal column-major dense matrix / vector types for the host facade.
// The reference passes Eigen::MatrixXd / VectorXd by const& (include/QPSolver.h:13-37); Eigen is not
// available in this image, so the facade uses these stand-ins with the same storage order
// (column-major, `.data()` contiguous) and the handful of accessors the call sites need.
#pragma once
#include <cstddef>
#include <initializer_list>
#include <vector>

namespace mpcb200 {
namespace host {

class MatrixXd {
public:
    MatrixXd() : r_(0), c_(0) {}
    MatrixXd(int rows, int cols) : r_(rows), c_(cols), v_((size_t)rows * cols, 0.0) {}
    static MatrixXd Zero(int rows, int cols) { return MatrixXd(rows, cols); }
    static MatrixXd Identity(int rows, int cols) {
        MatrixXd m(rows, cols);
        for (int i = 0; i < rows && i < cols; ++i) m(i, i) = 1.0;
        return m;
    }
    // row-major initialiser for readability at call sites:  MatrixXd::FromRows(2, 2, {a, b, c, d})
    static MatrixXd FromRows(int rows, int cols, std::initializer_list<double> vals) {
        MatrixXd m(rows, cols);
        int k = 0;
        for (double x : vals) { m(k / cols, k % cols) = x; ++k; }
        return m;
    }
    void resize(int rows, int cols) { r_ = rows; c_ = cols; v_.assign((size_t)rows * cols, 0.0); }
    int rows() const { return r_; }
    int cols() const { return c_; }
    double& operator()(int i, int j) { return v_[(size_t)i + (size_t)r_ * j]; }
    double operator()(int i, int j) const { return v_[(size_t)i + (size_t)r_ * j]; }
    double* data() { return v_.data(); }
    const double* data() const { return v_.data(); }
    MatrixXd operator*(double s) const { MatrixXd m(*this); for (double& x : m.v_) x *= s; return m; }

private:
    int r_, c_;
    std::vector<double> v_;
};
inline MatrixXd operator*(double s, const MatrixXd& m) { return m * s; }

class VectorXd {
public:
    VectorXd() {}
    explicit VectorXd(int n) : v_((size_t)n, 0.0) {}
    VectorXd(std::initializer_list<double> vals) : v_(vals) {}
    static VectorXd Zero(int n) { return VectorXd(n); }
    static VectorXd Constant(int n, double x) { VectorXd v(n); for (double& e : v.v_) e = x; return v; }
    void resize(int n) { v_.assign((size_t)n, 0.0); }
    int size() const { return (int)v_.size(); }
    double& operator()(int i) { return v_[(size_t)i]; }
    double operator()(int i) const { return v_[(size_t)i]; }
    double& operator[](int i) { return v_[(size_t)i]; }
    double operator[](int i) const { return v_[(size_t)i]; }
    double* data() { return v_.data(); }
    const double* data() const { return v_.data(); }
    VectorXd operator-() const { VectorXd v(*this); for (double& e : v.v_) e = -e; return v; }
    // diag(v) as a dense matrix (Eigen's .asDiagonal() at the reference call sites)
    MatrixXd asDiagonal() const {
        MatrixXd m(size(), size());
        for (int i = 0; i < size(); ++i) m(i, i) = v_[(size_t)i];
        return m;
    }

private:
    std::vector<double> v_;
};

}  // namespace host
}  // namespace mpcb200
