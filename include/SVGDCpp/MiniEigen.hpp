/*
 * MiniEigen.hpp — the sliver of Eigen the SVGDCpp user-facing API touches, for builds without Eigen.
 *
 * The reference API passes particles and parameters as Eigen::MatrixXd / Eigen::VectorXd
 * (SVGD.hpp:27-52, MultivariateNormal.hpp:39).  When <Eigen/Dense> is available the facade uses the
 * real thing (define nothing); otherwise this header supplies layout-compatible stand-ins in
 * namespace Eigen: column-major dynamic double matrices with the handful of operations the
 * reference's examples and tests use (comma initialiser, Random, Constant, Identity, scalar scaling,
 * +=, transpose, replicate-free accessors, stream output in Eigen's default format).
 * Nothing here runs on the hot path: it only carries data to the C ABI.
 */
#ifndef SVGDCPP_MINI_EIGEN_HPP
#define SVGDCPP_MINI_EIGEN_HPP

#if defined(SVGDCPP_USE_EIGEN) || (__has_include(<Eigen/Dense>) && !defined(SVGDCPP_NO_EIGEN))
#include <Eigen/Dense>
#else

#include <cmath>
#include <cstdlib>
#include <initializer_list>
#include <iomanip>
#include <ostream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace Eigen {

using Index = std::ptrdiff_t;

class MatrixXd {
public:
    MatrixXd() : rows_(0), cols_(0) {}
    MatrixXd(Index r, Index c) : rows_(r), cols_(c), v_(static_cast<size_t>(r * c), 0.0) {}

    Index rows() const { return rows_; }
    Index cols() const { return cols_; }
    Index size() const { return rows_ * cols_; }
    double *data() { return v_.data(); }
    const double *data() const { return v_.data(); }
    void resize(Index r, Index c) { rows_ = r; cols_ = c; v_.assign(static_cast<size_t>(r * c), 0.0); }

    double &operator()(Index r, Index c) { return v_[static_cast<size_t>(c * rows_ + r)]; } // column-major
    double operator()(Index r, Index c) const { return v_[static_cast<size_t>(c * rows_ + r)]; }
    double &operator()(Index i) { return v_[static_cast<size_t>(i)]; }
    double operator()(Index i) const { return v_[static_cast<size_t>(i)]; }

    static MatrixXd Zero(Index r, Index c) { return MatrixXd(r, c); }
    static MatrixXd Constant(Index r, Index c, double value)
    {
        MatrixXd m(r, c);
        for (auto &x : m.v_) x = value;
        return m;
    }
    static MatrixXd Identity(Index r, Index c)
    {
        MatrixXd m(r, c);
        for (Index i = 0; i < (r < c ? r : c); ++i) m(i, i) = 1.0;
        return m;
    }
    // Eigen's unseeded Random: -1 + 2 rand()/RAND_MAX in storage order
    static MatrixXd Random(Index r, Index c)
    {
        MatrixXd m(r, c);
        for (auto &x : m.v_) x = -1.0 + 2.0 * static_cast<double>(std::rand()) / static_cast<double>(RAND_MAX);
        return m;
    }

    MatrixXd transpose() const
    {
        MatrixXd t(cols_, rows_);
        for (Index r = 0; r < rows_; ++r)
            for (Index c = 0; c < cols_; ++c) t(c, r) = (*this)(r, c);
        return t;
    }
    MatrixXd col(Index c) const
    {
        MatrixXd out(rows_, 1);
        for (Index r = 0; r < rows_; ++r) out(r, 0) = (*this)(r, c);
        return out;
    }

    MatrixXd &operator*=(double s) { for (auto &x : v_) x *= s; return *this; }
    MatrixXd &operator+=(const MatrixXd &o)
    {
        if (o.rows_ != rows_ || o.cols_ != cols_) throw std::invalid_argument("MiniEigen: size mismatch in +=");
        for (size_t i = 0; i < v_.size(); ++i) v_[i] += o.v_[i];
        return *this;
    }
    bool operator==(const MatrixXd &o) const { return rows_ == o.rows_ && cols_ == o.cols_ && v_ == o.v_; }
    bool isApprox(const MatrixXd &o, double prec = 1e-12) const
    {
        if (rows_ != o.rows_ || cols_ != o.cols_) return false;
        double diff = 0.0, na = 0.0, nb = 0.0;
        for (size_t i = 0; i < v_.size(); ++i) { diff += (v_[i] - o.v_[i]) * (v_[i] - o.v_[i]); na += v_[i] * v_[i]; nb += o.v_[i] * o.v_[i]; }
        return diff <= prec * prec * (na < nb ? na : nb);
    }

    // comma initialiser: m << a, b, c, d;  (row-major fill, like Eigen)
    class CommaInit {
    public:
        CommaInit(MatrixXd &m, double first) : m_(m), i_(0) { put(first); }
        CommaInit &operator,(double v) { put(v); return *this; }
    private:
        void put(double v)
        {
            if (i_ >= m_.size()) throw std::out_of_range("MiniEigen: too many coefficients");
            m_(i_ / m_.cols(), i_ % m_.cols()) = v;
            ++i_;
        }
        MatrixXd &m_;
        Index i_;
    };
    CommaInit operator<<(double first) { return CommaInit(*this, first); }

protected:
    Index rows_, cols_;
    std::vector<double> v_;
};

inline MatrixXd operator*(double s, const MatrixXd &m) { MatrixXd o = m; o *= s; return o; }
inline MatrixXd operator*(const MatrixXd &m, double s) { MatrixXd o = m; o *= s; return o; }

// Eigen's default IOFormat: 6 significant digits, right-aligned to the widest coefficient
inline std::ostream &operator<<(std::ostream &os, const MatrixXd &m)
{
    std::vector<std::string> cell(static_cast<size_t>(m.size()));
    size_t width = 0;
    for (Index r = 0; r < m.rows(); ++r)
        for (Index c = 0; c < m.cols(); ++c) {
            std::ostringstream ss;
            ss << std::setprecision(6) << m(r, c);
            cell[static_cast<size_t>(r * m.cols() + c)] = ss.str();
            if (ss.str().size() > width) width = ss.str().size();
        }
    for (Index r = 0; r < m.rows(); ++r) {
        if (r) os << "\n";
        for (Index c = 0; c < m.cols(); ++c) {
            if (c) os << " ";
            os << std::setw(static_cast<int>(width)) << cell[static_cast<size_t>(r * m.cols() + c)];
        }
    }
    return os;
}

class VectorXd : public MatrixXd {
public:
    VectorXd() : MatrixXd() {}
    explicit VectorXd(Index n) : MatrixXd(n, 1) {}
    VectorXd(const MatrixXd &m) : MatrixXd(m)
    {
        if (m.cols() != 1 && m.rows() == 1) { MatrixXd t = m.transpose(); static_cast<MatrixXd &>(*this) = t; }
    }
    VectorXd(std::initializer_list<double> init) : MatrixXd(static_cast<Index>(init.size()), 1)
    {
        Index i = 0;
        for (double v : init) (*this)(i++) = v;
    }
    static VectorXd Constant(Index n, double value) { return VectorXd(MatrixXd::Constant(n, 1, value)); }
    static VectorXd Zero(Index n) { return VectorXd(n); }
    void resize(Index n) { MatrixXd::resize(n, 1); }
};

class Vector2d : public VectorXd {
public:
    Vector2d() : VectorXd(2) {}
    Vector2d(double a, double b) : VectorXd(2) { (*this)(0) = a; (*this)(1) = b; }
};

class Matrix2d : public MatrixXd {
public:
    Matrix2d() : MatrixXd(2, 2) {}
};

} // namespace Eigen
#endif // Eigen available?
#endif
