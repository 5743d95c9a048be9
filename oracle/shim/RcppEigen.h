// TEST INFRASTRUCTURE ONLY (oracle/): minimal stand-ins for the parts of Eigen and Rcpp
// that /root/reference/src/vbnmf_update.cpp uses, so that the reference's OWN source file
// can be compiled unmodified (in place, never copied) into oracle/_ref/ in a container
// that has no R, Rcpp, RcppEigen or Eigen installed.
//
// What is covered (every expression form in src/vbnmf_update.cpp:16-102):
//   Eigen::MatrixXd  : (rows, cols) ctor, rows(), cols(), (i,j), operator* (GEMM), + - unary -,
//                      .array(), .transpose(), ::Constant, .rowwise().sum(), .colwise().sum(),
//                      .row(i) / .col(j) as assignable views
//   Eigen::ArrayXXd  : coefficient-wise * and /, .log()
//   Rcpp::List       : string-keyed read (["lw"] -> MatrixXd or double), List::create(Named(..)=..)
//   Rcpp::NumericVector : operator[]
// Semantics are plain IEEE double, column-major, eager evaluation.  Eigen's lazy expression
// templates evaluate the same scalar formulas; only summation order inside the GEMMs may
// differ from a real Eigen build (Eigen blocks its GEMM), which perturbs results at the
// 1e-16 relative level.
#pragma once
#include <cmath>
#include <cstddef>
#include <map>
#include <string>
#include <utility>
#include <vector>

namespace Eigen {

class ArrayXXd;
class MatrixXd;

class MatrixXd {
  public:
    int nr_, nc_;
    std::vector<double> d_;  // column-major
    MatrixXd() : nr_(0), nc_(0) {}
    MatrixXd(int r, int c) : nr_(r), nc_(c), d_((size_t)r * (size_t)c, 0.0) {}
    MatrixXd(const ArrayXXd &a);  // implicit, as Eigen allows Matrix = Array-expression
    int rows() const { return nr_; }
    int cols() const { return nc_; }
    double &operator()(int i, int j) { return d_[(size_t)j * nr_ + i]; }
    double operator()(int i, int j) const { return d_[(size_t)j * nr_ + i]; }
    double *data() { return d_.data(); }
    const double *data() const { return d_.data(); }
    ArrayXXd array() const;
    MatrixXd transpose() const {
        MatrixXd t(nc_, nr_);
        for (int j = 0; j < nc_; j++)
            for (int i = 0; i < nr_; i++) t(j, i) = (*this)(i, j);
        return t;
    }
    static MatrixXd Constant(int r, int c, double v) {
        MatrixXd m(r, c);
        for (auto &x : m.d_) x = v;
        return m;
    }
    struct RowwiseOp {
        const MatrixXd &m;
        MatrixXd sum() const {  // r x 1 : sum of each row
            MatrixXd s(m.nr_, 1);
            for (int j = 0; j < m.nc_; j++)
                for (int i = 0; i < m.nr_; i++) s(i, 0) += m(i, j);
            return s;
        }
    };
    struct ColwiseOp {
        const MatrixXd &m;
        MatrixXd sum() const {  // 1 x c : sum of each column
            MatrixXd s(1, m.nc_);
            for (int j = 0; j < m.nc_; j++) {
                double a = 0;
                for (int i = 0; i < m.nr_; i++) a += m(i, j);
                s(0, j) = a;
            }
            return s;
        }
    };
    RowwiseOp rowwise() const { return RowwiseOp{*this}; }
    ColwiseOp colwise() const { return ColwiseOp{*this}; }
    struct RowRef {
        MatrixXd &m;
        int i;
        RowRef &operator=(const MatrixXd &v) {  // v is 1 x c
            for (int j = 0; j < m.nc_; j++) m(i, j) = v(0, j);
            return *this;
        }
        MatrixXd operator+(const MatrixXd &v) const {
            MatrixXd o(1, m.nc_);
            for (int j = 0; j < m.nc_; j++) o(0, j) = m(i, j) + v(0, j);
            return o;
        }
    };
    struct ColRef {
        MatrixXd &m;
        int j;
        ColRef &operator=(const MatrixXd &v) {  // v is r x 1
            for (int i = 0; i < m.nr_; i++) m(i, j) = v(i, 0);
            return *this;
        }
        MatrixXd operator+(const MatrixXd &v) const {
            MatrixXd o(m.nr_, 1);
            for (int i = 0; i < m.nr_; i++) o(i, 0) = m(i, j) + v(i, 0);
            return o;
        }
    };
    RowRef row(int i) { return RowRef{*this, i}; }
    ColRef col(int j) { return ColRef{*this, j}; }
};

// dense product, column-major, inner dimension innermost per output column (axpy form)
inline MatrixXd operator*(const MatrixXd &a, const MatrixXd &b) {
    MatrixXd c(a.nr_, b.nc_);
    const int n = a.nr_, kk = a.nc_, m = b.nc_;
#pragma omp parallel for schedule(static)
    for (int j = 0; j < m; j++) {
        double *cj = &c.d_[(size_t)j * n];
        for (int k = 0; k < kk; k++) {
            const double bkj = b(k, j);
            const double *ak = &a.d_[(size_t)k * n];
            for (int i = 0; i < n; i++) cj[i] += ak[i] * bkj;
        }
    }
    return c;
}
inline MatrixXd operator+(const MatrixXd &a, const MatrixXd &b) {
    MatrixXd c(a.nr_, a.nc_);
    for (size_t t = 0; t < c.d_.size(); t++) c.d_[t] = a.d_[t] + b.d_[t];
    return c;
}
inline MatrixXd operator-(const MatrixXd &a, const MatrixXd &b) {
    MatrixXd c(a.nr_, a.nc_);
    for (size_t t = 0; t < c.d_.size(); t++) c.d_[t] = a.d_[t] - b.d_[t];
    return c;
}
inline MatrixXd operator-(const MatrixXd &a) {
    MatrixXd c(a.nr_, a.nc_);
    for (size_t t = 0; t < c.d_.size(); t++) c.d_[t] = -a.d_[t];
    return c;
}

class ArrayXXd {
  public:
    MatrixXd m_;
    ArrayXXd() {}
    explicit ArrayXXd(const MatrixXd &m) : m_(m) {}
    ArrayXXd log() const {
        ArrayXXd o(m_);
        for (auto &x : o.m_.d_) x = std::log(x);
        return o;
    }
};
inline ArrayXXd operator*(const ArrayXXd &a, const ArrayXXd &b) {
    ArrayXXd c(a.m_);
    for (size_t t = 0; t < c.m_.d_.size(); t++) c.m_.d_[t] = a.m_.d_[t] * b.m_.d_[t];
    return c;
}
inline ArrayXXd operator/(const ArrayXXd &a, const ArrayXXd &b) {
    ArrayXXd c(a.m_);
    for (size_t t = 0; t < c.m_.d_.size(); t++) c.m_.d_[t] = a.m_.d_[t] / b.m_.d_[t];
    return c;
}
inline MatrixXd::MatrixXd(const ArrayXXd &a) : nr_(a.m_.nr_), nc_(a.m_.nc_), d_(a.m_.d_) {}
inline ArrayXXd MatrixXd::array() const { return ArrayXXd(*this); }

}  // namespace Eigen

namespace Rcpp {

struct Value {
    bool is_matrix = false;
    double scalar = 0.0;
    Eigen::MatrixXd matrix;
};

struct NamedValue {
    std::string name;
    Value value;
};

struct Named {
    std::string name;
    explicit Named(const char *n) : name(n) {}
    NamedValue operator=(const Eigen::MatrixXd &m) const {
        NamedValue nv;
        nv.name = name;
        nv.value.is_matrix = true;
        nv.value.matrix = m;
        return nv;
    }
    NamedValue operator=(double s) const {
        NamedValue nv;
        nv.name = name;
        nv.value.scalar = s;
        return nv;
    }
};

class List {
  public:
    std::map<std::string, Value> items;
    struct Proxy {
        const Value &v;
        operator Eigen::MatrixXd() const { return v.matrix; }
        operator double() const { return v.scalar; }
    };
    Proxy operator[](const std::string &key) const { return Proxy{items.at(key)}; }
    void set(const std::string &key, const Eigen::MatrixXd &m) {
        Value v;
        v.is_matrix = true;
        v.matrix = m;
        items[key] = v;
    }
    void set(const std::string &key, double s) {
        Value v;
        v.scalar = s;
        items[key] = v;
    }
    template <typename... Args>
    static List create(const Args &...args) {
        List l;
        const NamedValue all[] = {args...};
        for (const auto &nv : all) l.items[nv.name] = nv.value;
        return l;
    }
};

class NumericVector {
  public:
    std::vector<double> v;
    NumericVector() {}
    explicit NumericVector(std::initializer_list<double> il) : v(il) {}
    double operator[](int i) const { return v[i]; }
};

}  // namespace Rcpp
