// TEST INFRASTRUCTURE ONLY (oracle/): a minimal stand-in for <Rcpp.h>.
//
// Rcpp is not installed in this image, so the reference's
// src/cocons_full.cpp cannot be compiled against the real headers.  This
// header supplies just enough of the Rcpp container surface that
// /root/reference/src/cocons_full.cpp, cocons_taper.cpp + cocons_types.h compile UNMODIFIED,
// from where they lie, into oracle/_ref/ (see oracle/Makefile).  Semantics
// mirrored: NumericVector/NumericMatrix are reference-counted handles
// (copy = alias, clone() = deep copy), matrices are column-major and
// zero-initialised, `m(i,_)` yields a row view that converts to a fresh
// vector, and `scalar * vec`, `vec + vec`, `row - row` are element-wise.
// No arithmetic of the reference is re-implemented here.
#ifndef COCONS_ORACLE_RCPP_SHIM_H
#define COCONS_ORACLE_RCPP_SHIM_H

#include <cmath>
#include <cstddef>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#ifndef M_PI
#define M_PI 3.141592653589793238462643383280
#endif

namespace Rcpp {

struct Placeholder {};
static const Placeholder _ = Placeholder();

class NumericVector;

// strided read-only view of one matrix row
struct ConstRow {
  const double* base;
  long stride;
  long n;
  double operator()(long k) const { return base[k * stride]; }
  double operator[](long k) const { return base[k * stride]; }
  long size() const { return n; }
};

// element-wise difference of two rows (what `locs(i,_) - locs(j,_)` yields)
struct RowDiff {
  ConstRow a, b;
  long size() const { return a.n; }
  double operator[](long k) const { return a[k] - b[k]; }
};
inline RowDiff operator-(const ConstRow& a, const ConstRow& b) { return RowDiff{a, b}; }

class NumericVector {
 public:
  NumericVector() : buf_(std::make_shared<std::vector<double>>()) {}
  explicit NumericVector(int n) : buf_(std::make_shared<std::vector<double>>((size_t)n, 0.0)) {}
  explicit NumericVector(long n) : buf_(std::make_shared<std::vector<double>>((size_t)n, 0.0)) {}
  NumericVector(const double* p, long n) : buf_(std::make_shared<std::vector<double>>(p, p + n)) {}
  NumericVector(const ConstRow& r) : buf_(std::make_shared<std::vector<double>>((size_t)r.n)) {
    for (long k = 0; k < r.n; ++k) (*buf_)[k] = r[k];
  }
  NumericVector(const RowDiff& r) : buf_(std::make_shared<std::vector<double>>((size_t)r.size())) {
    for (long k = 0; k < r.size(); ++k) (*buf_)[k] = r[k];
  }
  // sugar assignment writes through into the existing storage when sizes agree
  NumericVector& operator=(const RowDiff& r) {
    if ((long)buf_->size() != r.size()) buf_ = std::make_shared<std::vector<double>>((size_t)r.size());
    for (long k = 0; k < r.size(); ++k) (*buf_)[k] = r[k];
    return *this;
  }
  long length() const { return (long)buf_->size(); }
  long size() const { return (long)buf_->size(); }
  double& operator()(long k) { return (*buf_)[k]; }
  double operator()(long k) const { return (*buf_)[k]; }
  double& operator[](long k) { return (*buf_)[k]; }
  double operator[](long k) const { return (*buf_)[k]; }
  NumericVector deep_copy() const {
    NumericVector out;
    out.buf_ = std::make_shared<std::vector<double>>(*buf_);
    return out;
  }

 private:
  std::shared_ptr<std::vector<double>> buf_;
};

inline NumericVector clone(const NumericVector& v) { return v.deep_copy(); }

inline NumericVector operator*(double s, const NumericVector& v) {
  NumericVector out(v.size());
  for (long k = 0; k < v.size(); ++k) out[k] = s * v[k];
  return out;
}
inline NumericVector operator*(int s, const NumericVector& v) { return (double)s * v; }
inline NumericVector operator+(const NumericVector& a, const NumericVector& b) {
  NumericVector out(a.size());
  for (long k = 0; k < a.size(); ++k) out[k] = a[k] + b[k];
  return out;
}

// `colindices - 1` (src/cocons_taper.cpp:76-77, 213-214): element-wise, fresh storage
inline NumericVector operator-(const NumericVector& a, int s) {
  NumericVector out(a.size());
  for (long k = 0; k < a.size(); ++k) out[k] = a[k] - s;
  return out;
}

class NumericMatrix {
 public:
  NumericMatrix() : nr_(0), nc_(0), buf_(std::make_shared<std::vector<double>>()) {}
  explicit NumericMatrix(int n) : nr_(n), nc_(n), buf_(std::make_shared<std::vector<double>>((size_t)n * n, 0.0)) {}
  NumericMatrix(int nr, int nc) : nr_(nr), nc_(nc), buf_(std::make_shared<std::vector<double>>((size_t)nr * nc, 0.0)) {}
  NumericMatrix(long nr, long nc, const double* colmajor)
      : nr_(nr), nc_(nc), buf_(std::make_shared<std::vector<double>>(colmajor, colmajor + nr * nc)) {}
  int nrow() const { return (int)nr_; }
  int ncol() const { return (int)nc_; }
  double& operator()(long i, long j) { return (*buf_)[(size_t)j * nr_ + i]; }
  double operator()(long i, long j) const { return (*buf_)[(size_t)j * nr_ + i]; }
  ConstRow operator()(long i, Placeholder) const { return ConstRow{buf_->data() + i, nr_, nc_}; }
  const double* data() const { return buf_->data(); }

 private:
  long nr_, nc_;
  std::shared_ptr<std::vector<double>> buf_;
};

// named list of numeric vectors; lookup by name, missing name is an error
class List {
 public:
  void set(const std::string& name, const NumericVector& v) { items_[name] = v; }
  NumericVector operator[](const char* name) const {
    auto it = items_.find(name);
    if (it == items_.end()) throw std::runtime_error(std::string("Index out of bounds: [index='") + name + "'].");
    return it->second;
  }

 private:
  std::map<std::string, NumericVector> items_;
};

}  // namespace Rcpp

#endif
