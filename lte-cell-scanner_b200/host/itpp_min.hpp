// itpp_min.hpp - the slice of IT++ that crosses the searcher.h boundary.
//
// The reference passes itpp::cvec / vec / mat / imat / cmat by reference through
// include/searcher.h:22-124.  IT++ is not installed in this image, so the drop-in is built and
// tested against this header-only stand-in: same type names, same element access, Mat stored
// COLUMN-MAJOR like IT++ (verified from the .it payloads).  When the real IT++ is
// present, compile searcher_dropin.cpp with -DLCS_USE_REAL_ITPP and this header is not used.
#pragma once
#include <complex>
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <vector>

namespace itpp {

template <class T>
class Vec {
 public:
  Vec() {}
  explicit Vec(int n) : d_(n) {}
  int length() const { return (int)d_.size(); }
  int size() const { return (int)d_.size(); }
  void set_size(int n, bool copy = false) { (void)copy; d_.resize(n); }
  void set_length(int n, bool copy = false) { set_size(n, copy); }
  T& operator()(int i) { return d_[i]; }
  const T& operator()(int i) const { return d_[i]; }
  T& operator[](int i) { return d_[i]; }
  const T& operator[](int i) const { return d_[i]; }
  T* _data() { return d_.data(); }
  const T* _data() const { return d_.data(); }
  Vec& operator=(const T& v) { for (auto& x : d_) x = v; return *this; }
 private:
  std::vector<T> d_;
};

template <class T>
class Mat {  // column-major, element (r,c) at [c*rows+r]
 public:
  Mat() : r_(0), c_(0) {}
  Mat(int r, int c) : r_(r), c_(c), d_((size_t)r * c) {}
  int rows() const { return r_; }
  int cols() const { return c_; }
  void set_size(int r, int c, bool copy = false) { (void)copy; r_ = r; c_ = c; d_.resize((size_t)r * c); }
  T& operator()(int r, int c) { return d_[(size_t)c * r_ + r]; }
  const T& operator()(int r, int c) const { return d_[(size_t)c * r_ + r]; }
  T* _data() { return d_.data(); }
  const T* _data() const { return d_.data(); }
 private:
  int r_, c_;
  std::vector<T> d_;
};

typedef Vec<double> vec;
typedef Vec<std::complex<double> > cvec;
typedef Vec<int> ivec;
typedef Mat<double> mat;
typedef Mat<std::complex<double> > cmat;
typedef Mat<int> imat;

template <class T> inline int length(const Vec<T>& v) { return v.length(); }

}  // namespace itpp
