// mini_eigen.h — stand-in for the parts of Eigen 3 that the reference's mapper layer (src/sfm/sequential_mapper.{h,cc},
// src/mapper.cc and the headers they include) and the drop-in shim touch.  Eigen is not in the image; this header exists
// so that the UNMODIFIED reference callers can be compiled and linked against the shim (tests/mapper_harness).
// One run-time sized dense matrix class behind the Eigen names; small and slow on purpose: test infrastructure.
#ifndef MAPPER_HARNESS_MINI_EIGEN_H_
#define MAPPER_HARNESS_MINI_EIGEN_H_
#include <cassert>
#include <cmath>
#include <cstddef>
#include <iostream>
#include <vector>
#include <math.h>

namespace Eigen {
const int Dynamic = -1;
enum { ComputeFullU = 1, ComputeFullV = 2, ComputeThinU = 4, ComputeThinV = 8 };

enum { ColMajor = 0, RowMajor = 1, AutoAlign = 0, DontAlign = 2 };
template <typename S, int R, int C> class MatrixImpl;
// (the storage-order / size-limit parameters of Eigen::Matrix are accepted and ignored: storage is always row-major here)
template <typename S, int R, int C, int Options = 0, int MaxR = R, int MaxC = C> using Matrix = MatrixImpl<S, R, C>;
template <typename S> class BlockRef;

template <typename S> struct CommaInit {
  Matrix<S, Dynamic, Dynamic>* m; std::size_t k;
  CommaInit& operator,(S v);
};

template <typename S, int R, int C>
class MatrixImpl {
 public:
  std::vector<S> d; std::size_t nr, nc;      // row-major
  MatrixImpl() : d((R > 0 ? R : 0) * (C > 0 ? C : 0), S(0)), nr(R > 0 ? R : 0), nc(C > 0 ? C : 0) {}
  explicit MatrixImpl(std::size_t n) : nr(R > 0 ? R : n), nc(C > 0 ? C : (R > 0 ? n : 1)) { if (R < 0 && C < 0) { nr = n; nc = 1; } d.assign(nr * nc, S(0)); }
  // two arguments: the coefficients of a fixed 2-vector, otherwise (rows, cols)
  template <typename A, typename B> MatrixImpl(A a, B b) {
    if ((R == 2 && C == 1) || (R == 1 && C == 2)) { nr = R; nc = C; d.resize(2); d[0] = S(a); d[1] = S(b); }
    else { nr = (std::size_t)a; nc = (std::size_t)b; d.assign(nr * nc, S(0)); }
  }
  MatrixImpl(S a, S b, S c) : d(3), nr(R == 1 ? 1 : 3), nc(R == 1 ? 3 : 1) { d[0] = a; d[1] = b; d[2] = c; }
  MatrixImpl(S a, S b, S c, S e) : d(4), nr(R == 1 ? 1 : 4), nc(R == 1 ? 4 : 1) { d[0] = a; d[1] = b; d[2] = c; d[3] = e; }
  template <int R2, int C2> MatrixImpl(const Matrix<S, R2, C2>& o) : d(o.d), nr(o.nr), nc(o.nc) { fix(); }
  MatrixImpl(const BlockRef<S>& b);
  template <int R2, int C2> MatrixImpl& operator=(const Matrix<S, R2, C2>& o) { d = o.d; nr = o.nr; nc = o.nc; fix(); return *this; }
  MatrixImpl& operator=(const BlockRef<S>& b);
  void fix() { if ((R == 1 && nr != 1 && nc == 1) || (C == 1 && nc != 1 && nr == 1)) std::swap(nr, nc); }   // row <-> column vector views
  static MatrixImpl Zero() { return MatrixImpl(); }
  static MatrixImpl Zero(std::size_t r, std::size_t c = 1) { MatrixImpl m; m.nr = r; m.nc = c; m.d.assign(r * c, S(0)); return m; }
  static MatrixImpl Ones() { MatrixImpl m; for (auto& v : m.d) v = S(1); return m; }
  static MatrixImpl Identity() { MatrixImpl m; for (std::size_t i = 0; i < m.nr && i < m.nc; ++i) m(i, i) = S(1); return m; }
  static MatrixImpl Identity(std::size_t r, std::size_t c) { MatrixImpl m = Zero(r, c); for (std::size_t i = 0; i < r && i < c; ++i) m(i, i) = S(1); return m; }
  static MatrixImpl UnitX() { MatrixImpl m; m.d[0] = S(1); return m; }
  static MatrixImpl UnitY() { MatrixImpl m; m.d[1] = S(1); return m; }
  static MatrixImpl UnitZ() { MatrixImpl m; m.d[2] = S(1); return m; }
  void resize(std::size_t r, std::size_t c) { nr = r; nc = c; d.assign(r * c, S(0)); }
  void setZero() { for (auto& v : d) v = S(0); }
  void fill(S x) { for (auto& v : d) v = x; }
  void setConstant(S x) { fill(x); }
  void setOnes() { fill(S(1)); }
  const MatrixImpl& eval() const { return *this; }
  template <int R2, int C2> MatrixImpl cwiseProduct(const Matrix<S, R2, C2>& o) const { MatrixImpl m(*this); for (std::size_t i = 0; i < m.d.size(); ++i) m.d[i] *= o.d[i]; return m; }
  template <int R2, int C2> MatrixImpl cwiseQuotient(const Matrix<S, R2, C2>& o) const { MatrixImpl m(*this); for (std::size_t i = 0; i < m.d.size(); ++i) m.d[i] /= o.d[i]; return m; }
  MatrixImpl cwiseProduct(const BlockRef<S>& o) const { return cwiseProduct(Matrix<S, Dynamic, Dynamic>(o)); }
  MatrixImpl cwiseSqrt() const { MatrixImpl m(*this); for (auto& v : m.d) v = std::sqrt(v); return m; }
  MatrixImpl cwiseInverse() const { MatrixImpl m(*this); for (auto& v : m.d) v = S(1) / v; return m; }
  void setIdentity() { setZero(); for (std::size_t i = 0; i < nr && i < nc; ++i) (*this)(i, i) = S(1); }
  std::size_t rows() const { return nr; }
  std::size_t cols() const { return nc; }
  std::size_t size() const { return d.size(); }
  S* data() { return d.data(); }
  const S* data() const { return d.data(); }
  S& operator()(std::size_t r, std::size_t c) { return d[r * nc + c]; }
  const S& operator()(std::size_t r, std::size_t c) const { return d[r * nc + c]; }
  S& operator()(std::size_t i) { return d[i]; }
  const S& operator()(std::size_t i) const { return d[i]; }
  S& operator[](std::size_t i) { return d[i]; }
  const S& operator[](std::size_t i) const { return d[i]; }
  S& x() { return d[0]; } S& y() { return d[1]; } S& z() { return d[2]; } S& w() { return d[3]; }
  const S& x() const { return d[0]; } const S& y() const { return d[1]; } const S& z() const { return d[2]; } const S& w() const { return d[3]; }
  // blocks
  BlockRef<S> block(std::size_t r0, std::size_t c0, std::size_t r, std::size_t c);
  Matrix<S, Dynamic, Dynamic> block(std::size_t r0, std::size_t c0, std::size_t r, std::size_t c) const;
  template <int BR, int BC> BlockRef<S> block(std::size_t r0, std::size_t c0);
  template <int BR, int BC> Matrix<S, BR, BC> block(std::size_t r0, std::size_t c0) const { return Matrix<S, BR, BC>(block(r0, c0, BR, BC)); }
  BlockRef<S> row(std::size_t r);
  BlockRef<S> col(std::size_t c);
  Matrix<S, 1, Dynamic> row(std::size_t r) const { return Matrix<S, 1, Dynamic>(block(r, 0, 1, nc)); }
  Matrix<S, Dynamic, 1> col(std::size_t c) const { return Matrix<S, Dynamic, 1>(block(0, c, nr, 1)); }
  BlockRef<S> head(std::size_t n);
  BlockRef<S> tail(std::size_t n);
  template <int N> BlockRef<S> head();
  template <int N> BlockRef<S> tail();
  Matrix<S, Dynamic, 1> head(std::size_t n) const { return Matrix<S, Dynamic, 1>(vec_block(0, n)); }
  Matrix<S, Dynamic, 1> tail(std::size_t n) const { return Matrix<S, Dynamic, 1>(vec_block(d.size() - n, n)); }
  template <int N> Matrix<S, N, 1> head() const { return Matrix<S, N, 1>(vec_block(0, N)); }
  template <int N> Matrix<S, N, 1> tail() const { return Matrix<S, N, 1>(vec_block(d.size() - N, N)); }
  Matrix<S, Dynamic, Dynamic> vec_block(std::size_t i0, std::size_t n) const { Matrix<S, Dynamic, Dynamic> m; m.nr = n; m.nc = 1; m.d.assign(d.begin() + i0, d.begin() + i0 + n); return m; }
  BlockRef<S> topLeftCorner(std::size_t r, std::size_t c);
  BlockRef<S> topRightCorner(std::size_t r, std::size_t c);
  BlockRef<S> leftCols(std::size_t c);
  BlockRef<S> rightCols(std::size_t c);
  BlockRef<S> topRows(std::size_t r);
  BlockRef<S> bottomRows(std::size_t r);
  // algebra
  Matrix<S, C, R> transpose() const { Matrix<S, C, R> t; t.nr = nc; t.nc = nr; t.d.resize(d.size()); for (std::size_t i = 0; i < nr; ++i) for (std::size_t j = 0; j < nc; ++j) t.d[j * nr + i] = d[i * nc + j]; return t; }
  S squaredNorm() const { S s = 0; for (auto v : d) s += v * v; return s; }
  S norm() const { return std::sqrt(squaredNorm()); }
  S sum() const { S s = 0; for (auto v : d) s += v; return s; }
  S mean() const { return sum() / S(d.size()); }
  S trace() const { S s = 0; for (std::size_t i = 0; i < nr && i < nc; ++i) s += (*this)(i, i); return s; }
  S maxCoeff() const { S m = d[0]; for (auto v : d) if (v > m) m = v; return m; }
  S minCoeff() const { S m = d[0]; for (auto v : d) if (v < m) m = v; return m; }
  MatrixImpl normalized() const { MatrixImpl m(*this); const S n = norm(); if (n > 0) for (auto& v : m.d) v /= n; return m; }
  void normalize() { const S n = norm(); if (n > 0) for (auto& v : d) v /= n; }
  MatrixImpl cwiseAbs() const { MatrixImpl m(*this); for (auto& v : m.d) v = std::fabs(v); return m; }
  Matrix<S, Dynamic, Dynamic> array() const { return Matrix<S, Dynamic, Dynamic>(*this); }
  Matrix<S, Dynamic, Dynamic> matrix() const { return Matrix<S, Dynamic, Dynamic>(*this); }
  Matrix<S, Dynamic, 1> rowwise_mean() const { Matrix<S, Dynamic, 1> m; m.nr = nr; m.nc = 1; m.d.assign(nr, S(0)); for (std::size_t i = 0; i < nr; ++i) { for (std::size_t j = 0; j < nc; ++j) m.d[i] += (*this)(i, j); m.d[i] /= S(nc); } return m; }
  Matrix<S, 1, Dynamic> colwise_mean() const { Matrix<S, 1, Dynamic> m; m.nr = 1; m.nc = nc; m.d.assign(nc, S(0)); for (std::size_t j = 0; j < nc; ++j) { for (std::size_t i = 0; i < nr; ++i) m.d[j] += (*this)(i, j); m.d[j] /= S(nr); } return m; }
  template <int R2, int C2> S dot(const Matrix<S, R2, C2>& o) const { S s = 0; for (std::size_t i = 0; i < d.size(); ++i) s += d[i] * o.d[i]; return s; }
  template <int R2, int C2> MatrixImpl cross(const Matrix<S, R2, C2>& o) const { MatrixImpl m(*this); m.d[0] = d[1] * o.d[2] - d[2] * o.d[1]; m.d[1] = d[2] * o.d[0] - d[0] * o.d[2]; m.d[2] = d[0] * o.d[1] - d[1] * o.d[0]; return m; }
  S determinant() const;
  MatrixImpl inverse() const;
  Matrix<S, Dynamic, 1> homogeneous() const { Matrix<S, Dynamic, 1> m; m.nr = d.size() + 1; m.nc = 1; m.d = d; m.d.push_back(S(1)); return m; }
  Matrix<S, Dynamic, 1> hnormalized() const { Matrix<S, Dynamic, 1> m; m.nr = d.size() - 1; m.nc = 1; m.d.assign(d.begin(), d.end() - 1); for (auto& v : m.d) v /= d.back(); return m; }
  template <typename T> Matrix<T, R, C> cast() const { Matrix<T, R, C> m; m.nr = nr; m.nc = nc; m.d.assign(d.begin(), d.end()); return m; }
  bool operator==(const MatrixImpl& o) const { return nr * nc == o.nr * o.nc && d == o.d; }
  bool operator!=(const MatrixImpl& o) const { return !(*this == o); }
  bool isApprox(const MatrixImpl& o, S eps = S(1e-12)) const { return (*this - o).norm() <= eps * std::min(norm(), o.norm()); }
  MatrixImpl operator-() const { MatrixImpl m(*this); for (auto& v : m.d) v = -v; return m; }
  MatrixImpl& operator+=(const MatrixImpl& o) { for (std::size_t i = 0; i < d.size(); ++i) d[i] += o.d[i]; return *this; }
  MatrixImpl& operator-=(const MatrixImpl& o) { for (std::size_t i = 0; i < d.size(); ++i) d[i] -= o.d[i]; return *this; }
  MatrixImpl& operator*=(S s) { for (auto& v : d) v *= s; return *this; }
  MatrixImpl& operator/=(S s) { for (auto& v : d) v /= s; return *this; }
  CommaInit<S> operator<<(S v) { d[0] = v; CommaInit<S> c; c.m = reinterpret_cast<Matrix<S, Dynamic, Dynamic>*>(this); c.k = 1; return c; }
};

template <typename S> CommaInit<S>& CommaInit<S>::operator,(S v) { m->d[k++] = v; return *this; }

// writable view of a rectangular part of a matrix
template <typename S>
class BlockRef {
 public:
  std::vector<S>* d; std::size_t ld, r0, c0, nr, nc;
  S& at(std::size_t i, std::size_t j) const { return (*d)[(r0 + i) * ld + c0 + j]; }
  template <int R, int C> BlockRef& operator=(const Matrix<S, R, C>& m) { std::size_t k = 0; for (std::size_t i = 0; i < nr; ++i) for (std::size_t j = 0; j < nc; ++j) at(i, j) = m.d[k++]; return *this; }
  BlockRef& operator=(const BlockRef& o) { Matrix<S, Dynamic, Dynamic> t(o); return *this = t; }
  S& operator()(std::size_t i, std::size_t j) { return at(i, j); }
  S& operator()(std::size_t i) { return nr == 1 ? at(0, i) : at(i, 0); }
  Matrix<S, Dynamic, Dynamic> eval() const { return Matrix<S, Dynamic, Dynamic>(*this); }
  Matrix<S, Dynamic, Dynamic> transpose() const { return eval().transpose(); }
  S norm() const { return eval().norm(); }
  Matrix<S, Dynamic, Dynamic> inverse() const { return eval().inverse(); }
  S determinant() const { return eval().determinant(); }
  S squaredNorm() const { return eval().squaredNorm(); }
  S sum() const { return eval().sum(); }
  S mean() const { return eval().mean(); }
  Matrix<S, Dynamic, Dynamic> normalized() const { return eval().normalized(); }
  void setZero() { fill(S(0)); }
  void fill(S x) { for (std::size_t i = 0; i < nr; ++i) for (std::size_t j = 0; j < nc; ++j) at(i, j) = x; }
  void setConstant(S x) { fill(x); }
  template <int R, int C> Matrix<S, Dynamic, Dynamic> cwiseProduct(const Matrix<S, R, C>& o) const { return eval().cwiseProduct(o); }
  Matrix<S, Dynamic, Dynamic> cwiseProduct(const BlockRef& o) const { return eval().cwiseProduct(o.eval()); }
  template <int R, int C> S dot(const Matrix<S, R, C>& o) const { return eval().dot(o); }
  Matrix<S, Dynamic, Dynamic> operator-() const { return -eval(); }
  template <int R, int C> BlockRef& operator+=(const Matrix<S, R, C>& m) { std::size_t k = 0; for (std::size_t i = 0; i < nr; ++i) for (std::size_t j = 0; j < nc; ++j) at(i, j) += m.d[k++]; return *this; }
  template <int R, int C> BlockRef& operator-=(const Matrix<S, R, C>& m) { std::size_t k = 0; for (std::size_t i = 0; i < nr; ++i) for (std::size_t j = 0; j < nc; ++j) at(i, j) -= m.d[k++]; return *this; }
  BlockRef& operator*=(S s) { for (std::size_t i = 0; i < nr; ++i) for (std::size_t j = 0; j < nc; ++j) at(i, j) *= s; return *this; }
  BlockRef& operator/=(S s) { for (std::size_t i = 0; i < nr; ++i) for (std::size_t j = 0; j < nc; ++j) at(i, j) /= s; return *this; }
};

template <typename S, int R, int C> MatrixImpl<S, R, C>::MatrixImpl(const BlockRef<S>& b) : d(b.nr * b.nc), nr(b.nr), nc(b.nc) { std::size_t k = 0; for (std::size_t i = 0; i < nr; ++i) for (std::size_t j = 0; j < nc; ++j) d[k++] = b.at(i, j); fix(); }
template <typename S, int R, int C> Matrix<S, R, C>& MatrixImpl<S, R, C>::operator=(const BlockRef<S>& b) { return *this = Matrix<S, Dynamic, Dynamic>(b); }
template <typename S, int R, int C> BlockRef<S> MatrixImpl<S, R, C>::block(std::size_t r0, std::size_t c0, std::size_t r, std::size_t c) { BlockRef<S> b; b.d = &d; b.ld = nc; b.r0 = r0; b.c0 = c0; b.nr = r; b.nc = c; return b; }
template <typename S, int R, int C> Matrix<S, Dynamic, Dynamic> MatrixImpl<S, R, C>::block(std::size_t r0, std::size_t c0, std::size_t r, std::size_t c) const { Matrix<S, Dynamic, Dynamic> m; m.nr = r; m.nc = c; m.d.resize(r * c); for (std::size_t i = 0; i < r; ++i) for (std::size_t j = 0; j < c; ++j) m.d[i * c + j] = (*this)(r0 + i, c0 + j); return m; }
template <typename S, int R, int C> template <int BR, int BC> BlockRef<S> MatrixImpl<S, R, C>::block(std::size_t r0, std::size_t c0) { return block(r0, c0, (std::size_t)BR, (std::size_t)BC); }
template <typename S, int R, int C> BlockRef<S> MatrixImpl<S, R, C>::row(std::size_t r) { return block(r, 0, 1, nc); }
template <typename S, int R, int C> BlockRef<S> MatrixImpl<S, R, C>::col(std::size_t c) { return block(0, c, nr, 1); }
template <typename S, int R, int C> BlockRef<S> MatrixImpl<S, R, C>::head(std::size_t n) { return nc == 1 ? block(0, 0, n, 1) : block(0, 0, 1, n); }
template <typename S, int R, int C> BlockRef<S> MatrixImpl<S, R, C>::tail(std::size_t n) { return nc == 1 ? block(nr - n, 0, n, 1) : block(0, nc - n, 1, n); }
template <typename S, int R, int C> template <int N> BlockRef<S> MatrixImpl<S, R, C>::head() { return head((std::size_t)N); }
template <typename S, int R, int C> template <int N> BlockRef<S> MatrixImpl<S, R, C>::tail() { return tail((std::size_t)N); }
template <typename S, int R, int C> BlockRef<S> MatrixImpl<S, R, C>::topLeftCorner(std::size_t r, std::size_t c) { return block(0, 0, r, c); }
template <typename S, int R, int C> BlockRef<S> MatrixImpl<S, R, C>::topRightCorner(std::size_t r, std::size_t c) { return block(0, nc - c, r, c); }
template <typename S, int R, int C> BlockRef<S> MatrixImpl<S, R, C>::leftCols(std::size_t c) { return block(0, 0, nr, c); }
template <typename S, int R, int C> BlockRef<S> MatrixImpl<S, R, C>::rightCols(std::size_t c) { return block(0, nc - c, nr, c); }
template <typename S, int R, int C> BlockRef<S> MatrixImpl<S, R, C>::topRows(std::size_t r) { return block(0, 0, r, nc); }
template <typename S, int R, int C> BlockRef<S> MatrixImpl<S, R, C>::bottomRows(std::size_t r) { return block(nr - r, 0, r, nc); }

// products / sums: the result type keeps the outer fixed dimensions so that e.g. Matrix3d * Vector3d is a Vector3d
template <typename S, int R, int K, int K2, int C> Matrix<S, R, C> operator*(const Matrix<S, R, K>& a, const Matrix<S, K2, C>& b) {
  Matrix<S, R, C> m; m.nr = a.nr; m.nc = b.nc; m.d.assign(a.nr * b.nc, S(0));
  for (std::size_t i = 0; i < a.nr; ++i) for (std::size_t k = 0; k < a.nc; ++k) { const S v = a(i, k); for (std::size_t j = 0; j < b.nc; ++j) m.d[i * b.nc + j] += v * b(k, j); }
  return m;
}
template <typename S, int R, int K> Matrix<S, R, Dynamic> operator*(const Matrix<S, R, K>& a, const BlockRef<S>& b) { return a * Matrix<S, Dynamic, Dynamic>(b); }
template <typename S, int K, int C> Matrix<S, Dynamic, C> operator*(const BlockRef<S>& a, const Matrix<S, K, C>& b) { return Matrix<S, Dynamic, Dynamic>(a) * b; }
template <typename S, int R, int C> Matrix<S, R, C> operator+(const Matrix<S, R, C>& a, const Matrix<S, R, C>& b) { Matrix<S, R, C> m(a); m += b; return m; }
template <typename S, int R, int C> Matrix<S, R, C> operator-(const Matrix<S, R, C>& a, const Matrix<S, R, C>& b) { Matrix<S, R, C> m(a); m -= b; return m; }
template <typename S, int R, int C, int R2, int C2> Matrix<S, R, C> operator+(const Matrix<S, R, C>& a, const Matrix<S, R2, C2>& b) { Matrix<S, R, C> m(a); for (std::size_t i = 0; i < m.d.size(); ++i) m.d[i] += b.d[i]; return m; }
template <typename S, int R, int C, int R2, int C2> Matrix<S, R, C> operator-(const Matrix<S, R, C>& a, const Matrix<S, R2, C2>& b) { Matrix<S, R, C> m(a); for (std::size_t i = 0; i < m.d.size(); ++i) m.d[i] -= b.d[i]; return m; }
template <typename S, int R, int C> Matrix<S, R, C> operator+(const Matrix<S, R, C>& a, const BlockRef<S>& b) { return a + Matrix<S, R, C>(b); }
template <typename S, int R, int C> Matrix<S, R, C> operator-(const Matrix<S, R, C>& a, const BlockRef<S>& b) { return a - Matrix<S, R, C>(b); }
template <typename S, int R, int C> Matrix<S, R, C> operator+(const BlockRef<S>& a, const Matrix<S, R, C>& b) { return Matrix<S, R, C>(a) + b; }
template <typename S, int R, int C> Matrix<S, R, C> operator-(const BlockRef<S>& a, const Matrix<S, R, C>& b) { return Matrix<S, R, C>(a) - b; }
template <typename S> Matrix<S, Dynamic, Dynamic> operator-(const BlockRef<S>& a, const BlockRef<S>& b) { return a.eval() - b.eval(); }
template <typename S> Matrix<S, Dynamic, Dynamic> operator+(const BlockRef<S>& a, const BlockRef<S>& b) { return a.eval() + b.eval(); }
template <typename S, int R, int C> Matrix<S, R, C> operator*(const Matrix<S, R, C>& a, double s) { Matrix<S, R, C> m(a); m *= S(s); return m; }
template <typename S, int R, int C> Matrix<S, R, C> operator*(double s, const Matrix<S, R, C>& a) { Matrix<S, R, C> m(a); m *= S(s); return m; }
template <typename S, int R, int C> Matrix<S, R, C> operator/(const Matrix<S, R, C>& a, double s) { Matrix<S, R, C> m(a); m /= S(s); return m; }
template <typename S> Matrix<S, Dynamic, Dynamic> operator*(const BlockRef<S>& a, double s) { return a.eval() * s; }
template <typename S> Matrix<S, Dynamic, Dynamic> operator*(double s, const BlockRef<S>& a) { return a.eval() * s; }
template <typename S> Matrix<S, Dynamic, Dynamic> operator/(const BlockRef<S>& a, double s) { return a.eval() / s; }
template <typename S, int R, int C> std::ostream& operator<<(std::ostream& os, const Matrix<S, R, C>& m) { for (std::size_t i = 0; i < m.nr; ++i) { for (std::size_t j = 0; j < m.nc; ++j) os << (j ? " " : "") << m(i, j); if (i + 1 < m.nr) os << "\n"; } return os; }
template <typename S> std::ostream& operator<<(std::ostream& os, const BlockRef<S>& b) { return os << b.eval(); }

// Gauss-Jordan with partial pivoting (inverse, determinant)
template <typename S, int R, int C> Matrix<S, R, C> MatrixImpl<S, R, C>::inverse() const {
  const std::size_t n = nr; Matrix<S, Dynamic, Dynamic> a(*this); Matrix<S, R, C> inv; inv.nr = inv.nc = n; inv.d.assign(n * n, S(0));
  for (std::size_t i = 0; i < n; ++i) inv(i, i) = S(1);
  for (std::size_t c = 0; c < n; ++c) {
    std::size_t p = c; for (std::size_t r = c + 1; r < n; ++r) if (std::fabs(a(r, c)) > std::fabs(a(p, c))) p = r;
    if (p != c) for (std::size_t j = 0; j < n; ++j) { std::swap(a(p, j), a(c, j)); std::swap(inv(p, j), inv(c, j)); }
    const S piv = a(c, c);
    for (std::size_t j = 0; j < n; ++j) { a(c, j) /= piv; inv(c, j) /= piv; }
    for (std::size_t r = 0; r < n; ++r) if (r != c) { const S f = a(r, c); if (f != S(0)) for (std::size_t j = 0; j < n; ++j) { a(r, j) -= f * a(c, j); inv(r, j) -= f * inv(c, j); } }
  }
  return inv;
}
template <typename S, int R, int C> S MatrixImpl<S, R, C>::determinant() const {
  const std::size_t n = nr; Matrix<S, Dynamic, Dynamic> a(*this); S det = S(1);
  for (std::size_t c = 0; c < n; ++c) {
    std::size_t p = c; for (std::size_t r = c + 1; r < n; ++r) if (std::fabs(a(r, c)) > std::fabs(a(p, c))) p = r;
    if (a(p, c) == S(0)) return S(0);
    if (p != c) { for (std::size_t j = 0; j < n; ++j) std::swap(a(p, j), a(c, j)); det = -det; }
    det *= a(c, c);
    for (std::size_t r = c + 1; r < n; ++r) { const S f = a(r, c) / a(c, c); for (std::size_t j = c; j < n; ++j) a(r, j) -= f * a(c, j); }
  }
  return det;
}

typedef Matrix<double, 2, 1> Vector2d; typedef Matrix<double, 3, 1> Vector3d; typedef Matrix<double, 4, 1> Vector4d;
typedef Matrix<float, 2, 1> Vector2f; typedef Matrix<float, 3, 1> Vector3f; typedef Matrix<int, 2, 1> Vector2i; typedef Matrix<int, 3, 1> Vector3i;
typedef Matrix<double, Dynamic, 1> VectorXd; typedef Matrix<double, 1, Dynamic> RowVectorXd; typedef Matrix<double, 1, 3> RowVector3d; typedef Matrix<double, 1, 2> RowVector2d;
typedef Matrix<double, 2, 2> Matrix2d; typedef Matrix<double, 3, 3> Matrix3d; typedef Matrix<double, 4, 4> Matrix4d;
typedef Matrix<double, Dynamic, Dynamic> MatrixXd; typedef Matrix<float, Dynamic, Dynamic> MatrixXf; typedef Matrix<int, Dynamic, Dynamic> MatrixXi;
typedef Matrix<float, 3, 3> Matrix3f; typedef Matrix<int, Dynamic, 1> VectorXi;

// one-sided Jacobi SVD (full V, thin U), singular values descending
template <typename M>
class JacobiSVD {
 public:
  MatrixXd U, V; VectorXd s;
  JacobiSVD() {}
  template <typename A> JacobiSVD(const A& a_in, unsigned = 0) { compute(a_in); }
  template <typename A> JacobiSVD& compute(const A& a_in, unsigned = 0) {
    MatrixXd a(a_in); const std::size_t m = a.rows(), n = a.cols();
    const bool wide = n > m; if (wide) a = MatrixXd(a.transpose());
    const std::size_t rr = a.rows(), cc = a.cols();
    MatrixXd v = MatrixXd::Identity(cc, cc);
    for (int sweep = 0; sweep < 60; ++sweep) {
      double off = 0;
      for (std::size_t p = 0; p + 1 < cc; ++p) for (std::size_t q = p + 1; q < cc; ++q) {
        double al = 0, be = 0, ga = 0;
        for (std::size_t i = 0; i < rr; ++i) { al += a(i, p) * a(i, p); be += a(i, q) * a(i, q); ga += a(i, p) * a(i, q); }
        if (std::fabs(ga) <= 1e-300 || std::fabs(ga) <= 1e-15 * std::sqrt(al * be)) continue;
        off = std::max(off, std::fabs(ga) / std::sqrt(al * be));
        const double zeta = (be - al) / (2 * ga), t = (zeta >= 0 ? 1.0 : -1.0) / (std::fabs(zeta) + std::sqrt(1 + zeta * zeta)), c = 1 / std::sqrt(1 + t * t), sn = c * t;
        for (std::size_t i = 0; i < rr; ++i) { const double x = a(i, p), y = a(i, q); a(i, p) = c * x - sn * y; a(i, q) = sn * x + c * y; }
        for (std::size_t i = 0; i < cc; ++i) { const double x = v(i, p), y = v(i, q); v(i, p) = c * x - sn * y; v(i, q) = sn * x + c * y; }
      }
      if (off < 1e-15) break;
    }
    std::vector<double> sv(cc); std::vector<std::size_t> ord(cc);
    for (std::size_t j = 0; j < cc; ++j) { double t = 0; for (std::size_t i = 0; i < rr; ++i) t += a(i, j) * a(i, j); sv[j] = std::sqrt(t); ord[j] = j; }
    for (std::size_t i = 0; i < cc; ++i) for (std::size_t j = i + 1; j < cc; ++j) if (sv[ord[j]] > sv[ord[i]]) std::swap(ord[i], ord[j]);
    MatrixXd u(rr, cc), vv(cc, cc); s = VectorXd::Zero(cc);
    for (std::size_t j = 0; j < cc; ++j) { const std::size_t o = ord[j]; s(j) = sv[o]; for (std::size_t i = 0; i < rr; ++i) u(i, j) = sv[o] > 0 ? a(i, o) / sv[o] : 0.0; for (std::size_t i = 0; i < cc; ++i) vv(i, j) = v(i, o); }
    if (wide) { U = vv; V = u; } else { U = u; V = vv; }
    return *this;
  }
  const MatrixXd& matrixU() const { return U; }
  const MatrixXd& matrixV() const { return V; }
  const VectorXd& singularValues() const { return s; }
};

template <typename S> class AngleAxis {
 public:
  S a; Matrix<S, 3, 1> ax;
  AngleAxis() : a(0), ax(S(1), S(0), S(0)) {}
  template <int R, int C> AngleAxis(S angle, const Matrix<S, R, C>& axis) : a(angle), ax(axis) {}
  template <int R, int C> AngleAxis(const Matrix<S, R, C>& m) { from_matrix(Matrix<S, 3, 3>(m)); }
  AngleAxis(const BlockRef<S>& b) { from_matrix(Matrix<S, 3, 3>(b)); }
  void from_matrix(const Matrix<S, 3, 3>& m) {
    const S c = std::min(S(1), std::max(S(-1), (m.trace() - S(1)) / S(2))); a = std::acos(c);
    Matrix<S, 3, 1> v(m(2, 1) - m(1, 2), m(0, 2) - m(2, 0), m(1, 0) - m(0, 1));
    const S n = v.norm();
    if (n > S(1e-12)) ax = v / n;
    else if (a < S(1)) ax = Matrix<S, 3, 1>(S(1), S(0), S(0));
    else { for (int i = 0; i < 3; ++i) ax(i) = std::sqrt(std::max(S(0), (m(i, i) + S(1)) / S(2))); if (m(0, 1) < 0) ax(1) = -ax(1); if (m(0, 2) < 0) ax(2) = -ax(2); }
  }
  S angle() const { return a; }
  S& angle() { return a; }
  const Matrix<S, 3, 1>& axis() const { return ax; }
  Matrix<S, 3, 1>& axis() { return ax; }
  Matrix<S, 3, 3> toRotationMatrix() const {
    const S c = std::cos(a), s = std::sin(a), t = S(1) - c, x = ax(0), y = ax(1), z = ax(2);
    Matrix<S, 3, 3> m; m(0, 0) = t * x * x + c; m(0, 1) = t * x * y - s * z; m(0, 2) = t * x * z + s * y; m(1, 0) = t * x * y + s * z; m(1, 1) = t * y * y + c; m(1, 2) = t * y * z - s * x;
    m(2, 0) = t * x * z - s * y; m(2, 1) = t * y * z + s * x; m(2, 2) = t * z * z + c; return m;
  }
  Matrix<S, 3, 3> matrix() const { return toRotationMatrix(); }
  operator Matrix<S, 3, 3>() const { return toRotationMatrix(); }
  AngleAxis inverse() const { AngleAxis r(*this); r.a = -a; return r; }
  template <int R, int C> Matrix<S, 3, 1> operator*(const Matrix<S, R, C>& v) const { return toRotationMatrix() * Matrix<S, 3, 1>(v); }
};
typedef AngleAxis<double> AngleAxisd;

template <typename S> class Quaternion {
 public:
  S qw, qx, qy, qz;
  Quaternion() : qw(1), qx(0), qy(0), qz(0) {}
  Quaternion(S w_, S x_, S y_, S z_) : qw(w_), qx(x_), qy(y_), qz(z_) {}
  template <int R, int C> Quaternion(const Matrix<S, R, C>& m) { AngleAxis<S> aa(m); *this = Quaternion(aa); }
  Quaternion(const AngleAxis<S>& aa) { const S h = aa.angle() / S(2), s = std::sin(h); qw = std::cos(h); qx = s * aa.axis()(0); qy = s * aa.axis()(1); qz = s * aa.axis()(2); }
  S w() const { return qw; } S x() const { return qx; } S y() const { return qy; } S z() const { return qz; }
  S& w() { return qw; } S& x() { return qx; } S& y() { return qy; } S& z() { return qz; }
  void normalize() { const S n = std::sqrt(qw * qw + qx * qx + qy * qy + qz * qz); qw /= n; qx /= n; qy /= n; qz /= n; }
  Quaternion normalized() const { Quaternion q(*this); q.normalize(); return q; }
  Matrix<S, 3, 3> toRotationMatrix() const {
    Matrix<S, 3, 3> m; m(0, 0) = 1 - 2 * (qy * qy + qz * qz); m(0, 1) = 2 * (qx * qy - qz * qw); m(0, 2) = 2 * (qx * qz + qy * qw); m(1, 0) = 2 * (qx * qy + qz * qw); m(1, 1) = 1 - 2 * (qx * qx + qz * qz);
    m(1, 2) = 2 * (qy * qz - qx * qw); m(2, 0) = 2 * (qx * qz - qy * qw); m(2, 1) = 2 * (qy * qz + qx * qw); m(2, 2) = 1 - 2 * (qx * qx + qy * qy); return m;
  }
  Matrix<S, 3, 3> matrix() const { return toRotationMatrix(); }
};
typedef Quaternion<double> Quaterniond;

enum { Affine = 1, Isometry = 2, Projective = 3 };
template <typename S, int Dim, int Mode> class Transform {
 public:
  Matrix<S, Dim + 1, Dim + 1> m;
  Transform() { m.setIdentity(); }
  template <int R, int C> Transform(const Matrix<S, R, C>& a) : m(a) {}
  Matrix<S, Dim + 1, Dim + 1>& matrix() { return m; }
  const Matrix<S, Dim + 1, Dim + 1>& matrix() const { return m; }
  Transform inverse() const { Transform t; t.m = m.inverse(); return t; }
  Matrix<S, Dim, Dim> linear() const { return Matrix<S, Dim, Dim>(m.block(0, 0, Dim, Dim)); }
  Matrix<S, Dim, Dim> rotation() const { Matrix<S, Dim, Dim> l = linear(); JacobiSVD<Matrix<S, Dim, Dim>> svd(l); return Matrix<S, Dim, Dim>(svd.matrixU() * svd.matrixV().transpose()); }
  Matrix<S, Dim, 1> translation() const { return Matrix<S, Dim, 1>(m.block(0, Dim, Dim, 1)); }
  template <int R, int C> Matrix<S, Dim, 1> operator*(const Matrix<S, R, C>& v) const { Matrix<S, Dim, 1> r; for (int i = 0; i < Dim; ++i) { S a = m(i, Dim); for (int j = 0; j < Dim; ++j) a += m(i, j) * v.d[j]; r.d[i] = a; } return r; }
  Transform operator*(const Transform& o) const { Transform t; t.m = m * o.m; return t; }
};
typedef Transform<double, 3, Affine> Affine3d;

// least-squares similarity transform dst ~ c R src + t (columns are points), after Umeyama 1991
template <int R, int C, int R2, int C2>
Matrix<double, Dynamic, Dynamic> umeyama(const Matrix<double, R, C>& src, const Matrix<double, R2, C2>& dst, bool with_scaling = true) {
  const std::size_t dim = src.rows(), n = src.cols();
  MatrixXd mu_s = MatrixXd::Zero(dim, 1), mu_d = MatrixXd::Zero(dim, 1);
  for (std::size_t j = 0; j < n; ++j) for (std::size_t i = 0; i < dim; ++i) { mu_s(i, 0) += src(i, j) / n; mu_d(i, 0) += dst(i, j) / n; }
  MatrixXd sigma = MatrixXd::Zero(dim, dim); double var_s = 0;
  for (std::size_t j = 0; j < n; ++j) for (std::size_t i = 0; i < dim; ++i) { const double a = src(i, j) - mu_s(i, 0); var_s += a * a / n; for (std::size_t k = 0; k < dim; ++k) sigma(k, i) += (dst(k, j) - mu_d(k, 0)) * a / n; }
  JacobiSVD<MatrixXd> svd(sigma);
  MatrixXd Sg = MatrixXd::Identity(dim, dim);
  if (sigma.determinant() < 0) Sg(dim - 1, dim - 1) = -1;
  MatrixXd Rm = svd.matrixU() * Sg * svd.matrixV().transpose();
  double c = 1.0;
  if (with_scaling) { double tr = 0; for (std::size_t i = 0; i < dim; ++i) tr += svd.singularValues()(i) * Sg(i, i); c = var_s > 0 ? tr / var_s : 1.0; }
  MatrixXd T = MatrixXd::Identity(dim + 1, dim + 1);
  MatrixXd t = mu_d - (Rm * mu_s) * c;
  for (std::size_t i = 0; i < dim; ++i) { for (std::size_t j = 0; j < dim; ++j) T(i, j) = c * Rm(i, j); T(i, dim) = t(i, 0); }
  return T;
}

template <typename S, int R, int C> struct aligned_allocator : std::allocator<Matrix<S, R, C>> {};
}  // namespace Eigen
#endif
