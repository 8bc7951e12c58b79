// ORACLE — TEST INFRASTRUCTURE ONLY. Not part of the shipped product path.
//
// Restatement of the *third-party* arithmetic the reference's hot path calls
// into (Eigen3; version unpinned by the reference: `find_package(Eigen3 REQUIRED)`
// cpp/trg_planner/CMakeLists.txt:23, apt libeigen3-dev => 3.3.7 or 3.4.0).
// Eigen is ABSENT from /root/reference and from this image, so everything in
// this header is restated from Eigen's published algorithm (recalled from
// Eigen 3.4.0: Eigen/src/SVD/JacobiSVD.h, Eigen/src/Jacobi/Jacobi.h,
// Eigen/src/misc/RealSvd2x2.h).
//
//   *** PARITY UNPINNED at this boundary: the reference holds no golden
//   *** vectors / tests for the edge weight, and Eigen's float summation
//   *** order (vectorised redux / GEBP product) cannot be reproduced here.
//   *** Everything else in the oracle (set membership, medians, ids, order)
//   *** is plain IEEE float arithmetic and is bit-reproducible.
//
// Call sites in the reference this header serves (trg.cpp):
//   :271,276  Vector2f::norm()            -> v2_norm
//   :277      Vector2f::normalized()      -> v2_normalized
//   :337      A.rowwise() - A.colwise().mean()
//   :338      (centered.adjoint()*centered)/double(n-1)
//   :339      JacobiSVD<MatrixXf>(cov, ComputeFullU)
//   :340      matrixU().normalized()   (Frobenius norm of the 3x3 => U/sqrt(3))
#ifndef ORACLE_EIGEN_RESTATE_H_
#define ORACLE_EIGEN_RESTATE_H_

#include <cmath>
#include <limits>
#include <utility>
#include <vector>

namespace erst {

// ---- Vector2f helpers (Eigen/src/Core/Dot.h: squaredNorm, norm, normalized) ----
inline float v2_sqnorm(float x, float y) { return x * x + y * y; }
inline float v2_norm(float x, float y) { return std::sqrt(v2_sqnorm(x, y)); }
// normalized(): z = squaredNorm(); if (z > 0) return v / sqrt(z); else return v;
inline void v2_normalized(float x, float y, float* ox, float* oy) {
  float z = v2_sqnorm(x, y);
  if (z > 0.0f) {
    float s = std::sqrt(z);
    *ox = x / s;
    *oy = y / s;
  } else {
    *ox = x;
    *oy = y;
  }
}

// ---- Jacobi rotation (Eigen/src/Jacobi/Jacobi.h) ----
template <typename T>
struct Rot {
  T c, s;
  Rot() : c(1), s(0) {}
  Rot(T c_, T s_) : c(c_), s(s_) {}
  // operator* for real scalars: (c*oc - s*os, c*os + s*oc)
  Rot operator*(const Rot& o) const { return Rot(c * o.c - s * o.s, c * o.s + s * o.c); }
  Rot transpose() const { return Rot(c, -s); }
  // makeJacobi(x, y, z): rotation J such that J^T [x y; y z] J is diagonal
  bool makeJacobi(T x, T y, T z) {
    T deno = T(2) * std::abs(y);
    if (deno < (std::numeric_limits<T>::min)()) {
      c = T(1);
      s = T(0);
      return false;
    }
    T tau = (x - z) / deno;
    T w = std::sqrt(tau * tau + T(1));
    T t;
    if (tau > T(0)) t = T(1) / (tau + w);
    else t = T(1) / (tau - w);
    T sign_t = t > T(0) ? T(1) : T(-1);
    T n = T(1) / std::sqrt(t * t + T(1));
    s = -sign_t * (y / std::abs(y)) * std::abs(t) * n;
    c = n;
    return true;
  }
};

// apply_rotation_in_the_plane(x, y, j): x_i' = c x_i + s y_i ; y_i' = -s x_i + c y_i
template <typename T>
inline void rot_plane(T* x, int incx, T* y, int incy, int n, const Rot<T>& j) {
  if (j.c == T(1) && j.s == T(0)) return;
  for (int i = 0; i < n; ++i) {
    T xi = x[i * incx], yi = y[i * incy];
    x[i * incx] = j.c * xi + j.s * yi;
    y[i * incy] = -j.s * xi + j.c * yi;
  }
}

// M is row-major 3x3: M[r*3+c]
template <typename T>
inline void apply_left(T* M, int p, int q, const Rot<T>& j) {  // rows p,q
  rot_plane(M + p * 3, 1, M + q * 3, 1, 3, j);
}
template <typename T>
inline void apply_right(T* M, int p, int q, const Rot<T>& j) {  // cols p,q with j.transpose()
  rot_plane(M + p, 3, M + q, 3, 3, j.transpose());
}

// real_2x2_jacobi_svd (Eigen/src/misc/RealSvd2x2.h)
template <typename T>
inline void real_2x2_jacobi_svd(const T* M, int p, int q, Rot<T>* j_left, Rot<T>* j_right) {
  T m[4] = {M[p * 3 + p], M[p * 3 + q], M[q * 3 + p], M[q * 3 + q]};  // [m00 m01; m10 m11]
  Rot<T> rot1;
  T t = m[0] + m[3];
  T d = m[2] - m[1];
  if (std::abs(d) < (std::numeric_limits<T>::min)()) {
    rot1.s = T(0);
    rot1.c = T(1);
  } else {
    T u = t / d;
    T tmp = std::sqrt(T(1) + u * u);
    rot1.s = T(1) / tmp;
    rot1.c = u / tmp;
  }
  // m.applyOnTheLeft(0,1,rot1)
  rot_plane(m + 0, 1, m + 2, 1, 2, rot1);
  j_right->makeJacobi(m[0], m[1], m[3]);
  *j_left = rot1 * j_right->transpose();
}

// JacobiSVD<MatrixXf>(A, ComputeFullU) for a square real 3x3 (no QR preconditioning
// is run for square input). A row-major; U row-major; sv descending.
template <typename T>
inline void jacobi_svd3(const T* A, T* U, T* sv) {
  const T precision = T(2) * std::numeric_limits<T>::epsilon();
  const T considerAsZero = (std::numeric_limits<T>::min)();
  T scale = T(0);
  for (int i = 0; i < 9; ++i) scale = std::max(scale, std::abs(A[i]));
  if (!(std::isfinite(scale))) {  // InvalidInput: Eigen leaves U unspecified; we return identity
    for (int i = 0; i < 9; ++i) U[i] = (i % 4 == 0) ? T(1) : T(0);
    sv[0] = sv[1] = sv[2] = std::numeric_limits<T>::quiet_NaN();
    return;
  }
  if (scale == T(0)) scale = T(1);
  T W[9];
  for (int i = 0; i < 9; ++i) W[i] = A[i] / scale;
  for (int i = 0; i < 9; ++i) U[i] = (i % 4 == 0) ? T(1) : T(0);

  T maxDiag = std::max(std::abs(W[0]), std::max(std::abs(W[4]), std::abs(W[8])));
  bool finished = false;
  int guard = 0;
  while (!finished && guard++ < 1000) {
    finished = true;
    for (int p = 1; p < 3; ++p) {
      for (int q = 0; q < p; ++q) {
        T threshold = std::max(considerAsZero, precision * maxDiag);
        if (std::abs(W[p * 3 + q]) > threshold || std::abs(W[q * 3 + p]) > threshold) {
          finished = false;
          Rot<T> jl, jr;
          real_2x2_jacobi_svd(W, p, q, &jl, &jr);
          apply_left(W, p, q, jl);
          // m_matrixU.applyOnTheRight(p,q,j_left.transpose())
          apply_right(U, p, q, jl.transpose());
          apply_right(W, p, q, jr);
          maxDiag = std::max(maxDiag, std::max(std::abs(W[p * 3 + p]), std::abs(W[q * 3 + q])));
        }
      }
    }
  }
  for (int i = 0; i < 3; ++i) {
    T a = W[i * 3 + i];
    sv[i] = std::abs(a);
    if (a < T(0))
      for (int r = 0; r < 3; ++r) U[r * 3 + i] = -U[r * 3 + i];
  }
  for (int i = 0; i < 3; ++i) sv[i] *= scale;
  // selection sort, descending, swapping U columns (first max wins on ties)
  for (int i = 0; i < 3; ++i) {
    int pos = 0;
    T best = sv[i];
    for (int k = i + 1; k < 3; ++k)
      if (sv[k] > best) {
        best = sv[k];
        pos = k - i;
      }
    if (best == T(0)) break;
    if (pos) {
      pos += i;
      std::swap(sv[i], sv[pos]);
      for (int r = 0; r < 3; ++r) std::swap(U[r * 3 + i], U[r * 3 + pos]);
    }
  }
}

// trg.cpp:332-363 — rows (x_local, y_local, z) -> edge weight.
// T = float follows the reference's float pipeline with a plain sequential
// summation order; T = double is the conditioning probe of SURVEY.md A.4.
template <typename T>
inline T edge_weight_from_rows(const std::vector<float>& rows /* n*3 */) {
  const int n = static_cast<int>(rows.size() / 3);
  T mean[3] = {0, 0, 0};
  for (int i = 0; i < n; ++i)
    for (int c = 0; c < 3; ++c) mean[c] += T(rows[i * 3 + c]);
  for (int c = 0; c < 3; ++c) mean[c] /= T(n);  // colwise().mean() = sum / size
  T cov[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (int i = 0; i < n; ++i) {
    T cx[3];
    for (int c = 0; c < 3; ++c) cx[c] = T(rows[i * 3 + c]) - mean[c];
    for (int a = 0; a < 3; ++a)
      for (int b = 0; b < 3; ++b) cov[a * 3 + b] += cx[a] * cx[b];
  }
  const T denom = T(static_cast<double>(n - 1));  // double(n-1) converted to Scalar
  for (int i = 0; i < 9; ++i) cov[i] /= denom;
  T U[9], sv[3];
  jacobi_svd3<T>(cov, U, sv);
  // matrixU().normalized(): divide by the Frobenius norm of the whole matrix
  T sq = 0;
  for (int i = 0; i < 9; ++i) sq += U[i] * U[i];
  T nrm = std::sqrt(sq);
  T ev20 = U[2 * 3 + 0], ev21 = U[2 * 3 + 1];
  if (sq > T(0)) {
    ev20 /= nrm;
    ev21 /= nrm;
  }
  // col.dot(-gravity) with gravity=(0,0,-1) is the z component; sign-flip form :347-354
  T hor = ev20 < T(0) ? -ev20 : ev20;
  T ver = ev21 < T(0) ? -ev21 : ev21;
  const float ratio_f = 0.8;  // `float ratio = 0.8;`
  T ratio = T(ratio_f);
  T one_minus = T(static_cast<float>(1 - ratio_f));
  T weight = ratio * hor + one_minus * ver;
  if (static_cast<double>(weight) < 0.1) weight = T(0);
  return weight;
}

}  // namespace erst
#endif  // ORACLE_EIGEN_RESTATE_H_
