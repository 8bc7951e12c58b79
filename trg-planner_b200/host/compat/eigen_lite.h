// Minimal stand-ins for the Eigen value types at the TRG API boundary (trg.h:18-98), used only
// when the real Eigen3 headers are absent (this image has none). With Eigen installed,
// trg_types.h includes <Eigen/Core> instead and this file is not compiled.
// Semantics follow Eigen: norm() = sqrt(x*x + y*y); normalized() divides by sqrt(squaredNorm)
// when it is > 0 (Eigen/src/Core/Dot.h).
#pragma once
#include <cmath>

namespace Eigen {

struct Vector2f {
  float v[2];
  Vector2f() : v{0.f, 0.f} {}
  Vector2f(float x, float y) : v{x, y} {}
  static Vector2f Zero() { return Vector2f(0.f, 0.f); }
  float& x() { return v[0]; }
  float& y() { return v[1]; }
  float x() const { return v[0]; }
  float y() const { return v[1]; }
  float& operator[](int i) { return v[i]; }
  float operator[](int i) const { return v[i]; }
  float& operator()(int i) { return v[i]; }
  float operator()(int i) const { return v[i]; }
  float squaredNorm() const { return v[0] * v[0] + v[1] * v[1]; }
  float norm() const { return std::sqrt(squaredNorm()); }
  Vector2f normalized() const {
    float z = squaredNorm();
    if (z > 0.f) { float s = std::sqrt(z); return Vector2f(v[0] / s, v[1] / s); }
    return *this;
  }
  void normalize() { *this = normalized(); }
  Vector2f operator+(const Vector2f& o) const { return Vector2f(v[0] + o.v[0], v[1] + o.v[1]); }
  Vector2f operator-(const Vector2f& o) const { return Vector2f(v[0] - o.v[0], v[1] - o.v[1]); }
  Vector2f operator-() const { return Vector2f(-v[0], -v[1]); }
  Vector2f operator*(float s) const { return Vector2f(v[0] * s, v[1] * s); }
  Vector2f operator/(float s) const { return Vector2f(v[0] / s, v[1] / s); }
};
inline Vector2f operator*(float s, const Vector2f& a) { return Vector2f(s * a.v[0], s * a.v[1]); }

struct Vector3f {
  float v[3];
  Vector3f() : v{0.f, 0.f, 0.f} {}
  Vector3f(float x, float y, float z) : v{x, y, z} {}
  static Vector3f Zero() { return Vector3f(0.f, 0.f, 0.f); }
  float& x() { return v[0]; }
  float& y() { return v[1]; }
  float& z() { return v[2]; }
  float x() const { return v[0]; }
  float y() const { return v[1]; }
  float z() const { return v[2]; }
  float& operator[](int i) { return v[i]; }
  float operator[](int i) const { return v[i]; }
  float& operator()(int i) { return v[i]; }
  float operator()(int i) const { return v[i]; }
  Vector2f head(int) const { return Vector2f(v[0], v[1]); }  // read-only head(2)
  float norm() const { return std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); }
  Vector3f operator+(const Vector3f& o) const { return Vector3f(v[0] + o.v[0], v[1] + o.v[1], v[2] + o.v[2]); }
  Vector3f operator-(const Vector3f& o) const { return Vector3f(v[0] - o.v[0], v[1] - o.v[1], v[2] - o.v[2]); }
  Vector3f& operator+=(const Vector3f& o) { v[0] += o.v[0]; v[1] += o.v[1]; v[2] += o.v[2]; return *this; }
  Vector3f operator/(float s) const { return Vector3f(v[0] / s, v[1] / s, v[2] / s); }
  Vector3f operator*(float s) const { return Vector3f(v[0] * s, v[1] * s, v[2] * s); }
};

}  // namespace Eigen
