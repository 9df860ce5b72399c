// Header-compatible stand-in for the reference's math3d.h (VerStarting/math3d.h:9-315): same names
// (math3d::V3D_Base / M4D_Base / V3D / M4D / Deg2Rad / V3DStr / M4DStr), same public members, same operation
// order, so code written against the reference compiles unchanged against mythtracer_b200.
// Host-side convenience only: the renderer's arithmetic runs on the GPU (mythtracer_b200/csrc).
#pragma once
#include <cmath>
#include <cstddef>
#include <iomanip>
#include <iostream>
#include <sstream>
#include <string>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

namespace math3d {

template <typename T>
class V3D_Base {
 public:
  typedef T basetype;

  T v[3]{};

#define MTB_V3_BINARY(op)                                                                                      \
  V3D_Base operator op(const V3D_Base &o) const { return V3D_Base{v[0] op o.v[0], v[1] op o.v[1], v[2] op o.v[2]}; } \
  V3D_Base &operator op##=(const V3D_Base &o) {                                                                  \
    for (int i = 0; i < 3; i++) v[i] op## = o.v[i];                                                              \
    return *this;                                                                                                \
  }
  MTB_V3_BINARY(+)
  MTB_V3_BINARY(-)
  MTB_V3_BINARY(*)  // element-wise, as in the reference (math3d.h:57-75)
  MTB_V3_BINARY(/)
#undef MTB_V3_BINARY

  V3D_Base operator-() const { return V3D_Base{-v[0], -v[1], -v[2]}; }
  V3D_Base operator+() const { return *this; }

  V3D_Base operator*(T s) const { return V3D_Base{v[0] * s, v[1] * s, v[2] * s}; }
  V3D_Base operator/(T s) const { return V3D_Base{v[0] / s, v[1] / s, v[2] / s}; }
  V3D_Base &operator*=(T s) {
    for (int i = 0; i < 3; i++) v[i] *= s;
    return *this;
  }
  V3D_Base &operator/=(T s) {
    for (int i = 0; i < 3; i++) v[i] /= s;
    return *this;
  }

  T SqrLength() const { return v[0] * v[0] + v[1] * v[1] + v[2] * v[2]; }
  T Length() const { return std::sqrt(SqrLength()); }
  T SqrDistance(const V3D_Base &a) const {
    const T dx = a.v[0] - v[0], dy = a.v[1] - v[1], dz = a.v[2] - v[2];
    return dx * dx + dy * dy + dz * dz;
  }
  T Distance(const V3D_Base &a) const { return std::sqrt(SqrDistance(a)); }
  T Dot(const V3D_Base &a) const { return a.v[0] * v[0] + a.v[1] * v[1] + a.v[2] * v[2]; }
  V3D_Base Cross(const V3D_Base &a) const {
    return V3D_Base{v[1] * a.v[2] - v[2] * a.v[1], v[2] * a.v[0] - v[0] * a.v[2], v[0] * a.v[1] - v[1] * a.v[0]};
  }
  void Norm() {
    const T len = Length();
    for (int i = 0; i < 3; i++) v[i] /= len;
  }
  V3D_Base DupNorm() const {
    V3D_Base r(*this);
    r.Norm();
    return r;
  }

  T &x() { return v[0]; }
  T &y() { return v[1]; }
  T &z() { return v[2]; }
  T &r() { return v[0]; }
  T &g() { return v[1]; }
  T &b() { return v[2]; }
  const T &x() const { return v[0]; }
  const T &y() const { return v[1]; }
  const T &z() const { return v[2]; }
  const T &r() const { return v[0]; }
  const T &g() const { return v[1]; }
  const T &b() const { return v[2]; }
};

template <typename T>
std::ostream &operator<<(std::ostream &os, const V3D_Base<T> &a) {
  return os << std::fixed << std::setprecision(5) << a.v[0] << ", " << a.v[1] << ", " << a.v[2];
}

template <typename T>
std::string ToStr(const T &a) {
  std::ostringstream s;
  s << a;
  return s.str();
}
#define V3DStr(a) math3d::ToStr(a).c_str()
#define M4DStr(a) math3d::ToStr(a).c_str()

inline double Deg2Rad(double angle) { return (angle * M_PI) / 180.0; }

template <typename T>
class M4D_Base {
 public:
  typedef T basetype;

  T m[4][4]{};

  M4D_Base operator*(const M4D_Base &a) {
    M4D_Base r;
    for (size_t row = 0; row < 4; row++) {
      for (size_t col = 0; col < 4; col++) {
        r.m[row][col] = m[row][0] * a.m[0][col] + m[row][1] * a.m[1][col] + m[row][2] * a.m[2][col] + m[row][3] * a.m[3][col];
      }
    }
    return r;
  }
  M4D_Base &operator*=(const M4D_Base &a) {
    *this = *this * a;
    return *this;
  }
  // The reference adds m[0][3] to all three rows (math3d.h:210-216); kept, it is 0 for every rotation.
  template <typename U>
  V3D_Base<U> operator*(const V3D_Base<U> &a) {
    V3D_Base<U> r;
    for (size_t row = 0; row < 3; row++) r.v[row] = m[row][0] * a.v[0] + m[row][1] * a.v[1] + m[row][2] * a.v[2] + m[0][3];
    return r;
  }

  void ResetIdentity() {
    for (size_t j = 0; j < 4; j++)
      for (size_t i = 0; i < 4; i++) m[j][i] = (i == j) ? 1.0 : 0.0;
  }
  void ResetRotationXRad(T a) {
    ResetIdentity();
    m[1][1] = cos(a);
    m[1][2] = -sin(a);
    m[2][1] = sin(a);
    m[2][2] = cos(a);
  }
  void ResetRotationYRad(T a) {
    ResetIdentity();
    m[0][0] = cos(a);
    m[0][2] = sin(a);
    m[2][0] = -sin(a);
    m[2][2] = cos(a);
  }
  void ResetRotationZRad(T a) {
    ResetIdentity();
    m[0][0] = cos(a);
    m[0][1] = -sin(a);
    m[1][0] = sin(a);
    m[1][1] = cos(a);
  }
#define MTB_M4_ROT(AXIS)                                   \
  static M4D_Base<T> Rotation##AXIS##Rad(T a) {            \
    M4D_Base<T> r;                                         \
    r.ResetRotation##AXIS##Rad(a);                         \
    return r;                                              \
  }                                                        \
  static M4D_Base<T> Rotation##AXIS##Deg(T a) { return Rotation##AXIS##Rad(Deg2Rad(a)); }
  MTB_M4_ROT(X)
  MTB_M4_ROT(Y)
  MTB_M4_ROT(Z)
#undef MTB_M4_ROT
};

template <typename T>
std::ostream &operator<<(std::ostream &os, const M4D_Base<T> &a) {
  os << std::fixed << std::setprecision(5);
  for (int row = 0; row < 4; row++) {
    os << (row == 0 ? "[  " : "   ") << a.m[row][0] << ", " << a.m[row][1] << ", " << a.m[row][2] << ", " << a.m[row][3]
       << (row == 3 ? "  ]\n" : "   \n");
  }
  return os;
}

typedef V3D_Base<double> V3D;
typedef M4D_Base<double> M4D;

}  // namespace math3d
