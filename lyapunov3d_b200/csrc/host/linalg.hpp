// linalg.hpp -- float vector / quaternion helpers for the host scene layer.
//
// Semantics follow the reference's vec3.hpp / quat.hpp, including the places where
// the reference promotes to double (epsilon compares, QUAT_PI) and its quirks
// (Quat::normalize multiplies by the magnitude, quat.hpp:148-165).  Results must be
// bit-identical to the reference host build, so every sum keeps the reference's
// left-to-right association; compile with -ffp-contract=off.
#pragma once
#include <cmath>

#include "lyap/types.h"

namespace lyap_host {

struct V3 {
    float x, y, z;
    V3() : x(0), y(0), z(0) {}
    V3(float x_, float y_, float z_) : x(x_), y(y_), z(z_) {}
    V3(const lyap_vec3 &v) : x(v.x), y(v.y), z(v.z) {}
    operator lyap_vec3() const { return lyap_vec3{x, y, z}; }
};

inline V3 operator+(V3 a, V3 b) { return V3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline V3 operator-(V3 a, V3 b) { return V3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline V3 operator*(V3 a, float s) { return V3(a.x * s, a.y * s, a.z * s); }
inline V3 operator/(V3 a, float s) { return V3(a.x / s, a.y / s, a.z / s); }
inline float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline V3 cross(V3 a, V3 b) { return V3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }

// vec3.hpp:128-154
inline V3 unit(V3 a)
{
    const float m2 = dot(a, a);
    if ((double)m2 < 1e-12) return V3();
    if (m2 == 1.0f || ((double)m2 > (double)1.0f - 1e-12 && (double)m2 < (double)1.0f + 1e-12)) return a;
    return a / std::sqrt(m2);
}

struct Q4 {
    float x, y, z, w;
    Q4() : x(0), y(0), z(0), w(1.0f) {}
    Q4(float x_, float y_, float z_, float w_) : x(x_), y(y_), z(z_), w(w_) {}
    Q4(const lyap_quat &q) : x(q.x), y(q.y), z(q.z), w(q.w) {}
    void store(lyap_quat &q) const { q.x = x; q.y = y; q.z = z; q.w = w; }
};

// quat.hpp:148-165 -- scales BY the magnitude when off unit length (reference quirk)
inline Q4 ref_normalize(Q4 q)
{
    const float m2 = q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w;
    if ((double)m2 < 1e-6) return Q4();
    if (m2 == 1.0f || ((double)m2 > (double)1.0f - 1e-6 && (double)m2 < (double)1.0f + 1e-6)) return q;
    const float m = std::sqrt(m2);
    return Q4(q.x * m, q.y * m, q.z * m, q.w * m);
}

// quat.hpp:227-238
inline Q4 from_axis_angle(V3 axis, float ang, bool degrees)
{
    ang = degrees ? (float)((double)ang * 3.14159265358979323846264338327950288 / (double)360.0f) : ang * 0.5f;
    const float s = std::sin(ang);
    return ref_normalize(Q4(axis.x * s, axis.y * s, axis.z * s, std::cos(ang)));
}

// quat.hpp:194-220
inline Q4 rotation_between(V3 p, V3 q, float scale)
{
    float cosa = dot(p, q);
    if ((double)cosa < -1.0) cosa = -1.0f;
    else if ((double)cosa > 1.0) cosa = 1.0f;
    if (cosa == 0 || ((double)cosa >= -1e-6 && (double)cosa <= 1e-6)) return Q4();
    const float ang = std::acos(cosa);
    const V3 axis = cross(p, q);
    const float half = (float)((double)ang * 0.5 * (double)scale);
    const float k = std::sin(half) / std::sin(ang);
    return Q4(axis.x * k, axis.y * k, axis.z * k, std::cos(half));
}

// quat.hpp:172-192
inline Q4 nlerp(Q4 a, Q4 b, float t)
{
    if ((double)t == 0.0 || (double)t < 1e-6) return a;
    if (t == 1.0f || (double)t > (double)1.0f - 1e-6) return b;
    const float d = a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w;
    const float tA = d >= 0 ? t : -t;
    const float tI = 1.0f - t;
    return ref_normalize(Q4(a.x * tI + b.x * tA, a.y * tI + b.y * tA, a.z * tI + b.z * tA, a.w * tI + b.w * tA));
}

// quat.hpp:279-284
inline V3 rotate(Q4 q, V3 v)
{
    const float x = q.x, y = q.y, z = q.z, w = q.w;
    return V3(w * w * v.x + 2 * y * w * v.z - 2 * z * w * v.y + x * x * v.x + 2 * y * x * v.y + 2 * z * x * v.z - z * z * v.x - y * y * v.x,
              2 * x * y * v.x + y * y * v.y + 2 * z * y * v.z + 2 * w * z * v.x - z * z * v.y + w * w * v.y - 2 * x * w * v.z - x * x * v.y,
              2 * x * z * v.x + 2 * y * z * v.y + z * z * v.z - 2 * w * y * v.x - y * y * v.z + 2 * w * x * v.y - x * x * v.z + w * w * v.z);
}

} // namespace lyap_host
