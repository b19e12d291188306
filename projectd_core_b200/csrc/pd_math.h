/*
 * pd_math.h -- small fp32 vector / quaternion helpers shared by every device function of the
 * batched Car::step path.  Operation order follows the reference's Core/Math.h:91-207 (vec3f, row-vector
 * mat44f) and ODE's odemath.h (dMultiply0_331 / dMultiply1_331, dCalcVectorDot3) so that single-tick
 * results stay within fp32 round-off of the reference.
 *
 * Everything is `PD_HD` (= __host__ __device__ under nvcc, nothing under g++): the same source is
 * compiled by nvcc for sm_100a (the product) and by g++ for tests/hostsim (a debugging aid that lets the
 * kernels' arithmetic be single-stepped on a box without a GPU; it is not a product path).
 */
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define PD_HD __host__ __device__ __forceinline__
#define PD_HDN __host__ __device__ __noinline__
#define PD_UNROLL _Pragma("unroll")
#define PD_NOUNROLL _Pragma("unroll 1")
#define PD_UNROLL4 _Pragma("unroll 4")
#else
#define PD_UNROLL
#define PD_UNROLL4
#define PD_NOUNROLL
#define PD_HD inline
#define PD_HDN inline
#endif

namespace pd {

/* out-of-line copies of the math-library functions: every inlined call site would carry its own copy of the slow path
 * (argument reduction etc.), and this kernel is instruction-fetch sensitive */
PD_HDN float m_sin(float x) { return sinf(x); }
PD_HDN float m_cos(float x) { return cosf(x); }
PD_HDN float m_tan(float x) { return tanf(x); }
PD_HDN float m_pow(float x, float y) { return powf(x, y); }
PD_HDN float m_acos(float x) { return acosf(x); }
PD_HDN float m_asin(float x) { return asinf(x); }
PD_HDN float m_atan2(float y, float x) { return atan2f(y, x); }


struct V3 {
    float x, y, z;
};
PD_HD V3 v3(float x, float y, float z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
PD_HD V3 operator+(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
PD_HD V3 operator-(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
PD_HD V3 operator*(V3 a, float f) { return v3(a.x * f, a.y * f, a.z * f); }
PD_HD V3 operator*(float f, V3 a) { return v3(a.x * f, a.y * f, a.z * f); }
PD_HD V3 operator/(V3 a, float f) { return v3(a.x / f, a.y / f, a.z / f); }
PD_HD void operator+=(V3& a, V3 b) { a.x += b.x; a.y += b.y; a.z += b.z; }
PD_HD void operator-=(V3& a, V3 b) { a.x -= b.x; a.y -= b.y; a.z -= b.z; }
PD_HD float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
PD_HD V3 cross(V3 a, V3 b) { return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
PD_HD float sqlen(V3 a) { return a.x * a.x + a.y * a.y + a.z * a.z; }
PD_HD float len(V3 a) { return sqrtf(sqlen(a)); }
/* vec3f::norm(l): scale by 1/l unless l == 0 (Core/Math.h:118-121) */
PD_HD V3 norm_l(V3 a, float l) { if (l != 0.0f) { float s = 1.0f / l; return a * s; } return a; }
PD_HD V3 norm(V3 a) { return norm_l(a, len(a)); }
PD_HD V3 neg(V3 a) { return v3(-a.x, -a.y, -a.z); }

PD_HD float tminf(float a, float b) { return a < b ? a : b; }
PD_HD float tmaxf(float a, float b) { return a > b ? a : b; }
PD_HD float tclampf(float x, float a, float b) { return x < a ? a : (x > b ? b : x); }
PD_HD double tclampd(double x, double a, double b) { return x < a ? a : (x > b ? b : x); }
PD_HD float signf_(float x) { return (x > 0.0f) ? 1.0f : ((x < 0.0f) ? -1.0f : 0.0f); }
PD_HD float linscalef(float x, float x0, float x1, float r0, float r1) {
    x = tclampf(x, x0, x1);
    return ((r1 - r0) * (x - x0)) / (x1 - x0) + r0;
}
PD_HD bool finitef(float x) { return (x - x) == 0.0f; }

struct Quat { float w, x, y, z; };

/* rigid-body frame: position + rotation held as the three body axes in world coordinates
 * (ax = first column of ODE's R = mat44f row 1: M11,M12,M13 ...) */
struct Frame {
    V3 p;
    V3 ax, ay, az;
};
/* dMultiply0_331: R * v */
PD_HD V3 rot(const Frame& f, V3 v) {
    return v3(f.ax.x * v.x + f.ay.x * v.y + f.az.x * v.z,
              f.ax.y * v.x + f.ay.y * v.y + f.az.y * v.z,
              f.ax.z * v.x + f.ay.z * v.y + f.az.z * v.z);
}
/* dMultiply1_331: R^T * v */
PD_HD V3 irot(const Frame& f, V3 v) { return v3(dot(f.ax, v), dot(f.ay, v), dot(f.az, v)); }
/* dBodyGetRelPointPos / dBodyGetPosRelPoint */
PD_HD V3 to_world(const Frame& f, V3 p) { V3 r = rot(f, p); return v3(r.x + f.p.x, r.y + f.p.y, r.z + f.p.z); }
PD_HD V3 to_local(const Frame& f, V3 p) { return irot(f, p - f.p); }

/* rotation.cpp dQtoR -> axes */
PD_HD void quat_to_axes(Quat q, V3& ax, V3& ay, V3& az) {
    float qq1 = 2 * q.x * q.x, qq2 = 2 * q.y * q.y, qq3 = 2 * q.z * q.z;
    float r00 = 1 - qq2 - qq3, r01 = 2 * (q.x * q.y - q.w * q.z), r02 = 2 * (q.x * q.z + q.w * q.y);
    float r10 = 2 * (q.x * q.y + q.w * q.z), r11 = 1 - qq1 - qq3, r12 = 2 * (q.y * q.z - q.w * q.x);
    float r20 = 2 * (q.x * q.z - q.w * q.y), r21 = 2 * (q.y * q.z + q.w * q.x), r22 = 1 - qq1 - qq2;
    ax = v3(r00, r10, r20); ay = v3(r01, r11, r21); az = v3(r02, r12, r22);
}
/* rotation.cpp dRtoQ from axes (R(i,j) = axis_j component i) + dNormalize4 */
PD_HD Quat axes_to_quat(V3 ax, V3 ay, V3 az) {
    float r00 = ax.x, r01 = ay.x, r02 = az.x, r10 = ax.y, r11 = ay.y, r12 = az.y, r20 = ax.z, r21 = ay.z, r22 = az.z;
    Quat q; float tr = r00 + r11 + r22, s;
    if (tr >= 0) {
        s = sqrtf(tr + 1); q.w = 0.5f * s; s = 0.5f * (1.0f / s);
        q.x = (r21 - r12) * s; q.y = (r02 - r20) * s; q.z = (r10 - r01) * s;
    } else {
        int c = 0;
        if (r11 > r00) { c = (r22 > r11) ? 2 : 1; } else if (r22 > r00) c = 2;
        if (c == 0) { s = sqrtf((r00 - (r11 + r22)) + 1); q.x = 0.5f * s; s = 0.5f * (1.0f / s); q.y = (r01 + r10) * s; q.z = (r20 + r02) * s; q.w = (r21 - r12) * s; }
        else if (c == 1) { s = sqrtf((r11 - (r22 + r00)) + 1); q.y = 0.5f * s; s = 0.5f * (1.0f / s); q.z = (r12 + r21) * s; q.x = (r01 + r10) * s; q.w = (r02 - r20) * s; }
        else { s = sqrtf((r22 - (r00 + r11)) + 1); q.z = 0.5f * s; s = 0.5f * (1.0f / s); q.x = (r20 + r02) * s; q.y = (r12 + r21) * s; q.w = (r10 - r01) * s; }
    }
    float l = q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z;
    if (l > 0) { l = 1.0f / sqrtf(l); q.w *= l; q.x *= l; q.y *= l; q.z *= l; } else { q.w = 1; q.x = q.y = q.z = 0; }
    return q;
}
/* ode.cpp dBodySetRotation: R = dOrthogonalizeR(input) (Gram-Schmidt on the ROWS of ODE's R, i.e. on
 * (ax.x,ay.x,az.x), (ax.y,ay.y,az.y); third row = row0 x row1), q = dRtoQ(raw input) normalised */
PD_HD void set_rotation(V3 ax, V3 ay, V3 az, V3& oax, V3& oay, V3& oaz, Quat& q) {
    q = axes_to_quat(ax, ay, az);
    V3 r0 = v3(ax.x, ay.x, az.x), r1 = v3(ax.y, ay.y, az.y);
    const float n0 = dot(r0, r0);
    if (n0 != 1.0f) { float l = n0; if (l > 0) { l = 1.0f / sqrtf(l); r0 = r0 * l; } else r0 = v3(1, 0, 0); }
    const float proj = dot(r0, r1);
    if (proj != 0) { r1.x -= proj * r0.x; r1.y -= proj * r0.y; r1.z -= proj * r0.z; }
    const float n1 = dot(r1, r1);
    if (n1 != 1.0f) { float l = n1; if (l > 0) { l = 1.0f / sqrtf(l); r1 = r1 * l; } else r1 = v3(1, 0, 0); }
    const V3 r2 = cross(r0, r1);
    oax = v3(r0.x, r1.x, r2.x); oay = v3(r0.y, r1.y, r2.y); oaz = v3(r0.z, r1.z, r2.z);
}

/* dQMultiply0: b*c ; dQMultiply1: inv(b)*c ; dQMultiply2: b*inv(c) */
PD_HD Quat qmul0(Quat b, Quat c) {
    Quat a;
    a.w = b.w * c.w - b.x * c.x - b.y * c.y - b.z * c.z;
    a.x = b.w * c.x + b.x * c.w + b.y * c.z - b.z * c.y;
    a.y = b.w * c.y + b.y * c.w + b.z * c.x - b.x * c.z;
    a.z = b.w * c.z + b.z * c.w + b.x * c.y - b.y * c.x;
    return a;
}
PD_HD Quat qmul1(Quat b, Quat c) {
    Quat a;
    a.w = b.w * c.w + b.x * c.x + b.y * c.y + b.z * c.z;
    a.x = b.w * c.x - b.x * c.w - b.y * c.z + b.z * c.y;
    a.y = b.w * c.y - b.y * c.w - b.z * c.x + b.x * c.z;
    a.z = b.w * c.z - b.z * c.w - b.x * c.y + b.y * c.x;
    return a;
}
PD_HD Quat qmul2(Quat b, Quat c) {
    Quat a;
    a.w = b.w * c.w + b.x * c.x + b.y * c.y + b.z * c.z;
    a.x = -b.w * c.x + b.x * c.w - b.y * c.z + b.z * c.y;
    a.y = -b.w * c.y + b.y * c.w - b.z * c.x + b.x * c.z;
    a.z = -b.w * c.z + b.z * c.w - b.x * c.y + b.y * c.x;
    return a;
}

/* mat44f::createFromAxisAngle (Core/Math.cpp:88-117) applied to a vector as `v * M` (row vector, no translation) */
struct M33 { float m11, m12, m13, m21, m22, m23, m31, m32, m33; };
PD_HD M33 axis_angle(V3 a, float angle) {
    M33 r; float s = m_sin(angle), c = m_cos(angle), o = 1.0f - c;
    r.m11 = ((a.x * a.x) * o) + c; r.m22 = ((a.y * a.y) * o) + c; r.m33 = ((a.z * a.z) * o) + c;
    r.m12 = (a.z * s) + (a.y * a.x) * o; r.m23 = (a.x * s) + (a.z * a.y) * o; r.m31 = (a.y * s) + (a.z * a.x) * o;
    r.m13 = (a.z * a.x) * o - (a.y * s); r.m21 = (a.y * a.x) * o - (a.z * s); r.m32 = (a.z * a.y) * o - (a.x * s);
    return r;
}

} // namespace pd
