// cb2_math.cuh — unfused f32 vector / quaternion / Decomposed arithmetic for sm_100a.
//
// Bit-exactness contract: the reference (rustc + cgmath 0.18) evaluates every
// + - * / sqrt as a separate IEEE-754 binary32 round-to-nearest operation and
// never contracts a*b+c.  All arithmetic here therefore goes through
// __fadd_rn/__fsub_rn/__fmul_rn/__fdiv_rn/__fsqrt_rn, which nvcc never fuses
// into FFMA and which are IEEE-correct regardless of -fmad/-prec-div/-ftz
// flags.  Operation ORDER follows cgmath (restated in SURVEY.md §8c):
//   dot(a,b)   = (a.x*b.x + a.y*b.y) + a.z*b.z
//   cross(a,b) = (a.y*b.z - a.z*b.y, a.z*b.x - a.x*b.z, a.x*b.y - a.y*b.x)
//   normalize_to(v,m) = v * (m / sqrt(dot(v,v)))
//   q*x        = cross(q.v, cross(q.v,x) + x*q.s) * 2 + x
//   invert(q)  = conjugate(q) / (q.s*q.s + dot(q.v,q.v))
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "cb2_hostsim.h"

namespace cb2 {

#define CB2_DI __device__ __forceinline__

CB2_DI float fadd(float a, float b) { return __fadd_rn(a, b); }
CB2_DI float fsub(float a, float b) { return __fsub_rn(a, b); }
CB2_DI float fmul(float a, float b) { return __fmul_rn(a, b); }
CB2_DI float fdiv(float a, float b) { return __fdiv_rn(a, b); }
CB2_DI float fsqrt(float a) { return __fsqrt_rn(a); }

#define CB2_F32_EPSILON 1.1920929e-7f
// approx::ulps_eq!(x, 0.0) with default epsilon (f32::EPSILON) and max_ulps (4):
// true iff |x| <= EPSILON (NaN -> false).
CB2_DI bool ulps_eq_zero(float x) { return fabsf(x) <= CB2_F32_EPSILON; }

struct f3 {
    float x, y, z;
};
CB2_DI f3 mk3(float x, float y, float z) {
    f3 r;
    r.x = x;
    r.y = y;
    r.z = z;
    return r;
}
CB2_DI f3 zero3() { return mk3(0.0f, 0.0f, 0.0f); }
CB2_DI f3 operator+(f3 a, f3 b) { return mk3(fadd(a.x, b.x), fadd(a.y, b.y), fadd(a.z, b.z)); }
CB2_DI f3 operator-(f3 a, f3 b) { return mk3(fsub(a.x, b.x), fsub(a.y, b.y), fsub(a.z, b.z)); }
CB2_DI f3 operator*(f3 a, float s) { return mk3(fmul(a.x, s), fmul(a.y, s), fmul(a.z, s)); }
CB2_DI f3 operator/(f3 a, float s) { return mk3(fdiv(a.x, s), fdiv(a.y, s), fdiv(a.z, s)); }
CB2_DI f3 operator-(f3 a) { return mk3(-a.x, -a.y, -a.z); }
CB2_DI float dot(f3 a, f3 b) { return fadd(fadd(fmul(a.x, b.x), fmul(a.y, b.y)), fmul(a.z, b.z)); }
CB2_DI f3 cross(f3 a, f3 b) {
    return mk3(fsub(fmul(a.y, b.z), fmul(a.z, b.y)), fsub(fmul(a.z, b.x), fmul(a.x, b.z)),
               fsub(fmul(a.x, b.y), fmul(a.y, b.x)));
}
CB2_DI float magnitude2(f3 a) { return dot(a, a); }
CB2_DI float magnitude(f3 a) { return fsqrt(magnitude2(a)); }
CB2_DI f3 normalize_to(f3 a, float m) { return a * fdiv(m, magnitude(a)); }
CB2_DI f3 normalize(f3 a) { return normalize_to(a, 1.0f); }
CB2_DI bool ulps_eq_zero(f3 a) {
    return ulps_eq_zero(a.x) && ulps_eq_zero(a.y) && ulps_eq_zero(a.z);
}
CB2_DI f3 lerp(f3 a, f3 b, float t) { return a + ((b - a) * t); }
CB2_DI f3 cross_aba(f3 a, f3 b) { return cross(cross(a, b), a); }  // simplex3d.rs:172-177

struct quat {
    float s, x, y, z;
};
CB2_DI f3 rotate(const quat& q, f3 v) {
    const f3 qv = mk3(q.x, q.y, q.z);
    f3 tmp = cross(qv, v) + (v * q.s);
    return (cross(qv, tmp) * 2.0f) + v;
}
CB2_DI quat invert(const quat& q) {
    float m2 = fadd(fmul(q.s, q.s), dot(mk3(q.x, q.y, q.z), mk3(q.x, q.y, q.z)));
    quat r;
    r.s = fdiv(q.s, m2);
    r.x = fdiv(-q.x, m2);
    r.y = fdiv(-q.y, m2);
    r.z = fdiv(-q.z, m2);
    return r;
}

// Decomposed<Vector3<f32>, Quaternion<f32>> plus per-pair invariants.  The inverse
// quaternion is the same value on every inverse_transform_vector call of a pair
// (cgmath recomputes it each time), so computing it once is bit-identical.
// unit == (scale == 1.0f): x/1 and x*1 are exact identities in IEEE-754, so the
// divides/multiplies by scale are skipped without changing any bit.
struct xform {
    quat q;
    quat qi;
    f3 disp;
    float scale;
    bool unit;
};
// raw = the 8 floats of a cb2_xform {scale, qs, qx, qy, qz, dx, dy, dz} as two float4
CB2_DI xform make_xform(float4 a, float4 b) {
    xform t;
    t.scale = a.x;
    t.q.s = a.y;
    t.q.x = a.z;
    t.q.y = a.w;
    t.q.z = b.x;
    t.disp = mk3(b.y, b.z, b.w);
    t.qi = invert(t.q);
    t.unit = (t.scale == 1.0f);
    return t;
}
CB2_DI f3 transform_point(const xform& t, f3 p) {
    if (!t.unit) p = p * t.scale;
    return rotate(t.q, p) + t.disp;
}
// caller has already checked !ulps_eq_zero(scale) (CB2_PANIC_INV_SCALE otherwise)
CB2_DI f3 inverse_transform_vector(const xform& t, f3 v) {
    if (!t.unit) v = v / t.scale;
    return rotate(t.qi, v);
}

}  // namespace cb2
