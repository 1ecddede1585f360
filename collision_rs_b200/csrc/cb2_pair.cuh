// cb2_pair.cuh — one pair of the narrow phase in registers: both shapes, both transforms, the
// Minkowski support (SupportPoint::from_minkowski, minkowski/mod.rs:39-60) and GJK::intersect
// (gjk/mod.rs:101-146).  Shared by the thread-per-pair kernels (cb2_narrow.cu), the lane-per-pair
// EPA (cb2_epa_lane.cuh) and the host simulation of the latter (tests/host_sim).
#pragma once
#include "cb2_gjk.cuh"
#include "cb2_kernels.h"

namespace cb2 {

CB2_DI shape_table table_of(const shape_table_args& a) {
    shape_table t;
    t.shapes = a.shapes;
    t.poly.verts = a.verts;
    t.poly.cells = a.cells;
    t.poly.ext = a.ext;
    return t;
}

struct pair_in {
    shape_table T;  // launch-uniform base pointers
    shape ls, rs;
    xform lt, rt;
};

template <bool WITH_A, bool ROLLED = false>
CB2_DI void minkowski(const pair_in& p, f3 dir, f3& v, f3& a) {
    const f3 l = support_point<ROLLED>(p.T, p.ls, p.lt, dir);
    const f3 r = support_point<ROLLED>(p.T, p.rs, p.rt, -dir);
    v = l - r;
    if (WITH_A) a = l;
}

CB2_DI xform load_xform(const float4* __restrict__ t, uint64_t i) {
    const float4 a = __ldg(t + 2 * i);
    const float4 b = __ldg(t + 2 * i + 1);
    return make_xform(a, b);
}

// seed direction shared by intersect and distance (gjk/mod.rs:117-122, :288-294)
CB2_DI f3 seed_direction(const xform& lt, const xform& rt) {
    const f3 left_pos = transform_point(lt, zero3());
    const f3 right_pos = transform_point(rt, zero3());
    f3 d = right_pos - left_pos;
    if (ulps_eq_zero(d)) d = mk3(1.0f, 1.0f, 1.0f);
    return d;
}

// GJK::intersect, gjk/mod.rs:101-146
template <bool WITH_A, bool WITH_B>
CB2_DI bool gjk_intersect(const pair_in& p, uint32_t max_iterations, simplex<WITH_A, WITH_B>& s) {
    f3 d = seed_direction(p.lt, p.rt);
    f3 v, a, b = zero3();
    auto point = [&](f3 dir) {  // SupportPoint::from_minkowski
        a = support_point(p.T, p.ls, p.lt, dir);
        b = support_point(p.T, p.rs, p.rt, -dir);
        v = a - b;
    };
    point(d);
    if (dot(v, d) <= 0.0f) return false;
    s.n = 0;
    s.push(v, a, b);
    d = -d;
    for (uint32_t i = 0; i < max_iterations; ++i) {
        point(d);
        if (dot(v, d) <= 0.0f) return false;
        s.push(v, a, b);
        if (reduce_to_closest_feature(s, d)) return true;
    }
    return false;
}

// Returns 0 or the pair's final status: ST_BAD_INDEX for a shape index outside the table (a Rust
// slice index would panic; nothing is read through it), ST_PANIC_INV_SCALE for a zero scale.
CB2_DI uint8_t load_pair(const narrow_args& A, uint64_t i, const float4* lt, const float4* rt,
                         pair_in& p) {
    uint32_t li, ri;
    uint64_t lx, rx;
    if (A.pair_body) {  // body-indexed pairs (SweepAndPrune3 output feeding the narrow phase)
        const uint2 b = __ldg(reinterpret_cast<const uint2*>(A.pair_body) + i);
        lx = b.x;
        rx = b.y;
        if (A.n_bodies && (lx >= A.n_bodies || rx >= A.n_bodies)) {
            if (A.bad_index) *A.bad_index = 1u;
            return ST_BAD_INDEX;
        }
        li = __ldg(A.l_shape + lx);
        ri = __ldg(A.l_shape + rx);
    } else {
        lx = rx = i;
        li = __ldg(A.l_shape + i);
        ri = __ldg(A.r_shape + i);
    }
    p.T = table_of(A.T);
    if (li >= A.n_shapes || ri >= A.n_shapes) {
        if (A.bad_index) *A.bad_index = 1u;
        return ST_BAD_INDEX;
    }
    p.ls = load_shape(p.T, li);
    p.rs = load_shape(p.T, ri);
    p.lt = load_xform(lt, lx);
    p.rt = load_xform(rt, rx);
    // inverse_transform_vector(..).unwrap() panics when ulps_eq(scale, 0); every entry point
    // reaches a support call before anything else can return (util.rs:18, sphere.rs:36)
    return ((needs_inverse(p.ls.kind) && ulps_eq_zero(p.lt.scale)) ||
            (needs_inverse(p.rs.kind) && ulps_eq_zero(p.rt.scale)))
               ? ST_PANIC_INV_SCALE
               : ST_NONE;
}

}  // namespace cb2
