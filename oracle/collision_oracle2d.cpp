// collision_oracle2d.cpp — CPU ORACLE for the 2-D narrow phase.  TEST INFRASTRUCTURE ONLY.
//
// Scalar-f32 restatement of GJK2 / EPA2 (rustgd/collision-rs 0.20.1): the generic GJK of
// algorithm/minkowski/gjk/mod.rs instantiated with SimplexProcessor2 (gjk/simplex/simplex2d.rs) and
// EPA2 (epa/epa2d.rs) over Primitive2 shapes and Decomposed<Vector2<f32>, Basis2<f32>> transforms.
// Same rules as collision_oracle.cpp: only tests/, __graft_entry__.smoke() and bench.py's CPU legs
// may load it; the product never does.
//
// Parity status: no oracle/_ref (no Rust toolchain).  PINNED by the reference's own 2-D known-answer
// tests (tests/test_oracle2d_golden.py): gjk/mod.rs:623-779 (GJK2 cases), epa/epa2d.rs:166-263,
// gjk/simplex/simplex2d.rs:107-195, primitive/util.rs:160-241, primitive/circle.rs:106-128,
// primitive/polygon.rs:210-243, primitive/line.rs:53-62.  Unpinned by reference tests: non-unit
// scale, the polygon neighbour walk beyond the one 17-gon, distance / time of impact with circles,
// panic paths.
//
// cgmath 0.18.0 formulas restated from its published source (not in /root/reference):
//   Vector2 dot(a,b)     = a.x*b.x + a.y*b.y
//   Matrix2 (column major, c0 c1) * v = (row0 . v, row1 . v) = (c0.x*v.x + c1.x*v.y, c0.y*v.x + c1.y*v.y)
//   Basis2::rotate_vector(v) = mat * v
//   Basis2::invert()     = Matrix2::invert().unwrap(): det = c0.x*c1.y - c1.x*c0.y; det == 0 => None
//                          (panic); else [c0' = (c1.y/det, -c0.y/det), c1' = (-c1.x/det, c0.x/det)]
//   Decomposed: transform_point(p) = rot*(p*scale) + disp;
//               inverse_transform_vector(x) = None if ulps_eq(scale,0) else invert(rot)*(x/scale)
//   normalize_to(v,m) = v * (m / sqrt(dot(v,v)));  lerp(a,b,t) = a + (b - a)*t
//   ulps_eq!(a, b) on f32 (approx: epsilon = f32::EPSILON, max_ulps = 4):
//       |a-b| <= eps, else signs differ => false, else |bits(a) - bits(b)| <= 4
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <thread>
#include <vector>

#include "../include/collision_b200.h"

namespace orc2 {

struct R {
    float f;
    R() : f(0.0f) {}
    R(float x) : f(x) {}
};
static inline R operator+(R a, R b) { return R(a.f + b.f); }
static inline R operator-(R a, R b) { return R(a.f - b.f); }
static inline R operator*(R a, R b) { return R(a.f * b.f); }
static inline R operator/(R a, R b) { return R(a.f / b.f); }
static inline R operator-(R a) { return R(-a.f); }
static inline bool operator>(R a, R b) { return a.f > b.f; }
static inline bool operator<(R a, R b) { return a.f < b.f; }
static inline bool operator>=(R a, R b) { return a.f >= b.f; }
static inline bool operator<=(R a, R b) { return a.f <= b.f; }
static inline bool operator==(R a, R b) { return a.f == b.f; }
static inline R sqrt_ieee(R a) { return R(std::sqrt(a.f)); }

static const float F32_EPSILON = 1.1920929e-07f;
static inline bool ulps_eq_zero(R x) { return std::fabs(x.f) <= F32_EPSILON; }
static inline uint32_t bits(float f) {
    uint32_t u;
    std::memcpy(&u, &f, 4);
    return u;
}
static inline float signum(float x) {  // f32::signum
    if (x != x) return x;
    return std::signbit(x) ? -1.0f : 1.0f;
}
// approx::UlpsEq for f32 with the default epsilon / max_ulps
static bool ulps_eq(float a, float b) {
    if (std::fabs(a - b) <= F32_EPSILON) return true;
    if (!(signum(a) == signum(b))) return false;  // NaN signum compares unequal
    const uint32_t ia = bits(a), ib = bits(b);
    return (ia <= ib ? ib - ia : ia - ib) <= 4u;
}

struct V2 {
    R x, y;
    V2() {}
    V2(R a, R b) : x(a), y(b) {}
};
static inline V2 operator+(V2 a, V2 b) { return V2(a.x + b.x, a.y + b.y); }
static inline V2 operator-(V2 a, V2 b) { return V2(a.x - b.x, a.y - b.y); }
static inline V2 operator*(V2 a, R s) { return V2(a.x * s, a.y * s); }
static inline V2 operator/(V2 a, R s) { return V2(a.x / s, a.y / s); }
static inline V2 operator-(V2 a) { return V2(-a.x, -a.y); }
static inline R dot(V2 a, V2 b) { return a.x * b.x + a.y * b.y; }
static inline R magnitude2(V2 a) { return dot(a, a); }
static inline R magnitude(V2 a) { return sqrt_ieee(magnitude2(a)); }
static inline V2 normalize_to(V2 a, R m) { return a * (m / magnitude(a)); }
static inline V2 normalize(V2 a) { return normalize_to(a, R(1.0f)); }
static inline bool ulps_eq_zero(V2 a) { return ulps_eq_zero(a.x) && ulps_eq_zero(a.y); }
static inline V2 lerp(V2 a, V2 b, R t) { return a + ((b - a) * t); }
static inline V2 zero2() { return V2(R(0.0f), R(0.0f)); }

struct Panic {
    uint8_t status;
};

struct Xf {  // Decomposed<Vector2<f32>, Basis2<f32>>
    R scale;
    V2 c0, c1;
    V2 disp;
};
static inline Xf load_xf(const cb2_xform2& t) {
    Xf x;
    x.scale = R(t.scale);
    x.c0 = V2(R(t.c0x), R(t.c0y));
    x.c1 = V2(R(t.c1x), R(t.c1y));
    x.disp = V2(R(t.dx), R(t.dy));
    return x;
}
static inline V2 mat_mul(V2 c0, V2 c1, V2 v) {
    return V2(dot(V2(c0.x, c1.x), v), dot(V2(c0.y, c1.y), v));
}
static inline V2 transform_point(const Xf& t, V2 p) { return mat_mul(t.c0, t.c1, p * t.scale) + t.disp; }
static V2 inverse_transform_vector_unwrap(const Xf& t, V2 v) {
    if (ulps_eq_zero(t.scale)) throw Panic{CB2_PANIC_INV_SCALE};
    const R det = t.c0.x * t.c1.y - t.c1.x * t.c0.y;
    if (det == R(0.0f)) throw Panic{(uint8_t)CB2_PANIC_INV_ROT};
    const V2 i0(t.c1.y / det, -t.c0.y / det);
    const V2 i1(-t.c1.x / det, t.c0.x / det);
    return mat_mul(i0, i1, v / t.scale);
}
static inline Xf translation_interpolate(const Xf& self, const Xf& other, R amount) {  // traits.rs:233-246
    Xf r;
    r.disp = lerp(self.disp, other.disp, amount);
    r.c0 = other.c0;
    r.c1 = other.c1;
    r.scale = other.scale;
    return r;
}

// ---- Primitive2 support points ------------------------------------------------------------------
struct Shape {
    uint32_t kind;
    R p[4];
    const float* verts;  // x,y pairs
    uint32_t vcnt;
};
static Shape load_shape(const cb2_shape2& s, const float* verts_xy) {
    Shape o;
    o.kind = s.kind;
    for (int i = 0; i < 4; ++i) o.p[i] = R(s.p[i]);
    o.verts = verts_xy ? verts_xy + 2 * (size_t)s.vtx_off : nullptr;
    o.vcnt = s.vtx_cnt;
    return o;
}
static inline V2 vtx(const Shape& s, uint32_t i) { return V2(R(s.verts[2 * i]), R(s.verts[2 * i + 1])); }

// util.rs:11-30 over an explicit list
static V2 get_max_point(const V2* pts, uint32_t n, V2 direction, const Xf& t) {
    const V2 d = inverse_transform_vector_unwrap(t, direction);
    V2 best = zero2();
    R best_dot(-std::numeric_limits<float>::infinity());
    for (uint32_t i = 0; i < n; ++i) {
        const R dt = dot(pts[i], d);
        if (dt > best_dot) {
            best = pts[i];
            best_dot = dt;
        }
    }
    return transform_point(t, best);
}

// polygon.rs:170-184
static inline R dot_index(const Shape& s, int32_t index, V2 d) {
    const uint32_t iu = (uint32_t)index;
    if (iu == s.vcnt) return dot(vtx(s, 0), d);
    if (index == -1) return dot(vtx(s, s.vcnt - 1), d);
    return dot(vtx(s, iu), d);
}
// polygon.rs:122-168: the neighbour walk used from 10 vertices up.  Reproduced INCLUDING its quirk:
// after choosing a direction the reference pairs `current_dot = dot(start)` with
// `index = start + add`, so the first neighbour's dot is never evaluated and the walk can stop on a
// vertex that is not the extreme one (tests/test_oracle_properties.py shows it happening and checks
// this restatement against an independent re-reading of the Rust).
static V2 polygon_walk(const Shape& s, V2 direction, const Xf& t) {
    const V2 d = inverse_transform_vector_unwrap(t, direction);
    const int32_t len = (int32_t)s.vcnt;
    int32_t start_index = 0;
    R max_dot = dot(vtx(s, 0), d);
    if (max_dot < R(0.0f)) {
        start_index = len / 2;
        max_dot = dot_index(s, start_index, d);
    }
    const R left_dot = dot_index(s, start_index - 1, d);
    const R right_dot = dot_index(s, start_index + 1, d);
    V2 p;
    if (start_index == 0 && max_dot > left_dot && max_dot > right_dot) {
        p = vtx(s, 0);
    } else {
        int32_t add = 1;
        R previous_dot = left_dot;
        if (left_dot > max_dot && left_dot > right_dot) {
            add = -1;
            previous_dot = right_dot;
        }
        int32_t index = start_index + add;
        R current_dot = max_dot;
        if (index == len) index = 0;
        if (index == -1) index = len - 1;
        while (index != start_index) {
            const R next_dot = dot_index(s, index + add, d);
            if (current_dot > previous_dot && current_dot > next_dot) break;
            previous_dot = current_dot;
            current_dot = next_dot;
            index += add;
            if (index == len) index = 0;
            if (index == -1) index = len - 1;
        }
        p = vtx(s, (uint32_t)index);
    }
    return transform_point(t, p);
}

static V2 support_point(const Shape& s, V2 direction, const Xf& t) {  // primitive2.rs:85-103
    switch (s.kind) {
        case CB2_PARTICLE2:
            return transform_point(t, zero2());
        case CB2_LINE2: {  // line.rs:13-25
            const V2 d = inverse_transform_vector_unwrap(t, direction);
            const V2 origin(s.p[0], s.p[1]), dest(s.p[2], s.p[3]);
            const R tt = dot(d, dest - origin);
            return (tt >= R(0.0f)) ? transform_point(t, dest) : transform_point(t, origin);
        }
        case CB2_CIRCLE: {  // circle.rs:34-40
            const V2 d = inverse_transform_vector_unwrap(t, direction);
            return transform_point(t, normalize_to(d, s.p[0]));
        }
        case CB2_RECTANGLE:
        case CB2_SQUARE: {  // rectangle.rs:29-60, 71-76, 130-134
            const V2 dim = s.kind == CB2_SQUARE ? V2(s.p[0], s.p[0]) : V2(s.p[0], s.p[1]);
            const V2 h = dim / (R(1.0f) + R(1.0f));
            const V2 c[4] = {V2(h.x, h.y), V2(-h.x, h.y), V2(-h.x, -h.y), V2(h.x, -h.y)};
            return get_max_point(c, 4, direction, t);
        }
        default: {  // CB2_POLYGON, polygon.rs:33-42
            if (s.vcnt < 10) {
                V2 pts[10];
                for (uint32_t i = 0; i < s.vcnt; ++i) pts[i] = vtx(s, i);
                return get_max_point(pts, s.vcnt, direction, t);
            }
            return polygon_walk(s, direction, t);
        }
    }
}

// ---- SupportPoint, minkowski/mod.rs:16-76 -------------------------------------------------------
struct Sup {
    V2 v, sup_a;
};
static Sup from_minkowski(const Shape& l, const Xf& lt, const Shape& r, const Xf& rt, V2 dir) {
    Sup s;
    const V2 a = support_point(l, dir, lt);
    const V2 b = support_point(r, -dir, rt);
    s.v = a - b;
    s.sup_a = a;
    return s;
}

typedef std::vector<Sup> Simplex;  // SmallVec<[SupportPoint; 5]>: same sequence semantics

// util.rs:42-49
static inline V2 triple_product(V2 a, V2 b, V2 c) {
    const R ac = a.x * c.x + a.y * c.y;
    const R bc = b.x * c.x + b.y * c.y;
    return V2(b.x * ac - a.x * bc, b.y * ac - a.y * bc);
}

// SimplexProcessor2::reduce_to_closest_feature, simplex2d.rs:24-66
static bool reduce_to_closest_feature(Simplex& simplex, V2& d) {
    if (simplex.size() == 3) {
        const V2 a = simplex[2].v, b = simplex[1].v, c = simplex[0].v;
        const V2 ao = -a, ab = b - a, ac = c - a;
        const V2 abp = triple_product(ac, ab, ab);
        if (dot(abp, ao) > R(0.0f)) {
            simplex.erase(simplex.begin());
            d = abp;
        } else {
            const V2 acp = triple_product(ab, ac, ac);
            if (dot(acp, ao) > R(0.0f)) {
                simplex.erase(simplex.begin() + 1);
                d = acp;
            } else {
                return true;
            }
        }
    } else if (simplex.size() == 2) {
        const V2 a = simplex[1].v, b = simplex[0].v;
        const V2 ao = -a, ab = b - a;
        d = triple_product(ab, ao, ab);
        if (ulps_eq_zero(d)) d = V2(-ab.y, ab.x);
    }
    return false;
}

// util.rs:98-114
static V2 get_closest_point_on_edge(V2 start, V2 end, V2 point) {
    const V2 line = end - start;
    const V2 line_dir = normalize(line);
    const V2 v = point - start;
    const R d = dot(v, line_dir);
    if (d < R(0.0f)) return start;
    if ((d * d) > magnitude2(line)) return end;
    return start + line_dir * d;
}

// simplex2d.rs:71-89
static V2 get_closest_point_to_origin(Simplex& simplex) {
    V2 d = zero2();
    if (reduce_to_closest_feature(simplex, d)) return d;
    if (simplex.size() == 1) return simplex[0].v;
    return get_closest_point_on_edge(simplex[1].v, simplex[0].v, zero2());
}

// ---- EPA2, epa2d.rs -----------------------------------------------------------------------------
struct Edge {
    V2 normal;
    R distance;
    size_t index;
};
// :136-157.  false = None
static bool closest_edge(const std::vector<Sup>& simplex, Edge& edge) {
    if (simplex.size() < 3) return false;
    edge.normal = zero2();
    edge.distance = R(std::numeric_limits<float>::infinity());
    edge.index = 0;
    for (size_t i = 0; i < simplex.size(); ++i) {
        const size_t j = (i + 1 == simplex.size()) ? 0 : i + 1;
        const V2 a = simplex[i].v, b = simplex[j].v;
        const V2 e = b - a;
        const V2 n = normalize(triple_product(e, a, e));
        const R d = dot(n, a);
        if (d < edge.distance) {
            edge.normal = n;
            edge.distance = d;
            edge.index = j;
        }
    }
    // assert_ulps_ne!(S::infinity(), edge.distance)
    if (ulps_eq(std::numeric_limits<float>::infinity(), edge.distance.f))
        throw Panic{(uint8_t)CB2_PANIC_EPA2_EDGE};
    return true;
}

// :87-110
static V2 epa_point(const std::vector<Sup>& simplex, const Edge& edge) {
    const Sup& b = simplex[edge.index];
    const Sup& a = edge.index == 0 ? simplex[simplex.size() - 1] : simplex[edge.index - 1];
    const V2 oa = -a.v;
    const V2 ab = b.v - a.v;
    const R t = dot(oa, ab) / magnitude2(ab);
    if (t < R(0.0f)) return a.sup_a;
    if (t > R(1.0f)) return b.sup_a;
    return a.sup_a + (b.sup_a - a.sup_a) * t;
}

static void contact_zero(cb2_contact2* c, uint32_t strategy) {
    std::memset(c, 0, sizeof(*c));
    c->strategy = strategy;
}

// EPA2::process, :27-67.  false = None
static bool epa_process(std::vector<Sup>& simplex, const Shape& l, const Xf& lt, const Shape& r,
                        const Xf& rt, R tolerance, uint32_t max_iterations, cb2_contact2* out,
                        uint32_t* iters) {
    Edge e;
    if (!closest_edge(simplex, e)) return false;
    uint32_t it = 0;
    for (uint32_t k = 0; k < max_iterations; ++k) {
        ++it;
        const Sup p = from_minkowski(l, lt, r, rt, e.normal);
        const R d = dot(p.v, e.normal);
        if (d - e.distance < tolerance) break;
        simplex.insert(simplex.begin() + e.index, p);
        closest_edge(simplex, e);
    }
    if (iters) *iters = it;
    const V2 pt = epa_point(simplex, e);
    out->strategy = CB2_FULL_RESOLUTION;
    out->normal[0] = e.normal.x.f;
    out->normal[1] = e.normal.y.f;
    out->penetration_depth = e.distance.f;
    out->contact_point[0] = pt.x.f;
    out->contact_point[1] = pt.y.f;
    out->time_of_impact = 0.0f;
    return true;
}

// ---- GJK, gjk/mod.rs ----------------------------------------------------------------------------
static bool gjk_intersect(const Shape& l, const Xf& lt, const Shape& r, const Xf& rt,
                          uint32_t max_iterations, Simplex& simplex, uint32_t* iters) {  // :101-146
    const V2 left_pos = transform_point(lt, zero2());
    const V2 right_pos = transform_point(rt, zero2());
    V2 d = right_pos - left_pos;
    if (ulps_eq_zero(d)) d = V2(R(1.0f), R(1.0f));
    const Sup a = from_minkowski(l, lt, r, rt, d);
    if (iters) *iters = 0;
    if (dot(a.v, d) <= R(0.0f)) return false;
    simplex.clear();
    simplex.push_back(a);
    d = -d;
    for (uint32_t it = 0; it < max_iterations; ++it) {
        if (iters) *iters = it + 1;
        const Sup a2 = from_minkowski(l, lt, r, rt, d);
        if (dot(a2.v, d) <= R(0.0f)) return false;
        simplex.push_back(a2);
        if (reduce_to_closest_feature(simplex, d)) return true;
    }
    return false;
}

static uint8_t gjk_intersection(uint32_t strategy, const cb2_settings& st, const Shape& l,
                                const Xf& lt, const Shape& r, const Xf& rt, cb2_contact2* out,
                                uint32_t* gjk_iters, uint32_t* epa_iters) {  // :362-391
    contact_zero(out, strategy);
    try {
        Simplex simplex;
        if (!gjk_intersect(l, lt, r, rt, st.max_iterations, simplex, gjk_iters)) return CB2_NONE;
        if (strategy == CB2_COLLISION_ONLY) return CB2_SOME;
        if (!epa_process(simplex, l, lt, r, rt, R(st.epa_tolerance), st.max_iterations, out,
                         epa_iters)) {
            contact_zero(out, strategy);
            return CB2_NONE;
        }
        return CB2_SOME;
    } catch (Panic& p) {
        contact_zero(out, strategy);
        return p.status;
    }
}

static uint8_t gjk_distance(const cb2_settings& st, const Shape& l, const Xf& lt, const Shape& r,
                            const Xf& rt, float* out, uint32_t* iters) {  // :271-321
    *out = 0.0f;
    if (iters) *iters = 0;
    try {
        const V2 right_pos = transform_point(rt, zero2());
        const V2 left_pos = transform_point(lt, zero2());
        Simplex simplex;
        V2 d = right_pos - left_pos;
        if (ulps_eq_zero(d)) d = V2(R(1.0f), R(1.0f));
        simplex.push_back(from_minkowski(l, lt, r, rt, d));
        simplex.push_back(from_minkowski(l, lt, r, rt, -d));
        for (uint32_t it = 0; it < st.max_iterations; ++it) {
            if (iters) *iters = it + 1;
            V2 dd = get_closest_point_to_origin(simplex);
            if (ulps_eq_zero(dd)) return CB2_NONE;
            dd = -dd;
            const Sup p = from_minkowski(l, lt, r, rt, dd);
            const R dp = dot(p.v, dd);
            const R d0 = dot(simplex[0].v, dd);
            if (dp - d0 < R(st.distance_tolerance)) {
                *out = magnitude(dd).f;
                return CB2_SOME;
            }
            simplex.push_back(p);
        }
        return CB2_NONE;
    } catch (Panic& p) {
        return p.status;
    }
}

static uint8_t gjk_toi(const cb2_settings& st, const Shape& l, const Xf& lt0, const Xf& lt1,
                       const Shape& r, const Xf& rt0, const Xf& rt1, uint32_t iter_cap,
                       cb2_contact2* out, uint32_t* iters) {  // :163-256 (+ iter_cap)
    contact_zero(out, CB2_FULL_RESOLUTION);
    if (iters) *iters = 0;
    try {
        const V2 left_lin_vel = transform_point(lt1, zero2()) - transform_point(lt0, zero2());
        const V2 right_lin_vel = transform_point(rt1, zero2()) - transform_point(rt0, zero2());
        const V2 ray = right_lin_vel - left_lin_vel;
        R lambda(0.0f);
        V2 normal = zero2();
        V2 ray_origin = zero2();
        Simplex simplex;
        const Sup p0 = from_minkowski(l, lt0, r, rt0, -ray);
        V2 v = p0.v;
        uint32_t it = 0;
        while (magnitude2(v) > R(st.continuous_tolerance)) {
            if (it >= iter_cap) return CB2_NONCONVERGED;
            ++it;
            if (iters) *iters = it;
            const Sup p = from_minkowski(l, lt0, r, rt0, -v);
            const R vp = dot(v, p.v);
            const R vr = dot(v, ray);
            if (vp > lambda * vr) {
                if (vr > R(0.0f)) {
                    lambda = vp / vr;
                    if (lambda > R(1.0f)) return CB2_NONE;
                    ray_origin = ray * lambda;
                    simplex.clear();
                    normal = -v;
                } else {
                    return CB2_NONE;
                }
            }
            Sup q = p;
            q.v = p.v - ray_origin;
            simplex.push_back(q);
            v = get_closest_point_to_origin(simplex);
        }
        if (magnitude2(v) <= R(st.continuous_tolerance)) {
            const Xf t = translation_interpolate(rt0, rt1, lambda);
            const V2 n = -normalize(normal);
            const V2 cp = transform_point(t, ray_origin);
            out->strategy = CB2_FULL_RESOLUTION;
            out->normal[0] = n.x.f;
            out->normal[1] = n.y.f;
            out->penetration_depth = magnitude(v).f;
            out->contact_point[0] = cp.x.f;
            out->contact_point[1] = cp.y.f;
            out->time_of_impact = lambda.f;
            return CB2_SOME;
        }
        return CB2_NONE;
    } catch (Panic& p) {
        contact_zero(out, CB2_FULL_RESOLUTION);
        return p.status;
    }
}

// ---- bounds: ComputeBound<Aabb2> for Primitive2 (primitive2.rs:68-80) + Aabb2::transform ------------
static inline R rmin(R l, R r) { return (l > r) ? r : l; }  // volume/aabb/mod.rs:21-26: Less|Equal|None => lhs
static inline R rmax(R l, R r) { return (l < r) ? r : l; }  // :28-33
struct Box2 {
    V2 mn, mx;
};
static inline Box2 box_new(V2 p1, V2 p2) {  // aabb2.rs:30-35
    Box2 b;
    b.mn = V2(rmin(p1.x, p2.x), rmin(p1.y, p2.y));
    b.mx = V2(rmax(p1.x, p2.x), rmax(p1.y, p2.y));
    return b;
}
static inline Box2 box_grow(const Box2& a, V2 p) {  // aabb/mod.rs:116-118: new(min(self.min, p), max(self.max, p))
    return box_new(V2(rmin(a.mn.x, p.x), rmin(a.mn.y, p.y)), V2(rmax(a.mx.x, p.x), rmax(a.mx.y, p.y)));
}
static Box2 compute_bound(const Shape& s) {
    switch (s.kind) {
        case CB2_PARTICLE2:
            return box_new(zero2(), zero2());  // Aabb2::zero(), primitive2.rs:70
        case CB2_LINE2:
            return box_new(V2(s.p[0], s.p[1]), V2(s.p[2], s.p[3]));  // primitive/line.rs:32-34
        case CB2_CIRCLE:
            return box_new(V2(-s.p[0], -s.p[0]), V2(s.p[0], s.p[0]));  // circle.rs:47-52
        case CB2_RECTANGLE:
        case CB2_SQUARE: {  // rectangle.rs:83-88
            const V2 dim = s.kind == CB2_SQUARE ? V2(s.p[0], s.p[0]) : V2(s.p[0], s.p[1]);
            const V2 h = dim / (R(1.0f) + R(1.0f));
            return box_new(-h, h);
        }
        default: {  // polygon.rs:50-52: get_bound = fold(Aabb2::zero(), grow), util.rs:32-38
            Box2 b = box_new(zero2(), zero2());
            for (uint32_t i = 0; i < s.vcnt; ++i) b = box_grow(b, vtx(s, i));
            return b;
        }
    }
}
static Box2 box_transform(const Box2& a, const Xf& t) {  // aabb2.rs:39-46, 78-88
    const V2 c[4] = {a.mn, V2(a.mx.x, a.mn.y), V2(a.mn.x, a.mx.y), a.mx};
    const V2 first = transform_point(t, c[0]);
    Box2 b = box_new(first, first);
    for (int i = 1; i < 4; ++i) b = box_grow(b, transform_point(t, c[i]));
    return b;
}

// ---- composite shapes: gjk/mod.rs:412-588 with the 2-D instantiation ---------------------------------
// Decomposed::concat (cgmath 0.18 transform.rs): {scale: s*os, rot: rot*orot, disp: rot.rotate(odisp*s) + disp};
// Basis2 * Basis2 = Matrix2 * Matrix2 = [lhs * rhs.c0, lhs * rhs.c1] (each a row-dot matrix-vector product)
static inline Xf concat(const Xf& self, const Xf& other) {
    Xf r;
    r.scale = self.scale * other.scale;
    r.c0 = mat_mul(self.c0, self.c1, other.c0);
    r.c1 = mat_mul(self.c0, self.c1, other.c1);
    r.disp = mat_mul(self.c0, self.c1, other.disp * self.scale) + self.disp;
    return r;
}
static inline int partial_cmp(float a, float b) {  // -1 Less, 0 Equal, 1 Greater, 2 None
    if (a < b) return -1;
    if (a > b) return 1;
    if (a == b) return 0;
    return 2;
}
struct Composite {  // &[(Primitive, local transform)]
    const uint32_t* shape;
    const cb2_xform2* local;
    uint32_t n;
};
#define CB2_PANIC_CMP_NAN 8  // max_by(partial_cmp(..).unwrap()) on a NaN depth, gjk/mod.rs:455-459

// GJK::intersection_complex, :412-461
static uint8_t gjk_intersection_complex(uint32_t strategy, const cb2_settings& st, const cb2_shape2* shapes,
                                        const float* verts, const Composite& L, const Xf& lt,
                                        const Composite& Rc, const Xf& rt, cb2_contact2* out) {
    contact_zero(out, strategy);
    bool have = false;
    cb2_contact2 best;
    for (uint32_t a = 0; a < L.n; ++a) {
        const Xf left_transform = concat(lt, load_xf(L.local[a]));
        for (uint32_t b = 0; b < Rc.n; ++b) {
            const Xf right_transform = concat(rt, load_xf(Rc.local[b]));
            cb2_contact2 c;
            const uint8_t s = gjk_intersection(strategy, st, load_shape(shapes[L.shape[a]], verts), left_transform,
                                               load_shape(shapes[Rc.shape[b]], verts), right_transform, &c,
                                               nullptr, nullptr);
            if (s == CB2_NONE) continue;
            if (s != CB2_SOME) return s;  // the reference panics right here
            if (strategy == CB2_COLLISION_ONLY) {
                *out = c;
                return CB2_SOME;
            }
            if (!have) {
                best = c;
                have = true;
            } else {
                // Iterator::max_by: compare(best, c) == Greater keeps best, otherwise takes c
                const int o = partial_cmp(best.penetration_depth, c.penetration_depth);
                if (o == 2) return CB2_PANIC_CMP_NAN;
                if (o != 1) best = c;
            }
        }
    }
    if (!have) return CB2_NONE;
    *out = best;
    return CB2_SOME;
}

// GJK::distance_complex, :477-516
static uint8_t gjk_distance_complex(const cb2_settings& st, const cb2_shape2* shapes, const float* verts,
                                    const Composite& L, const Xf& lt, const Composite& Rc, const Xf& rt,
                                    float* out) {
    *out = 0.0f;
    bool have = false;
    float m = 0.0f;
    for (uint32_t a = 0; a < L.n; ++a) {
        const Xf left_transform = concat(lt, load_xf(L.local[a]));
        for (uint32_t b = 0; b < Rc.n; ++b) {
            const Xf right_transform = concat(rt, load_xf(Rc.local[b]));
            float d;
            const uint8_t s = gjk_distance(st, load_shape(shapes[L.shape[a]], verts), left_transform,
                                           load_shape(shapes[Rc.shape[b]], verts), right_transform, &d, nullptr);
            if (s != CB2_SOME) return s;  // None => return None (colliding); panic => panic
            m = have ? std::fmin(d, m) : d;  // f32::min: a NaN operand yields the other one
            have = true;
        }
    }
    if (!have) return CB2_NONE;
    *out = m;
    return CB2_SOME;
}

// GJK::intersection_complex_time_of_impact, :538-588
static uint8_t gjk_toi_complex(uint32_t strategy, const cb2_settings& st, const cb2_shape2* shapes,
                               const float* verts, const Composite& L, const Xf& lt0, const Xf& lt1,
                               const Composite& Rc, const Xf& rt0, const Xf& rt1, uint32_t iter_cap,
                               cb2_contact2* out) {
    contact_zero(out, strategy);
    bool have = false;
    cb2_contact2 best;
    for (uint32_t a = 0; a < L.n; ++a) {
        const Xf l0 = concat(lt0, load_xf(L.local[a]));
        const Xf l1 = concat(lt1, load_xf(L.local[a]));
        for (uint32_t b = 0; b < Rc.n; ++b) {
            const Xf r0 = concat(rt0, load_xf(Rc.local[b]));
            const Xf r1 = concat(rt1, load_xf(Rc.local[b]));
            cb2_contact2 c;
            const uint8_t s = gjk_toi(st, load_shape(shapes[L.shape[a]], verts), l0, l1,
                                      load_shape(shapes[Rc.shape[b]], verts), r0, r1, iter_cap, &c, nullptr);
            if (s == CB2_NONE) continue;
            if (s != CB2_SOME) return s;
            if (strategy == CB2_COLLISION_ONLY) {
                c.strategy = CB2_COLLISION_ONLY;
                *out = c;
                return CB2_SOME;
            }
            if (!have) {
                best = c;
                have = true;
            } else {
                // Iterator::min_by with unwrap_or(Equal): only Greater replaces the current best
                if (partial_cmp(best.time_of_impact, c.time_of_impact) == 1) best = c;
            }
        }
    }
    if (!have) return CB2_NONE;
    *out = best;
    return CB2_SOME;
}

template <typename F>
static void parallel_for(uint64_t n, int nthreads, F fn) {
    if (nthreads <= 1 || n < 1024) {
        fn((uint64_t)0, n);
        return;
    }
    std::vector<std::thread> th;
    const uint64_t per = (n + nthreads - 1) / nthreads;
    for (int t = 0; t < nthreads; ++t) {
        const uint64_t b = (uint64_t)t * per, e = b + per < n ? b + per : n;
        if (b >= e) break;
        th.emplace_back([=] { fn(b, e); });
    }
    for (auto& x : th) x.join();
}

}  // namespace orc2

using namespace orc2;

extern "C" {

// ---- unit-level hooks for the golden tests -------------------------------------------------------
void orc2_support_point(const cb2_shape2* shape, const float* verts_xy, const cb2_xform2* t,
                        const float* dir, float* out, uint8_t* status) {
    *status = CB2_SOME;
    try {
        const V2 p = support_point(load_shape(*shape, verts_xy), V2(R(dir[0]), R(dir[1])), load_xf(*t));
        out[0] = p.x.f;
        out[1] = p.y.f;
    } catch (Panic& p) {
        *status = p.status;
    }
}
static Simplex simplex_from(const float* pts, uint32_t n) {
    Simplex s;
    for (uint32_t i = 0; i < n; ++i) {
        Sup p;
        p.v = V2(R(pts[2 * i]), R(pts[2 * i + 1]));
        p.sup_a = zero2();
        s.push_back(p);
    }
    return s;
}
static void simplex_to(const Simplex& s, float* pts, uint32_t* n) {
    *n = (uint32_t)s.size();
    for (size_t i = 0; i < s.size(); ++i) {
        pts[2 * i] = s[i].v.x.f;
        pts[2 * i + 1] = s[i].v.y.f;
    }
}
int orc2_reduce_to_closest_feature(float* pts, uint32_t* n, float* d) {
    Simplex s = simplex_from(pts, *n);
    V2 dd = V2(R(d[0]), R(d[1]));
    const bool hit = reduce_to_closest_feature(s, dd);
    simplex_to(s, pts, n);
    d[0] = dd.x.f;
    d[1] = dd.y.f;
    return hit ? 1 : 0;
}
void orc2_closest_point_to_origin(float* pts, uint32_t* n, float* out) {
    Simplex s = simplex_from(pts, *n);
    const V2 p = get_closest_point_to_origin(s);
    simplex_to(s, pts, n);
    out[0] = p.x.f;
    out[1] = p.y.f;
}
void orc2_closest_point_on_edge(const float* start, const float* end, const float* point, float* out) {
    const V2 p = get_closest_point_on_edge(V2(R(start[0]), R(start[1])), V2(R(end[0]), R(end[1])),
                                           V2(R(point[0]), R(point[1])));
    out[0] = p.x.f;
    out[1] = p.y.f;
}
// returns 0 None, 1 Some, or the panic status; out = normal.xy, distance; *index
uint8_t orc2_closest_edge(const float* pts, uint32_t n, float* out, uint32_t* index) {
    try {
        Edge e;
        if (!closest_edge(simplex_from(pts, n), e)) return CB2_NONE;
        out[0] = e.normal.x.f;
        out[1] = e.normal.y.f;
        out[2] = e.distance.f;
        *index = (uint32_t)e.index;
        return CB2_SOME;
    } catch (Panic& p) {
        return p.status;
    }
}
uint8_t orc2_epa2_process(const float* simplex_v, uint32_t n, const cb2_shape2* l, const cb2_xform2* lt,
                          const cb2_shape2* r, const cb2_xform2* rt, const float* verts_xy,
                          const cb2_settings* st, cb2_contact2* out) {
    contact_zero(out, CB2_FULL_RESOLUTION);
    try {
        std::vector<Sup> s = simplex_from(simplex_v, n);
        return epa_process(s, load_shape(*l, verts_xy), load_xf(*lt), load_shape(*r, verts_xy),
                           load_xf(*rt), R(st->epa_tolerance), st->max_iterations, out, nullptr)
                   ? CB2_SOME
                   : CB2_NONE;
    } catch (Panic& p) {
        return p.status;
    }
}
uint8_t orc2_gjk2_intersect(const cb2_shape2* l, const cb2_xform2* lt, const cb2_shape2* r,
                            const cb2_xform2* rt, const float* verts_xy, const cb2_settings* st,
                            float* simplex_v, uint32_t* n) {
    try {
        Simplex s;
        if (!gjk_intersect(load_shape(*l, verts_xy), load_xf(*lt), load_shape(*r, verts_xy),
                           load_xf(*rt), st->max_iterations, s, nullptr))
            return CB2_NONE;
        simplex_to(s, simplex_v, n);
        return CB2_SOME;
    } catch (Panic& p) {
        return p.status;
    }
}

// ---- batch entry points: same array contract as include/collision_b200.h --------------------------
// iters (optional): per pair u32 x 2 = {gjk iterations, epa iterations}
void orc2_gjk2_intersection_batch(uint32_t strategy, const cb2_settings* st, const cb2_shape2* shapes,
                                  const float* verts_xy, const uint32_t* l_shape,
                                  const uint32_t* r_shape, const cb2_xform2* l_t,
                                  const cb2_xform2* r_t, uint64_t n, cb2_contact2* out,
                                  uint8_t* status, uint32_t* iters, int nthreads) {
    parallel_for(n, nthreads, [=](uint64_t b, uint64_t e) {
        for (uint64_t i = b; i < e; ++i) {
            uint32_t gi = 0, ei = 0;
            status[i] = gjk_intersection(strategy, *st, load_shape(shapes[l_shape[i]], verts_xy),
                                         load_xf(l_t[i]), load_shape(shapes[r_shape[i]], verts_xy),
                                         load_xf(r_t[i]), &out[i], &gi, &ei);
            if (iters) {
                iters[2 * i] = gi;
                iters[2 * i + 1] = ei;
            }
        }
    });
}
void orc2_gjk2_distance_batch(const cb2_settings* st, const cb2_shape2* shapes, const float* verts_xy,
                              const uint32_t* l_shape, const uint32_t* r_shape,
                              const cb2_xform2* l_t, const cb2_xform2* r_t, uint64_t n, float* out,
                              uint8_t* status, uint32_t* iters, int nthreads) {
    parallel_for(n, nthreads, [=](uint64_t b, uint64_t e) {
        for (uint64_t i = b; i < e; ++i) {
            uint32_t it = 0;
            status[i] = gjk_distance(*st, load_shape(shapes[l_shape[i]], verts_xy), load_xf(l_t[i]),
                                     load_shape(shapes[r_shape[i]], verts_xy), load_xf(r_t[i]),
                                     &out[i], &it);
            if (iters) iters[i] = it;
        }
    });
}
void orc2_gjk2_toi_batch(const cb2_settings* st, const cb2_shape2* shapes, const float* verts_xy,
                         const uint32_t* l_shape, const uint32_t* r_shape, const cb2_xform2* l_t0,
                         const cb2_xform2* l_t1, const cb2_xform2* r_t0, const cb2_xform2* r_t1,
                         uint64_t n, uint32_t iter_cap, cb2_contact2* out, uint8_t* status,
                         uint32_t* iters, int nthreads) {
    if (iter_cap == 0) iter_cap = 1024;
    parallel_for(n, nthreads, [=](uint64_t b, uint64_t e) {
        for (uint64_t i = b; i < e; ++i) {
            uint32_t it = 0;
            status[i] = gjk_toi(*st, load_shape(shapes[l_shape[i]], verts_xy), load_xf(l_t0[i]),
                                load_xf(l_t1[i]), load_shape(shapes[r_shape[i]], verts_xy),
                                load_xf(r_t0[i]), load_xf(r_t1[i]), iter_cap, &out[i], &it);
            if (iters) iters[i] = it;
        }
    });
}

// composite batch entry point: comp_off[n_comp+1] indexes comp_shape / comp_local; op 0/1/2
void orc2_gjk2_complex_batch(uint32_t op, uint32_t strategy, const cb2_settings* st, const cb2_shape2* shapes,
                             const float* verts_xy, const uint32_t* comp_off, const uint32_t* comp_shape,
                             const cb2_xform2* comp_local, const uint32_t* l_comp, const uint32_t* r_comp,
                             const cb2_xform2* l_t0, const cb2_xform2* l_t1, const cb2_xform2* r_t0,
                             const cb2_xform2* r_t1, uint64_t n, uint32_t iter_cap, cb2_contact2* out_c,
                             float* out_d, uint8_t* status, int nthreads) {
    if (iter_cap == 0) iter_cap = 1024;
    parallel_for(n, nthreads, [=](uint64_t b, uint64_t e) {
        for (uint64_t i = b; i < e; ++i) {
            const uint32_t lc = l_comp[i], rc = r_comp[i];
            const Composite L = {comp_shape + comp_off[lc], comp_local + comp_off[lc], comp_off[lc + 1] - comp_off[lc]};
            const Composite Rc = {comp_shape + comp_off[rc], comp_local + comp_off[rc], comp_off[rc + 1] - comp_off[rc]};
            if (op == 0)
                status[i] = gjk_intersection_complex(strategy, *st, shapes, verts_xy, L, load_xf(l_t0[i]), Rc,
                                                     load_xf(r_t0[i]), &out_c[i]);
            else if (op == 1)
                status[i] = gjk_distance_complex(*st, shapes, verts_xy, L, load_xf(l_t0[i]), Rc, load_xf(r_t0[i]),
                                                 &out_d[i]);
            else
                status[i] = gjk_toi_complex(strategy, *st, shapes, verts_xy, L, load_xf(l_t0[i]), load_xf(l_t1[i]),
                                            Rc, load_xf(r_t0[i]), load_xf(r_t1[i]), iter_cap, &out_c[i]);
        }
    });
}

// out[i] = shape.compute_bound().transform_volume(&t[i]) as min.xy, max.xy
void orc2_world_aabbs(const cb2_shape2* shapes, const float* verts_xy, const uint32_t* shape, const cb2_xform2* t,
                      uint64_t n, cb2_aabb2* out) {
    for (uint64_t i = 0; i < n; ++i) {
        const Box2 b = box_transform(compute_bound(load_shape(shapes[shape[i]], verts_xy)), load_xf(t[i]));
        out[i].min[0] = b.mn.x.f;
        out[i].min[1] = b.mn.y.f;
        out[i].max[0] = b.mx.x.f;
        out[i].max[1] = b.mx.y.f;
    }
}

}  // extern "C"
