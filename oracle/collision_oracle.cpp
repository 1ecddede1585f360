// collision_oracle.cpp — CPU ORACLE.  TEST INFRASTRUCTURE ONLY.
//
// A scalar-f32, single-pair-at-a-time restatement of the reference algorithms
// (rustgd/collision-rs 0.20.1) on the GJK3/EPA3 + SweepAndPrune3 path.  It is
// the checker for the CUDA kernels: only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs may load it.  The product
// (collision_rs_b200/csrc) never includes, links or calls anything in oracle/.
//
// Parity status: the Rust crate cannot be built here (no rustc/cargo; cgmath
// 0.18.0 and smallvec 1.6.1 are not vendored), so there is no oracle/_ref.
// The restatement is PINNED by the reference's own known-answer tests
// (tests/test_oracle_golden.py): gjk/mod.rs:634-825, epa/epa3d.rs:230-341,
// gjk/simplex/simplex3d.rs:234-373, primitive/sphere.rs:114-152,
// primitive/util.rs:160-250, tests/aabb.rs:246-283, sweep_prune.rs:317-361.
// Unpinned by reference tests (the restatement is the authority): non-unit
// scale, general rotations end-to-end, SweepAndPrune3, polyhedra inside GJK,
// every panic / cap path.
//
// Arithmetic: every + - * / sqrt is a separate IEEE-754 binary32 operation in
// the reference's order (rustc never contracts to FMA).  Build with
//   g++ -O2 -ffp-contract=off -fno-fast-math   (see oracle/Makefile)
// cgmath 0.18.0 formulas (source not in /root/reference; Cargo.toml:39) are
// restated from its published source:
//   dot(a,b)          = (a.x*b.x + a.y*b.y) + a.z*b.z
//   cross(a,b)        = (a.y*b.z - a.z*b.y, a.z*b.x - a.x*b.z, a.x*b.y - a.y*b.x)
//   normalize_to(v,m) = v * (m / sqrt(dot(v,v)))
//   q * x             = cross(q.v, cross(q.v,x) + x*q.s) * 2 + x
//   invert(q)         = conjugate(q) / (q.s*q.s + dot(q.v,q.v))
//   Decomposed: transform_point(p) = rot*(p*scale) + disp
//               inverse_transform_vector(x) = None if ulps_eq(scale,0)
//                                             else invert(rot) * (x / scale)
//   ulps_eq!(x, 0) (eps = f32::EPSILON, 4 ulps) == |x| <= 1.1920929e-7, NaN => false
//   lerp(a,b,t) = a + (b - a) * t
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <thread>
#include <map>
#include <vector>

#include <map>
#include <mutex>

#include "../include/collision_b200.h"

namespace orc {

// ---- scalar with optional op counting (algorithmic flops: + - * / sqrt) -----
#ifdef COUNT_OPS
static thread_local uint64_t t_ops = 0;
#define OP() (++t_ops)
#else
#define OP() ((void)0)
#endif

struct R {
    float f;
    R() : f(0.0f) {}
    R(float x) : f(x) {}
};
static inline R operator+(R a, R b) { OP(); return R(a.f + b.f); }
static inline R operator-(R a, R b) { OP(); return R(a.f - b.f); }
static inline R operator*(R a, R b) { OP(); return R(a.f * b.f); }
static inline R operator/(R a, R b) { OP(); return R(a.f / b.f); }
static inline R operator-(R a) { return R(-a.f); }
static inline bool operator>(R a, R b) { return a.f > b.f; }
static inline bool operator<(R a, R b) { return a.f < b.f; }
static inline bool operator>=(R a, R b) { return a.f >= b.f; }
static inline bool operator<=(R a, R b) { return a.f <= b.f; }
static inline bool operator==(R a, R b) { return a.f == b.f; }
static inline R sqrt_ieee(R a) { OP(); return R(std::sqrt(a.f)); }  // f32::sqrt

static const float F32_EPSILON = 1.1920929e-7f;
// approx::ulps_eq!(x, 0.0) with default epsilon / max_ulps
static inline bool ulps_eq_zero(R x) { return std::fabs(x.f) <= F32_EPSILON; }

struct V3 {
    R x, y, z;
    V3() {}
    V3(R a, R b, R c) : x(a), y(b), z(c) {}
};
static inline V3 operator+(V3 a, V3 b) { return V3(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline V3 operator-(V3 a, V3 b) { return V3(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline V3 operator*(V3 a, R s) { return V3(a.x * s, a.y * s, a.z * s); }
static inline V3 operator/(V3 a, R s) { return V3(a.x / s, a.y / s, a.z / s); }
static inline V3 operator-(V3 a) { return V3(-a.x, -a.y, -a.z); }
static inline R dot(V3 a, V3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
static inline V3 cross(V3 a, V3 b) {
    return V3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
static inline R magnitude2(V3 a) { return dot(a, a); }
static inline R magnitude(V3 a) { return sqrt_ieee(magnitude2(a)); }
static inline V3 normalize_to(V3 a, R m) { return a * (m / magnitude(a)); }
static inline V3 normalize(V3 a) { return normalize_to(a, R(1.0f)); }
static inline bool ulps_eq_zero(V3 a) {
    return ulps_eq_zero(a.x) && ulps_eq_zero(a.y) && ulps_eq_zero(a.z);
}
static inline V3 lerp(V3 a, V3 b, R t) { return a + ((b - a) * t); }
static inline V3 zero3() { return V3(R(0.0f), R(0.0f), R(0.0f)); }
static inline V3 v3p(const float* p) { return V3(R(p[0]), R(p[1]), R(p[2])); }

struct Quat {
    R s;
    V3 v;
};
static inline V3 rotate(const Quat& q, V3 x) {
    V3 tmp = cross(q.v, x) + (x * q.s);
    return (cross(q.v, tmp) * R(2.0f)) + x;
}
static inline Quat invert(const Quat& q) {
    R m2 = q.s * q.s + dot(q.v, q.v);
    Quat r;
    r.s = q.s / m2;
    r.v = (-q.v) / m2;
    return r;
}

struct Xf {  // Decomposed<Vector3<f32>, Quaternion<f32>>
    R scale;
    Quat rot;
    V3 disp;
};
static inline Xf load_xf(const cb2_xform& t) {
    Xf x;
    x.scale = R(t.scale);
    x.rot.s = R(t.qs);
    x.rot.v = V3(R(t.qx), R(t.qy), R(t.qz));
    x.disp = V3(R(t.dx), R(t.dy), R(t.dz));
    return x;
}
static inline V3 transform_point(const Xf& t, V3 p) { return rotate(t.rot, p * t.scale) + t.disp; }
// returns false for None
static inline bool inverse_transform_vector(const Xf& t, V3 v, V3* out) {
    if (ulps_eq_zero(t.scale)) return false;
    *out = rotate(invert(t.rot), v / t.scale);
    return true;
}
// TranslationInterpolate for Decomposed, traits.rs:233-246
static inline Xf translation_interpolate(const Xf& self, const Xf& other, R amount) {
    Xf r;
    r.disp = lerp(self.disp, other.disp, amount);
    r.rot = other.rot;
    r.scale = other.scale;
    return r;
}

// Quaternion * Quaternion, cgmath 0.18 quaternion.rs (restated; left-to-right evaluation)
static inline Quat quat_mul(const Quat& a, const Quat& b) {
    Quat r;
    r.s = a.s * b.s - a.v.x * b.v.x - a.v.y * b.v.y - a.v.z * b.v.z;
    r.v.x = a.s * b.v.x + a.v.x * b.s + a.v.y * b.v.z - a.v.z * b.v.y;
    r.v.y = a.s * b.v.y + a.v.y * b.s + a.v.z * b.v.x - a.v.x * b.v.z;
    r.v.z = a.s * b.v.z + a.v.z * b.s + a.v.x * b.v.y - a.v.y * b.v.x;
    return r;
}
// Transform::concat for Decomposed (cgmath 0.18 transform.rs): self then-applied-after other
static inline Xf concat(const Xf& self, const Xf& other) {
    Xf r;
    r.scale = self.scale * other.scale;
    r.rot = quat_mul(self.rot, other.rot);
    r.disp = rotate(self.rot, other.disp * self.scale) + self.disp;
    return r;
}

struct Panic {
    uint8_t status;
};

// ---- primitives ----------------------------------------------------------------
struct Shape {
    uint32_t kind;
    V3 p;                 // Sphere: p.x = radius; Cuboid: dim
    const float* verts;   // polyhedron vertices (xyz interleaved)
    uint32_t nverts;
    uint32_t voff;        // first vertex in the global table (HalfEdge ring lookup)
};

// HalfEdge polyhedra (ConvexPolyhedron::new_with_faces): the half-edge structure of every such
// shape, built once by orc_build_half_edges exactly as build_half_edges does (polyhedron.rs:
// 282-380).  Only what hill_climb_support_point reads is kept: per vertex its `edge`, per edge
// its target_vertex / twin_edge / next_edge.  Indexed by the GLOBAL vertex number.
struct HalfEdges {
    std::vector<uint32_t> vertex_edge;            // per global vertex: index into the shape's edges
    std::vector<uint32_t> shape_edge_base;        // per global vertex: first edge of its shape
    std::vector<uint32_t> target, twin, next;     // per edge (all shapes concatenated)
};
static HalfEdges g_he;

static Shape load_shape(const cb2_shape& s, const float* verts_xyz) {
    Shape r;
    r.kind = s.kind;
    r.p = V3(R(s.p[0]), R(s.p[1]), R(s.p[2]));
    r.verts = verts_xyz ? verts_xyz + 3ull * s.vtx_off : nullptr;
    r.nverts = s.vtx_cnt;
    r.voff = s.vtx_off;
    return r;
}

// Cuboid::generate_corners, cuboid.rs:54-65 (half_dim = dim / 2, cuboid.rs:36)
static void cuboid_corners(const Shape& s, V3 c[8]) {
    V3 h = s.p / (R(1.0f) + R(1.0f));
    c[0] = V3(h.x, h.y, h.z);
    c[1] = V3(-h.x, h.y, h.z);
    c[2] = V3(-h.x, -h.y, h.z);
    c[3] = V3(h.x, -h.y, h.z);
    c[4] = V3(h.x, h.y, -h.z);
    c[5] = V3(-h.x, h.y, -h.z);
    c[6] = V3(-h.x, -h.y, -h.z);
    c[7] = V3(h.x, -h.y, -h.z);
}

// Primitive::support_point, Primitive3 dispatch (primitive3.rs:153-176):
//   Particle -> transform_point(origin)                         (primitive3.rs:164)
//   Cuboid / Cube -> util.rs:11-30 get_max_point over the 8 cached corners (cuboid.rs:74-79,170-182)
//   Quad     -> get_max_point over 4 corners in the z = 0 plane  (quad.rs:46-57, 65-77)
//   Sphere   -> sphere.rs:32-38
//   Cylinder -> cylinder.rs:45-78;  Capsule -> capsule.rs:45-61
//   ConvexPolyhedron (VertexOnly) -> polyhedron.rs:160-176, 388-400
static inline R signum(R x) {  // f32::signum: 1 for +0 and positive, -1 for -0 and negative, NaN
    if (x.f != x.f) return x;
    return R(std::signbit(x.f) ? -1.0f : 1.0f);
}
static V3 support_point(const Shape& s, V3 direction, const Xf& t) {
    if (s.kind == CB2_PARTICLE) return transform_point(t, zero3());
    V3 d;
    if (!inverse_transform_vector(t, direction, &d)) throw Panic{CB2_PANIC_INV_SCALE};
    if (s.kind == CB2_SPHERE) {
        return transform_point(t, normalize_to(d, s.p.x));
    }
    if (s.kind == CB2_CAPSULE) {  // p.x = half_height, p.y = radius
        V3 result = zero3();
        result.y = signum(d.y) * s.p.x;
        return transform_point(t, result + normalize_to(d, s.p.y));
    }
    if (s.kind == CB2_CYLINDER) {
        V3 result = d;
        const bool negative = std::signbit(result.y.f);  // is_sign_negative
        result.y = R(0.0f);
        if (magnitude2(result).f == 0.0f) {
            result = zero3();
        } else {
            result = normalize(result);
            if (result.x.f == 0.0f && result.y.f == 0.0f && result.z.f == 0.0f) result = zero3();
            else result = result * s.p.y;
        }
        result.y = negative ? -s.p.x : s.p.x;
        return transform_point(t, result);
    }
    if (s.kind == CB2_POLYHEDRON_HE) {  // hill_climb_support_point, polyhedron.rs:179-209
        auto pos = [&](uint32_t i) { return V3(R(s.verts[3 * i]), R(s.verts[3 * i + 1]), R(s.verts[3 * i + 2])); };
        const uint32_t ebase = g_he.shape_edge_base[s.voff];
        uint32_t best_index = 0;
        R best_dot = dot(pos(0), d);
        for (uint64_t guard = 0;; ++guard) {
            // the dot strictly grows each round, so a climb makes fewer than nverts rounds; only
            // a NaN best_dot (NaN != NaN) keeps the reference's loop alive forever
            if (guard > (uint64_t)s.nverts + 1) throw Panic{CB2_NONCONVERGED};
            const R previous_best_dot = best_dot;
            uint32_t edge_index = g_he.vertex_edge[s.voff + best_index];
            const uint32_t start_edge_index = edge_index;
            for (;;) {
                const uint32_t vertex_index = g_he.target[ebase + edge_index];
                const R dt = dot(pos(vertex_index), d);
                if (dt > best_dot) {
                    best_index = vertex_index;
                    best_dot = dt;
                }
                edge_index = g_he.next[ebase + g_he.twin[ebase + edge_index]];
                if (start_edge_index == edge_index) break;
            }
            if (best_dot == previous_best_dot) break;
        }
        return transform_point(t, pos(best_index));
    }
    V3 best = zero3();  // P::origin()
    R best_dot = R(-std::numeric_limits<float>::infinity());
    if (s.kind == CB2_CUBOID || s.kind == CB2_CUBE) {
        Shape cs = s;
        if (s.kind == CB2_CUBE) cs.p = V3(s.p.x, s.p.x, s.p.x);
        V3 c[8];
        cuboid_corners(cs, c);
        for (int i = 0; i < 8; ++i) {
            R dt = dot(c[i], d);
            if (dt > best_dot) {
                best = c[i];
                best_dot = dt;
            }
        }
    } else if (s.kind == CB2_QUAD) {
        const R hx = s.p.x / (R(1.0f) + R(1.0f)), hy = s.p.y / (R(1.0f) + R(1.0f));
        const V3 c[4] = {V3(hx, hy, R(0.0f)), V3(-hx, hy, R(0.0f)), V3(-hx, -hy, R(0.0f)),
                         V3(hx, -hy, R(0.0f))};
        for (int i = 0; i < 4; ++i) {
            R dt = dot(c[i], d);
            if (dt > best_dot) {
                best = c[i];
                best_dot = dt;
            }
        }
    } else {
        for (uint32_t i = 0; i < s.nverts; ++i) {
            V3 v(R(s.verts[3 * i]), R(s.verts[3 * i + 1]), R(s.verts[3 * i + 2]));
            R dt = dot(v, d);
            if (dt > best_dot) {
                best = v;
                best_dot = dt;
            }
        }
    }
    return transform_point(t, best);
}

// Aabb3 + ComputeBound + Aabb::transform (aabb3.rs:30-50, 90-100; aabb/mod.rs:21-33,116-118)
static inline R rmin(R l, R r) { return (l > r) ? r : l; }  // Less|Equal|None => lhs
static inline R rmax(R l, R r) { return (l < r) ? r : l; }  // Greater|Equal|None => lhs
struct Aabb {
    V3 mn, mx;
};
static inline Aabb aabb_new(V3 p1, V3 p2) {
    Aabb a;
    a.mn = V3(rmin(p1.x, p2.x), rmin(p1.y, p2.y), rmin(p1.z, p2.z));
    a.mx = V3(rmax(p1.x, p2.x), rmax(p1.y, p2.y), rmax(p1.z, p2.z));
    return a;
}
static inline Aabb aabb_grow(const Aabb& a, V3 p) {
    V3 mn(rmin(a.mn.x, p.x), rmin(a.mn.y, p.y), rmin(a.mn.z, p.z));
    V3 mx(rmax(a.mx.x, p.x), rmax(a.mx.y, p.y), rmax(a.mx.z, p.z));
    return aabb_new(mn, mx);
}
static Aabb compute_bound(const Shape& s) {
    if (s.kind == CB2_SPHERE) {  // sphere.rs:41-51
        R r = s.p.x;
        return aabb_new(V3(-r, -r, -r), V3(r, r, r));
    }
    if (s.kind == CB2_CUBOID || s.kind == CB2_CUBE) {  // cuboid.rs:82-92, 184-191
        V3 dim = s.kind == CB2_CUBE ? V3(s.p.x, s.p.x, s.p.x) : s.p;
        V3 h = dim / (R(1.0f) + R(1.0f));
        return aabb_new(-h, h);
    }
    if (s.kind == CB2_PARTICLE) return aabb_new(zero3(), zero3());  // Aabb3::zero(), primitive3.rs:120
    if (s.kind == CB2_QUAD) {  // quad.rs:80-90
        const R hx = s.p.x / (R(1.0f) + R(1.0f)), hy = s.p.y / (R(1.0f) + R(1.0f));
        return aabb_new(V3(-hx, -hy, R(0.0f)), V3(hx, hy, R(0.0f)));
    }
    if (s.kind == CB2_CYLINDER) {  // cylinder.rs:81-91
        const R hh = s.p.x, r = s.p.y;
        return aabb_new(V3(-r, -hh, -r), V3(r, hh, r));
    }
    if (s.kind == CB2_CAPSULE) {  // capsule.rs:64-74
        const R hh = s.p.x, r = s.p.y;
        return aabb_new(V3(-r, -hh - r, -r), V3(r, hh + r, r));
    }
    // polyhedron.rs:113-115: fold(Aabb3::zero(), grow) — always contains the origin
    Aabb b = aabb_new(zero3(), zero3());
    for (uint32_t i = 0; i < s.nverts; ++i)
        b = aabb_grow(b, V3(R(s.verts[3 * i]), R(s.verts[3 * i + 1]), R(s.verts[3 * i + 2])));
    return b;
}
static Aabb aabb_transform(const Aabb& a, const Xf& t) {  // aabb3.rs:39-50, 90-100
    V3 c[8] = {a.mn,
               V3(a.mx.x, a.mn.y, a.mn.z),
               V3(a.mn.x, a.mx.y, a.mn.z),
               V3(a.mx.x, a.mx.y, a.mn.z),
               V3(a.mn.x, a.mn.y, a.mx.z),
               V3(a.mx.x, a.mn.y, a.mx.z),
               V3(a.mn.x, a.mx.y, a.mx.z),
               a.mx};
    V3 f = transform_point(t, c[0]);
    Aabb b = aabb_new(f, f);
    for (int i = 1; i < 8; ++i) b = aabb_grow(b, transform_point(t, c[i]));
    return b;
}
// Discrete<Aabb3> for Aabb3, aabb3.rs:246-253 (strict)
static inline bool aabb_intersects(const float* a, const float* b) {
    return a[3] > b[0] && a[0] < b[3] && a[4] > b[1] && a[1] < b[4] && a[5] > b[2] && a[2] < b[5];
}

// ---- Minkowski support point, minkowski/mod.rs:16-76 ------------------------------
struct Sup {
    V3 v, sup_a, sup_b;
};
static Sup from_minkowski(const Shape& l, const Xf& lt, const Shape& r, const Xf& rt, V3 dir) {
    Sup s;
    V3 pl = support_point(l, dir, lt);
    V3 pr = support_point(r, -dir, rt);
    s.v = pl - pr;
    s.sup_a = pl;
    s.sup_b = pr;
    return s;
}

// SmallVec<[SupportPoint; 5]> (simplex/mod.rs:12): five points inline on the stack, the heap only
// beyond that — the crate's GJK loops never allocate, and neither may a baseline that is timed
// against it (a std::vector here cost a malloc per pair).
class Simplex {
  public:
    Simplex() : n_(0) {}
    size_t size() const { return n_; }
    Sup& operator[](size_t i) { return data()[i]; }
    const Sup& operator[](size_t i) const { return data()[i]; }
    Sup* begin() { return data(); }
    Sup* end() { return data() + n_; }
    const Sup* begin() const { return data(); }
    const Sup* end() const { return data() + n_; }
    void clear() { n_ = 0; }
    void push_back(const Sup& p) {
        if (spill_.empty() && n_ == 5) spill_.assign(inl_, inl_ + 5);  // SmallVec spills to the heap
        if (!spill_.empty()) {
            spill_.resize(n_ + 1);
            spill_[n_++] = p;
        } else {
            inl_[n_++] = p;
        }
    }
    void erase(Sup* at) {  // SmallVec::remove: shift the tail down
        for (Sup* q = at; q + 1 < end(); ++q) *q = *(q + 1);
        --n_;
    }

  private:
    Sup* data() { return spill_.empty() ? inl_ : spill_.data(); }
    const Sup* data() const { return spill_.empty() ? inl_ : spill_.data(); }
    Sup inl_[5];
    size_t n_;
    std::vector<Sup> spill_;  // empty (no allocation) unless a sixth point is ever pushed
};

// ---- SimplexProcessor3, simplex3d.rs ----------------------------------------------
static inline V3 cross_aba(V3 a, V3 b) { return cross(cross(a, b), a); }  // :172-177

static void check_side(V3 abc, V3 ab, V3 ac, V3 ao, Simplex& simplex, V3& v, bool above,
                       bool ignore_ab) {  // :180-220
    V3 ab_perp = cross(ab, abc);
    if (!ignore_ab && dot(ab_perp, ao) > R(0.0f)) {
        simplex.erase(simplex.begin() + 0);
        v = cross_aba(ab, ao);
        return;
    }
    V3 ac_perp = cross(abc, ac);
    if (dot(ac_perp, ao) > R(0.0f)) {
        simplex.erase(simplex.begin() + 1);
        v = cross_aba(ac, ao);
        return;
    }
    if (above || dot(abc, ao) > R(0.0f)) {
        v = abc;
    } else {
        std::swap(simplex[0], simplex[1]);
        v = -abc;
    }
}

static bool reduce_to_closest_feature(Simplex& simplex, V3& v) {  // :24-98
    if (simplex.size() == 4) {
        V3 a = simplex[3].v, b = simplex[2].v, c = simplex[1].v, d = simplex[0].v;
        V3 ao = -a, ab = b - a, ac = c - a;
        V3 abc = cross(ab, ac);
        if (dot(abc, ao) > R(0.0f)) {
            simplex.erase(simplex.begin() + 0);
            check_side(abc, ab, ac, ao, simplex, v, true, false);
        } else {
            V3 ad = d - a;
            V3 acd = cross(ac, ad);
            if (dot(acd, ao) > R(0.0f)) {
                simplex.erase(simplex.begin() + 2);
                check_side(acd, ac, ad, ao, simplex, v, true, true);
            } else {
                V3 adb = cross(ad, ab);
                if (dot(adb, ao) > R(0.0f)) {
                    simplex.erase(simplex.begin() + 1);
                    std::swap(simplex[0], simplex[1]);
                    v = adb;
                } else {
                    return true;
                }
            }
        }
    } else if (simplex.size() == 3) {
        V3 a = simplex[2].v, b = simplex[1].v, c = simplex[0].v;
        V3 ao = -a, ab = b - a, ac = c - a;
        check_side(cross(ab, ac), ab, ac, ao, simplex, v, false, false);
    } else if (simplex.size() == 2) {
        V3 a = simplex[1].v, b = simplex[0].v;
        V3 ao = -a, ab = b - a;
        v = cross_aba(ab, ao);
        if (ulps_eq_zero(v)) v.x = R(0.1f);
    }
    return false;
}

// util.rs:53-72
static void barycentric_vector(V3 p, V3 a, V3 b, V3 c, R& u, R& v, R& w) {
    V3 v0 = b - a, v1 = c - a, v2 = p - a;
    R d00 = dot(v0, v0), d01 = dot(v0, v1), d11 = dot(v1, v1), d20 = dot(v2, v0),
      d21 = dot(v2, v1);
    R inv_denom = R(1.0f) / (d00 * d11 - d01 * d01);
    v = (d11 * d20 - d01 * d21) * inv_denom;
    w = (d00 * d21 - d01 * d20) * inv_denom;
    u = R(1.0f) - v - w;
}

// util.rs:98-114
static V3 get_closest_point_on_edge(V3 start, V3 end, V3 point) {
    V3 line = end - start;
    V3 line_dir = normalize(line);
    V3 v = point - start;
    R d = dot(v, line_dir);
    if (d < R(0.0f)) return start;
    if ((d * d) > magnitude2(line)) return end;
    return start + line_dir * d;
}

static inline bool in_range(R v) { return v >= R(0.0f) && v <= R(1.0f); }  // :164-169

// simplex3d.rs:131-161 with Plane::from_points (plane.rs:78-95) and
// Plane ∩ Ray3 (plane.rs:176-188) written out
static V3 get_closest_point_on_face(V3 a, V3 b, V3 c, V3 normal) {
    V3 ray_origin = zero3();
    V3 ray_dir = -normal;
    V3 v0 = b - a, v1 = c - a;
    V3 n = cross(v0, v1);
    if (ulps_eq_zero(n)) throw Panic{CB2_PANIC_UNWRAP_PLANE};
    n = normalize(n);
    R pd = -dot(a, n);
    R t = -(pd + dot(ray_origin, n)) / dot(ray_dir, n);
    if (t < R(0.0f)) return zero3();
    V3 point = ray_origin + ray_dir * t;
    R u, v, w;
    barycentric_vector(point, a, b, c, u, v, w);
    if (!(in_range(u) && in_range(v) && in_range(w))) throw Panic{CB2_PANIC_ASSERT_RANGE};
    return point;
}

static V3 get_closest_point_to_origin(Simplex& simplex) {  // :103-121
    V3 d = zero3();
    if (reduce_to_closest_feature(simplex, d)) return d;
    if (simplex.size() == 1) return simplex[0].v;
    if (simplex.size() == 2)
        return get_closest_point_on_edge(simplex[1].v, simplex[0].v, zero3());
    return get_closest_point_on_face(simplex[2].v, simplex[1].v, simplex[0].v, d);
}

// ---- EPA3, epa3d.rs -----------------------------------------------------------------
struct Face {
    size_t vertices[3];
    V3 normal;
    R distance;
};
static Face face_new_impl(const std::vector<Sup>& s, size_t a, size_t b, size_t c) {  // :189-199
    V3 ab = s[b].v - s[a].v;
    V3 ac = s[c].v - s[a].v;
    Face f;
    f.normal = normalize(cross(ab, ac));
    f.distance = dot(f.normal, s[a].v);
    f.vertices[0] = a;
    f.vertices[1] = b;
    f.vertices[2] = c;
    return f;
}
static std::vector<Face> face_new(const std::vector<Sup>& s) {  // :201-208
    std::vector<Face> f;
    f.push_back(face_new_impl(s, 3, 2, 1));
    f.push_back(face_new_impl(s, 3, 1, 0));
    f.push_back(face_new_impl(s, 3, 0, 2));
    f.push_back(face_new_impl(s, 2, 0, 1));
    return f;
}
static thread_local uint32_t t_peak_edges = 0;  // instrumentation only
typedef std::pair<size_t, size_t> Edge;
static void remove_or_add_edge(std::vector<Edge>& edges, Edge e) {  // :212-219
    for (size_t i = 0; i < edges.size(); ++i) {
        if (e.first == edges[i].second && e.second == edges[i].first) {
            edges.erase(edges.begin() + i);
            return;
        }
    }
    edges.push_back(e);
}
struct Polytope {
    std::vector<Sup>* vertices;
    std::vector<Face> faces;
};
static size_t closest_face_to_origin(const Polytope& p) {  // :137-145
    if (p.faces.empty()) throw Panic{CB2_PANIC_EPA_EMPTY};
    size_t face = 0;
    for (size_t i = 1; i < p.faces.size(); ++i)
        if (p.faces[i].distance < p.faces[face].distance) face = i;
    return face;
}
static void polytope_add(Polytope& p, const Sup& sup) {  // :147-175
    std::vector<Edge> edges;
    size_t i = 0;
    while (i < p.faces.size()) {
        R dt = dot(p.faces[i].normal, sup.v - (*p.vertices)[p.faces[i].vertices[0]].v);
        if (dt > R(0.0f)) {
            Face face = p.faces[i];
            p.faces[i] = p.faces.back();  // swap_remove
            p.faces.pop_back();
            remove_or_add_edge(edges, Edge(face.vertices[0], face.vertices[1]));
            remove_or_add_edge(edges, Edge(face.vertices[1], face.vertices[2]));
            remove_or_add_edge(edges, Edge(face.vertices[2], face.vertices[0]));
            if (edges.size() > t_peak_edges) t_peak_edges = (uint32_t)edges.size();
        } else {
            ++i;
        }
    }
    size_t n = p.vertices->size();
    p.vertices->push_back(sup);
    std::vector<Face> new_faces;
    for (size_t k = 0; k < edges.size(); ++k)
        new_faces.push_back(face_new_impl(*p.vertices, n, edges[k].first, edges[k].second));
    p.faces.insert(p.faces.end(), new_faces.begin(), new_faces.end());
}
static void epa_contact(const Polytope& p, const Face& face, cb2_contact* out) {  // :84-114
    R u, v, w;
    const std::vector<Sup>& vs = *p.vertices;
    barycentric_vector(face.normal * face.distance, vs[face.vertices[0]].v,
                       vs[face.vertices[1]].v, vs[face.vertices[2]].v, u, v, w);
    V3 pt = vs[face.vertices[0]].sup_a * u + vs[face.vertices[1]].sup_a * v +
            vs[face.vertices[2]].sup_a * w;
    out->strategy = CB2_FULL_RESOLUTION;
    out->normal[0] = face.normal.x.f;
    out->normal[1] = face.normal.y.f;
    out->normal[2] = face.normal.z.f;
    out->penetration_depth = face.distance.f;
    out->contact_point[0] = pt.x.f;
    out->contact_point[1] = pt.y.f;
    out->contact_point[2] = pt.z.f;
    out->time_of_impact = 0.0f;
}
struct EpaStats {
    uint32_t iterations, faces, verts;
    uint32_t peak_faces, peak_edges;  // scratch the device kernels must provision
};
// EPA3::process, :27-65. Returns false for None.
static bool epa_process(std::vector<Sup>& simplex, const Shape& l, const Xf& lt, const Shape& r,
                        const Xf& rt, R tolerance, uint32_t max_iterations, cb2_contact* out,
                        EpaStats* st) {
    if (simplex.size() < 4) return false;
    Polytope poly;
    poly.vertices = &simplex;
    poly.faces = face_new(simplex);
    uint32_t i = 1;
    uint32_t peak_faces = 4;
    t_peak_edges = 0;
    for (;;) {
        if (poly.faces.size() > peak_faces) peak_faces = (uint32_t)poly.faces.size();
        size_t fi = closest_face_to_origin(poly);
        Face face = poly.faces[fi];
        Sup p = from_minkowski(l, lt, r, rt, face.normal);
        R d = dot(p.v, face.normal);
        if (d - face.distance < tolerance || i >= max_iterations) {
            epa_contact(poly, face, out);
            if (st) {
                st->iterations = i;
                st->faces = (uint32_t)poly.faces.size();
                st->verts = (uint32_t)simplex.size();
                st->peak_faces = peak_faces;
                st->peak_edges = t_peak_edges;
            }
            return true;
        }
        polytope_add(poly, p);
        i += 1;
    }
}

// ---- GJK, gjk/mod.rs -----------------------------------------------------------------
// GJK::intersect, :101-146. Returns false for None.
static bool gjk_intersect(const Shape& l, const Xf& lt, const Shape& r, const Xf& rt,
                          uint32_t max_iterations, Simplex& simplex, uint32_t* iters) {
    V3 left_pos = transform_point(lt, zero3());
    V3 right_pos = transform_point(rt, zero3());
    V3 d = right_pos - left_pos;
    if (ulps_eq_zero(d)) d = V3(R(1.0f), R(1.0f), R(1.0f));
    Sup a = from_minkowski(l, lt, r, rt, d);
    if (iters) *iters = 0;
    if (dot(a.v, d) <= R(0.0f)) return false;
    simplex.clear();
    simplex.push_back(a);
    d = -d;
    for (uint32_t it = 0; it < max_iterations; ++it) {
        if (iters) *iters = it + 1;
        Sup a2 = from_minkowski(l, lt, r, rt, d);
        if (dot(a2.v, d) <= R(0.0f)) return false;
        simplex.push_back(a2);
        if (reduce_to_closest_feature(simplex, d)) return true;
    }
    return false;
}

static void contact_zero(cb2_contact* c, uint32_t strategy) {  // Contact::new, contact.rs:54-56
    std::memset(c, 0, sizeof(*c));
    c->strategy = strategy;
}

// GJK::intersection, :362-391
static uint8_t gjk_intersection(uint32_t strategy, const cb2_settings& st, const Shape& l,
                                const Xf& lt, const Shape& r, const Xf& rt, cb2_contact* out,
                                uint32_t* gjk_iters, EpaStats* est) {
    contact_zero(out, strategy);
    try {
        Simplex simplex;
        if (!gjk_intersect(l, lt, r, rt, st.max_iterations, simplex, gjk_iters)) return CB2_NONE;
        if (strategy == CB2_COLLISION_ONLY) return CB2_SOME;
        std::vector<Sup> v(simplex.begin(), simplex.end());  // into_vec()
        if (!epa_process(v, l, lt, r, rt, R(st.epa_tolerance), st.max_iterations, out, est)) {
            contact_zero(out, strategy);
            return CB2_NONE;
        }
        return CB2_SOME;
    } catch (Panic& p) {
        contact_zero(out, strategy);
        return p.status;
    }
}

// GJK::distance, :271-321
static uint8_t gjk_distance(const cb2_settings& st, const Shape& l, const Xf& lt, const Shape& r,
                            const Xf& rt, float* out, uint32_t* iters) {
    *out = 0.0f;
    if (iters) *iters = 0;
    try {
        V3 right_pos = transform_point(rt, zero3());
        V3 left_pos = transform_point(lt, zero3());
        Simplex simplex;
        V3 d = right_pos - left_pos;
        if (ulps_eq_zero(d)) d = V3(R(1.0f), R(1.0f), R(1.0f));
        simplex.push_back(from_minkowski(l, lt, r, rt, d));
        simplex.push_back(from_minkowski(l, lt, r, rt, -d));
        for (uint32_t it = 0; it < st.max_iterations; ++it) {
            if (iters) *iters = it + 1;
            V3 dd = get_closest_point_to_origin(simplex);
            if (ulps_eq_zero(dd)) return CB2_NONE;
            dd = -dd;
            Sup p = from_minkowski(l, lt, r, rt, dd);
            R dp = dot(p.v, dd);
            R d0 = dot(simplex[0].v, dd);
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

// GJK::intersection_time_of_impact, :163-256 (+ iter_cap: the reference loop is unbounded)
static uint8_t gjk_toi(const cb2_settings& st, const Shape& l, const Xf& lt0, const Xf& lt1,
                       const Shape& r, const Xf& rt0, const Xf& rt1, uint32_t iter_cap,
                       cb2_contact* out, uint32_t* iters) {
    contact_zero(out, CB2_FULL_RESOLUTION);
    if (iters) *iters = 0;
    try {
        V3 left_lin_vel = transform_point(lt1, zero3()) - transform_point(lt0, zero3());
        V3 right_lin_vel = transform_point(rt1, zero3()) - transform_point(rt0, zero3());
        V3 ray = right_lin_vel - left_lin_vel;
        R lambda = R(0.0f);
        V3 normal = zero3();
        V3 ray_origin = zero3();
        Simplex simplex;
        Sup p0 = from_minkowski(l, lt0, r, rt0, -ray);
        V3 v = p0.v;
        uint32_t it = 0;
        while (magnitude2(v) > R(st.continuous_tolerance)) {
            if (it >= iter_cap) return CB2_NONCONVERGED;
            ++it;
            if (iters) *iters = it;
            Sup p = from_minkowski(l, lt0, r, rt0, -v);
            R vp = dot(v, p.v);
            R vr = dot(v, ray);
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
            Sup q = p;  // Sub<P> for SupportPoint, minkowski/mod.rs:63-76
            q.v = p.v - ray_origin;
            simplex.push_back(q);
            v = get_closest_point_to_origin(simplex);
        }
        if (magnitude2(v) <= R(st.continuous_tolerance)) {
            Xf t = translation_interpolate(rt0, rt1, lambda);
            V3 n = -normalize(normal);
            V3 cp = transform_point(t, ray_origin);
            out->strategy = CB2_FULL_RESOLUTION;
            out->normal[0] = n.x.f;
            out->normal[1] = n.y.f;
            out->normal[2] = n.z.f;
            out->penetration_depth = magnitude(v).f;
            out->contact_point[0] = cp.x.f;
            out->contact_point[1] = cp.y.f;
            out->contact_point[2] = cp.z.f;
            out->time_of_impact = lambda.f;
            return CB2_SOME;
        }
        return CB2_NONE;
    } catch (Panic& p) {
        contact_zero(out, CB2_FULL_RESOLUTION);
        return p.status;
    }
}

// ---- SweepAndPrune3::find_collider_pairs, sweep_prune.rs:76-136 ------------------------
// partial_cmp on f32: -1 Less, 0 Equal, 1 Greater, 2 None
static inline int partial_cmp(float a, float b) {
    if (a < b) return -1;
    if (a > b) return 1;
    if (a == b) return 0;
    return 2;
}
// ---- composite shapes: gjk/mod.rs:412-588 ------------------------------------------------------
struct Composite {  // &[(Primitive, local transform)]
    const uint32_t* shape;
    const cb2_xform* local;
    uint32_t n;
};
#define CB2_PANIC_CMP_NAN 8  // max_by(partial_cmp(..).unwrap()) on a NaN depth, gjk/mod.rs:455-459

// GJK::intersection_complex, :412-461
static uint8_t gjk_intersection_complex(uint32_t strategy, const cb2_settings& st,
                                        const cb2_shape* shapes, const float* verts_xyz,
                                        const Composite& L, const Xf& lt, const Composite& Rc,
                                        const Xf& rt, cb2_contact* out) {
    contact_zero(out, strategy);
    bool have = false;
    cb2_contact best;
    for (uint32_t a = 0; a < L.n; ++a) {
        const Xf left_transform = concat(lt, load_xf(L.local[a]));
        for (uint32_t b = 0; b < Rc.n; ++b) {
            const Xf right_transform = concat(rt, load_xf(Rc.local[b]));
            cb2_contact c;
            const uint8_t s = gjk_intersection(strategy, st, load_shape(shapes[L.shape[a]], verts_xyz),
                                               left_transform, load_shape(shapes[Rc.shape[b]], verts_xyz),
                                               right_transform, &c, nullptr, nullptr);
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
static uint8_t gjk_distance_complex(const cb2_settings& st, const cb2_shape* shapes,
                                    const float* verts_xyz, const Composite& L, const Xf& lt,
                                    const Composite& Rc, const Xf& rt, float* out) {
    *out = 0.0f;
    bool have = false;
    float m = 0.0f;
    for (uint32_t a = 0; a < L.n; ++a) {
        const Xf left_transform = concat(lt, load_xf(L.local[a]));
        for (uint32_t b = 0; b < Rc.n; ++b) {
            const Xf right_transform = concat(rt, load_xf(Rc.local[b]));
            float d;
            const uint8_t s = gjk_distance(st, load_shape(shapes[L.shape[a]], verts_xyz), left_transform,
                                           load_shape(shapes[Rc.shape[b]], verts_xyz), right_transform,
                                           &d, nullptr);
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
static uint8_t gjk_toi_complex(uint32_t strategy, const cb2_settings& st, const cb2_shape* shapes,
                               const float* verts_xyz, const Composite& L, const Xf& lt0,
                               const Xf& lt1, const Composite& Rc, const Xf& rt0, const Xf& rt1,
                               uint32_t iter_cap, cb2_contact* out) {
    contact_zero(out, strategy);
    bool have = false;
    cb2_contact best;
    for (uint32_t a = 0; a < L.n; ++a) {
        const Xf l0 = concat(lt0, load_xf(L.local[a]));
        const Xf l1 = concat(lt1, load_xf(L.local[a]));
        for (uint32_t b = 0; b < Rc.n; ++b) {
            const Xf r0 = concat(rt0, load_xf(Rc.local[b]));
            const Xf r1 = concat(rt1, load_xf(Rc.local[b]));
            cb2_contact c;
            const uint8_t s = gjk_toi(st, load_shape(shapes[L.shape[a]], verts_xyz), l0, l1,
                                      load_shape(shapes[Rc.shape[b]], verts_xyz), r0, r1, iter_cap, &c,
                                      nullptr);
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

static int sap_cmp(const float* a, const float* b, uint32_t axis) {  // :87-97, returns Ordering
    int c = partial_cmp(a[axis], b[axis]);
    if (c == 0) {
        int m = partial_cmp(a[3 + axis], b[3 + axis]);
        return m == 2 ? 0 : m;
    }
    if (c == 2) return 0;
    return c;
}

template <typename F>
static void parallel_for(uint64_t n, int nthreads, F fn) {
    if (nthreads <= 1 || n < 2) {
        fn(0, n, 0);
        return;
    }
    // dynamic chunks: per-pair cost varies by 100x
    std::atomic<uint64_t> next(0);
    const uint64_t chunk = std::max<uint64_t>(64, std::min<uint64_t>(4096, n / (16ull * nthreads) + 1));
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; ++t) {
        th.emplace_back([&, t]() {
            for (;;) {
                uint64_t b = next.fetch_add(chunk);
                if (b >= n) break;
                fn(b, std::min(n, b + chunk), t);
            }
        });
    }
    for (auto& x : th) x.join();
}

}  // namespace orc

using namespace orc;

// =========================== extern "C" surface (ctypes) ===============================
extern "C" {

int orc_hw_threads(void) { return (int)std::thread::hardware_concurrency(); }

// build_half_edges (polyhedron.rs:282-380) for every CB2_POLYHEDRON_HE shape of the table; faces
// hold vertex indices local to the shape; the shape's face range is (bits of p[0], bits of p[1]).
// Returns 0, or -1 where the reference panics at construction (degenerate face) / a ring of the
// structure does not close.
int orc_build_half_edges(const cb2_shape* shapes, uint32_t n_shapes, const float* verts_xyz,
                         uint64_t n_verts, const uint32_t* faces) {
    g_he = HalfEdges();
    g_he.vertex_edge.assign(n_verts, 0);
    g_he.shape_edge_base.assign(n_verts, 0);
    for (uint32_t si = 0; si < n_shapes; ++si) {
        const cb2_shape& s = shapes[si];
        if (s.kind != CB2_POLYHEDRON_HE) continue;
        uint32_t f_off, f_cnt;
        std::memcpy(&f_off, &s.p[0], 4);
        std::memcpy(&f_cnt, &s.p[1], 4);
        const uint32_t ebase = (uint32_t)g_he.target.size();
        const uint32_t V = s.vtx_cnt;
        if (V == 0 || f_cnt == 0) return -1;
        std::vector<uint32_t> vedge(V, 0);
        std::vector<char> vready(V, 0), eready;
        std::map<std::pair<uint32_t, uint32_t>, uint32_t> edge_map;
        std::vector<uint32_t> target, twin, next;
        for (uint32_t f = 0; f < f_cnt; ++f) {
            const uint32_t fv[3] = {faces[3 * (f_off + f)], faces[3 * (f_off + f) + 1],
                                    faces[3 * (f_off + f) + 2]};
            for (int k = 0; k < 3; ++k)
                if (fv[k] >= V) return -1;
            {   // Plane::from_points(a, b, c).unwrap(), plane.rs:78-95
                const float* b0 = verts_xyz + 3ull * s.vtx_off;
                V3 a = v3p(b0 + 3 * fv[0]), b = v3p(b0 + 3 * fv[1]), c = v3p(b0 + 3 * fv[2]);
                V3 n = cross(b - a, c - a);
                if (ulps_eq_zero(n)) return -1;
            }
            uint32_t face_edge[3] = {0, 0, 0};
            for (int j = 0; j < 3; ++j) {
                const int i = j == 0 ? 2 : j - 1;
                const uint32_t v0 = fv[i], v1 = fv[j];
                uint32_t e01, e10;
                auto it = edge_map.find({v0, v1});
                if (it != edge_map.end()) {
                    e01 = it->second;
                    e10 = twin[e01];
                } else {
                    e01 = (uint32_t)target.size();
                    e10 = e01 + 1;
                    target.push_back(v1); twin.push_back(e10); next.push_back(0); eready.push_back(0);
                    target.push_back(v0); twin.push_back(e01); next.push_back(0); eready.push_back(0);
                }
                edge_map[{v0, v1}] = e01;
                edge_map[{v1, v0}] = e10;
                if (!eready[e01]) eready[e01] = 1;
                if (!vready[v0]) {
                    vedge[v0] = e01;
                    vready[v0] = 1;
                }
                face_edge[i] = e01;
            }
            for (int j = 0; j < 3; ++j) {
                const int i = j == 0 ? 2 : j - 1;
                next[face_edge[i]] = face_edge[j];
            }
        }
        // every ring the hill climb can walk must close (else the reference never returns)
        const size_t E = target.size();
        for (uint32_t v = 0; v < V; ++v) {
            uint32_t e = vedge[v];
            size_t steps = 0;
            do {
                e = next[twin[e]];
                if (++steps > E) return -1;
            } while (e != vedge[v]);
        }
        for (uint32_t v = 0; v < V; ++v) {
            g_he.vertex_edge[s.vtx_off + v] = vedge[v];
            g_he.shape_edge_base[s.vtx_off + v] = ebase;
        }
        g_he.target.insert(g_he.target.end(), target.begin(), target.end());
        g_he.twin.insert(g_he.twin.end(), twin.begin(), twin.end());
        g_he.next.insert(g_he.next.end(), next.begin(), next.end());
    }
    return 0;
}

// ---- unit-level hooks for the golden tests ---------------------------------------------
void orc_support_point(const cb2_shape* shape, const float* verts_xyz, const cb2_xform* t,
                       const float* dir, float* out, uint8_t* status) {
    *status = CB2_SOME;
    try {
        V3 p = support_point(load_shape(*shape, verts_xyz), v3p(dir), load_xf(*t));
        out[0] = p.x.f;
        out[1] = p.y.f;
        out[2] = p.z.f;
    } catch (Panic& p) {
        *status = p.status;
    }
}

static Simplex simplex_from(const float* pts, uint32_t n) {
    Simplex s;
    for (uint32_t i = 0; i < n; ++i) {
        Sup p;
        p.v = V3(R(pts[3 * i]), R(pts[3 * i + 1]), R(pts[3 * i + 2]));
        p.sup_a = zero3();
        p.sup_b = zero3();
        s.push_back(p);
    }
    return s;
}
static void simplex_to(const Simplex& s, float* pts, uint32_t* n) {
    *n = (uint32_t)s.size();
    for (size_t i = 0; i < s.size(); ++i) {
        pts[3 * i] = s[i].v.x.f;
        pts[3 * i + 1] = s[i].v.y.f;
        pts[3 * i + 2] = s[i].v.z.f;
    }
}
// simplex points in/out (n x 3 floats, only .v), v in/out
int orc_reduce_to_closest_feature(float* pts, uint32_t* n, float* v) {
    Simplex s = simplex_from(pts, *n);
    V3 vv = v3p(v);
    bool hit = reduce_to_closest_feature(s, vv);
    simplex_to(s, pts, n);
    v[0] = vv.x.f;
    v[1] = vv.y.f;
    v[2] = vv.z.f;
    return hit ? 1 : 0;
}
// mirrors the reference test helper test_check_side (simplex3d.rs:381-393)
void orc_check_side(float* pts, uint32_t* n, int above, int ignore_ab, float* v) {
    Simplex s = simplex_from(pts, *n);
    V3 ab = s[1].v - s[2].v, ac = s[0].v - s[2].v, ao = -s[2].v;
    V3 abc = cross(ab, ac);
    V3 vv = zero3();
    check_side(abc, ab, ac, ao, s, vv, above != 0, ignore_ab != 0);
    simplex_to(s, pts, n);
    v[0] = vv.x.f;
    v[1] = vv.y.f;
    v[2] = vv.z.f;
}
uint8_t orc_closest_point_to_origin(float* pts, uint32_t* n, float* out) {
    try {
        Simplex s = simplex_from(pts, *n);
        V3 p = get_closest_point_to_origin(s);
        simplex_to(s, pts, n);
        out[0] = p.x.f;
        out[1] = p.y.f;
        out[2] = p.z.f;
        return CB2_SOME;
    } catch (Panic& p) {
        return p.status;
    }
}
void orc_closest_point_on_edge(const float* start, const float* end, const float* point,
                               float* out) {
    V3 p = get_closest_point_on_edge(v3p(start), v3p(end), v3p(point));
    out[0] = p.x.f;
    out[1] = p.y.f;
    out[2] = p.z.f;
}
void orc_barycentric(const float* p, const float* a, const float* b, const float* c, float* uvw) {
    R u, v, w;
    barycentric_vector(v3p(p), v3p(a), v3p(b), v3p(c), u, v, w);
    uvw[0] = u.f;
    uvw[1] = v.f;
    uvw[2] = w.f;
}
// Plane::from_points + Plane ∩ Ray3 hooks (tests/plane.rs)
int orc_plane_from_points(const float* a, const float* b, const float* c, float* n_d) {
    V3 A = v3p(a), B = v3p(b), C = v3p(c);
    V3 n = cross(B - A, C - A);
    if (ulps_eq_zero(n)) return 0;
    n = normalize(n);
    R d = -dot(A, n);
    n_d[0] = n.x.f;
    n_d[1] = n.y.f;
    n_d[2] = n.z.f;
    n_d[3] = d.f;
    return 1;
}

// Polytope hooks (epa3d.rs:246-316): vertices n x 3 in; faces out as (a,b,c) + (nx,ny,nz,d)
static void faces_out(const std::vector<Face>& f, uint32_t* idx, float* nd, uint32_t* nf) {
    *nf = (uint32_t)f.size();
    for (size_t i = 0; i < f.size(); ++i) {
        idx[3 * i] = (uint32_t)f[i].vertices[0];
        idx[3 * i + 1] = (uint32_t)f[i].vertices[1];
        idx[3 * i + 2] = (uint32_t)f[i].vertices[2];
        nd[4 * i] = f[i].normal.x.f;
        nd[4 * i + 1] = f[i].normal.y.f;
        nd[4 * i + 2] = f[i].normal.z.f;
        nd[4 * i + 3] = f[i].distance.f;
    }
}
// adds n_add points one after another to the polytope built from the first 4 vertices
void orc_polytope(const float* verts, uint32_t n_init, const float* add_pts, uint32_t n_add,
                  uint32_t* face_idx, float* face_nd, uint32_t* n_faces, uint32_t* closest,
                  uint32_t* n_verts) {
    const Simplex s0 = simplex_from(verts, n_init);
    std::vector<Sup> s(s0.begin(), s0.end());
    Polytope p;
    p.vertices = &s;
    p.faces = face_new(s);
    for (uint32_t k = 0; k < n_add; ++k) {
        Sup q;
        q.v = V3(R(add_pts[3 * k]), R(add_pts[3 * k + 1]), R(add_pts[3 * k + 2]));
        q.sup_a = zero3();
        q.sup_b = zero3();
        polytope_add(p, q);
    }
    faces_out(p.faces, face_idx, face_nd, n_faces);
    *closest = (uint32_t)closest_face_to_origin(p);
    *n_verts = (uint32_t)s.size();
}
// remove_or_add_edge hook (epa3d.rs:230-243): edges as 2 x n u32 in/out
void orc_remove_or_add_edge(uint32_t* edges, uint32_t* n, uint32_t a, uint32_t b) {
    std::vector<Edge> e;
    for (uint32_t i = 0; i < *n; ++i) e.push_back(Edge(edges[2 * i], edges[2 * i + 1]));
    remove_or_add_edge(e, Edge(a, b));
    *n = (uint32_t)e.size();
    for (size_t i = 0; i < e.size(); ++i) {
        edges[2 * i] = (uint32_t)e[i].first;
        edges[2 * i + 1] = (uint32_t)e[i].second;
    }
}
// EPA3::process on a caller-supplied simplex (only .v set), epa3d.rs:319-341
uint8_t orc_epa3_process(const float* simplex_v, uint32_t n, const cb2_shape* l,
                         const cb2_xform* lt, const cb2_shape* r, const cb2_xform* rt,
                         const float* verts_xyz, const cb2_settings* st, cb2_contact* out) {
    try {
        const Simplex s0 = simplex_from(simplex_v, n);
        std::vector<Sup> s(s0.begin(), s0.end());
        bool ok = epa_process(s, load_shape(*l, verts_xyz), load_xf(*lt), load_shape(*r, verts_xyz),
                              load_xf(*rt), R(st->epa_tolerance), st->max_iterations, out, nullptr);
        return ok ? CB2_SOME : CB2_NONE;
    } catch (Panic& p) {
        return p.status;
    }
}
// GJK::intersect: returns the simplex (v only), :101-146
uint8_t orc_gjk3_intersect(const cb2_shape* l, const cb2_xform* lt, const cb2_shape* r,
                           const cb2_xform* rt, const float* verts_xyz, const cb2_settings* st,
                           float* simplex_v, uint32_t* n) {
    try {
        Simplex s;
        *n = 0;
        if (!gjk_intersect(load_shape(*l, verts_xyz), load_xf(*lt), load_shape(*r, verts_xyz),
                           load_xf(*rt), st->max_iterations, s, nullptr))
            return CB2_NONE;
        simplex_to(s, simplex_v, n);
        return CB2_SOME;
    } catch (Panic& p) {
        return p.status;
    }
}
int orc_aabb3_intersects(const cb2_aabb3* a, const cb2_aabb3* b) {
    return aabb_intersects(a->min, b->min) ? 1 : 0;
}

// ---- batch entry points: same array contract as include/collision_b200.h ----------------
// stats (optional, may be NULL): per pair u32 x 4 = {gjk iterations, epa iterations, peak faces,
// verts | peak horizon edges << 16}
// GJK::intersect over a batch, whole SupportPoints out (gjk/mod.rs:101-146)
void orc_gjk3_intersect_batch(const cb2_settings* st, const cb2_shape* shapes, const float* verts_xyz,
                              const uint32_t* l_shape, const uint32_t* r_shape, const cb2_xform* l_t,
                              const cb2_xform* r_t, uint64_t n, cb2_support_point* simplex_out,
                              uint8_t* status, int nthreads) {
    parallel_for(n, nthreads, [&](uint64_t b, uint64_t e, int) {
        for (uint64_t i = b; i < e; ++i) {
            std::memset(simplex_out + 4 * i, 0, 4 * sizeof(cb2_support_point));
            try {
                Simplex s;
                bool hit = gjk_intersect(load_shape(shapes[l_shape[i]], verts_xyz), load_xf(l_t[i]),
                                         load_shape(shapes[r_shape[i]], verts_xyz), load_xf(r_t[i]),
                                         st->max_iterations, s, nullptr);
                status[i] = hit ? CB2_SOME : CB2_NONE;
                if (hit)
                    for (size_t k = 0; k < s.size() && k < 4; ++k) {
                        cb2_support_point& o = simplex_out[4 * i + k];
                        o.v[0] = s[k].v.x.f; o.v[1] = s[k].v.y.f; o.v[2] = s[k].v.z.f;
                        o.sup_a[0] = s[k].sup_a.x.f; o.sup_a[1] = s[k].sup_a.y.f; o.sup_a[2] = s[k].sup_a.z.f;
                        o.sup_b[0] = s[k].sup_b.x.f; o.sup_b[1] = s[k].sup_b.y.f; o.sup_b[2] = s[k].sup_b.z.f;
                    }
            } catch (Panic& p) {
                status[i] = p.status;
            }
        }
    });
}
// GJK::get_contact_manifold over a batch (gjk/mod.rs:326-344 -> EPA3::process, epa3d.rs:27-65)
void orc_gjk3_manifold_batch(const cb2_settings* st, const cb2_support_point* simplex,
                             const uint32_t* simplex_len, const cb2_shape* shapes, const float* verts_xyz,
                             const uint32_t* l_shape, const uint32_t* r_shape, const cb2_xform* l_t,
                             const cb2_xform* r_t, uint64_t n, cb2_contact* out, uint8_t* status,
                             int nthreads) {
    parallel_for(n, nthreads, [&](uint64_t b, uint64_t e, int) {
        for (uint64_t i = b; i < e; ++i) {
            contact_zero(&out[i], CB2_FULL_RESOLUTION);
            try {
                const uint32_t len = simplex_len ? simplex_len[i] : 4u;
                std::vector<Sup> v;
                for (uint32_t k = 0; k < len && k < 4; ++k) {
                    const cb2_support_point& q = simplex[4 * i + k];
                    Sup p;
                    p.v = V3(R(q.v[0]), R(q.v[1]), R(q.v[2]));
                    p.sup_a = V3(R(q.sup_a[0]), R(q.sup_a[1]), R(q.sup_a[2]));
                    p.sup_b = V3(R(q.sup_b[0]), R(q.sup_b[1]), R(q.sup_b[2]));
                    v.push_back(p);
                }
                bool ok = epa_process(v, load_shape(shapes[l_shape[i]], verts_xyz), load_xf(l_t[i]),
                                      load_shape(shapes[r_shape[i]], verts_xyz), load_xf(r_t[i]),
                                      R(st->epa_tolerance), st->max_iterations, &out[i], nullptr);
                if (!ok) contact_zero(&out[i], CB2_FULL_RESOLUTION);
                status[i] = ok ? CB2_SOME : CB2_NONE;
            } catch (Panic& p) {
                contact_zero(&out[i], CB2_FULL_RESOLUTION);
                status[i] = p.status;
            }
        }
    });
}
void orc_gjk3_intersection_batch(uint32_t strategy, const cb2_settings* st, const cb2_shape* shapes,
                                 const float* verts_xyz, const uint32_t* l_shape,
                                 const uint32_t* r_shape, const cb2_xform* l_t,
                                 const cb2_xform* r_t, uint64_t n, cb2_contact* out,
                                 uint8_t* status, uint32_t* stats, int nthreads) {
    parallel_for(n, nthreads, [&](uint64_t b, uint64_t e, int) {
        for (uint64_t i = b; i < e; ++i) {
            uint32_t gi = 0;
            EpaStats es = {0, 0, 0, 0, 0};
            status[i] = gjk_intersection(strategy, *st, load_shape(shapes[l_shape[i]], verts_xyz),
                                         load_xf(l_t[i]), load_shape(shapes[r_shape[i]], verts_xyz),
                                         load_xf(r_t[i]), &out[i], &gi, &es);
            if (stats) {
                stats[4 * i] = gi;
                stats[4 * i + 1] = es.iterations;
                stats[4 * i + 2] = es.peak_faces;
                stats[4 * i + 3] = es.verts | (es.peak_edges << 16);
            }
        }
    });
}
void orc_gjk3_distance_batch(const cb2_settings* st, const cb2_shape* shapes,
                             const float* verts_xyz, const uint32_t* l_shape,
                             const uint32_t* r_shape, const cb2_xform* l_t, const cb2_xform* r_t,
                             uint64_t n, float* out, uint8_t* status, uint32_t* iters,
                             int nthreads) {
    parallel_for(n, nthreads, [&](uint64_t b, uint64_t e, int) {
        for (uint64_t i = b; i < e; ++i) {
            uint32_t it = 0;
            status[i] = gjk_distance(*st, load_shape(shapes[l_shape[i]], verts_xyz), load_xf(l_t[i]),
                                     load_shape(shapes[r_shape[i]], verts_xyz), load_xf(r_t[i]),
                                     &out[i], &it);
            if (iters) iters[i] = it;
        }
    });
}
void orc_gjk3_toi_batch(const cb2_settings* st, const cb2_shape* shapes, const float* verts_xyz,
                        const uint32_t* l_shape, const uint32_t* r_shape, const cb2_xform* l_t0,
                        const cb2_xform* l_t1, const cb2_xform* r_t0, const cb2_xform* r_t1,
                        uint64_t n, uint32_t iter_cap, cb2_contact* out, uint8_t* status,
                        uint32_t* iters, int nthreads) {
    if (iter_cap == 0) iter_cap = 1024;
    parallel_for(n, nthreads, [&](uint64_t b, uint64_t e, int) {
        for (uint64_t i = b; i < e; ++i) {
            uint32_t it = 0;
            status[i] = gjk_toi(*st, load_shape(shapes[l_shape[i]], verts_xyz), load_xf(l_t0[i]),
                                load_xf(l_t1[i]), load_shape(shapes[r_shape[i]], verts_xyz),
                                load_xf(r_t0[i]), load_xf(r_t1[i]), iter_cap, &out[i], &it);
            if (iters) iters[i] = it;
        }
    });
}
// composite batch entry points: comp_off[n_comp+1] indexes comp_shape / comp_local
void orc_gjk3_complex_batch(uint32_t op, uint32_t strategy, const cb2_settings* st,
                            const cb2_shape* shapes, const float* verts_xyz,
                            const uint32_t* comp_off, const uint32_t* comp_shape,
                            const cb2_xform* comp_local, const uint32_t* l_comp,
                            const uint32_t* r_comp, const cb2_xform* l_t0, const cb2_xform* l_t1,
                            const cb2_xform* r_t0, const cb2_xform* r_t1, uint64_t n,
                            uint32_t iter_cap, cb2_contact* out_c, float* out_d, uint8_t* status,
                            int nthreads) {
    if (iter_cap == 0) iter_cap = 1024;
    parallel_for(n, nthreads, [&](uint64_t b, uint64_t e, int) {
        for (uint64_t i = b; i < e; ++i) {
            const uint32_t lc = l_comp[i], rc = r_comp[i];
            const Composite L = {comp_shape + comp_off[lc], comp_local + comp_off[lc],
                                 comp_off[lc + 1] - comp_off[lc]};
            const Composite Rc = {comp_shape + comp_off[rc], comp_local + comp_off[rc],
                                  comp_off[rc + 1] - comp_off[rc]};
            if (op == 0)
                status[i] = gjk_intersection_complex(strategy, *st, shapes, verts_xyz, L, load_xf(l_t0[i]),
                                                     Rc, load_xf(r_t0[i]), &out_c[i]);
            else if (op == 1)
                status[i] = gjk_distance_complex(*st, shapes, verts_xyz, L, load_xf(l_t0[i]), Rc,
                                                 load_xf(r_t0[i]), &out_d[i]);
            else
                status[i] = gjk_toi_complex(strategy, *st, shapes, verts_xyz, L, load_xf(l_t0[i]),
                                            load_xf(l_t1[i]), Rc, load_xf(r_t0[i]), load_xf(r_t1[i]),
                                            iter_cap, &out_c[i]);
        }
    });
}
void orc_world_aabbs(const cb2_shape* shapes, const float* verts_xyz, const uint32_t* shape,
                     const cb2_xform* t, uint64_t n, cb2_aabb3* out, int nthreads) {
    parallel_for(n, nthreads, [&](uint64_t b, uint64_t e, int) {
        for (uint64_t i = b; i < e; ++i) {
            Aabb a = aabb_transform(compute_bound(load_shape(shapes[shape[i]], verts_xyz)),
                                    load_xf(t[i]));
            out[i].min[0] = a.mn.x.f;
            out[i].min[1] = a.mn.y.f;
            out[i].min[2] = a.mn.z.f;
            out[i].max[0] = a.mx.x.f;
            out[i].max[1] = a.mx.y.f;
            out[i].max[2] = a.mx.z.f;
        }
    });
}

// SweepAndPrune3::find_collider_pairs, literal (sort, active list with retain, strict test).
// Single-threaded like the reference.  pairs_out may be NULL to only count.
// Returns the number of pairs (even if > cap; only the first cap are written).
uint64_t orc_sap3_find_collider_pairs(const cb2_aabb3* aabbs, uint64_t n, uint32_t* axis_inout,
                                      uint32_t* perm_out, uint32_t* pairs_out, uint64_t cap) {
    for (uint64_t i = 0; i < n; ++i) perm_out[i] = (uint32_t)i;
    if (n <= 1) return 0;  // :83-85, axis unchanged
    const uint32_t axis = *axis_inout;
    std::stable_sort(perm_out, perm_out + n, [&](uint32_t a, uint32_t b) {
        return sap_cmp(aabbs[a].min, aabbs[b].min, axis) < 0;
    });
    std::vector<cb2_aabb3> s(n);
    for (uint64_t i = 0; i < n; ++i) s[i] = aabbs[perm_out[i]];
    uint64_t np = 0;
    std::vector<uint32_t> active;
    active.push_back(0);
    for (uint64_t j = 1; j < n; ++j) {
        const float mn = s[j].min[axis];
        // active.retain(max_extent[axis] >= min_extent[axis]), :109-112
        size_t w = 0;
        for (size_t k = 0; k < active.size(); ++k)
            if (s[active[k]].max[axis] >= mn) active[w++] = active[k];
        active.resize(w);
        for (size_t k = 0; k < active.size(); ++k) {
            if (aabb_intersects(s[active[k]].min, s[j].min)) {
                if (pairs_out && np < cap) {
                    pairs_out[2 * np] = active[k];
                    pairs_out[2 * np + 1] = (uint32_t)j;
                }
                ++np;
            }
        }
        active.push_back((uint32_t)j);
    }
    // Variance3::add_bound discards its sums (sweep_prune.rs:254-261: add_element_wise
    // returns by value) => compute_axis sees zeros => axis 0 (:264-277, :130-133).
    *axis_inout = 0;
    return np;
}

// The same sweep split over threads (for the 1 M-box comparison of the GPU tests; the reference is
// single-threaded).  Thread t replays the loop of :100-122 literally over its range [j0, j1) of
// sorted positions; all it needs is the `active` list as the serial loop would hold it when it
// reaches j0.  With finite, ascending min extents `retain` is monotone — an entry dropped at step j
// stays dropped — so that list is {a < j0 - 1 : max[a] >= min[j0 - 1]} in ascending order, plus
// j0 - 1 itself.  Inputs that break the premise (NaN on the sweep axis) take the serial function.
uint64_t orc_sap3_find_collider_pairs_mt(const cb2_aabb3* aabbs, uint64_t n, uint32_t* axis_inout,
                                         uint32_t* perm_out, uint32_t* pairs_out, uint64_t cap,
                                         int nthreads) {
    const uint32_t axis = *axis_inout;
    bool plain = n > 1 && nthreads > 1;
    for (uint64_t i = 0; i < n && plain; ++i)
        plain = aabbs[i].min[axis] == aabbs[i].min[axis] && aabbs[i].max[axis] == aabbs[i].max[axis];
    if (!plain) return orc_sap3_find_collider_pairs(aabbs, n, axis_inout, perm_out, pairs_out, cap);
    for (uint64_t i = 0; i < n; ++i) perm_out[i] = (uint32_t)i;
    std::stable_sort(perm_out, perm_out + n, [&](uint32_t a, uint32_t b) {
        return sap_cmp(aabbs[a].min, aabbs[b].min, axis) < 0;
    });
    std::vector<cb2_aabb3> s(n);
    for (uint64_t i = 0; i < n; ++i) s[i] = aabbs[perm_out[i]];
    // parallel_for hands out dynamic chunks: results are kept per chunk and stitched in chunk order
    std::map<uint64_t, std::vector<uint32_t>> found;
    std::mutex mu;
    parallel_for(n - 1, nthreads, [&](uint64_t b, uint64_t e, int) {
        const uint64_t j0 = b + 1, j1 = e + 1;  // positions 1 .. n-1
        std::vector<uint32_t> active;
        if (j0 >= 2) {
            const float prev = s[j0 - 1].min[axis];
            for (uint64_t a = 0; a + 1 < j0; ++a)
                if (s[a].max[axis] >= prev) active.push_back((uint32_t)a);
        }
        active.push_back((uint32_t)(j0 - 1));
        std::vector<uint32_t> out;
        for (uint64_t j = j0; j < j1; ++j) {
            const float mn = s[j].min[axis];
            size_t w = 0;
            for (size_t k = 0; k < active.size(); ++k)
                if (s[active[k]].max[axis] >= mn) active[w++] = active[k];
            active.resize(w);
            for (size_t k = 0; k < active.size(); ++k)
                if (aabb_intersects(s[active[k]].min, s[j].min)) {
                    out.push_back(active[k]);
                    out.push_back((uint32_t)j);
                }
            active.push_back((uint32_t)j);
        }
        std::lock_guard<std::mutex> lock(mu);
        found[b].swap(out);
    });
    uint64_t np = 0;
    for (auto& kv : found) {
        const std::vector<uint32_t>& f = kv.second;
        for (size_t k = 0; k + 1 < f.size(); k += 2) {
            if (pairs_out && np < cap) {
                pairs_out[2 * np] = f[k];
                pairs_out[2 * np + 1] = f[k + 1];
            }
            ++np;
        }
    }
    *axis_inout = 0;
    return np;
}

// BruteForce::find_collider_pairs (brute_force.rs:20-39) as a cross-check of the SAP pair SET:
// counts pairs (i<j) with strict overlap, in the given order.
uint64_t orc_brute_force_count(const cb2_aabb3* aabbs, uint64_t n) {
    uint64_t c = 0;
    for (uint64_t i = 0; i < n; ++i)
        for (uint64_t j = i + 1; j < n; ++j)
            if (aabb_intersects(aabbs[i].min, aabbs[j].min)) ++c;
    return c;
}

// the pairs themselves, in the reference's order (left asc, right asc); returns the count
uint64_t orc_brute_force_pairs(const cb2_aabb3* aabbs, uint64_t n, uint32_t* pairs, uint64_t cap) {
    uint64_t c = 0;
    for (uint64_t i = 0; i + 1 < n; ++i)
        for (uint64_t j = i + 1; j < n; ++j)
            if (aabb_intersects(aabbs[i].min, aabbs[j].min)) {
                if (pairs && c < cap) {
                    pairs[2 * c] = (uint32_t)i;
                    pairs[2 * c + 1] = (uint32_t)j;
                }
                ++c;
            }
    return c;
}

// algorithmic flops of the last batch call on this thread (COUNT_OPS builds; single-thread)
uint64_t orc_ops_reset(void) {
#ifdef COUNT_OPS
    uint64_t v = t_ops;
    t_ops = 0;
    return v;
#else
    return 0;
#endif
}
int orc_counts_ops(void) {
#ifdef COUNT_OPS
    return 1;
#else
    return 0;
#endif
}

}  // extern "C"
