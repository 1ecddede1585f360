// cb2_narrow2d.cu — the 2-D narrow phase: GJK2 (+EPA2), SURVEY §8 row N4.
//
// The reference instantiates the SAME generic GJK (gjk/mod.rs:101-391) with SimplexProcessor2
// (gjk/simplex/simplex2d.rs:17-97) and EPA2 (epa/epa2d.rs:21-157); shapes are Primitive2 variants
// (primitive/primitive2.rs:20-103) under Decomposed<Vector2<f32>, Basis2<f32>>.  Arithmetic is
// unfused f32 in cgmath's order (cb2_math.cuh), so results are bit-identical to the reference.
//
// Design: a pair is cheap in 2-D (a support is 10-30 flops, the simplex holds at most 3 points), so
// one thread owns one pair and keeps the whole GJK state in registers:
//   k_gjk2_intersection      GJK::intersect; FullResolution hits go to two queues by shape kind
//   k_epa2<12, small>        EPA2 for pairs with <= 8 hull vertices, ring entirely in shared memory
//   k_epa2<104, large>       EPA2 for everything else (and the small tier's overflows)
//   k_gjk2_distance[_p]      GJK::distance: thread per pair, or persistent lanes for long-tailed tables
//   k_gjk2_time_of_impact    GJK::intersection_time_of_impact
//   k_world_aabbs2, k_lift_aabb2   bounds and the SweepAndPrune2 lift
// EPA2 differs from the reference in one respect that cannot change a bit: the reference recomputes
// EVERY edge's normal and distance (a sqrt and a divide each) after every insertion (closest_edge,
// epa2d.rs:136-157); an edge's (normal, distance) is a pure function of its two end points, and an
// insertion replaces exactly one edge by two, so the kernels cache the distances per edge (the scan
// becomes compares only) and keep the polygon as a linked ring (nothing is shifted).
#include <cstdlib>

#include "cb2_kernels.h"
#include "cb2_math.cuh"

namespace cb2 {

struct f2 {
    float x, y;
};
CB2_DI f2 mk2(float x, float y) {
    f2 r;
    r.x = x;
    r.y = y;
    return r;
}
CB2_DI f2 zero2() { return mk2(0.0f, 0.0f); }
CB2_DI f2 operator+(f2 a, f2 b) { return mk2(fadd(a.x, b.x), fadd(a.y, b.y)); }
CB2_DI f2 operator-(f2 a, f2 b) { return mk2(fsub(a.x, b.x), fsub(a.y, b.y)); }
CB2_DI f2 operator*(f2 a, float s) { return mk2(fmul(a.x, s), fmul(a.y, s)); }
CB2_DI f2 operator/(f2 a, float s) { return mk2(fdiv(a.x, s), fdiv(a.y, s)); }
CB2_DI f2 operator-(f2 a) { return mk2(-a.x, -a.y); }
CB2_DI float dot(f2 a, f2 b) { return fadd(fmul(a.x, b.x), fmul(a.y, b.y)); }
CB2_DI float magnitude2(f2 a) { return dot(a, a); }
CB2_DI float magnitude(f2 a) { return fsqrt(magnitude2(a)); }
CB2_DI f2 normalize_to(f2 a, float m) { return a * fdiv(m, magnitude(a)); }
CB2_DI f2 normalize(f2 a) { return normalize_to(a, 1.0f); }
CB2_DI bool ulps_eq_zero(f2 a) { return ulps_eq_zero(a.x) && ulps_eq_zero(a.y); }
CB2_DI f2 lerp(f2 a, f2 b, float t) { return a + ((b - a) * t); }
// util.rs:42-49
CB2_DI f2 triple_product(f2 a, f2 b, f2 c) {
    const float ac = fadd(fmul(a.x, c.x), fmul(a.y, c.y));
    const float bc = fadd(fmul(b.x, c.x), fmul(b.y, c.y));
    return mk2(fsub(fmul(b.x, ac), fmul(a.x, bc)), fsub(fmul(b.y, ac), fmul(a.y, bc)));
}

// Decomposed<Vector2, Basis2> plus the per-pair invariant inverse matrix (cgmath recomputes
// Matrix2::invert() on every inverse_transform_vector call; same inputs, same bits).
struct xform2 {
    f2 c0, c1;  // columns of the Basis2 matrix
    f2 i0, i1;  // columns of its inverse
    f2 disp;
    float scale;
    uint32_t bad;  // 0, CB2_PANIC_INV_SCALE (scale ~ 0) or CB2_PANIC_INV_ROT (det == 0)
};
CB2_DI f2 mat_mul(f2 c0, f2 c1, f2 v) { return mk2(dot(mk2(c0.x, c1.x), v), dot(mk2(c0.y, c1.y), v)); }
CB2_DI xform2 load_xform2(const float4* t, uint64_t i) {
    const float4 a = __ldg(t + 2 * i), b = __ldg(t + 2 * i + 1);
    xform2 x;
    x.scale = a.x;
    x.c0 = mk2(a.y, a.z);
    x.c1 = mk2(a.w, b.x);
    x.disp = mk2(b.y, b.z);
    const float det = fsub(fmul(x.c0.x, x.c1.y), fmul(x.c1.x, x.c0.y));
    x.i0 = mk2(fdiv(x.c1.y, det), fdiv(-x.c0.y, det));
    x.i1 = mk2(fdiv(-x.c1.x, det), fdiv(x.c0.x, det));
    x.bad = ulps_eq_zero(x.scale) ? (uint32_t)CB2_PANIC_INV_SCALE
                                  : (det == 0.0f ? (uint32_t)CB2_PANIC_INV_ROT : 0u);
    return x;
}
CB2_DI f2 transform_point(const xform2& t, f2 p) { return mat_mul(t.c0, t.c1, p * t.scale) + t.disp; }
CB2_DI f2 inverse_transform_vector(const xform2& t, f2 v) { return mat_mul(t.i0, t.i1, v / t.scale); }

struct shape2 {
    uint32_t kind;
    float p0, p1, p2, p3;
    const float2* verts;
    uint32_t vcnt;
};
CB2_DI shape2 load_shape2(const float4* shapes, const float2* verts, uint32_t i) {
    const float4 a = __ldg(shapes + 2 * (size_t)i), b = __ldg(shapes + 2 * (size_t)i + 1);
    shape2 s;
    s.kind = __float_as_uint(a.x);
    s.p0 = a.y;
    s.p1 = a.z;
    s.p2 = a.w;
    s.p3 = b.x;
    s.verts = verts + __float_as_uint(b.y);
    s.vcnt = __float_as_uint(b.z);
    return s;
}
CB2_DI f2 vtx(const shape2& s, uint32_t i) {
    const float2 v = __ldg(s.verts + i);
    return mk2(v.x, v.y);
}

// polygon.rs:170-184
CB2_DI float dot_index(const shape2& s, int index, f2 d) {
    const uint32_t iu = (uint32_t)index;
    if (iu == s.vcnt) return dot(vtx(s, 0), d);
    if (index == -1) return dot(vtx(s, s.vcnt - 1), d);
    return dot(vtx(s, iu), d);
}

// Primitive2::support_point (primitive2.rs:85-103).  The caller has ruled out t.bad for every
// non-Particle shape (the unwrap()s of util.rs:18, circle.rs:38, line.rs:17, polygon.rs:128).
__device__ __noinline__ f2 support_point2(const shape2& s, const xform2& t, f2 direction) {
    if (s.kind == CB2_PARTICLE2) return transform_point(t, zero2());
    const f2 d = inverse_transform_vector(t, direction);
    f2 best = zero2();
    if (s.kind == CB2_CIRCLE) {  // circle.rs:34-40
        best = normalize_to(d, s.p0);
    } else if (s.kind == CB2_LINE2) {  // line.rs:13-25
        const f2 origin = mk2(s.p0, s.p1), dest = mk2(s.p2, s.p3);
        best = dot(d, dest - origin) >= 0.0f ? dest : origin;
    } else if (s.kind == CB2_RECTANGLE || s.kind == CB2_SQUARE) {
        // rectangle.rs:29-60 corners (+,+) (-,+) (-,-) (+,-); get_max_point, util.rs:11-30
        const f2 h = mk2(s.p0, s.kind == CB2_SQUARE ? s.p0 : s.p1) / 2.0f;
        float bd = -INFINITY;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const f2 c = mk2((k == 1 || k == 2) ? -h.x : h.x, k >= 2 ? -h.y : h.y);
            const float dt = dot(c, d);
            if (dt > bd) {
                bd = dt;
                best = c;
            }
        }
    } else if (s.vcnt < 10) {  // polygon.rs:33-42: get_max_point
        float bd = -INFINITY;
#pragma unroll 3
        for (uint32_t k = 0; k < s.vcnt; ++k) {
            const f2 c = vtx(s, k);
            const float dt = dot(c, d);
            if (dt > bd) {
                bd = dt;
                best = c;
            }
        }
    } else {  // polygon.rs:122-168: the neighbour walk
        const int len = (int)s.vcnt;
        int start_index = 0;
        float max_dot = dot(vtx(s, 0), d);
        if (max_dot < 0.0f) {
            start_index = len / 2;
            max_dot = dot_index(s, start_index, d);
        }
        const float left_dot = dot_index(s, start_index - 1, d);
        const float right_dot = dot_index(s, start_index + 1, d);
        if (start_index == 0 && max_dot > left_dot && max_dot > right_dot) {
            best = vtx(s, 0);
        } else {
            int add = 1;
            float previous_dot = left_dot;
            if (left_dot > max_dot && left_dot > right_dot) {
                add = -1;
                previous_dot = right_dot;
            }
            int index = start_index + add;
            float current_dot = max_dot;
            if (index == len) index = 0;
            if (index == -1) index = len - 1;
            // The walk visits consecutive vertices (mod len) and every step needs only the NEXT
            // vertex's dot; the vertices of the next four steps are fetched together, so the chain
            // of dependent L2 round trips is a quarter as long.  Same dots, same comparisons.
            auto step_of = [&](int i) { i += add; return i == len ? 0 : (i == -1 ? len - 1 : i); };
            float nb0 = 0.0f, nb1 = 0.0f, nb2 = 0.0f, nb3 = 0.0f;
            int avail = 0;
            while (index != start_index) {
                if (avail == 0) {
                    const int i1 = step_of(index), i2 = step_of(i1), i3 = step_of(i2), i4 = step_of(i3);
                    const f2 v1 = vtx(s, (uint32_t)i1), v2 = vtx(s, (uint32_t)i2), v3 = vtx(s, (uint32_t)i3),
                             v4 = vtx(s, (uint32_t)i4);
                    nb0 = dot(v1, d);
                    nb1 = dot(v2, d);
                    nb2 = dot(v3, d);
                    nb3 = dot(v4, d);
                    avail = 4;
                }
                const float next_dot = nb0;  // == dot_index(vertices, index + add, direction)
                if (current_dot > previous_dot && current_dot > next_dot) break;
                previous_dot = current_dot;
                current_dot = next_dot;
                index = step_of(index);
                nb0 = nb1;
                nb1 = nb2;
                nb2 = nb3;
                --avail;
            }
            best = vtx(s, (uint32_t)index);
        }
    }
    return transform_point(t, best);
}

struct sup2 {
    f2 v, a;  // SupportPoint.v, .sup_a (sup_b is never read)
};
// SupportPoint::from_minkowski, minkowski/mod.rs:39-60
CB2_DI sup2 from_minkowski2(const shape2& l, const xform2& lt, const shape2& r, const xform2& rt, f2 dir) {
    sup2 s;
    s.a = support_point2(l, lt, dir);
    s.v = s.a - support_point2(r, rt, -dir);
    return s;
}

// the pair's unwrap() panics, in call order (left support first)
CB2_DI uint32_t pair_panic(const shape2& l, const xform2& lt, const shape2& r, const xform2& rt) {
    if (l.kind != CB2_PARTICLE2 && lt.bad) return lt.bad;
    if (r.kind != CB2_PARTICLE2 && rt.bad) return rt.bad;
    return 0u;
}

// Simplex<Point2>: at most 3 points on this path, statically indexed (registers)
struct simplex2 {
    sup2 p0, p1, p2;
    int len;
};
CB2_DI void sx_push(simplex2& s, const sup2& p) {
    if (s.len == 0) s.p0 = p;
    else if (s.len == 1) s.p1 = p;
    else s.p2 = p;
    ++s.len;
}
// SimplexProcessor2::reduce_to_closest_feature, simplex2d.rs:24-66
CB2_DI bool reduce2(simplex2& s, f2& d) {
    if (s.len == 3) {
        const f2 a = s.p2.v, b = s.p1.v, c = s.p0.v;
        const f2 ao = -a, ab = b - a, ac = c - a;
        const f2 abp = triple_product(ac, ab, ab);
        if (dot(abp, ao) > 0.0f) {
            s.p0 = s.p1;  // remove(0)
            s.p1 = s.p2;
            s.len = 2;
            d = abp;
        } else {
            const f2 acp = triple_product(ab, ac, ac);
            if (dot(acp, ao) > 0.0f) {
                s.p1 = s.p2;  // remove(1)
                s.len = 2;
                d = acp;
            } else {
                return true;
            }
        }
    } else if (s.len == 2) {
        const f2 a = s.p1.v, b = s.p0.v;
        const f2 ao = -a, ab = b - a;
        d = triple_product(ab, ao, ab);
        if (ulps_eq_zero(d)) d = mk2(-ab.y, ab.x);
    }
    return false;
}
// util.rs:98-114 with point = origin
CB2_DI f2 closest_on_edge_to_origin(f2 start, f2 end) {
    const f2 line = end - start;
    const f2 line_dir = normalize(line);
    const f2 v = zero2() - start;
    const float d = dot(v, line_dir);
    if (d < 0.0f) return start;
    if (fmul(d, d) > magnitude2(line)) return end;
    return start + line_dir * d;
}
// simplex2d.rs:71-89
CB2_DI f2 closest_point_to_origin2(simplex2& s) {
    f2 d = zero2();
    if (reduce2(s, d)) return d;
    if (s.len == 1) return s.p0.v;
    return closest_on_edge_to_origin(s.p1.v, s.p0.v);
}

struct narrow2_args_dev {
    const float4* shapes;
    const float2* verts;
    const uint32_t* l_shape;
    const uint32_t* r_shape;
    const float4* l_t;
    const float4* r_t;
    const float4* l_t1;
    const float4* r_t1;
    const uint32_t* pair_body;  // optional 2 x n body indices: then l_shape / l_t are per BODY
    uint64_t n;
    cb2_settings settings;
    uint32_t iter_cap;
    uint32_t strategy;
    cb2_contact2* out_contact;
    float* out_distance;
    uint8_t* out_status;
    // GJK -> EPA2 hand-off (FullResolution intersection only)
    uint32_t* counters;  // [0] hits; [2],[3] small queue length, cursor; [4],[5] large queue; zeroed per batch
    uint32_t* hits;      // pair index per hit (slot)
    uint32_t* queue_small;  // hit slots for the small-polygon EPA2 tier
    uint32_t* queue_large;  // ... and for the 104-point tier (also fed by small-tier overflows)
    float4* simplex;     // 3 x (v.xy, sup_a.xy) per hit
};

CB2_DI void store_contact2(cb2_contact2* out, uint64_t i, uint32_t strategy, f2 n, float depth, f2 pt,
                           float toi) {
    float* o = reinterpret_cast<float*>(out + i);
    o[0] = __uint_as_float(strategy);
    o[1] = n.x;
    o[2] = n.y;
    o[3] = depth;
    o[4] = pt.x;
    o[5] = pt.y;
    o[6] = toi;
}

// approx ulps_eq!(inf, x) (epa2d.rs:154 assert_ulps_ne!): x is +inf or within 4 ulps below it
CB2_DI bool ulps_eq_inf(float x) { return x > 0.0f && __float_as_uint(x) >= 0x7f800000u - 4u && x == x; }

constexpr int EPA2_CAP = 3 + (int)CB2_MAX_ITERATIONS_2D + 1;

// edge (a -> b) of the EPA2 polygon: normal.xy, distance (epa2d.rs:141-147)
CB2_DI float4 edge_of(f2 a, f2 b) {
    const f2 e = b - a;
    const f2 n = normalize(triple_product(e, a, e));
    return make_float4(n.x, n.y, dot(n, a), 0.0f);
}

// pair i of an intersection batch: direct (shape / transform per pair and side) or through bodies
// (pair_body[2i], pair_body[2i+1] index one body table: SweepAndPrune2 output feeding the narrow phase)
CB2_DI void load_pair2(const narrow2_args_dev& A, uint64_t i, shape2& l, shape2& r, xform2& lt, xform2& rt) {
    if (A.pair_body) {
        const uint2 b = __ldg(reinterpret_cast<const uint2*>(A.pair_body) + i);
        l = load_shape2(A.shapes, A.verts, __ldg(A.l_shape + b.x));
        r = load_shape2(A.shapes, A.verts, __ldg(A.l_shape + b.y));
        lt = load_xform2(A.l_t, b.x);
        rt = load_xform2(A.l_t, b.y);
    } else {
        l = load_shape2(A.shapes, A.verts, __ldg(A.l_shape + i));
        r = load_shape2(A.shapes, A.verts, __ldg(A.r_shape + i));
        lt = load_xform2(A.l_t, i);
        rt = load_xform2(A.r_t, i);
    }
}

// vertices a shape can contribute to the Minkowski difference's hull (a circle: unbounded)
CB2_DI uint32_t hull_vertices(const shape2& s) {
    return s.kind == CB2_PARTICLE2 ? 1u
         : s.kind == CB2_LINE2 ? 2u
         : (s.kind == CB2_RECTANGLE || s.kind == CB2_SQUARE) ? 4u
         : s.kind == CB2_POLYGON ? s.vcnt : 1000u;
}

// GJK2::intersection (gjk/mod.rs:362-391), stage 1: GJK::intersect (:101-146), one thread per pair.
// Misses, panics and CollisionOnly hits are final here; a FullResolution hit hands its 3-point
// simplex to k_epa2 through the hit list (one warp-aggregated reservation per warp).
__global__ void __launch_bounds__(128) k_gjk2_intersection(narrow2_args_dev A) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = i < A.n;
    bool to_epa = false, small = false;
    simplex2 s;
    s.len = 0;
    if (valid) {
        shape2 l, r;
        xform2 lt, rt;
        load_pair2(A, i, l, r, lt, rt);
        small = hull_vertices(l) + hull_vertices(r) <= 8u;
        uint8_t status = CB2_NONE;
        const uint32_t panic = pair_panic(l, lt, r, rt);
        if (panic) {
            status = (uint8_t)panic;
        } else {
            f2 d = transform_point(rt, zero2()) - transform_point(lt, zero2());
            if (ulps_eq_zero(d)) d = mk2(1.0f, 1.0f);
            sup2 a = from_minkowski2(l, lt, r, rt, d);
            bool hit = false;
            if (!(dot(a.v, d) <= 0.0f)) {
                sx_push(s, a);
                d = -d;
                for (uint32_t it = 0; it < A.settings.max_iterations; ++it) {
                    a = from_minkowski2(l, lt, r, rt, d);
                    if (dot(a.v, d) <= 0.0f) break;
                    sx_push(s, a);
                    if (reduce2(s, d)) {
                        hit = true;
                        break;
                    }
                }
            }
            if (hit) status = CB2_SOME;
            to_epa = hit && A.strategy == CB2_FULL_RESOLUTION;
        }
        if (!to_epa) {
            store_contact2(A.out_contact, i, A.strategy, zero2(), 0.0f, zero2(), 0.0f);
            A.out_status[i] = status;
        }
    }
    const unsigned m = __ballot_sync(0xffffffffu, to_epa);
    if (m) {
        const int lane = threadIdx.x & 31;
        uint32_t base = 0;
        if (lane == __ffs(m) - 1) base = atomicAdd(&A.counters[0], (uint32_t)__popc(m));
        base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
        uint32_t slot = 0;
        if (to_epa) {
            slot = base + __popc(m & ((1u << lane) - 1u));
            A.hits[slot] = (uint32_t)i;
            float4* sx = A.simplex + 3 * (size_t)slot;
            sx[0] = make_float4(s.p0.v.x, s.p0.v.y, s.p0.a.x, s.p0.a.y);
            sx[1] = make_float4(s.p1.v.x, s.p1.v.y, s.p1.a.x, s.p1.a.y);
            sx[2] = make_float4(s.p2.v.x, s.p2.v.y, s.p2.a.x, s.p2.a.y);
        }
        // two EPA2 queues: pairs whose Minkowski difference has at most 8 vertices (rectangles,
        // squares, lines, particles, tiny polygons) go to the small-polygon tier
        const unsigned ms = __ballot_sync(0xffffffffu, to_epa && small);
        const unsigned ml = m & ~ms;
        uint32_t bs = 0, bl = 0;
        if (ms && lane == __ffs(ms) - 1) bs = atomicAdd(&A.counters[2], (uint32_t)__popc(ms));
        if (ml && lane == __ffs(ml) - 1) bl = atomicAdd(&A.counters[4], (uint32_t)__popc(ml));
        if (ms) bs = __shfl_sync(0xffffffffu, bs, __ffs(ms) - 1);
        if (ml) bl = __shfl_sync(0xffffffffu, bl, __ffs(ml) - 1);
        if (to_epa) {
            if (small) A.queue_small[bs + __popc(ms & ((1u << lane) - 1u))] = slot;
            else A.queue_large[bl + __popc(ml & ((1u << lane) - 1u))] = slot;
        }
    }
}

// Stage 2: EPA2::process (epa2d.rs:27-67).  Persistent lanes: a lane owns one hit at a time, every
// round of the loop is ONE EPA iteration for every lane of the warp (closest edge, one Minkowski
// support, insert), and a lane whose pair finished pulls the next hit from the device queue — the
// 100-iteration pairs (circles) no longer hold 31 idle lanes hostage.
//
// The reference keeps the polygon as a Vec and inserts in the middle (simplex.insert(index, p),
// :53); here the polygon is a linked ring over append-only storage, so nothing is ever shifted:
//   P[k]    point k in insertion order (v.xy, sup_a.xy)                       — local memory
//   nx[k]   the point that follows k around the ring                          — local memory
//   D[k]    distance of the edge k -> nx[k] (its normal is recomputed for the — shared memory,
//           one winning edge; edge_of is a pure function of the end points)     [k][thread]
// closest_edge's "first strictly smallest distance in Vec order" is the minimum over D in ANY
// order unless two edges tie exactly; the scan counts ties and only then walks the ring from the
// head (the point at Vec index 0: an insertion on the wrap-around edge becomes the new head).
// Two tiers.  SMALL (12 points): pairs whose Minkowski difference has at most 8 hull vertices —
// points, successors and distances all in shared memory ([k][thread], 252 B per lane: 7 blocks per
// SM, no local memory at all); a polygon that outgrows it anyway (near-duplicate support points)
// is re-queued for the large tier and restarts there from its simplex.  LARGE (104 points: the
// 100-iteration cap): distances in shared memory, points and successors in local memory.
constexpr int EPA2_THREADS = 128;
constexpr int EPA2_SMALL_CAP = 12;
template <int CAP, bool SMALL>
__global__ void __launch_bounds__(EPA2_THREADS) k_epa2(narrow2_args_dev A) {
    extern __shared__ float4 smem4[];
    // SMALL: float4 P[CAP][T] | float D[CAP][T] | uint8 nx[CAP][T];   LARGE: float D[CAP][T]
    float4* Ps = smem4 + threadIdx.x;
    float* D = (SMALL ? reinterpret_cast<float*>(smem4 + CAP * EPA2_THREADS) : reinterpret_cast<float*>(smem4)) +
               threadIdx.x;
    uint8_t* nxs = reinterpret_cast<uint8_t*>(reinterpret_cast<float*>(smem4 + CAP * EPA2_THREADS) +
                                              CAP * EPA2_THREADS) + threadIdx.x;
    float4 Pl[SMALL ? 1 : CAP];
    uint8_t nxl[SMALL ? 1 : CAP];
    auto P = [&](int q) -> float4& { return SMALL ? Ps[q * EPA2_THREADS] : Pl[SMALL ? 0 : q]; };
    auto nx = [&](int q) -> uint8_t& { return SMALL ? nxs[q * EPA2_THREADS] : nxl[SMALL ? 0 : q]; };
    const int lane = threadIdx.x & 31;
    const uint32_t* queue = SMALL ? A.queue_small : A.queue_large;
    const uint32_t n_hits = A.counters[SMALL ? 2 : 4];
    uint32_t* cursor = &A.counters[SMALL ? 3 : 5];
    enum { FETCH, RUN, EXIT };
    int state = FETCH;
    int len = 0, head = 0;
    uint32_t pair = 0, k = 0, slot = 0;
    shape2 l, r;
    xform2 lt, rt;
    l.kind = r.kind = CB2_PARTICLE2;
    l.p0 = l.p1 = l.p2 = l.p3 = r.p0 = r.p1 = r.p2 = r.p3 = 0.0f;
    l.verts = r.verts = A.verts;
    l.vcnt = r.vcnt = 0;
    lt = load_xform2(A.l_t, 0);
    rt = lt;
    for (;;) {
        // ---- refill ----
        const unsigned need = __ballot_sync(0xffffffffu, state == FETCH);
        if (need) {
            uint32_t base = 0;
            if (lane == __ffs(need) - 1) base = atomicAdd(cursor, (uint32_t)__popc(need));
            base = __shfl_sync(0xffffffffu, base, __ffs(need) - 1);
            if (state == FETCH) {
                const uint32_t qi = base + __popc(need & ((1u << lane) - 1u));
                if (qi >= n_hits) {
                    state = EXIT;
                    len = 0;
                } else {
                    slot = queue[qi];
                    pair = A.hits[slot];
                    load_pair2(A, pair, l, r, lt, rt);
                    const float4* sx = A.simplex + 3 * (size_t)slot;
                    const float4 a = sx[0], b = sx[1], c = sx[2];
                    P(0) = a;
                    P(1) = b;
                    P(2) = c;
                    nx(0) = 1;
                    nx(1) = 2;
                    nx(2) = 0;
                    D[0 * EPA2_THREADS] = edge_of(mk2(a.x, a.y), mk2(b.x, b.y)).z;
                    D[1 * EPA2_THREADS] = edge_of(mk2(b.x, b.y), mk2(c.x, c.y)).z;
                    D[2 * EPA2_THREADS] = edge_of(mk2(c.x, c.y), mk2(a.x, a.y)).z;
                    len = 3;
                    head = 0;
                    k = 0;
                    state = RUN;
                }
            }
        }
        if (__all_sync(0xffffffffu, state == EXIT)) break;
        const bool run = state == RUN;
        // ---- closest_edge (:136-157) over the cached distances ----
        float dmin = INFINITY;
        int arg = 0, ties = 0;
        {
            const int wlen = __reduce_max_sync(0xffffffffu, run ? len : 0);
            for (int q = 0; q < wlen; ++q) {
                if (q < len) {
                    const float d = D[q * EPA2_THREADS];
                    if (d < dmin) {
                        dmin = d;
                        arg = q;
                        ties = 1;
                    } else if (d == dmin) {
                        ++ties;
                    }
                }
            }
        }
        if (run && ties > 1) {  // exact tie: the reference takes the first in Vec order
            int q = head;
            while (!(D[q * EPA2_THREADS] == dmin)) q = nx(q);
            arg = q;
            dmin = D[q * EPA2_THREADS];  // -0.0 == +0.0: keep the winner's own bits
        }
        const bool panicked = run && ulps_eq_inf(dmin);  // assert_ulps_ne!(inf, edge.distance)
        const bool step = run && !panicked && k < A.settings.max_iterations;
        bool finished = run && !step;
        // the winning edge a -> b (b is the reference's edge.index)
        const int ea = arg, eb = run ? nx(arg) : 0;
        float4 pa4 = make_float4(0.f, 0.f, 0.f, 0.f), pb4 = pa4;
        f2 en = zero2();
        if (run && !panicked) {
            pa4 = P(ea);
            pb4 = P(eb);
            const float4 e = edge_of(mk2(pa4.x, pa4.y), mk2(pb4.x, pb4.y));
            en = mk2(e.x, e.y);
        }
        if (step) {
            ++k;
            const sup2 p = from_minkowski2(l, lt, r, rt, en);
            if (fsub(dot(p.v, en), dmin) < A.settings.epa_tolerance) {
                finished = true;
            } else if (SMALL && len >= CAP) {
                // outgrew the small tier: the large tier redoes this pair from its simplex
                A.queue_large[atomicAdd(&A.counters[4], 1u)] = slot;
                state = FETCH;
                len = 0;
            } else {
                // simplex.insert(index, p): p goes between a and b
                P(len) = make_float4(p.v.x, p.v.y, p.a.x, p.a.y);
                nx(len) = (uint8_t)eb;
                nx(ea) = (uint8_t)len;
                // insertion at Vec index 0 (the wrap-around edge tail -> head): p is the new head
                if (eb == head) head = len;
                D[ea * EPA2_THREADS] = edge_of(mk2(pa4.x, pa4.y), p.v).z;
                D[len * EPA2_THREADS] = edge_of(p.v, mk2(pb4.x, pb4.y)).z;
                ++len;
            }
        }
        if (finished) {
            if (panicked) {
                store_contact2(A.out_contact, pair, CB2_FULL_RESOLUTION, zero2(), 0.0f, zero2(), 0.0f);
                A.out_status[pair] = (uint8_t)CB2_PANIC_EPA2_EDGE;
            } else {
                // point(), :87-110
                const f2 av = mk2(pa4.x, pa4.y), bv = mk2(pb4.x, pb4.y);
                const f2 asup = mk2(pa4.z, pa4.w), bsup = mk2(pb4.z, pb4.w);
                const f2 ab = bv - av;
                const float t = fdiv(dot(-av, ab), magnitude2(ab));
                const f2 pt = t < 0.0f ? asup : (t > 1.0f ? bsup : asup + (bsup - asup) * t);
                store_contact2(A.out_contact, pair, CB2_FULL_RESOLUTION, en, dmin, pt, 0.0f);
                A.out_status[pair] = CB2_SOME;
            }
            state = FETCH;
            len = 0;
        }
    }
}

template <int CAP, bool SMALL>
static cudaError_t launch_epa2(const narrow2_args_dev& A, uint64_t n, int sm_count, cudaStream_t st) {
    auto kern = k_epa2<CAP, SMALL>;
    const size_t smem = (size_t)CAP * EPA2_THREADS * (SMALL ? sizeof(float4) + sizeof(float) + 1 : sizeof(float));
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, EPA2_THREADS, smem) != cudaSuccess ||
        per_sm < 1)
        per_sm = 1;
    unsigned eb = (unsigned)(per_sm * sm_count);
    const unsigned want = (unsigned)((n + EPA2_THREADS - 1) / EPA2_THREADS);
    if (eb > want) eb = want;
    kern<<<eb, EPA2_THREADS, smem, st>>>(A);
    return cudaGetLastError();
}

// GJK2::distance, gjk/mod.rs:271-321
__global__ void __launch_bounds__(128) k_gjk2_distance(narrow2_args_dev A) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= A.n) return;
    const shape2 l = load_shape2(A.shapes, A.verts, __ldg(A.l_shape + i));
    const shape2 r = load_shape2(A.shapes, A.verts, __ldg(A.r_shape + i));
    const xform2 lt = load_xform2(A.l_t, i), rt = load_xform2(A.r_t, i);
    uint8_t status = CB2_NONE;
    float out = 0.0f;
    const uint32_t panic = pair_panic(l, lt, r, rt);
    if (panic) {
        status = (uint8_t)panic;
    } else {
        f2 d = transform_point(rt, zero2()) - transform_point(lt, zero2());
        if (ulps_eq_zero(d)) d = mk2(1.0f, 1.0f);
        simplex2 s;
        s.len = 0;
        sx_push(s, from_minkowski2(l, lt, r, rt, d));
        sx_push(s, from_minkowski2(l, lt, r, rt, -d));
        for (uint32_t it = 0; it < A.settings.max_iterations; ++it) {
            f2 dd = closest_point_to_origin2(s);
            if (ulps_eq_zero(dd)) break;
            dd = -dd;
            const sup2 p = from_minkowski2(l, lt, r, rt, dd);
            if (fsub(dot(p.v, dd), dot(s.p0.v, dd)) < A.settings.distance_tolerance) {
                out = magnitude(dd);
                status = CB2_SOME;
                break;
            }
            sx_push(s, p);
        }
    }
    A.out_distance[i] = out;
    A.out_status[i] = status;
}

// GJK2::distance with persistent lanes: every round of the loop is ONE Minkowski support for every
// lane (the two seed supports of gjk/mod.rs:296-297 and the loop body's support share the slot), and
// a lane whose pair finished pulls the next pair.  Circle pairs routinely run to the 100-iteration
// cap while boxes finish in 2-4 iterations; without the refill one such lane holds 31 others.
__global__ void __launch_bounds__(128) k_gjk2_distance_p(narrow2_args_dev A) {
    const int lane = threadIdx.x & 31;
    enum { FETCH, SEED1, SEED2, ITER, EXIT };
    int state = FETCH;
    uint64_t i = 0;
    uint32_t it = 0;
    shape2 l, r;
    xform2 lt, rt;
    l.kind = r.kind = CB2_PARTICLE2;
    l.p0 = l.p1 = l.p2 = l.p3 = r.p0 = r.p1 = r.p2 = r.p3 = 0.0f;
    l.verts = r.verts = A.verts;
    l.vcnt = r.vcnt = 0;
    lt = load_xform2(A.l_t, 0);
    rt = lt;
    simplex2 s;
    s.len = 0;
    s.p0.v = s.p0.a = s.p1.v = s.p1.a = s.p2.v = s.p2.a = zero2();
    f2 d0 = zero2();
    for (;;) {
        const unsigned need = __ballot_sync(0xffffffffu, state == FETCH);
        if (need) {
            unsigned long long base = 0;
            if (lane == __ffs(need) - 1) base = atomicAdd(&A.counters[6], (uint32_t)__popc(need));
            base = __shfl_sync(0xffffffffu, base, __ffs(need) - 1);
            if (state == FETCH) {
                i = base + __popc(need & ((1u << lane) - 1u));
                if (i >= A.n) {
                    state = EXIT;
                } else {
                    l = load_shape2(A.shapes, A.verts, __ldg(A.l_shape + i));
                    r = load_shape2(A.shapes, A.verts, __ldg(A.r_shape + i));
                    lt = load_xform2(A.l_t, i);
                    rt = load_xform2(A.r_t, i);
                    const uint32_t panic = pair_panic(l, lt, r, rt);
                    if (panic) {
                        A.out_distance[i] = 0.0f;
                        A.out_status[i] = (uint8_t)panic;  // stays in FETCH: refilled next round
                    } else {
                        d0 = transform_point(rt, zero2()) - transform_point(lt, zero2());
                        if (ulps_eq_zero(d0)) d0 = mk2(1.0f, 1.0f);
                        s.len = 0;
                        state = SEED1;
                    }
                }
            }
        }
        if (__all_sync(0xffffffffu, state == EXIT)) break;
        // ---- the direction of this round's support ----
        f2 dir = zero2();
        bool sup = false, done = false;
        uint8_t status = CB2_NONE;
        float out = 0.0f;
        if (state == SEED1) {
            dir = d0;
            sup = true;
        } else if (state == SEED2) {
            dir = -d0;
            sup = true;
        } else if (state == ITER) {
            const f2 dd = closest_point_to_origin2(s);
            if (ulps_eq_zero(dd)) {
                done = true;  // None: the shapes touch or overlap
            } else {
                dir = -dd;
                sup = true;
            }
        }
        sup2 p;
        p.v = p.a = zero2();
        if (sup) p = from_minkowski2(l, lt, r, rt, dir);
        if (sup) {
            if (state == SEED1) {
                sx_push(s, p);
                state = SEED2;
            } else if (state == SEED2) {
                sx_push(s, p);
                it = 0;
                state = ITER;
                if (A.settings.max_iterations == 0) done = true;
            } else {
                if (fsub(dot(p.v, dir), dot(s.p0.v, dir)) < A.settings.distance_tolerance) {
                    out = magnitude(dir);
                    status = CB2_SOME;
                    done = true;
                } else {
                    sx_push(s, p);
                    if (++it >= A.settings.max_iterations) done = true;
                }
            }
        }
        if (done) {
            A.out_distance[i] = out;
            A.out_status[i] = status;
            state = FETCH;
        }
    }
}

// GJK2::intersection_time_of_impact, gjk/mod.rs:163-256 (+ iter_cap on the uncapped loop)
__global__ void __launch_bounds__(128) k_gjk2_time_of_impact(narrow2_args_dev A) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= A.n) return;
    const shape2 l = load_shape2(A.shapes, A.verts, __ldg(A.l_shape + i));
    const shape2 r = load_shape2(A.shapes, A.verts, __ldg(A.r_shape + i));
    const xform2 lt = load_xform2(A.l_t, i), rt = load_xform2(A.r_t, i);
    const xform2 lt1 = load_xform2(A.l_t1, i), rt1 = load_xform2(A.r_t1, i);
    uint8_t status = CB2_NONE;
    f2 out_n = zero2(), out_p = zero2();
    float out_d = 0.0f, out_toi = 0.0f;
    const uint32_t panic = pair_panic(l, lt, r, rt);
    if (panic) {
        status = (uint8_t)panic;
    } else {
        const f2 left_lin_vel = transform_point(lt1, zero2()) - transform_point(lt, zero2());
        const f2 right_lin_vel = transform_point(rt1, zero2()) - transform_point(rt, zero2());
        const f2 ray = right_lin_vel - left_lin_vel;
        float lambda = 0.0f;
        f2 normal = zero2(), ray_origin = zero2();
        simplex2 s;
        s.len = 0;
        f2 v = from_minkowski2(l, lt, r, rt, -ray).v;
        uint32_t it = 0;
        bool done = false;
        while (magnitude2(v) > A.settings.continuous_tolerance) {
            if (it >= A.iter_cap) {
                status = CB2_NONCONVERGED;
                done = true;
                break;
            }
            ++it;
            sup2 p = from_minkowski2(l, lt, r, rt, -v);
            const float vp = dot(v, p.v), vr = dot(v, ray);
            if (vp > fmul(lambda, vr)) {
                if (vr > 0.0f) {
                    lambda = fdiv(vp, vr);
                    if (lambda > 1.0f) {
                        done = true;
                        break;
                    }
                    ray_origin = ray * lambda;
                    s.len = 0;
                    normal = -v;
                } else {
                    done = true;
                    break;
                }
            }
            p.v = p.v - ray_origin;
            sx_push(s, p);
            v = closest_point_to_origin2(s);
        }
        if (!done && magnitude2(v) <= A.settings.continuous_tolerance) {
            // translation_interpolate (traits.rs:233-246): rot and scale of the END transform
            xform2 t = rt1;
            t.disp = lerp(rt.disp, rt1.disp, lambda);
            out_n = -normalize(normal);
            out_d = magnitude(v);
            out_p = transform_point(t, ray_origin);
            out_toi = lambda;
            status = CB2_SOME;
        }
    }
    store_contact2(A.out_contact, i, CB2_FULL_RESOLUTION, out_n, out_d, out_p, out_toi);
    A.out_status[i] = status;
}

// ---- world bounds: shape.compute_bound().transform_volume(&t) for Primitive2 ---------------------------
// ComputeBound<Aabb2>: primitive2.rs:68-80 (circle.rs:47-52, rectangle.rs:83-88, line.rs:32-34,
// polygon.rs:50-52 = fold from Aabb2::zero()); Aabb2::transform: aabb2.rs:39-46, 78-88.
CB2_DI float rmin2(float l, float r) { return (l > r) ? r : l; }  // aabb/mod.rs:21-26
CB2_DI float rmax2(float l, float r) { return (l < r) ? r : l; }  // aabb/mod.rs:28-33
__global__ void __launch_bounds__(256) k_world_aabbs2(const float4* __restrict__ shapes,
                                                      const float2* __restrict__ verts,
                                                      const uint32_t* __restrict__ shape,
                                                      const float4* __restrict__ t, uint64_t n,
                                                      cb2_aabb2* __restrict__ out) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const shape2 s = load_shape2(shapes, verts, __ldg(shape + i));
    const xform2 x = load_xform2(t, i);
    f2 mn = zero2(), mx = zero2();
    auto ordered = [&](f2 a, f2 b) {  // Aabb2::new
        mn = mk2(rmin2(a.x, b.x), rmin2(a.y, b.y));
        mx = mk2(rmax2(a.x, b.x), rmax2(a.y, b.y));
    };
    auto grow = [&](f2 p) {  // Aabb::grow: new(min(self.min, p), max(self.max, p))
        ordered(mk2(rmin2(mn.x, p.x), rmin2(mn.y, p.y)), mk2(rmax2(mx.x, p.x), rmax2(mx.y, p.y)));
    };
    if (s.kind == CB2_LINE2) {
        ordered(mk2(s.p0, s.p1), mk2(s.p2, s.p3));
    } else if (s.kind == CB2_CIRCLE) {
        ordered(mk2(-s.p0, -s.p0), mk2(s.p0, s.p0));
    } else if (s.kind == CB2_RECTANGLE || s.kind == CB2_SQUARE) {
        const f2 h = mk2(s.p0, s.kind == CB2_SQUARE ? s.p0 : s.p1) / 2.0f;
        ordered(-h, h);
    } else if (s.kind == CB2_POLYGON) {
        for (uint32_t k = 0; k < s.vcnt; ++k) grow(vtx(s, k));
    }  // Particle: Aabb2::zero()
    const f2 c1 = mk2(mx.x, mn.y), c2 = mk2(mn.x, mx.y), c0 = mn, c3 = mx;
    const f2 first = transform_point(x, c0);
    ordered(first, first);
    grow(transform_point(x, c1));
    grow(transform_point(x, c2));
    grow(transform_point(x, c3));
    out[i].min[0] = mn.x;
    out[i].min[1] = mn.y;
    out[i].max[0] = mx.x;
    out[i].max[1] = mx.y;
}
cudaError_t launch_world_aabbs2(const cb2_shape2* shapes, const float* verts, const uint32_t* shape,
                                const cb2_xform2* t, uint64_t n, cb2_aabb2* out, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    k_world_aabbs2<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(
        reinterpret_cast<const float4*>(shapes), reinterpret_cast<const float2*>(verts), shape,
        reinterpret_cast<const float4*>(t), n, out);
    return cudaGetLastError();
}

// SweepAndPrune2 through the 3-D machinery: boxes lifted to z in [0, 1] (the strict z test always holds)
__global__ void __launch_bounds__(256) k_lift_aabb2(const cb2_aabb2* __restrict__ in, uint64_t n,
                                                    cb2_aabb3* __restrict__ out) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 b = __ldg(reinterpret_cast<const float4*>(in) + i);
    out[i].min[0] = b.x;
    out[i].min[1] = b.y;
    out[i].min[2] = 0.0f;
    out[i].max[0] = b.z;
    out[i].max[1] = b.w;
    out[i].max[2] = 1.0f;
}
cudaError_t launch_lift_aabb2(const cb2_aabb2* in, uint64_t n, cb2_aabb3* out, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    k_lift_aabb2<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(in, n, out);
    return cudaGetLastError();
}

cudaError_t launch_narrow2(int op, const narrow2_args& H, int sm_count, cudaStream_t st) {
    narrow2_args_dev A;
    A.shapes = reinterpret_cast<const float4*>(H.shapes);
    A.verts = reinterpret_cast<const float2*>(H.verts);
    A.l_shape = H.l_shape;
    A.r_shape = H.r_shape;
    A.l_t = reinterpret_cast<const float4*>(H.l_t);
    A.r_t = reinterpret_cast<const float4*>(H.r_t);
    A.l_t1 = reinterpret_cast<const float4*>(H.l_t1);
    A.r_t1 = reinterpret_cast<const float4*>(H.r_t1);
    A.pair_body = H.pair_body;
    A.n = H.n;
    A.settings = H.settings;
    A.iter_cap = H.iter_cap;
    A.strategy = H.strategy;
    A.out_contact = H.out_contact;
    A.out_distance = H.out_distance;
    A.out_status = H.out_status;
    A.counters = H.counters;
    A.hits = H.hits;
    A.queue_small = H.hits ? H.hits + H.n : nullptr;
    A.queue_large = H.hits ? H.hits + 2 * H.n : nullptr;
    A.simplex = reinterpret_cast<float4*>(H.simplex);
    if (H.n == 0) return cudaSuccess;
    const unsigned blocks = (unsigned)((H.n + 127) / 128);
    if (op == 0) {
        const bool full = H.strategy == CB2_FULL_RESOLUTION;
        if (full) {
            cudaError_t e = cudaMemsetAsync(H.counters, 0, 8 * sizeof(uint32_t), st);
            if (e != cudaSuccess) return e;
        }
        k_gjk2_intersection<<<blocks, 128, 0, st>>>(A);
        if (full) {
            // persistent lanes refilling from the two hit queues; the large tier runs second because
            // the small tier may hand pairs over
            cudaError_t e = launch_epa2<EPA2_SMALL_CAP, true>(A, H.n, sm_count, st);
            if (e != cudaSuccess) return e;
            e = launch_epa2<EPA2_CAP, false>(A, H.n, sm_count, st);
            if (e != cudaSuccess) return e;
        }
    } else if (op == 1) {
        static const bool plain = std::getenv("CB2_NO_REFILL") && std::getenv("CB2_NO_REFILL")[0] == '1';
        if (plain || !H.counters || !H.long_tails) {
            k_gjk2_distance<<<blocks, 128, 0, st>>>(A);
        } else {
            cudaError_t e = cudaMemsetAsync(H.counters, 0, 8 * sizeof(uint32_t), st);
            if (e != cudaSuccess) return e;
            int per_sm = 0;
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_gjk2_distance_p, 128, 0) != cudaSuccess ||
                per_sm < 1)
                per_sm = 1;
            unsigned pb = (unsigned)(per_sm * sm_count);
            if (pb > blocks) pb = blocks;
            k_gjk2_distance_p<<<pb, 128, 0, st>>>(A);
        }
    }
    else k_gjk2_time_of_impact<<<blocks, 128, 0, st>>>(A);
    return cudaGetLastError();
}

}  // namespace cb2
