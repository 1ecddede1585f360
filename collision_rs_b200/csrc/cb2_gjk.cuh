// cb2_gjk.cuh — per-thread GJK3 / SimplexProcessor3 / EPA3 building blocks.
//
// Not a translation of the reference's container code: the simplex lives in
// registers with compile-time indices (SmallVec::remove/swap become fixed register
// moves), SimplexProcessor3's four call sites of check_side are flattened into ONE
// predicated check (less divergence), support functions hoist per-pair invariants,
// and the Cuboid corner scan uses 9 flops instead of 40 by exploiting the sign
// symmetry of IEEE rounding — all while producing the same bits as the reference.
#pragma once
#include "cb2_math.cuh"
#include "cb2_cells.cuh"
#include "../../include/collision_b200.h"

namespace cb2 {

enum : uint32_t {
    K_SPHERE = 0, K_CUBOID = 1, K_POLY = 2, K_PARTICLE = 3, K_QUAD = 4, K_CYLINDER = 5, K_CAPSULE = 6,
    K_POLY_HE = 8
};  // a Cube (7) is uploaded as a Cuboid
enum : uint8_t {
    ST_NONE = 0,
    ST_SOME = 1,
    ST_PANIC_UNWRAP_PLANE = 2,
    ST_PANIC_ASSERT_RANGE = 3,
    ST_PANIC_INV_SCALE = 4,
    ST_NONCONVERGED = 5,
    ST_PANIC_EPA_EMPTY = 6,
    ST_SCRATCH_OVERFLOW = 7,
    ST_PANIC_CMP_NAN = 8,
    ST_BAD_INDEX = 11  // (9 and 10 belong to the 2-D path)
};

// Device shape-table entry (32 B, two 128-bit loads).  h = half_dim for a Cuboid
// (dim / 2, cuboid.rs:36 — exact) and (x, y) of a Quad, h.x = radius for a Sphere, (h.x, h.y) =
// (half_height, radius) for a Cylinder / Capsule, h.x = max |vertex| (rounded up) for a
// ConvexPolyhedron.  cell_off / cell_cfg locate the polyhedron's support cells
// (cb2_cells.cuh): cell_cfg = N | wide << 8.
struct __align__(16) dshape {
    uint32_t kind;
    float hx, hy, hz;
    uint32_t voff, vcnt, cell_off, cell_cfg;
};

// the device-resident shape table every narrow-phase kernel reads
struct shape_table {
    const dshape* shapes;
    poly_tables poly;      // vertices (+ HalfEdge ring records behind), support cells, extension pool
};

struct shape {
    uint32_t kind;
    f3 h;                   // poly: h.x = R
    poly_ref poly;
    mutable uint32_t hang;  // set when a support call hit the reference's endless hill climb
};

CB2_DI shape load_shape(const shape_table& T, uint32_t idx) {
    const float4* p = reinterpret_cast<const float4*>(T.shapes + idx);
    const float4 a = __ldg(p);
    const float4 b = __ldg(p + 1);
    shape s;
    s.kind = __float_as_uint(a.x);
    s.h = mk3(a.y, a.z, a.w);
    s.poly.voff = __float_as_uint(b.x);
    s.poly.vcnt = __float_as_uint(b.y);
    s.poly.cell_off = __float_as_uint(b.z);
    s.poly.cfg = __float_as_uint(b.w);
    s.hang = 0u;
    return s;
}

// ---- Primitive::support_point, local part (direction already in model space) ----------
// Cuboid: util.rs:11-30 get_max_point over corners in cuboid.rs:54-65 order
//   (+++)(-++)(--+)(+-+)(++-)(-+-)(---)(+--), first strictly greatest dot, start (origin,-inf).
// With px=hx*dx, py=hy*dy, pz=hz*dz the eight dots (cx*dx + cy*dy) + cz*dz are
//   d0=(px+py)+pz  d1=(py-px)+pz  d2=pz-(px+py)  d3=pz-(py-px)  d4=-d2 d5=-d3 d6=-d0 d7=-d1
// bit-for-bit, because negation commutes with round-to-nearest.
CB2_DI f3 cuboid_local_support(f3 h, f3 d) {
    const float px = fmul(h.x, d.x), py = fmul(h.y, d.y), pz = fmul(h.z, d.z);
    const float a = fadd(px, py);
    const float b = fsub(py, px);
    const float d0 = fadd(a, pz), d1 = fadd(b, pz), d2 = fsub(pz, a), d3 = fsub(pz, b);
    float best = -INFINITY;
    int idx = -1;
    if (d0 > best) { best = d0; idx = 0; }
    if (d1 > best) { best = d1; idx = 1; }
    if (d2 > best) { best = d2; idx = 2; }
    if (d3 > best) { best = d3; idx = 3; }
    if (-d2 > best) { best = -d2; idx = 4; }
    if (-d3 > best) { best = -d3; idx = 5; }
    if (-d0 > best) { best = -d0; idx = 6; }
    if (-d1 > best) { best = -d1; idx = 7; }
    if (idx < 0) return zero3();  // every dot NaN / -inf: P::origin()
    // sign masks of the corner table: x negative for {1,2,5,6}, y for {2,3,6,7}, z for {4..7}
    const bool nx = (0x66 >> idx) & 1, ny = (0xCC >> idx) & 1, nz = idx >= 4;
    return mk3(nx ? -h.x : h.x, ny ? -h.y : h.y, nz ? -h.z : h.z);
}
// Sphere: sphere.rs:32-38
CB2_DI f3 sphere_local_support(float r, f3 d) { return normalize_to(d, r); }
// Quad: quad.rs:46-57,65-77 — get_max_point over (+,+,0) (-,+,0) (-,-,0) (+,-,0).  With
// px = hx*dx, py = hy*dy, z = 0*dz the four dots ((cx*dx) + (cy*dy)) + 0*dz are
//   (px+py)+z   (py-px)+z   -(px+py)+z   -(py-px)+z      bit for bit.
CB2_DI f3 quad_local_support(f3 h, f3 d) {
    const float px = fmul(h.x, d.x), py = fmul(h.y, d.y), z = fmul(0.0f, d.z);
    const float a = fadd(px, py), b = fsub(py, px);
    const float d0 = fadd(a, z), d1 = fadd(b, z), d2 = fadd(-a, z), d3 = fadd(-b, z);
    float best = -INFINITY;
    int idx = -1;
    if (d0 > best) { best = d0; idx = 0; }
    if (d1 > best) { best = d1; idx = 1; }
    if (d2 > best) { best = d2; idx = 2; }
    if (d3 > best) { best = d3; idx = 3; }
    if (idx < 0) return zero3();
    const bool nx = idx == 1 || idx == 2, ny = idx >= 2;
    return mk3(nx ? -h.x : h.x, ny ? -h.y : h.y, 0.0f);
}
// Cylinder (axis y; h.x = half_height, h.y = radius): cylinder.rs:45-78
CB2_DI f3 cylinder_local_support(f3 h, f3 d) {
    const bool negative = (__float_as_uint(d.y) >> 31) != 0u;  // is_sign_negative
    f3 r = mk3(d.x, 0.0f, d.z);
    if (magnitude2(r) == 0.0f) {
        r = zero3();
    } else {
        r = normalize(r);
        if (r.x == 0.0f && r.y == 0.0f && r.z == 0.0f) r = zero3();
        else r = r * h.y;
    }
    r.y = negative ? -h.x : h.x;
    return r;
}
// Capsule (axis y; h.x = half_height, h.y = radius): capsule.rs:45-61
CB2_DI f3 capsule_local_support(f3 h, f3 d) {
    // f32::signum: 1 for +0 and positive, -1 for -0 and negative, NaN for NaN
    const float sg = (d.y != d.y) ? d.y : ((__float_as_uint(d.y) >> 31) ? -1.0f : 1.0f);
    const f3 p = mk3(0.0f, fmul(sg, h.x), 0.0f);
    return p + normalize_to(d, h.y);
}

// the less common variants stay out of line so that they do not bloat the hot loops
CB2_NOINLINE_DEVICE f3 other_local_support(uint32_t kind, f3 h, f3 d) {
    if (kind == K_QUAD) return quad_local_support(h, d);
    if (kind == K_CYLINDER) return cylinder_local_support(h, d);
    return capsule_local_support(h, d);
}

// Primitive3 dispatch (primitive3.rs:153-176); ConvexPolyhedron VertexOnly (polyhedron.rs:160-176)
// goes through the support cells (cb2_cells.cuh)
template <bool ROLLED = false>
CB2_DI f3 local_support(const shape_table& T, const shape& s, f3 d) {
    if (s.kind == K_CUBOID) return cuboid_local_support(s.h, d);
    if (s.kind == K_POLY) return poly_local_support<ROLLED>(T.poly, s.poly, s.h.x, d);
    if (s.kind == K_SPHERE) return sphere_local_support(s.h.x, d);
    if (s.kind == K_POLY_HE) {
        const float4 r = poly_hill_climb(T.poly.verts + s.poly.voff, s.poly.vcnt, d);
        if (r.w != 0.0f) s.hang = 1u;
        return mk3(r.x, r.y, r.z);
    }
    return other_local_support(s.kind, s.h, d);
}

template <bool ROLLED = false>
CB2_DI f3 support_point(const shape_table& T, const shape& s, const xform& t, f3 dir) {
    // Particle: transform.transform_point(origin) — no inverse transform (primitive3.rs:164)
    if (s.kind == K_PARTICLE) return transform_point(t, zero3());
    return transform_point(t, local_support<ROLLED>(T, s, inverse_transform_vector(t, dir)));
}
// support_point for kernels in which NP lanes (lane ^ STEP, lane ^ 2 STEP, ...) hold the same shape,
// transform and direction (the EPA groups: even lanes carry the left shape, odd lanes the right
// one, so STEP = 2).  A polyhedron's candidate list is split over those lanes
// (poly_local_support_part) and the parts are combined by a butterfly: greater dot wins, equal
// dots go to the lower part — the reference's first-strict-max over the whole list.  Must be
// called by the whole warp (the shuffles use the full mask); `part` = this lane's rank among its NP.
template <uint32_t NP, uint32_t STEP>
CB2_DI f3 support_point_split(const shape_table& T, const shape& s, const xform& t, f3 dir, uint32_t part) {
    f3 loc = zero3();
    float bd = 0.0f;
    bool whole = true;
    if (s.kind == K_POLY) {
        loc = poly_local_support_part<NP>(T.poly, s.poly, s.h.x, inverse_transform_vector(t, dir), part, bd, whole);
    } else if (s.kind != K_PARTICLE) {
        const f3 d = inverse_transform_vector(t, dir);
        if (s.kind == K_CUBOID) loc = cuboid_local_support(s.h, d);
        else if (s.kind == K_SPHERE) loc = sphere_local_support(s.h.x, d);
        else if (s.kind == K_POLY_HE) {
            const float4 r = poly_hill_climb(T.poly.verts + s.poly.voff, s.poly.vcnt, d);
            if (r.w != 0.0f) s.hang = 1u;
            loc = mk3(r.x, r.y, r.z);
        } else loc = other_local_support(s.kind, s.h, d);
    }
    if (NP > 1) {
#pragma unroll
        for (uint32_t k = 1; k < NP; k <<= 1) {
            const uint32_t off = k * STEP;
            const float od = __shfl_xor_sync(0xffffffffu, bd, off);
            const float ox = __shfl_xor_sync(0xffffffffu, loc.x, off);
            const float oy = __shfl_xor_sync(0xffffffffu, loc.y, off);
            const float oz = __shfl_xor_sync(0xffffffffu, loc.z, off);
            const bool upper = (part & k) != 0u;
            if (!whole && (od > bd || (upper && od == bd))) {
                bd = od;
                loc = mk3(ox, oy, oz);
            }
        }
    }
    return transform_point(t, loc);
}
// does support_point call inverse_transform_vector(..).unwrap() (the zero-scale panic site)?
CB2_DI bool needs_inverse(uint32_t kind) { return kind != K_PARTICLE; }
CB2_DI bool is_curved(uint32_t kind) { return kind == K_SPHERE || kind == K_CYLINDER || kind == K_CAPSULE; }

// ---- simplex in registers -----------------------------------------------------------------
// v[i] = SupportPoint.v, a[i] = SupportPoint.sup_a (only kept when the EPA contact needs it),
// b[i] = SupportPoint.sup_b (never read on the 3-D path: only kept by the exported GJK::intersect,
// whose caller receives whole SupportPoints).  Newest point is LAST, as in the reference.
template <bool WITH_A, bool WITH_B = false>
struct simplex {
    f3 v[4];
    f3 a[WITH_A ? 4 : 1];
    f3 b[WITH_B ? 4 : 1];
    int n;
    CB2_DI void set(int i, f3 pv, f3 pa, f3 pb = zero3()) {
        v[i] = pv;
        if (WITH_A) a[i] = pa;
        if (WITH_B) b[i] = pb;
    }
    CB2_DI void mov(int dst, int src) {
        v[dst] = v[src];
        if (WITH_A) a[dst] = a[src];
        if (WITH_B) b[dst] = b[src];
    }
    CB2_DI void swap01() {
        f3 t = v[0];
        v[0] = v[1];
        v[1] = t;
        if (WITH_A) {
            f3 u = a[0];
            a[0] = a[1];
            a[1] = u;
        }
        if (WITH_B) {
            f3 u = b[0];
            b[0] = b[1];
            b[1] = u;
        }
    }
    CB2_DI void push(f3 pv, f3 pa, f3 pb = zero3()) {  // n <= 3 on entry
        if (n == 0) set(0, pv, pa, pb);
        else if (n == 1) set(1, pv, pa, pb);
        else if (n == 2) set(2, pv, pa, pb);
        else set(3, pv, pa, pb);
        ++n;
    }
};

// SimplexProcessor3::reduce_to_closest_feature (simplex3d.rs:24-98) + check_side (:180-220).
// Returns true iff the origin is inside the tetrahedron.  The reference calls check_side from
// three places (ABC / ACD / triangle); here each place only selects operands and the single
// check below runs once for the whole warp.
template <bool WITH_A, bool WITH_B>
CB2_DI bool reduce_to_closest_feature(simplex<WITH_A, WITH_B>& s, f3& v) {
    bool check = false, above = false, ignore_ab = false;
    f3 N, E1, E2, ao;
    if (s.n == 4) {
        const f3 a = s.v[3], b = s.v[2], c = s.v[1], d = s.v[0];
        ao = -a;
        const f3 ab = b - a, ac = c - a;
        const f3 abc = cross(ab, ac);
        if (dot(abc, ao) > 0.0f) {
            // remove(0) -> [c, b, a]
            s.mov(0, 1);
            s.mov(1, 2);
            s.mov(2, 3);
            s.n = 3;
            check = true; above = true;
            N = abc; E1 = ab; E2 = ac;
        } else {
            const f3 ad = d - a;
            const f3 acd = cross(ac, ad);
            if (dot(acd, ao) > 0.0f) {
                // remove(2) -> [d, c, a]
                s.mov(2, 3);
                s.n = 3;
                check = true; above = true; ignore_ab = true;
                N = acd; E1 = ac; E2 = ad;
            } else {
                const f3 adb = cross(ad, ab);
                if (dot(adb, ao) > 0.0f) {
                    // remove(1) -> [d, b, a]; swap(0,1) -> [b, d, a]
                    s.mov(1, 0);
                    s.mov(0, 2);
                    s.mov(2, 3);
                    s.n = 3;
                    v = adb;
                } else {
                    return true;
                }
            }
        }
    } else if (s.n == 3) {
        const f3 a = s.v[2], b = s.v[1], c = s.v[0];
        ao = -a;
        E1 = b - a;
        E2 = c - a;
        N = cross(E1, E2);
        check = true;
    } else if (s.n == 2) {
        const f3 a = s.v[1], b = s.v[0];
        const f3 ao2 = -a, ab = b - a;
        v = cross_aba(ab, ao2);
        if (ulps_eq_zero(v)) v.x = 0.1f;
    }
    if (check) {  // check_side(N=abc, E1=ab, E2=ac, ao, above, ignore_ab) on [C, B, A]
        const f3 ab_perp = cross(E1, N);
        if (!ignore_ab && dot(ab_perp, ao) > 0.0f) {
            s.mov(0, 1);  // remove(0) -> [B, A]
            s.mov(1, 2);
            s.n = 2;
            v = cross_aba(E1, ao);
        } else {
            const f3 ac_perp = cross(N, E2);
            if (dot(ac_perp, ao) > 0.0f) {
                s.mov(1, 2);  // remove(1) -> [C, A]
                s.n = 2;
                v = cross_aba(E2, ao);
            } else if (above || dot(N, ao) > 0.0f) {
                v = N;
            } else {
                s.swap01();
                v = -N;
            }
        }
    }
    return false;
}

// util.rs:53-72
CB2_DI void barycentric_vector(f3 p, f3 a, f3 b, f3 c, float& u, float& v, float& w) {
    const f3 v0 = b - a, v1 = c - a, v2 = p - a;
    const float d00 = dot(v0, v0), d01 = dot(v0, v1), d11 = dot(v1, v1), d20 = dot(v2, v0),
                d21 = dot(v2, v1);
    const float inv_denom = fdiv(1.0f, fsub(fmul(d00, d11), fmul(d01, d01)));
    v = fmul(fsub(fmul(d11, d20), fmul(d01, d21)), inv_denom);
    w = fmul(fsub(fmul(d00, d21), fmul(d01, d20)), inv_denom);
    u = fsub(fsub(1.0f, v), w);
}

// util.rs:98-114 with point = origin
CB2_DI f3 closest_point_on_edge_to_origin(f3 start, f3 end) {
    const f3 line = end - start;
    const f3 line_dir = normalize(line);
    const f3 vv = zero3() - start;
    const float d = dot(vv, line_dir);
    if (d < 0.0f) return start;
    if (fmul(d, d) > magnitude2(line)) return end;
    return start + line_dir * d;
}

CB2_DI bool in_range(float x) { return x >= 0.0f && x <= 1.0f; }

// simplex3d.rs:131-161 with Plane::from_points (plane.rs:78-95) and Plane ∩ Ray3
// (plane.rs:176-188) inlined; ray = (origin, -normal).  status != 0 => reference panics.
CB2_DI f3 closest_point_on_face(f3 a, f3 b, f3 c, f3 normal, uint8_t& status) {
    const f3 origin = zero3();
    const f3 dir = -normal;
    const f3 v0 = b - a, v1 = c - a;
    f3 n = cross(v0, v1);
    if (ulps_eq_zero(n)) {
        status = ST_PANIC_UNWRAP_PLANE;
        return zero3();
    }
    n = normalize(n);
    const float pd = -dot(a, n);
    const float t = fdiv(-fadd(pd, dot(origin, n)), dot(dir, n));
    if (t < 0.0f) return zero3();
    const f3 point = origin + dir * t;
    float u, v, w;
    barycentric_vector(point, a, b, c, u, v, w);
    if (!(in_range(u) && in_range(v) && in_range(w))) {
        status = ST_PANIC_ASSERT_RANGE;
        return zero3();
    }
    return point;
}

// SimplexProcessor3::get_closest_point_to_origin, simplex3d.rs:103-121
CB2_DI f3 closest_point_to_origin(simplex<false>& s, uint8_t& status) {
    f3 d = zero3();
    if (reduce_to_closest_feature(s, d)) return d;
    if (s.n == 1) return s.v[0];
    if (s.n == 2) return closest_point_on_edge_to_origin(s.v[1], s.v[0]);
    return closest_point_on_face(s.v[2], s.v[1], s.v[0], d, status);
}

// ---- output record ----------------------------------------------------------------------------
struct contact_out {
    f3 normal;
    float depth;
    f3 point;
    float toi;
};
CB2_DI contact_out zero_contact() {
    contact_out c;
    c.normal = zero3();
    c.depth = 0.0f;
    c.point = zero3();
    c.toi = 0.0f;
    return c;
}
CB2_DI void store_contact(cb2_contact* __restrict__ out, uint64_t i, uint32_t strategy,
                          const contact_out& c) {
    float* o = reinterpret_cast<float*>(out + i);
    o[0] = __uint_as_float(strategy);
    o[1] = c.normal.x;
    o[2] = c.normal.y;
    o[3] = c.normal.z;
    o[4] = c.depth;
    o[5] = c.point.x;
    o[6] = c.point.y;
    o[7] = c.point.z;
    o[8] = c.toi;
}

}  // namespace cb2
