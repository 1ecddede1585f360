// cb2_cells.cuh — support cells: an exact accelerator for ConvexPolyhedron::support_point
// (polyhedron.rs:160-176, brute_force_support_point: first strictly greatest vertex·d).
//
// The reference scans all V vertices on every support call.  Here each polyhedron gets, once at
// cb2_shapes_upload, a cube map of 6·N² direction cells; a cell stores the (ascending) indices
// of every vertex that can be the argmax for SOME direction inside the (padded) cell.  A vertex
// v is left out of a cell only when one other vertex w beats it at all four cell corners c_i by
// a margin,  (w - v)·c_i > CELL_MARGIN · R · |c_i|   (R = max |vertex|, evaluated in f64),
// which by linearity holds for every direction d of the cell with CELL_MARGIN·R·|d| to spare.
// The f32 dot the reference computes carries at most 3·2^-24·R·|d| of rounding error, so such a
// v has fl(v·d) < fl(w·d) strictly: it can neither win nor tie, and scanning the cell's list in
// ascending index order with `dot > best` returns exactly the vertex the full scan returns.
// Directions the argument does not cover (NaN/Inf, zero, magnitudes where products under- or
// overflow) take the full scan.  Lists are short (2-10 entries for 32-256 vertex hulls), so one
// thread answers a support query from L2-resident tables with one 128-bit cell load plus a few
// vertex loads, and nothing is staged in shared memory.
#pragma once
#include "cb2_math.cuh"

namespace cb2 {

#define CB2_CELL_MARGIN 4.0e-6
#define CB2_CELL_PAD 1.0e-3

constexpr uint32_t CELL_INLINE8 = 15;    // u8 indices held in the 16-byte record (V <= 256)
constexpr uint32_t CELL_INLINE16 = 7;    // u16 indices (V <= 65535)
constexpr uint32_t CELL_EXT = 0xFFu;     // count code: list lives in the extension pool
constexpr uint32_t CELL_FULL = 0xFEu;    // count code: scan every vertex

// cube-map cell of a direction: face = 2*axis + (negative), then an N x N grid over the two
// minor components divided by the major one.  `m` receives the major magnitude.
CB2_DI uint32_t cell_index(f3 d, uint32_t N, float& m) {
    const float ax = fabsf(d.x), ay = fabsf(d.y), az = fabsf(d.z);
    float major, a, b;
    uint32_t axis;
    if (ax >= ay && ax >= az) { axis = 0; major = d.x; a = d.y; b = d.z; }
    else if (ay >= az) { axis = 1; major = d.y; a = d.z; b = d.x; }
    else { axis = 2; major = d.z; a = d.x; b = d.y; }
    m = fabsf(major);
    const float inv = __frcp_rn(m);
    const float fn = 0.5f * (float)N;
    int iu = (int)((a * inv + 1.0f) * fn);
    int iv = (int)((b * inv + 1.0f) * fn);
    iu = min(max(iu, 0), (int)N - 1);
    iv = min(max(iv, 0), (int)N - 1);
    const uint32_t face = 2u * axis + (major < 0.0f ? 1u : 0u);
    return (face * N + (uint32_t)iu) * N + (uint32_t)iv;
}

struct poly_ref {
    const float4* verts;     // (x,y,z,0)
    uint32_t vcnt;
    const uint4* cells;      // this shape's 6*N*N records (nullptr when N == 0)
    const uint16_t* ext;     // extension pool
    uint32_t N;              // 0 => no cells
    bool wide;               // u16 indices
    float R;                 // max |vertex|, rounded up
};

CB2_DI f3 poly_full_scan(const float4* __restrict__ verts, uint32_t n, f3 d) {
    f3 best = zero3();
    float best_dot = -INFINITY;
    for (uint32_t i = 0; i < n; ++i) {
        const float4 v = __ldg(verts + i);
        const f3 p = mk3(v.x, v.y, v.z);
        const float dt = dot(p, d);
        if (dt > best_dot) {
            best_dot = dt;
            best = p;
        }
    }
    return best;
}

// local-space support point of a polyhedron for direction d (already in model space)
CB2_DI f3 poly_local_support(const poly_ref& s, f3 d) {
    if (s.N == 0) return poly_full_scan(s.verts, s.vcnt, d);
    float m;
    const uint32_t cell = cell_index(d, s.N, m);
    const float scale = m * s.R;
    // outside this window (or NaN) the rounding argument does not hold: full scan
    if (!(scale >= 1.0e-30f && scale <= 1.0e30f)) return poly_full_scan(s.verts, s.vcnt, d);
    uint4 rec = __ldg(s.cells + cell);
    float best_dot = -INFINITY;
    uint32_t best = 0xFFFFFFFFu;
    const uint32_t bits = s.wide ? 16u : 8u;
    const uint32_t mask = s.wide ? 0xFFFFu : 0xFFu;
    uint32_t cnt = rec.x & mask;
    if (cnt <= (s.wide ? CELL_INLINE16 : CELL_INLINE8)) {
        // the record is a 128-bit little-endian string [count | idx0 | idx1 | ...]: shift it down
        // one entry per iteration (4 funnel shifts) instead of indexing it dynamically
        for (; cnt > 0; --cnt) {
            rec.x = __funnelshift_r(rec.x, rec.y, bits);
            rec.y = __funnelshift_r(rec.y, rec.z, bits);
            rec.z = __funnelshift_r(rec.z, rec.w, bits);
            rec.w >>= bits;
            const uint32_t idx = rec.x & mask;
            const float4 v = __ldg(s.verts + idx);
            const float dt = dot(mk3(v.x, v.y, v.z), d);
            if (dt > best_dot) {
                best_dot = dt;
                best = idx;
            }
        }
    } else if (cnt == (mask & (0xFF00u | CELL_FULL))) {
        return poly_full_scan(s.verts, s.vcnt, d);
    } else {
        // extension list: rec.y = offset into the pool, rec.z = count
        const uint16_t* __restrict__ e = s.ext + rec.y;
        for (uint32_t k = 0; k < rec.z; ++k) {
            const uint32_t idx = __ldg(e + k);
            const float4 v = __ldg(s.verts + idx);
            const float dt = dot(mk3(v.x, v.y, v.z), d);
            if (dt > best_dot) {
                best_dot = dt;
                best = idx;
            }
        }
    }
    if (best == 0xFFFFFFFFu) return zero3();  // no dot > -inf: P::origin() (polyhedron.rs:163)
    const float4 v = __ldg(s.verts + best);
    return mk3(v.x, v.y, v.z);
}

}  // namespace cb2
