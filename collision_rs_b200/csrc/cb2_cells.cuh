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

// the table-wide base pointers (uniform across a launch: they stay in uniform registers / the
// constant bank) ...
struct poly_tables {
    const float4* verts;     // (x, y, z, HalfEdge ring offset)
    const uint4* cells;      // support-cell records of all polyhedra
    const uint16_t* ext;     // extension pool
};
// ... and the four words a thread keeps per polyhedron
struct poly_ref {
    uint32_t voff;           // first vertex in the table
    uint32_t vcnt;
    uint32_t cell_off;       // first cell record
    uint32_t cfg;            // N | wide << 8 (N = 0: no cells)
};

// kept out of line: it is the rare path, and inlining it (three call sites) bloats the hot loops
CB2_NOINLINE_DEVICE f3 poly_full_scan(const float4* __restrict__ verts, uint32_t n, f3 d) {
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

CB2_NOINLINE_DEVICE f3 poly_ext_scan(const float4* __restrict__ verts,
                                         const uint16_t* __restrict__ e, uint32_t n, f3 d) {
    f3 best = zero3();
    float best_dot = -INFINITY;
    for (uint32_t k = 0; k < n; ++k) {
        const float4 v = __ldg(verts + __ldg(e + k));
        const f3 p = mk3(v.x, v.y, v.z);
        const float dt = dot(p, d);
        if (dt > best_dot) {
            best_dot = dt;
            best = p;
        }
    }
    return best;
}

// HalfEdge mode (ConvexPolyhedron::new_with_faces): hill_climb_support_point, polyhedron.rs:179-209.
// The ring of a vertex — the target vertices of the reference's walk edges[twin].next_edge from
// vertices[v].edge, in walk order — is a record [count, t0, t1, ...] of uint32 stored behind the
// vertex table; the vertex's own float4 carries its offset (in uint32 units from the shape's first
// vertex) in .w, so the climb needs no pointer beyond `verts`.  Returns (point, hang): the
// reference's loop `while best_dot != previous_best_dot` never ends once best_dot is NaN.
CB2_NOINLINE_DEVICE float4 poly_hill_climb(const float4* __restrict__ verts, uint32_t vcnt,
                                                      f3 d) {
    float4 bv = __ldg(verts);
    float best_dot = dot(mk3(bv.x, bv.y, bv.z), d);
    // the dot strictly grows every round, so a climb takes fewer than vcnt rounds
    for (uint32_t round = 0; round <= vcnt; ++round) {
        const float previous = best_dot;
        const uint32_t* __restrict__ ring = reinterpret_cast<const uint32_t*>(verts) + __float_as_uint(bv.w);
        const uint32_t cnt = __ldg(ring);
        for (uint32_t k = 1; k <= cnt; ++k) {
            const float4 v = __ldg(verts + __ldg(ring + k));
            const float dt = dot(mk3(v.x, v.y, v.z), d);
            if (dt > best_dot) {
                best_dot = dt;
                bv = v;
            }
        }
        if (best_dot == previous) return make_float4(bv.x, bv.y, bv.z, 0.0f);
    }
    return make_float4(bv.x, bv.y, bv.z, 1.0f);
}

// local-space support point of a polyhedron for direction d (already in model space);
// R = max |vertex| of the shape, rounded up
// ROLLED: keep the candidate loop rolled.  Unrolled by two (the compiler's choice) a warp in which
// one lane holds five candidates runs an eight-candidate block for everybody: k_gjk_hits_p gains
// 6 % rolled, k_time_of_impact_p loses 14 % — so the caller decides.
template <bool ROLLED = false>
CB2_DI f3 poly_local_support(const poly_tables& T, const poly_ref& s, float R, f3 d) {
    const float4* __restrict__ verts = T.verts + s.voff;
    const uint32_t N = s.cfg & 0xFFu;
    const bool wide = (s.cfg >> 8) & 1u;
    if (N == 0) return poly_full_scan(verts, s.vcnt, d);
    float m;
    const uint32_t cell = cell_index(d, N, m);
    const float scale = m * R;
    // outside this window (or NaN) the rounding argument does not hold: full scan
    if (!(scale >= 1.0e-30f && scale <= 1.0e30f)) return poly_full_scan(verts, s.vcnt, d);
    const uint4 rec = __ldg(T.cells + s.cell_off + cell);
    const uint32_t bits = wide ? 16u : 8u;
    const uint32_t mask = wide ? 0xFFFFu : 0xFFu;
    const uint32_t cnt = rec.x & mask;
    if (cnt > (wide ? CELL_INLINE16 : CELL_INLINE8)) {
        if (cnt == (mask & (0xFF00u | CELL_FULL))) return poly_full_scan(verts, s.vcnt, d);
        return poly_ext_scan(verts, T.ext + rec.y, rec.z, d);  // rec.y = pool offset, rec.z = count
    }
    // The record is a 128-bit little-endian string [count | idx0 | idx1 | ...].  Drop the count,
    // then take FOUR candidates per round with their vertex loads issued back to back, so a
    // round costs one memory latency instead of four (the loads are the support call's cost).
    // Unused slots hold index 0: a valid address whose value is ignored.
    uint32_t x = __funnelshift_r(rec.x, rec.y, bits), y = __funnelshift_r(rec.y, rec.z, bits),
             z = __funnelshift_r(rec.z, rec.w, bits), w = rec.w >> bits;
    f3 best = zero3();  // no dot > -inf: P::origin() (polyhedron.rs:163)
    float best_dot = -INFINITY;
    auto trip = [&](uint32_t base) {
        uint32_t i0, i1, i2, i3;
        if (!wide) {
            i0 = x & 0xFFu; i1 = (x >> 8) & 0xFFu; i2 = (x >> 16) & 0xFFu; i3 = x >> 24;
            x = y; y = z; z = w; w = 0u;
        } else {
            i0 = x & 0xFFFFu; i1 = x >> 16; i2 = y & 0xFFFFu; i3 = y >> 16;
            x = z; y = w; z = 0u; w = 0u;
        }
        const float4 v0 = __ldg(verts + i0), v1 = __ldg(verts + i1);
        const float4 v2 = __ldg(verts + i2), v3 = __ldg(verts + i3);
        const uint32_t rem = cnt - base;
        const float d0 = dot(mk3(v0.x, v0.y, v0.z), d), d1 = dot(mk3(v1.x, v1.y, v1.z), d);
        const float d2 = dot(mk3(v2.x, v2.y, v2.z), d), d3 = dot(mk3(v3.x, v3.y, v3.z), d);
        // ascending index order + strict '>' = the reference's first-strict-max rule
        if (d0 > best_dot) { best_dot = d0; best = mk3(v0.x, v0.y, v0.z); }
        if (rem > 1u && d1 > best_dot) { best_dot = d1; best = mk3(v1.x, v1.y, v1.z); }
        if (rem > 2u && d2 > best_dot) { best_dot = d2; best = mk3(v2.x, v2.y, v2.z); }
        if (rem > 3u && d3 > best_dot) { best_dot = d3; best = mk3(v3.x, v3.y, v3.z); }
    };
    if constexpr (ROLLED) {
#pragma unroll 1
        for (uint32_t base = 0; base < cnt; base += 4u) trip(base);
    } else {
        for (uint32_t base = 0; base < cnt; base += 4u) trip(base);
    }
    return best;
}

// The same query answered by NP lanes that hold the same shape and direction (the EPA kernel's
// groups keep 2 or 4 lanes per side): the cell's list is cut into chunks of four candidates and
// lane `part` scans chunks [part*q, (part+1)*q), q = ceil(chunks / NP) — a contiguous run, so every
// candidate of a lower part precedes every candidate of a higher one.  The caller combines the
// parts' (best_dot, best) with "greater dot wins, equal dots go to the lower part", which is the
// full scan's first-strict-max rule.  `whole` is set when this lane answered the whole query on
// its own (no cells, full scan, extension list): all parts then hold the same answer.
template <uint32_t NP>
CB2_DI f3 poly_local_support_part(const poly_tables& T, const poly_ref& s, float R, f3 d, uint32_t part,
                                  float& best_dot, bool& whole) {
    const float4* __restrict__ verts = T.verts + s.voff;
    const uint32_t N = s.cfg & 0xFFu;
    const bool wide = (s.cfg >> 8) & 1u;
    whole = true;
    best_dot = 0.0f;
    if (N == 0) return poly_full_scan(verts, s.vcnt, d);
    float m;
    const uint32_t cell = cell_index(d, N, m);
    const float scale = m * R;
    if (!(scale >= 1.0e-30f && scale <= 1.0e30f)) return poly_full_scan(verts, s.vcnt, d);
    const uint4 rec = __ldg(T.cells + s.cell_off + cell);
    const uint32_t bits = wide ? 16u : 8u;
    const uint32_t mask = wide ? 0xFFFFu : 0xFFu;
    const uint32_t cnt = rec.x & mask;
    if (cnt > (wide ? CELL_INLINE16 : CELL_INLINE8)) {
        if (cnt == (mask & (0xFF00u | CELL_FULL))) return poly_full_scan(verts, s.vcnt, d);
        return poly_ext_scan(verts, T.ext + rec.y, rec.z, d);
    }
    whole = false;
    const uint32_t x = __funnelshift_r(rec.x, rec.y, bits), y = __funnelshift_r(rec.y, rec.z, bits),
                   z = __funnelshift_r(rec.z, rec.w, bits), w = rec.w >> bits;
    const uint32_t nch = (cnt + 3u) >> 2;
    const uint32_t q = (nch + NP - 1u) / NP;
    uint32_t c = part * q;
    const uint32_t c_end = min(c + q, nch);
    f3 best = zero3();
    best_dot = -INFINITY;
    // (one chunk per lane is the rule: lists beyond 4 NP candidates are rare, keep the loop rolled)
#pragma unroll 1
    for (; c < c_end; ++c) {
        uint32_t i0, i1, i2, i3;
        if (!wide) {
            const uint32_t u = c == 0u ? x : (c == 1u ? y : (c == 2u ? z : w));
            i0 = u & 0xFFu; i1 = (u >> 8) & 0xFFu; i2 = (u >> 16) & 0xFFu; i3 = u >> 24;
        } else {
            const uint32_t u = c == 0u ? x : z, v = c == 0u ? y : w;
            i0 = u & 0xFFFFu; i1 = u >> 16; i2 = v & 0xFFFFu; i3 = v >> 16;
        }
        const float4 v0 = __ldg(verts + i0), v1 = __ldg(verts + i1);
        const float4 v2 = __ldg(verts + i2), v3 = __ldg(verts + i3);
        const uint32_t rem = cnt - 4u * c;
        const float d0 = dot(mk3(v0.x, v0.y, v0.z), d), d1 = dot(mk3(v1.x, v1.y, v1.z), d);
        const float d2 = dot(mk3(v2.x, v2.y, v2.z), d), d3 = dot(mk3(v3.x, v3.y, v3.z), d);
        if (d0 > best_dot) { best_dot = d0; best = mk3(v0.x, v0.y, v0.z); }
        if (rem > 1u && d1 > best_dot) { best_dot = d1; best = mk3(v1.x, v1.y, v1.z); }
        if (rem > 2u && d2 > best_dot) { best_dot = d2; best = mk3(v2.x, v2.y, v2.z); }
        if (rem > 3u && d3 > best_dot) { best_dot = d3; best = mk3(v3.x, v3.y, v3.z); }
    }
    return best;
}

}  // namespace cb2
