// cb2_epa.cu — sub-warp cooperative EPA3 (EPA3::process, epa3d.rs:27-65): the second stage of
// GJK3::intersection(FullResolution) (gjk/mod.rs:362-391) for every shape kind.
//
// Design (B200-first, not a translation of the reference's per-pair loop):
//  * k_gjk_hits (cb2_narrow.cu, thread per pair) has already run GJK::intersect and queued the
//    hits with their terminal simplex.  Here a GROUP of G lanes owns one hit; a warp holds 32/G
//    groups and is persistent: a group that finishes pulls the next hit from a device-side queue,
//    so lanes never idle behind the slowest pair of a warp.
//  * Support points cost one thread a cube-map lookup plus a handful of vertex dots (support
//    cells, cb2_cells.cuh), so nothing is staged in shared memory and the group splits the two
//    support calls of SupportPoint::from_minkowski (minkowski/mod.rs:39-60) by lane parity: even
//    lanes evaluate the left shape, odd lanes the right one, one shuffle exchanges the results.
//    Each lane therefore keeps only ITS shape and transform in registers.
//  * The O(F) work of Polytope::add (epa3d.rs:147-175) is split across the G lanes: visibility
//    tests, horizon extraction, Face::new of the new faces, and the closest-face search (:137-145).
//  * Polytope::add is reproduced in ORDER without replaying it serially: face visibility becomes
//    a register bit mask (ballot), the swap_remove sweep is a two-pointer walk over that mask
//    (which faces are examined in which order, which tail face lands in which hole), and the
//    horizon list is "edges of visible faces, in examination order, whose reverse edge is not an
//    edge of a visible face" — evaluated in parallel through a vertex-adjacency bit matrix in
//    shared memory.  That equals remove_or_add_edge's list (epa3d.rs:212-219) whenever no directed
//    edge occurs twice; the (non-manifold) exception is detected by the same atomicOr and the
//    pair goes to the serial second-chance kernel (cb2_narrow.cu).
//  * The polytope lives in per-group shared memory in three tiers: 40 faces / 24 vertices and
//    72 / 40 (4 and 8 lanes per pair; ~86 % and ~99.5 % of pairs) and 208 / 104 (16 lanes; the
//    manifold worst case 2V-4 at the 100-iteration cap).  Small tiers buy residency: tier 0 runs
//    20 warps per SM.  A pair that outgrows its tier dumps its polytope to global memory and the
//    next tier RESUMES it at the same iteration (nothing is recomputed).  sup_a of each polytope
//    vertex is only read by the final contact (epa3d.rs:84-114), so it lives in a global
//    scratch, not shared memory.
//  * The kernel is issue-bound at 20 warps per SM and its time follows its executed warp
//    instructions, so the per-round phases are written against their SASS (profiles/README.md,
//    round 2): running shared-memory addresses and unconditional loads in the per-face loops, a
//    branch-free walk step, the lanes of a side sharing a polyhedron's candidate list
//    (support_point_split), rare multi-trip loops kept rolled.
#include <type_traits>

#include "cb2_epa_team.cuh"
#include "cb2_tma.cuh"

#ifndef CB2_T0_F  // build-time overrides for A/B experiments (tools/ab_variants.sh)
#define CB2_T0_F 40
#define CB2_T0_V 24
#define CB2_T0_O 20   // (20: the record is then 1392 bytes = -16 mod 128, see group_layout)
#endif
#ifndef CB2_OPT_SPLIT   // a polyhedron's support candidates are split over the lanes of a side
#define CB2_OPT_SPLIT 1
#endif
#ifndef CB2_OPT_FOLD    // closest_face_to_origin folded into the visibility pass and Face::new: bit-identical,
#define CB2_OPT_FOLD 0  // but no faster (tier 0 14.03 ms with, 13.95 ms without: three more live registers)
#endif
#ifndef CB2_OPT_STAGED  // tier 0 adopts hits the warp prepared 32 at a time (staged fetch): bit-identical,
#define CB2_OPT_STAGED 0  // 10.75 ms against 10.61 ms without — the other warps already hide the fetch chain
#endif
#ifndef CB2_VIS2        // visibility pass: two faces per lane and trip
#define CB2_VIS2 1
#endif
#ifndef CB2_EPA_MIN_BLOCKS
#define CB2_EPA_MIN_BLOCKS 20
#endif
// loop-unroll factors of the per-face loops (A/B experiments): closest face, visibility, new faces
#define CB2_PRAGMA(x) _Pragma(#x)
#define CB2_UNROLL(n) CB2_PRAGMA(unroll n)
#ifndef CB2_U_CLOSEST
#define CB2_U_CLOSEST 1
#endif
#ifndef CB2_U_VIS
#define CB2_U_VIS 1
#endif
#ifndef CB2_U_FACE
#define CB2_U_FACE 1
#endif
namespace cb2 {

// per-group polytope scratch in shared memory.  ADJ_IN_PV (one adjacency word per vertex row, i.e.
// VCAP <= 32): the row of vertex v lives in the unused fourth word of pv[v] instead of an array of
// its own — 96 bytes per group less, which is what lets a 20th block fit an SM in tier 0.
template <int FCAP, int VCAP, int OCAP, bool ADJ_IN_PV = (VCAP <= 32)>
struct poly_smem {
    static constexpr int AW = (VCAP + 31) / 32;  // adjacency words per vertex row
    float4 fnd[FCAP];          // normal.xyz, distance
    float4 pv[VCAP];           // SupportPoint.v (x,y,z,-)
    uint32_t fidx[FCAP];       // v0 | v1<<8 | v2<<16
    uint32_t adj[VCAP * AW];   // directed-edge bit matrix (all zero between iterations)
    uint16_t elist[OCAP];      // horizon edges a | b<<8 in reference order
    uint8_t order[OCAP];       // visible faces in the order Polytope::add examines them
    uint16_t mv[OCAP];         // swap_remove: tail face (mv >> 8) lands in hole (mv & 255)
    uint64_t mbar;             // tracks the bulk copy that resumes a polytope from the previous tier
    CB2_DI uint32_t* adj_word_ptr(uint32_t w) { return &adj[w]; }
};
template <int FCAP, int VCAP, int OCAP>
struct poly_smem<FCAP, VCAP, OCAP, true> {
    static constexpr int AW = 1;
    float4 fnd[FCAP];
    float4 pv[VCAP];           // (x, y, z, adjacency row of this vertex: all zero between iterations)
    uint32_t fidx[FCAP];
    uint16_t elist[OCAP];
    uint8_t order[OCAP];
    uint16_t mv[OCAP];
    uint64_t mbar;
    CB2_DI uint32_t* adj_word_ptr(uint32_t w) { return reinterpret_cast<uint32_t*>(&pv[w]) + 3; }
};

// Where the groups of a warp keep their polytopes.  ncu (round 2, 4 M pairs) showed 60 % of tier 0's
// shared-memory wavefronts to be bank-conflict replays: the records were laid out back to back with
// a stride that is a multiple of 128 bytes, so the same field of every group sat in the same banks.
// A stride of -16 (mod 128) bytes for 8 groups — 32 (mod 128) for 4 — spreads a unit-stride 32-bit
// access of all groups over all 32 banks; and with 8 groups, the two groups that share a
// quarter-warp (which is what a 128-bit access is split into) are given slots 4 apart, i.e. 64
// bytes apart modulo 128, so their float4 reads do not collide either.
template <class PS, int NG>
struct group_layout {
    static constexpr size_t TARGET = NG == 8 ? 112 : (NG == 4 ? 32 : 0);
    // (16 groups: an odd multiple of 16 bytes, so the groups visit all eight 16-byte columns of a 128-byte row)
    // (8 groups: any ODD multiple of 16 bytes does what -16 does — eight distinct 16-byte columns,
    // slots 4 apart 64 bytes apart — so a record whose size already is one takes no padding)
    static constexpr size_t STRIDE = (NG == 16 || NG == 8) ? sizeof(PS) + ((sizeof(PS) / 16) % 2 ? 0 : 16)
                                     : (NG >= 4 ? sizeof(PS) + ((TARGET + 128 - sizeof(PS) % 128) % 128) : sizeof(PS));
    static constexpr size_t BYTES = STRIDE * NG;
    static_assert(STRIDE % 16 == 0, "float4 members");
    CB2_DI static int slot(int gw) { return NG == 8 ? ((gw & 1) * 4 + (gw >> 1)) : gw; }
};

CB2_DI int popc_any(uint32_t v) { return __popc(v); }
CB2_DI int popc_any(unsigned long long v) { return __popcll(v); }

// shared-memory loads the compiler may not move into a branch
CB2_DI float4 lds128(uint32_t a) {
    float4 r;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "r"(a));
    return r;
}
CB2_DI uint32_t lds32(uint32_t a) {
    uint32_t r;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(r) : "r"(a));
    return r;
}

template <class PS>
CB2_DI f3 ps_pv(const PS& P, uint32_t i) {
    const float4 v = P.pv[i];
    return mk3(v.x, v.y, v.z);
}

// Face::new_impl (epa3d.rs:189-199) into slot
template <class PS>
CB2_DI float ps_face_new(PS& P, int slot, uint32_t a, uint32_t b, uint32_t c) {
    const f3 va = ps_pv(P, a);
    const f3 ab = ps_pv(P, b) - va;
    const f3 ac = ps_pv(P, c) - va;
    const f3 n = normalize(cross(ab, ac));
    const float dist = dot(n, va);
    P.fnd[slot] = make_float4(n.x, n.y, n.z, dist);
    P.fidx[slot] = a | (b << 8) | (c << 16);
    return dist;
}

// fixed-size bit mask in registers (fully unrolled accessors)
template <int W>
struct bitmask {
    uint32_t w[W];
    CB2_DI void clear() {
#pragma unroll
        for (int k = 0; k < W; ++k) w[k] = 0u;
    }
    CB2_DI void or_bits(int pos, uint32_t bits) {  // pos is a multiple of the group width (<= 32)
#pragma unroll
        for (int k = 0; k < W; ++k)
            if ((pos >> 5) == k) w[k] |= bits << (pos & 31);
    }
    CB2_DI bool test(int i) const {
        uint32_t v = 0;
#pragma unroll
        for (int k = 0; k < W; ++k)
            if ((i >> 5) == k) v = w[k];
        return (v >> (i & 31)) & 1u;
    }
    CB2_DI void reset(int i) {
#pragma unroll
        for (int k = 0; k < W; ++k)
            if ((i >> 5) == k) w[k] &= ~(1u << (i & 31));
    }
    CB2_DI void set_word(int k, uint32_t v) { w[k] |= v; }
    CB2_DI bool any() const {
        uint32_t v = 0u;
#pragma unroll
        for (int k = 0; k < W; ++k) v |= w[k];
        return v != 0u;
    }
    CB2_DI int first() const { return next(0); }  // lowest set bit (any() must hold)
    // lowest set bit at position >= from, or -1
    CB2_DI int next(int from) const {
        int r = -1;
#pragma unroll
        for (int k = W - 1; k >= 0; --k) {
            uint32_t v = w[k];
            if ((from >> 5) == k) v &= 0xFFFFFFFFu << (from & 31);
            else if ((from >> 5) > k) v = 0u;
            if (v) r = k * 32 + (__ffs(v) - 1);
        }
        return r;
    }
};

// up to 64 faces: one 64-bit register pair (the generic version above ends up in local memory,
// nvcc turns its unrolled selects back into a dynamically indexed array)
template <>
struct bitmask<2> {
    unsigned long long m;
    CB2_DI void clear() { m = 0ull; }
    CB2_DI void or_bits(int pos, uint32_t bits) { m |= (unsigned long long)bits << pos; }
    CB2_DI bool test(int i) const { return (m >> i) & 1ull; }
    CB2_DI void reset(int i) { m &= ~(1ull << i); }
    CB2_DI bool any() const { return m != 0ull; }
    CB2_DI int first() const { return __ffsll((long long)m) - 1; }
    CB2_DI void set_word(int k, uint32_t v) { m |= (unsigned long long)v << (32 * k); }  // k is a constant
    CB2_DI uint32_t low_word() const { return (uint32_t)m; }
    CB2_DI int next(int from) const {
        const unsigned long long v = from < 64 ? (m >> from) << from : 0ull;
        return v ? __ffsll((long long)v) - 1 : -1;
    }
};
template <>
struct bitmask<1> : bitmask<2> {};
// up to 128 faces: two 64-bit halves
template <>
struct bitmask<4> {
    unsigned long long lo, hi;
    CB2_DI void clear() { lo = hi = 0ull; }
    CB2_DI void or_bits(int pos, uint32_t bits) {  // a group's bits never straddle bit 64
        if (pos < 64) lo |= (unsigned long long)bits << pos;
        else hi |= (unsigned long long)bits << (pos - 64);
    }
    CB2_DI bool test(int i) const { return ((i < 64 ? lo >> i : hi >> (i - 64)) & 1ull) != 0ull; }
    CB2_DI void reset(int i) {
        if (i < 64) lo &= ~(1ull << i);
        else hi &= ~(1ull << (i - 64));
    }
    CB2_DI void set_word(int k, uint32_t v) {  // k is a constant
        if (k < 2) lo |= (unsigned long long)v << (32 * k);
        else hi |= (unsigned long long)v << (32 * (k - 2));
    }
    CB2_DI bool any() const { return (lo | hi) != 0ull; }
    CB2_DI int first() const { return lo ? __ffsll((long long)lo) - 1 : 64 + __ffsll((long long)hi) - 1; }
    CB2_DI int next(int from) const {
        if (from < 64) {
            const unsigned long long v = (lo >> from) << from;
            if (v) return __ffsll((long long)v) - 1;
            return hi ? 64 + __ffsll((long long)hi) - 1 : -1;
        }
        const unsigned long long v = from < 128 ? (hi >> (from - 64)) << (from - 64) : 0ull;
        return v ? 64 + __ffsll((long long)v) - 1 : -1;
    }
};
template <>
struct bitmask<3> : bitmask<4> {};

// one-word mask (up to 32 faces)
struct bitmask32 {
    uint32_t m;
    CB2_DI bool test(int i) const { return (m >> i) & 1u; }
    CB2_DI void reset(int i) { m &= ~(1u << i); }
    CB2_DI bool any() const { return m != 0u; }
    CB2_DI int first() const { return __ffs((int)m) - 1; }
};

// The swap_remove sweep of Polytope::add (epa3d.rs:147-175) as a two-pointer walk over the mask of
// visible faces: `i` climbs from the front, `len` shrinks from the back; a visible face at slot i
// is examined, then the tail face dropped into slot i is examined next, and so on.  Examined
// faces leave the mask, so the next `i` is simply its lowest bit (everything below the old i and
// everything from len up has been examined or was never visible).  One walk step per warp
// iteration, branch-free: every lane of the warp executes it, idle groups change nothing.  Lane k
// of the group keeps the k-th examined face (`mine`) for the horizon passes; the examination
// order goes to P.order, the moves (tail face src lands in hole dst) to P.mv as dst | src << 8.
template <int G, int OCAP, class MK, class PS>
CB2_DI void epa_walk(MK& vis, PS& P, bool add, int gl, int& len, int& nvis, int& nmv, int& mine,
                     bool& overflow, int& ci) {
    bool walking = add && vis.any();
    bool step = walking && nvis < OCAP;  // (nvis is 0 here)
    int i = vis.first();
    int cur = i;
    while (__any_sync(0xffffffffu, step)) {
        mine = (step && nvis == gl) ? cur : mine;
        if (step && gl == 0) P.order[nvis] = (uint8_t)cur;
        if (step) vis.reset(cur);
        const int len1 = len - 1;
        const bool last = i == len1;                 // the examined face was the tail itself
        const bool tail_vis = vis.test(len1);        // the tail face dropped into slot i is examined next
        const bool mv = step && !last && !tail_vis;  // ... or stays there
        if (mv && gl == 0) P.mv[nmv] = (uint16_t)(i | (len1 << 8));
        ci = (mv && ci == len1) ? i : ci;  // (closest-face tracking: the tracked face moves with the swap)
        const int nxt = vis.first();
        i = mv ? nxt : i;
        cur = tail_vis ? len1 : nxt;
        nmv += mv ? 1 : 0;
        nvis += step ? 1 : 0;
        len = step ? len1 : len;
        walking = step && !last && (tail_vis || vis.any());
        step = walking && nvis < OCAP;
    }
    overflow = walking;  // faces left to examine, no room in the lists
}

// ---- staged fetch (tier 0) ------------------------------------------------------------------------------
// A group that finished used to fetch its next hit on the spot: work-list entry, shape indices,
// shape records, transforms and their inverse rotations, the simplex, the four initial faces — ~250
// instructions and four or five dependent memory round trips, executed by the 1.3 groups that
// happened to finish in that round while the rest of the warp waited (8 % of tier 0's instructions
// at 5 of 32 lanes).  Now the WARP prepares 32 hits at once — every lane one hit, full lanes, one
// chain of round trips per 32 hits — into a per-warp record array in global memory (L2 resident:
// 12 KB), and a finished group adopts a prepared record with a handful of 128-bit loads.
constexpr int SUP_HDR = 0, SUP_SHAPE = 1, SUP_XF = 5, SUP_PV = 11, SUP_PA = 15, SUP_FND = 19, SUP_FIDX = 23,
              SUP_SIZE = 24;  // float4 per record
CB2_NOINLINE_DEVICE void epa_prepare_hit(const narrow_args& A, const shape_table& T, const float4* simplex,
                                         uint32_t pair, float4* rec) {
    f3 pv[4];
#pragma unroll 1
    for (int side = 0; side < 2; ++side) {
        uint32_t si;
        uint64_t xi;
        if (A.pair_body) {
            const uint2 b = __ldg(reinterpret_cast<const uint2*>(A.pair_body) + pair);
            xi = side ? b.y : b.x;
            si = __ldg(A.l_shape + xi);
        } else {
            xi = pair;
            si = __ldg((side ? A.r_shape : A.l_shape) + pair);
        }
        const float4* sp = reinterpret_cast<const float4*>(T.shapes + si);
        rec[SUP_SHAPE + 2 * side] = __ldg(sp);
        rec[SUP_SHAPE + 2 * side + 1] = __ldg(sp + 1);
        const float4* tp = (side ? A.r_t : A.l_t) + 2 * xi;
        const xform t = make_xform(__ldg(tp), __ldg(tp + 1));
        rec[SUP_XF + 3 * side] = make_float4(t.q.s, t.q.x, t.q.y, t.q.z);
        rec[SUP_XF + 3 * side + 1] = make_float4(t.qi.s, t.qi.x, t.qi.y, t.qi.z);
        rec[SUP_XF + 3 * side + 2] = make_float4(t.disp.x, t.disp.y, t.disp.z, t.scale);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float4 v = simplex[(size_t)pair * 8 + j];
        rec[SUP_PV + j] = v;
        rec[SUP_PA + j] = simplex[(size_t)pair * 8 + 4 + j];
        pv[j] = mk3(v.x, v.y, v.z);
    }
    // Polytope::new: Face::new (3,2,1) (3,1,0) (3,0,2) (2,0,1), epa3d.rs:201-208, and the closest of
    // them (first strictly smallest distance; a NaN at index 0 sticks)
    uint32_t idx[4];
    float bd = INFINITY;
    uint32_t bi = 0xFFFFFFFFu;
    float d0 = 0.0f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const uint32_t tri = j == 0 ? 0x010203u : (j == 1 ? 0x000103u : (j == 2 ? 0x020003u : 0x010002u));
        const uint32_t a = tri & 255u, b = (tri >> 8) & 255u, c = tri >> 16;
        const f3 va = pv[a];
        const f3 n = normalize(cross(pv[b] - va, pv[c] - va));
        const float dist = dot(n, va);
        rec[SUP_FND + j] = make_float4(n.x, n.y, n.z, dist);
        idx[j] = a | (b << 8) | (c << 16);
        if (j == 0) d0 = dist;
        if (dist < bd) { bd = dist; bi = (uint32_t)j; }
    }
    rec[SUP_FIDX] = make_float4(__uint_as_float(idx[0]), __uint_as_float(idx[1]), __uint_as_float(idx[2]),
                                __uint_as_float(idx[3]));
    const uint32_t best = (d0 != d0 || bi == 0xFFFFFFFFu) ? 0u : bi;
    rec[SUP_HDR] = make_float4(__uint_as_float(pair), __uint_as_float(best), 0.f, 0.f);
}

enum : int { S_FETCH = 0, S_RUN = 1, S_EXIT = 2 };

// RF/RV: caps of the tier whose dumps this instantiation resumes (0 = none)
template <int G, int FCAP, int VCAP, int OCAP, int RF, int RV>
__global__ void __launch_bounds__(32, (RF == 0 ? CB2_EPA_MIN_BLOCKS : 20)) k_epa(epa_args C) {
    static_assert(G == 2 || G == 4 || G == 8 || G == 16, "3*G edge-keep bits must fit two words");
    static_assert(RF <= FCAP && RV <= VCAP, "a resumed polytope must fit");
    using PS = poly_smem<FCAP, VCAP, OCAP>;
    // The per-face loops (closest-face scan, lean visibility pass) take two faces per lane and trip
    // and load before they predicate: the last trip may read face slots up to FCAP + 2 G - 2.  Those
    // reads must stay inside the group's own record (their values are never used).
    static_assert((size_t)(FCAP + 2 * G - 1) * 16 <= sizeof(PS), "face slots read past nf leave the record");
    static_assert(offsetof(PS, fidx) + (size_t)(FCAP + 2 * G - 1) * sizeof(uint32_t) <= sizeof(PS),
                  "face index slots read past nf leave the record");
    constexpr int FW = (FCAP + 31) / 32;
    constexpr int AW = PS::AW;
    constexpr int NG = 32 / G;
    // adjacency bit (row a, column b): word a*AW + b/32, bit b%32
    auto adj_word = [](uint32_t a, uint32_t b) -> uint32_t { return AW == 1 ? a : a * AW + (b >> 5); };
    auto adj_bit = [](uint32_t b) -> uint32_t { return AW == 1 ? b : (b & 31u); };
    extern __shared__ __align__(16) float smem[];
    const int lane = threadIdx.x;
    const int gl = lane & (G - 1);
    const int gw = lane / G;
    const unsigned gmask = ((G == 32) ? 0xffffffffu : ((1u << G) - 1u)) << (gw * G);
    const uint32_t gbits = (1u << G) - 1u;
    const bool right = gl & 1;  // this lane evaluates the right shape's support
    const narrow_args& A = C.A;
    const shape_table T = table_of(A.T);
    using GL = group_layout<PS, NG>;
    PS& P = *reinterpret_cast<PS*>(reinterpret_cast<char*>(smem) + (size_t)GL::slot(gw) * GL::STRIDE);
    float4* pa = C.W.pa + ((size_t)blockIdx.x * NG + gw) * VCAP;

    const uint32_t* list;
    uint32_t count;
    uint32_t* over_list;
    uint32_t* over_count;
    list = C.W.list[C.tier];
    count = C.W.EC->n_list[C.tier];
    if (C.tier + 1 < EPA_TIERS) {
        over_list = C.W.list[C.tier + 1];
        over_count = &C.W.EC->n_list[C.tier + 1];
    } else {
        over_list = nullptr;
        over_count = nullptr;
    }
    uint32_t* cursor = &C.W.EC->next[C.tier];
    const uint32_t max_it = A.settings.max_iterations;
    const float epa_tol = A.settings.epa_tolerance;
    // the adjacency matrix is kept all-zero between iterations
    for (int k = gl; k < VCAP * AW; k += G) *P.adj_word_ptr(k) = 0u;
    uint32_t resume_phase = 0;  // parity of the group's mbarrier
    if (RF > 0 && gl == 0) mbar_init(&P.mbar, 1);
    __syncwarp(gmask);

    int state = S_FETCH;
    uint32_t pair = 0;
    // this lane's side of the pair
    shape sh;
    sh.kind = K_CUBOID;
    sh.h = zero3();
    sh.poly.voff = 0; sh.poly.vcnt = 0; sh.poly.cell_off = 0; sh.poly.cfg = 0;
    sh.hang = 0u;
    xform xf = make_xform(make_float4(1.f, 1.f, 0.f, 0.f), make_float4(0.f, 0.f, 0.f, 0.f));
    f3 d = zero3();
    uint32_t it = 0;
    int nf = 0, nv = 0;
    int best_face = -1;  // < 0: closest_face_to_origin has to scan (after a fetch, a resume, or a tie)

    // tier 0: the warp's supply of prepared hits (staged fetch above)
    constexpr bool STAGED = CB2_OPT_STAGED && RF == 0 && G == 4;
    float4* const sup = STAGED ? C.supply + (size_t)blockIdx.x * 32 * SUP_SIZE : nullptr;
    uint32_t sup_head = 0, sup_count = 0;  // warp-uniform: records [head, count) are ready
    bool sup_dry = false;                  // the work list is exhausted

    for (;;) {
        if constexpr (STAGED) {
            // ================= ADOPT a prepared hit (groups that finished) ==============================
            const bool want = state == S_FETCH;
            const unsigned need = __ballot_sync(0xffffffffu, want && gl == 0);  // one bit per wanting group
            if (need) {
                if (sup_head == sup_count && !sup_dry) {
                    uint32_t base = 0;
                    if (lane == 0) base = atomicAdd(cursor, 32u);
                    base = __shfl_sync(0xffffffffu, base, 0);
                    sup_count = base < count ? (count - base < 32u ? count - base : 32u) : 0u;
                    sup_head = 0u;
                    sup_dry = sup_count == 0u;
                    if ((uint32_t)lane < sup_count)
                        epa_prepare_hit(A, T, C.W.simplex, list[base + lane] & ~OVER_RESUME, sup + (size_t)lane * SUP_SIZE);
                    __syncwarp();
                }
                if (want) {
                    const uint32_t e = sup_head + (uint32_t)__popc(need & ((1u << (gw * G)) - 1u));
                    if (sup_dry && sup_head == sup_count) {
                        state = S_EXIT;
                        nf = 0;
                        best_face = -1;
                        sh.kind = K_CUBOID;
                        sh.h = zero3();
                    } else if (e < sup_count) {
                        // (L2 loads: the records were written by other lanes of this warp a moment ago)
                        const float4* rec = sup + (size_t)e * SUP_SIZE;
                        const float4 hdr = __ldcg(rec + SUP_HDR);
                        pair = __float_as_uint(hdr.x);
                        best_face = (int)__float_as_uint(hdr.y);
                        const int side = right ? 1 : 0;
                        const float4 s0 = __ldcg(rec + SUP_SHAPE + 2 * side), s1 = __ldcg(rec + SUP_SHAPE + 2 * side + 1);
                        sh.kind = __float_as_uint(s0.x);
                        sh.h = mk3(s0.y, s0.z, s0.w);
                        sh.poly.voff = __float_as_uint(s1.x);
                        sh.poly.vcnt = __float_as_uint(s1.y);
                        sh.poly.cell_off = __float_as_uint(s1.z);
                        sh.poly.cfg = __float_as_uint(s1.w);
                        sh.hang = 0u;
                        const float4 x0 = __ldcg(rec + SUP_XF + 3 * side), x1 = __ldcg(rec + SUP_XF + 3 * side + 1),
                                     x2 = __ldcg(rec + SUP_XF + 3 * side + 2);
                        xf.q.s = x0.x; xf.q.x = x0.y; xf.q.y = x0.z; xf.q.z = x0.w;
                        xf.qi.s = x1.x; xf.qi.x = x1.y; xf.qi.y = x1.z; xf.qi.z = x1.w;
                        xf.disp = mk3(x2.x, x2.y, x2.z);
                        xf.scale = x2.w;
                        xf.unit = (x2.w == 1.0f);
                        P.pv[gl] = __ldcg(rec + SUP_PV + gl);
                        pa[gl] = __ldcg(rec + SUP_PA + gl);
                        P.fnd[gl] = __ldcg(rec + SUP_FND + gl);
                        P.fidx[gl] = __ldcg(reinterpret_cast<const uint32_t*>(rec + SUP_FIDX) + gl);
                        nv = 4;
                        nf = 4;
                        it = 1;
                        state = S_RUN;
                    }  // else: the supply ran short this round; the group asks again in the next one
                }
                sup_head = sup_head + (uint32_t)__popc(need) < sup_count ? sup_head + (uint32_t)__popc(need) : sup_count;
            }
        } else
        // ================= FETCH (divergent: only groups that finished) ================================
        if (state == S_FETCH) {
            uint32_t k = 0;
            if (gl == 0) k = atomicAdd(cursor, 1u);
            k = __shfl_sync(gmask, k, gw * G);
            best_face = -1;
            if (k >= count) {
                state = S_EXIT;
                nf = 0;
                sh.kind = K_CUBOID;
                sh.h = zero3();
            } else {
                const uint32_t entry = list[k];
                pair = entry & ~OVER_RESUME;
                uint32_t si;
                uint64_t xi;
                if (A.pair_body) {
                    const uint2 b = __ldg(reinterpret_cast<const uint2*>(A.pair_body) + pair);
                    xi = right ? b.y : b.x;
                    si = __ldg(A.l_shape + xi);
                } else {
                    xi = pair;
                    si = __ldg((right ? A.r_shape : A.l_shape) + pair);
                }
                sh = load_shape(T, si);
                const float4* tp = (right ? A.r_t : A.l_t) + 2 * xi;
                xf = make_xform(__ldg(tp), __ldg(tp + 1));
                nv = 4;
                nf = 4;
                it = 1;
                if (RF > 0 && (entry & OVER_RESUME)) {
                    // resume a polytope that outgrew the previous tier, at the same iteration
                    using DL = dump_layout<RF, RV>;
                    const float4* rec = C.W.dump[C.tier - 1] + (size_t)k * DL::SIZE;
                    const float4 h = rec[0];
                    nf = (int)__float_as_uint(h.x);
                    nv = (int)__float_as_uint(h.y);
                    it = __float_as_uint(h.z);
                    // faces, vertices and face indices arrive as three bulk copies (TMA) tracked by
                    // the group's mbarrier; sup_a goes from the record to this group's scratch
                    const uint32_t b_fnd = (uint32_t)nf * 16u, b_pv = (uint32_t)nv * 16u,
                                   b_idx = ((uint32_t)nf * 4u + 15u) & ~15u;
                    if (gl == 0) {
                        mbar_arrive_expect_tx(&P.mbar, b_fnd + b_pv + b_idx);
                        bulk_load(P.fnd, rec + DL::FND, b_fnd, &P.mbar);
                        bulk_load(P.pv, rec + DL::PV, b_pv, &P.mbar);
                        bulk_load(P.fidx, rec + DL::FIDX, b_idx, &P.mbar);
                    }
                    for (int j = gl; j < nv; j += G) pa[j] = rec[DL::PA + j];
                    mbar_wait(&P.mbar, resume_phase);
                    resume_phase ^= 1u;
                } else {
                    // Polytope::new (epa3d.rs:41-45, 129-135): the 4 simplex points ...
                    for (int j = gl; j < 8; j += G) {
                        const float4 r = C.W.simplex[(size_t)pair * 8 + j];
                        if (j < 4) P.pv[j] = r;
                        else pa[j - 4] = r;
                    }
                    __syncwarp(gmask);
                    // ... and Face::new (3,2,1) (3,1,0) (3,0,2) (2,0,1), epa3d.rs:201-208
#if CB2_OPT_FOLD
                    if (G >= 4) {
                        // lanes 0..3 of the group each build one face and hold its distance: the first
                        // strictly smallest (NaN never smaller; a NaN at index 0 sticks) by a butterfly
                        float fd = INFINITY;
                        uint32_t fi = (uint32_t)gl;
                        if (gl < 4) {
                            const uint32_t tri = gl == 0 ? 0x010203u : (gl == 1 ? 0x000103u : (gl == 2 ? 0x020003u : 0x010002u));
                            fd = ps_face_new(P, gl, tri & 255u, (tri >> 8) & 255u, tri >> 16);
                        }
                        const float d0 = __shfl_sync(gmask, fd, gw * G);
                        if (fd != fd) fd = INFINITY;
#pragma unroll
                        for (int off = 2; off >= 1; off >>= 1) {
                            const float ob = __shfl_xor_sync(gmask, fd, off);
                            const uint32_t oi = __shfl_xor_sync(gmask, fi, off);
                            if (ob < fd || (ob == fd && oi < fi)) { fd = ob; fi = oi; }
                        }
                        fd = __shfl_sync(gmask, fd, gw * G);  // (groups wider than four: lane 0 holds the result)
                        fi = __shfl_sync(gmask, fi, gw * G);
                        best_face = (d0 != d0 || fd == INFINITY) ? 0 : (int)fi;
                    } else
#endif
                    for (int j = gl; j < 4; j += G) {
                        const uint32_t tri = j == 0 ? 0x010203u : (j == 1 ? 0x000103u : (j == 2 ? 0x020003u : 0x010002u));
                        ps_face_new(P, j, tri & 255u, (tri >> 8) & 255u, tri >> 16);
                    }
                }
                state = S_RUN;
            }
        }
        if (__all_sync(0xffffffffu, state == S_EXIT)) break;
        __syncwarp();
        // From here to the end of the iteration every phase is executed by the WHOLE warp in
        // lockstep (trip counts are warp maxima, per-group work is predicated), so the 32/G
        // groups share one instruction stream.
        const bool run = state == S_RUN;

        // ================= closest_face_to_origin (epa3d.rs:137-145) ====================================
        // first strictly smallest distance; a NaN at index 0 sticks.  Its normal is the next support
        // direction.  Only groups without a tracked closest face scan (fresh or resumed polytopes,
        // and the rare exact tie): otherwise the visibility pass and Face::new of the previous
        // iteration have already seen every distance of the polytope (see below).
        {
            const bool need = run && best_face < 0;
            const int nfn = need ? nf : 0;  // idle groups (and groups with a tracked face) scan nothing
            const int nf_max = __reduce_max_sync(0xffffffffu, nfn);
            if (nf_max > 0) {
                // two faces per lane and trip (f, f + G: still ascending per lane, so ties keep the
                // first); a NaN distance never compares smaller, so it needs no special case here
                float bd = INFINITY;
                uint32_t bi = 0xFFFFFFFFu;
                uint32_t a_d = (uint32_t)__cvta_generic_to_shared(&P) + (uint32_t)offsetof(PS, fnd) + (uint32_t)gl * 16u + 12u;
                for (int f = gl; f - gl < nf_max; f += 2 * G, a_d += 32u * G) {
                    const float d0 = __uint_as_float(lds32(a_d)), d1 = __uint_as_float(lds32(a_d + 16u * G));
                    const bool t0 = (f < nfn) & (d0 < bd);
                    bd = t0 ? d0 : bd;
                    bi = t0 ? (uint32_t)f : bi;
                    const bool t1 = (f + G < nfn) & (d1 < bd);
                    bd = t1 ? d1 : bd;
                    bi = t1 ? (uint32_t)(f + G) : bi;
                }
#pragma unroll
                for (int off = G / 2; off >= 1; off >>= 1) {
                    const float ob = __shfl_xor_sync(0xffffffffu, bd, off);
                    const uint32_t oi = __shfl_xor_sync(0xffffffffu, bi, off);
                    if (ob < bd || (ob == bd && oi < bi)) { bd = ob; bi = oi; }
                }
                if (need) {
                    const float d0 = P.fnd[0].w;
                    best_face = (d0 != d0 || bi == 0xFFFFFFFFu) ? 0 : (int)bi;
                }
            }
        }
        const float4 fb = run ? P.fnd[best_face] : make_float4(0.f, 0.f, 0.f, 0.f);
        const f3 n = mk3(fb.x, fb.y, fb.z);

        // ================= SupportPoint::from_minkowski(n), minkowski/mod.rs:39-60 ======================
        // even lanes evaluate the left shape, odd lanes the right one; the G/2 lanes of a side share
        // the candidate list of a polyhedron's support cell (support_point_split)
        f3 sup_v, sup_a;
        {
#if CB2_OPT_SPLIT
            const f3 w = support_point_split<G / 2, 2>(T, sh, xf, right ? -n : n, (uint32_t)gl >> 1);
#else
            const f3 w = support_point(T, sh, xf, right ? -n : n);
#endif
            const f3 o = mk3(__shfl_xor_sync(0xffffffffu, w.x, 1), __shfl_xor_sync(0xffffffffu, w.y, 1),
                             __shfl_xor_sync(0xffffffffu, w.z, 1));
            sup_a = right ? o : w;
            sup_v = sup_a - (right ? w : o);
        }
        // a HalfEdge polyhedron's hill climb met a NaN direction: the reference never returns
        // (the vote is taken by every lane BEFORE it is combined with `run`: `run && vote(..)`
        // would let the lanes of idle groups skip a full-mask vote)
        const unsigned hang_votes = __ballot_sync(0xffffffffu, sh.hang != 0u);
        const bool hung = run && (hang_votes & gmask) != 0u;
        // epa3d.rs:46-64: stop when the new point is within tolerance of the face, or at the cap
        const bool conv = run && !hung && (fsub(dot(sup_v, n), fb.w) < epa_tol || it >= max_it);
        const bool add = run && !hung && !conv;

        // ================= Polytope::add (epa3d.rs:147-175) =============================================
        // (1) visibility of every face -> bit mask (uniform across the group).  The same pass tracks
        //     the smallest distance among the faces that stay (cd, ci per lane): together with the
        //     distances Face::new computes below that is every face of the next polytope, so the next
        //     closest_face_to_origin is their plain minimum — unless two distances are exactly equal
        //     or one is NaN (ctie), where Vec order decides and the scan above runs instead.
        bitmask<FW> vis;
        vis.clear();
        float cd = INFINITY;
        int ci = -1;
        uint32_t ctie = 0u;
        const int vis_top = __reduce_max_sync(0xffffffffu, add ? nf : 0);  // most faces any group tests
        {
            const int nf_max = vis_top;
            if constexpr (FW <= 4) {
                // Lean trips for the small tiers: shared-memory addresses advance by constants, every
                // lane loads (slots past nf lie inside the record and are never used; the loads are
                // volatile asm so that the compiler keeps them out of a branch) and only the result is
                // predicated; the group's nibble of each ballot is appended to a 32-bit word by one
                // multiply-add (the fields are disjoint, so + is |).
                const uint32_t sb = (uint32_t)__cvta_generic_to_shared(&P);
                const uint32_t s_pv = sb + (uint32_t)offsetof(PS, pv);
                // (a slot past nf may hold anything: its vertex index is masked to stay inside the record)
                constexpr uint32_t VM = VCAP <= 32 ? 31u : (VCAP <= 64 ? 63u : 127u);
                static_assert(sizeof(PS) - offsetof(PS, pv) >= (VM + 1u) * 16u, "index mask of the lean visibility pass");
#pragma unroll
                for (int w = 0; w < FW; ++w) {
                    uint32_t acc = 0u, pw = 1u;
                    uint32_t a_f = sb + (uint32_t)offsetof(PS, fnd) + (uint32_t)(32 * w + gl) * 16u;
                    uint32_t a_i = sb + (uint32_t)offsetof(PS, fidx) + (uint32_t)(32 * w + gl) * 4u;
                    const int f_hi = min(nf_max, 32 * (w + 1));
#if CB2_VIS2
                    // two faces per lane and trip: two independent load -> dot chains in flight
                    for (int f = 32 * w + gl; f - gl < f_hi; f += 2 * G, a_f += 32u * G, a_i += 8u * G) {
                        const float4 fn0 = lds128(a_f), fn1 = lds128(a_f + 16u * G);
                        const uint32_t ix0 = lds32(a_i), ix1 = lds32(a_i + 4u * G);
                        const float4 p0 = lds128(s_pv + ((ix0 & VM) << 4)), p1 = lds128(s_pv + ((ix1 & VM) << 4));
                        const float dv0 = dot(mk3(fn0.x, fn0.y, fn0.z), sup_v - mk3(p0.x, p0.y, p0.z));
                        const float dv1 = dot(mk3(fn1.x, fn1.y, fn1.z), sup_v - mk3(p1.x, p1.y, p1.z));
                        const bool v0 = add & (f < nf) & (dv0 > 0.0f);
                        const bool v1 = add & (f + G < nf) & (dv1 > 0.0f);
#if CB2_OPT_FOLD
                        {   // smallest distance among the faces that stay (equal or NaN: Vec order decides)
                            const bool s0 = add & (f < nf) & !v0;
                            const bool l0 = s0 & (fn0.w < cd);
                            ctie |= (uint32_t)(s0 & !l0 & !(fn0.w > cd));
                            cd = l0 ? fn0.w : cd;
                            ci = l0 ? f : ci;
                            const bool s1 = add & (f + G < nf) & !v1;
                            const bool l1 = s1 & (fn1.w < cd);
                            ctie |= (uint32_t)(s1 & !l1 & !(fn1.w > cd));
                            cd = l1 ? fn1.w : cd;
                            ci = l1 ? f + G : ci;
                        }
#endif
                        const uint32_t n0 = (__ballot_sync(0xffffffffu, v0) >> (gw * G)) & gbits;
                        const uint32_t n1 = (__ballot_sync(0xffffffffu, v1) >> (gw * G)) & gbits;
                        acc += (n0 | (n1 << G)) * pw;
                        pw <<= 2 * G;
                    }
#else
                    for (int f = 32 * w + gl; f - gl < f_hi; f += G, a_f += 16u * G, a_i += 4u * G) {
                        const float4 fn = lds128(a_f);
                        const uint32_t idx = lds32(a_i);
                        const float4 p0 = lds128(s_pv + ((idx & VM) << 4));
                        const float dv = dot(mk3(fn.x, fn.y, fn.z), sup_v - mk3(p0.x, p0.y, p0.z));
                        const bool live = add & (f < nf);
                        const bool v = live & (dv > 0.0f);
#if CB2_OPT_FOLD
                        const bool stay = live & !v;
                        const bool lt = stay & (fn.w < cd);
                        ctie |= (uint32_t)(stay & !lt & !(fn.w > cd));  // equal or NaN: Vec order decides
                        cd = lt ? fn.w : cd;
                        ci = lt ? f : ci;
#endif
                        const uint32_t nib = (__ballot_sync(0xffffffffu, v) >> (gw * G)) & gbits;
                        acc += nib * pw;
                        pw <<= G;
                    }
#endif
                    vis.set_word(w, acc);
                }
            } else {
                CB2_UNROLL(CB2_U_VIS)
                for (int f0 = 0; f0 < nf_max; f0 += G) {
                    const int f = f0 + gl;
                    bool v = false;
                    if (add && f < nf) {
                        const float4 fn = P.fnd[f];
                        const uint32_t idx = P.fidx[f];
                        v = dot(mk3(fn.x, fn.y, fn.z), sup_v - ps_pv(P, idx & 255u)) > 0.0f;
#if CB2_OPT_FOLD
                        if (!v) {
                            if (fn.w < cd) { cd = fn.w; ci = f; }
                            else if (!(fn.w > cd)) ctie = 1u;
                        }
#endif
                    }
                    const uint32_t bits = (__ballot_sync(0xffffffffu, v) >> (gw * G)) & gbits;
                    vis.or_bits(f0, bits);
                }
            }
        }
        // (2) the swap_remove sweep (epa_walk above)
        int len = nf, nvis = 0, nmv = 0, mine = 0;
        bool overflow = false;
        if constexpr (FW <= 2) {
            // while no polytope of the warp has more than 32 faces the walk runs on one word
            if (vis_top <= 32) {
                bitmask32 v32;
                v32.m = vis.low_word();
                epa_walk<G, OCAP>(v32, P, add, gl, len, nvis, nmv, mine, overflow, ci);
            } else {
                epa_walk<G, OCAP>(vis, P, add, gl, len, nvis, nmv, mine, overflow, ci);
            }
        } else {
            epa_walk<G, OCAP>(vis, P, add, gl, len, nvis, nmv, mine, overflow, ci);
        }
        __syncwarp();
        // (3) horizon = edges of the examined faces, in examination order, whose reverse edge is not
        //     an edge of an examined face (remove_or_add_edge, epa3d.rs:212-219).  One examined face
        //     per lane; its edges are (v0,v1) (v1,v2) (v2,v0).
        const bool edges = add && !overflow;
        const int nvis_max = __reduce_max_sync(0xffffffffu, edges ? nvis : 0);
        bool dup = false;
        const uint32_t idx_mine = (edges && gl < nvis) ? P.fidx[mine] : 0u;
        // Three passes over the examined faces — set the directed-edge bits, keep the edges whose
        // reverse bit is clear, clear the bits again — each run straight-line for the first G faces
        // (lane k holds the k-th examined face in a register); more than G examined faces are
        // rare and take the rolled loops behind.
        auto pass_set = [&](uint32_t idx) {
            const uint32_t v0 = idx & 255u, v1 = (idx >> 8) & 255u, v2 = (idx >> 16) & 255u;
            // (one word per row when VCAP <= 32: no word index, no bit mask)
            const uint32_t o0 = atomicOr(P.adj_word_ptr(adj_word(v0, v1)), 1u << adj_bit(v1));
            const uint32_t o1 = atomicOr(P.adj_word_ptr(adj_word(v1, v2)), 1u << adj_bit(v2));
            const uint32_t o2 = atomicOr(P.adj_word_ptr(adj_word(v2, v0)), 1u << adj_bit(v0));
            dup |= ((o0 >> adj_bit(v1)) | (o1 >> adj_bit(v2)) | (o2 >> adj_bit(v0))) & 1u;
        };
        int ne = 0;
        auto pass_keep = [&](uint32_t idx, bool active) {
            uint32_t keep = 0, v0 = 0, v1 = 0, v2 = 0;
            if (active) {
                v0 = idx & 255u; v1 = (idx >> 8) & 255u; v2 = (idx >> 16) & 255u;
                keep = (~(*P.adj_word_ptr(adj_word(v1, v0)) >> adj_bit(v0)) & 1u) |
                       ((~(*P.adj_word_ptr(adj_word(v2, v1)) >> adj_bit(v1)) & 1u) << 1) |
                       ((~(*P.adj_word_ptr(adj_word(v0, v2)) >> adj_bit(v2)) & 1u) << 2);
            }
            // (three keep bits per lane: one word up to 8 lanes per pair, two beyond)
            using keep_t = typename std::conditional<(3 * G > 32), unsigned long long, uint32_t>::type;
            keep_t all = (keep_t)keep << (3 * gl);
#pragma unroll
            for (int off = G / 2; off >= 1; off >>= 1) all |= __shfl_xor_sync(0xffffffffu, all, off);
            int pos = ne + popc_any(all & (((keep_t)1 << (3 * gl)) - 1));
            if (keep & 1u) { if (pos < OCAP) P.elist[pos] = (uint16_t)(v0 | (v1 << 8)); ++pos; }
            if (keep & 2u) { if (pos < OCAP) P.elist[pos] = (uint16_t)(v1 | (v2 << 8)); ++pos; }
            if (keep & 4u) { if (pos < OCAP) P.elist[pos] = (uint16_t)(v2 | (v0 << 8)); }
            ne += popc_any(all);
        };
        auto pass_clear = [&](uint32_t idx) {
            const uint32_t v0 = idx & 255u, v1 = (idx >> 8) & 255u, v2 = (idx >> 16) & 255u;
#pragma unroll
            for (int q = 0; q < AW; ++q) {
                *P.adj_word_ptr(v0 * AW + q) = 0u;
                *P.adj_word_ptr(v1 * AW + q) = 0u;
                *P.adj_word_ptr(v2 * AW + q) = 0u;
            }
        };
        const bool act0 = edges && gl < nvis;
        if (act0) pass_set(idx_mine);
#pragma unroll 1
        for (int k = G + gl; k - gl < nvis_max; k += G)
            if (edges && k < nvis) pass_set(P.fidx[P.order[k]]);
        __syncwarp();
        pass_keep(idx_mine, act0);
#pragma unroll 1
        for (int k = G + gl; k - gl < nvis_max; k += G) {
            const bool act = edges && k < nvis;
            pass_keep(act ? P.fidx[P.order[k]] : 0u, act);
        }
        dup = (__ballot_sync(0xffffffffu, dup) & gmask) != 0u;
        __syncwarp();
        // (4) clear the adjacency bits again
        if (act0) pass_clear(idx_mine);
#pragma unroll 1
        for (int k = G + gl; k - gl < nvis_max; k += G)
            if (edges && k < nvis) pass_clear(P.fidx[P.order[k]]);
        const bool spill = add && (overflow || dup || ne > OCAP || nv >= VCAP || len + ne > FCAP);
        const bool grow = add && !spill;
        __syncwarp();
        // (5) apply the swap_remove moves, append the vertex
        if (grow) {
            for (int m = gl; m < nmv; m += G) {
                const uint32_t mvr = P.mv[m];
                const int dst = (int)(mvr & 255u), src = (int)(mvr >> 8);
                P.fnd[dst] = P.fnd[src];
                P.fidx[dst] = P.fidx[src];
            }
            if (gl == 0) {
                P.pv[nv] = make_float4(sup_v.x, sup_v.y, sup_v.z, 0.f);
                pa[nv] = make_float4(sup_a.x, sup_a.y, sup_a.z, 0.f);
            }
        }
        __syncwarp();
        // (6) build the new faces (one horizon edge per lane)
        {
            const int ne_max = __reduce_max_sync(0xffffffffu, grow ? ne : 0);
            CB2_UNROLL(CB2_U_FACE)
            for (int k = gl; k < ne_max; k += G) {
                if (grow && k < ne) {
                    const uint32_t e = P.elist[k];
                    const float dist = ps_face_new(P, len + k, (uint32_t)nv, e & 255u, e >> 8);
#if CB2_OPT_FOLD
                    if (dist < cd) { cd = dist; ci = len + k; }
                    else if (!(dist > cd)) ctie = 1u;
#else
                    (void)dist;
#endif
                }
            }
        }
#if CB2_OPT_FOLD
        // the group's minimum of the tracked distances = the next closest face, unless tied
#pragma unroll
        for (int off = G / 2; off >= 1; off >>= 1) {
            const float ob = __shfl_xor_sync(0xffffffffu, cd, off);
            const int oi = __shfl_xor_sync(0xffffffffu, ci, off);
            if (ob < cd) { cd = ob; ci = oi; }
            else if (ob == cd && (oi | ci) >= 0) ctie = 1u;
        }
        const bool gtie = (__ballot_sync(0xffffffffu, ctie != 0u) & gmask) != 0u;
        const int next_best = (grow && !gtie) ? ci : -1;
#else
        const int next_best = -1;
#endif
        uint8_t result = ST_NONE;
        if (grow) {
            ++nv;
            nf = len + ne;
            ++it;
            if (nf == 0) result = ST_PANIC_EPA_EMPTY;  // faces[0] on an empty Vec (epa3d.rs:138)
        }
        if (spill) result = ST_SCRATCH_OVERFLOW;
        if (conv) result = ST_SOME;
        if (hung) result = ST_NONCONVERGED;

        // ================= finished pairs (divergent) ====================================================
        if (result != ST_NONE) {
            contact_out co = zero_contact();
            if (conv) {
                // contact(), epa3d.rs:84-114
                const uint32_t idx = P.fidx[best_face];
                const uint32_t i0 = idx & 255u, i1 = (idx >> 8) & 255u, i2 = (idx >> 16) & 255u;
                float u, v, w;
                barycentric_vector(n * fb.w, ps_pv(P, i0), ps_pv(P, i1), ps_pv(P, i2), u, v, w);
                const float4 a0 = pa[i0], a1 = pa[i1], a2 = pa[i2];
                co.normal = n;
                co.depth = fb.w;
                co.point = (mk3(a0.x, a0.y, a0.z) * u + mk3(a1.x, a1.y, a1.z) * v) +
                           mk3(a2.x, a2.y, a2.z) * w;
            }
            if (result == ST_SCRATCH_OVERFLOW && over_list) {
                // re-queue for the next tier, with the polytope as it stands (this iteration has
                // not modified it) so that tier resumes instead of starting over
                using DL = dump_layout<FCAP, VCAP>;
                uint32_t h = 0;
                if (gl == 0) h = atomicAdd(over_count, 1u);
                h = __shfl_sync(gmask, h, gw * G);
                const bool room = h < C.W.dump_cap[C.tier];
                if (room) {
                    float4* rec = C.W.dump[C.tier] + (size_t)h * DL::SIZE;
                    if (gl == 0)
                        rec[0] = make_float4(__uint_as_float((uint32_t)nf), __uint_as_float((uint32_t)nv),
                                             __uint_as_float(it), 0.f);
                    // the polytope leaves shared memory as three bulk copies (TMA): ordinary writes
                    // are fenced towards the async proxy first, and the elected lane waits until the
                    // copies have read their source before the group reuses it
                    fence_proxy_async();
                    __syncwarp(gmask);
                    if (gl == 0) {
                        bulk_store(rec + DL::FND, P.fnd, (uint32_t)nf * 16u);
                        bulk_store(rec + DL::PV, P.pv, (uint32_t)nv * 16u);
                        bulk_store(rec + DL::FIDX, P.fidx, ((uint32_t)nf * 4u + 15u) & ~15u);
                        bulk_commit();
                    }
                    for (int j = gl; j < nv; j += G) rec[DL::PA + j] = pa[j];
                    if (gl == 0) bulk_wait_read();
                    __syncwarp(gmask);
                }
                if (gl == 0) over_list[h] = pair | (room ? OVER_RESUME : 0u);
            } else if (gl == 0) {
                store_contact(A.out_contact, pair, CB2_FULL_RESOLUTION, co);
                A.out_status[pair] = result;
            }
            state = S_FETCH;
            nf = 0;
        }
        best_face = next_best;
        __syncwarp();
    }
}

// ---- lane-per-pair tiers (cb2_epa_lane.cuh) ------------------------------------------------------------
#ifndef CB2_LANE_PREFETCH
#define CB2_LANE_PREFETCH 1
#endif
#ifndef CB2_LANE_GLOBAL   // 1: the polytopes live in global memory (L1/L2-backed) instead of shared memory
#define CB2_LANE_GLOBAL 0
#endif
#ifndef CB2_LANE_WARPS    // blocks (= warps) per SM when CB2_LANE_GLOBAL
#define CB2_LANE_WARPS 12
#endif
template <int FCAP, int VCAP, int ECAP>
constexpr int lane_blocks_per_sm() {
    if (CB2_LANE_GLOBAL) return CB2_LANE_WARPS;
    // 227 KB of shared memory per SM, 1 KB reserved per block
    return (int)((227u * 1024u) / (sizeof(lane_store<FCAP, VCAP, ECAP>) + 1024u));
}
// per-block scratch in global memory: [lane_store when CB2_LANE_GLOBAL] sup_a[VCAP][32]
template <int FCAP, int VCAP, int ECAP>
struct lane_scratch {
    static constexpr size_t BYTES =
        (CB2_LANE_GLOBAL ? sizeof(lane_store<FCAP, VCAP, ECAP>) : 0) + (size_t)VCAP * 32 * sizeof(float4);
};
template <int FCAP, int VCAP, int ECAP, int RF, int RV>
__global__ void __launch_bounds__(32, (lane_blocks_per_sm<FCAP, VCAP, ECAP>()))
k_epa_lane(epa_args C) {
    extern __shared__ __align__(16) float smem[];
    constexpr size_t SCRATCH = lane_scratch<FCAP, VCAP, ECAP>::BYTES;
    char* scratch = reinterpret_cast<char*>(C.W.pa) + (size_t)blockIdx.x * SCRATCH;
    auto& S = *reinterpret_cast<lane_store<FCAP, VCAP, ECAP>*>(CB2_LANE_GLOBAL ? (void*)scratch : (void*)smem);
    float4* pa = reinterpret_cast<float4*>(scratch + (CB2_LANE_GLOBAL ? sizeof(lane_store<FCAP, VCAP, ECAP>) : 0));
    epa_lane_run<FCAP, VCAP, ECAP, RF, RV, CB2_LANE_PREFETCH != 0>(C, S, pa, threadIdx.x);
}

// ---- warp-team tiers (cb2_epa_team.cuh) --------------------------------------------------------------
template <int W, int FCAP, int VCAP, int ECAP>
constexpr int team_blocks_per_sm() {
    return (int)((227u * 1024u) / (sizeof(team_store<W, FCAP, VCAP, ECAP>) + 1024u));
}
template <int W, int FCAP, int VCAP, int ECAP, int RF, int RV>
__global__ void __launch_bounds__(32 * W, (team_blocks_per_sm<W, FCAP, VCAP, ECAP>()))
k_epa_team(epa_args C) {
    using T = epa_team<W, FCAP, VCAP, ECAP, RF, RV>;
    extern __shared__ __align__(16) float smem[];
    auto& S = *reinterpret_cast<typename T::store*>(smem);
    float4* pa = C.W.pa + (size_t)blockIdx.x * VCAP * 32;
    const int w = threadIdx.x >> 5;
    const unsigned lane = threadIdx.x & 31u;
    typename T::regs R;
    T::init(R);
    for (;;) {
        T::part0(C, S, pa, R, w, lane);
        __syncthreads();
        T::part1(C, S, pa, R, w, lane);
        __syncthreads();
        T::part2(C, S, pa, R, w, lane);
        __syncthreads();
        T::part3(C, S, pa, R, w, lane);
        __syncthreads();
        T::part4(C, S, pa, R, w, lane);
        __syncthreads();
        T::part5(C, S, pa, R, w, lane);
        if (__syncthreads_and(R.state == L_EXIT)) break;
    }
}

// ---- host side ----------------------------------------------------------------------------------------
// Which kernel runs tier 0 (and, with CB2_LANE_TIER1, tier 1).  Measured on one B200, cfg 3 at 4 M
// pairs (profiles/README.md, round 2): the group kernel k_epa 14.6 ms, the lane kernel 29.1 ms, the
// warp-team kernel 26.7-29.3 ms (W = 3 / 4) — all three bit-identical to the oracle.  The group
// kernel stays the product path; the other two are kept as measured experiments (A/B builds:
// tools/ab_epa.sh name "-DCB2_TIER0_KIND=2 -DCB2_TEAM_W=3") and as the code the host simulation runs.
#ifndef CB2_TIER0_KIND    // 0: group kernel (4 lanes of one warp per pair), 1: lane kernel, 2: warp-team kernel
#define CB2_TIER0_KIND 0
#endif
#ifndef CB2_TEAM_W        // warps per 32 pairs of the warp-team kernel
#define CB2_TEAM_W 4
#endif
constexpr bool TIER0_GROUP = CB2_TIER0_KIND == 0, TIER0_TEAM = CB2_TIER0_KIND == 2;
// tier 0 (and optionally tier 1): lane per pair        faces  verts  edge/move list
#ifndef CB2_L0_F
#define CB2_L0_F 40
#define CB2_L0_V 24
#define CB2_L0_E 16
#endif
#ifndef CB2_LANE_TIER1   // 1: tier 1 runs the lane kernel as well (caps CB2_L1_*)
#define CB2_LANE_TIER1 0
#endif
#ifndef CB2_L1_F
#define CB2_L1_F 64
#define CB2_L1_V 32
#define CB2_L1_E 24
#endif
constexpr int L0_F = TIER0_GROUP ? CB2_T0_F : CB2_L0_F, L0_V = TIER0_GROUP ? CB2_T0_V : CB2_L0_V, L0_E = CB2_L0_E;
constexpr int LK_F = CB2_L0_F, LK_V = CB2_L0_V;  // caps the lane / team kernels are instantiated with
constexpr int L1_F = CB2_L1_F, L1_V = CB2_L1_V, L1_E = CB2_L1_E;
// group tiers (k_epa above):      G   faces  verts  examined/horizon cap
#ifndef CB2_T1_G
#define CB2_T1_G 8
#endif
#ifndef CB2_T1_F
#define CB2_T1_F 72
#define CB2_T1_V 40
#define CB2_T1_O 32
#endif
#ifndef CB2_T0_G
#define CB2_T0_G 4
#endif
constexpr int T0_G = CB2_T0_G, T0_F = CB2_T0_F, T0_V = CB2_T0_V, T0_O = CB2_T0_O;
constexpr bool LANE_T1 = CB2_LANE_TIER1 && !TIER0_GROUP;
constexpr int T1_G = CB2_T1_G, T1_F = LANE_T1 ? L1_F : CB2_T1_F, T1_V = LANE_T1 ? L1_V : CB2_T1_V,
              T1_O = CB2_T1_O;
#ifndef CB2_T2_G   // 16 lanes per pair in the last tier: it is bound by the latency of its longest pairs
#define CB2_T2_G 16  // (cfg 3: 0.48 -> 0.36 ms; cfg 2's curved pairs, which start here: 96 -> 128 M pairs/s)
#endif
constexpr int T2_G = CB2_T2_G, T2_F = 208, T2_V = 104, T2_O = 96;  // manifold worst case at 100 iterations

template <int G, int F, int V, int O, int RF, int RV>
static int epa_blocks(int sm_count) {
    auto kern = k_epa<G, F, V, O, RF, RV>;
    const size_t smem = group_layout<poly_smem<F, V, O>, 32 / G>::BYTES;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 32, smem) != cudaSuccess || per_sm < 1)
        per_sm = 1;
    return per_sm * sm_count;
}
template <int F, int V, int E, int RF, int RV>
static int lane_blocks(int sm_count) {
    auto kern = k_epa_lane<F, V, E, RF, RV>;
    const size_t smem = CB2_LANE_GLOBAL ? 0 : sizeof(lane_store<F, V, E>);
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 32, smem) != cudaSuccess || per_sm < 1)
        per_sm = 1;
    return per_sm * sm_count;
}

template <int G, int F, int V, int O, int RF, int RV>
static cudaError_t launch_tier(const epa_args& C, int sm_count, cudaStream_t st) {
    const size_t smem = group_layout<poly_smem<F, V, O>, 32 / G>::BYTES;
    k_epa<G, F, V, O, RF, RV><<<epa_blocks<G, F, V, O, RF, RV>(sm_count), 32, smem, st>>>(C);
    return cudaGetLastError();
}
template <int W, int F, int V, int E, int RF, int RV>
static int team_blocks(int sm_count) {
    auto kern = k_epa_team<W, F, V, E, RF, RV>;
    const size_t smem = sizeof(team_store<W, F, V, E>);
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 32 * W, smem) != cudaSuccess || per_sm < 1)
        per_sm = 1;
    return per_sm * sm_count;
}
template <int W, int F, int V, int E, int RF, int RV>
static cudaError_t launch_team_tier(const epa_args& C, int sm_count, cudaStream_t st) {
    k_epa_team<W, F, V, E, RF, RV><<<team_blocks<W, F, V, E, RF, RV>(sm_count), 32 * W,
                                     sizeof(team_store<W, F, V, E>), st>>>(C);
    return cudaGetLastError();
}
template <int F, int V, int E, int RF, int RV>
static cudaError_t launch_lane_tier(const epa_args& C, int sm_count, cudaStream_t st) {
    k_epa_lane<F, V, E, RF, RV><<<lane_blocks<F, V, E, RF, RV>(sm_count), 32,
                                  CB2_LANE_GLOBAL ? 0 : sizeof(lane_store<F, V, E>), st>>>(C);
    return cudaGetLastError();
}

static size_t epa_pa_region_bytes(int sm_count) {
    size_t m, b;
#if CB2_TIER0_KIND == 0
    m = (size_t)epa_blocks<T0_G, T0_F, T0_V, T0_O, 0, 0>(sm_count) * (32 / T0_G) * T0_V * sizeof(float4);
    b = (size_t)epa_blocks<T1_G, T1_F, T1_V, T1_O, L0_F, L0_V>(sm_count) * (32 / T1_G) * T1_V * sizeof(float4);
#else
    constexpr int TW = CB2_TEAM_W;
    if (TIER0_TEAM) m = (size_t)team_blocks<TW, LK_F, LK_V, L0_E, 0, 0>(sm_count) * 32 * LK_V * sizeof(float4);
    else m = (size_t)lane_blocks<LK_F, LK_V, L0_E, 0, 0>(sm_count) * lane_scratch<LK_F, LK_V, L0_E>::BYTES;
    if (CB2_LANE_TIER1 && TIER0_TEAM)
        b = (size_t)team_blocks<TW, L1_F, L1_V, L1_E, LK_F, LK_V>(sm_count) * 32 * L1_V * sizeof(float4);
    else if (CB2_LANE_TIER1)
        b = (size_t)lane_blocks<L1_F, L1_V, L1_E, LK_F, LK_V>(sm_count) * lane_scratch<L1_F, L1_V, L1_E>::BYTES;
    else
        b = (size_t)epa_blocks<T1_G, T1_F, T1_V, T1_O, L0_F, L0_V>(sm_count) * (32 / T1_G) * T1_V * sizeof(float4);
#endif
    const size_t c = (size_t)epa_blocks<T2_G, T2_F, T2_V, T2_O, T1_F, T1_V>(sm_count) * (32 / T2_G) * T2_V *
                     sizeof(float4);
    m = m > b ? m : b;
    m = m > c ? m : c;
    return m;
}
// tier 0's supply of prepared hits sits behind the sup_a scratch: 32 records per warp
static size_t epa_supply_bytes(int sm_count) {
#if CB2_TIER0_KIND == 0
    return (size_t)epa_blocks<T0_G, T0_F, T0_V, T0_O, 0, 0>(sm_count) * 32 * SUP_SIZE * sizeof(float4);
#else
    return 0;
#endif
}
size_t epa_pa_scratch_bytes(int sm_count) { return epa_pa_region_bytes(sm_count) + epa_supply_bytes(sm_count); }
size_t epa_dump_record_bytes(int tier) {
    return (tier == 0 ? dump_layout<L0_F, L0_V>::SIZE : dump_layout<T1_F, T1_V>::SIZE) * sizeof(float4);
}

cudaError_t launch_epa_tier(int tier, const narrow_args& A, const narrow_ws& W, int sm_count,
                            cudaStream_t st) {
    epa_args C;
    C.A = A;
    C.W = W;
    C.tier = (uint32_t)tier;
    C.sm_count = (uint32_t)sm_count;
    C.supply = reinterpret_cast<float4*>(reinterpret_cast<char*>(W.pa) + epa_pa_region_bytes(sm_count));
    if (tier == 0) {
#if CB2_TIER0_KIND == 0
        return launch_tier<T0_G, T0_F, T0_V, T0_O, 0, 0>(C, sm_count, st);
#else
        constexpr int TW = CB2_TEAM_W;
        if (TIER0_TEAM) return launch_team_tier<TW, LK_F, LK_V, L0_E, 0, 0>(C, sm_count, st);
        return launch_lane_tier<LK_F, LK_V, L0_E, 0, 0>(C, sm_count, st);
#endif
    }
    if (tier == 1) {
#if CB2_TIER0_KIND != 0
        constexpr int TW = CB2_TEAM_W;
        if (CB2_LANE_TIER1 && TIER0_TEAM) return launch_team_tier<TW, L1_F, L1_V, L1_E, LK_F, LK_V>(C, sm_count, st);
        if (CB2_LANE_TIER1) return launch_lane_tier<L1_F, L1_V, L1_E, LK_F, LK_V>(C, sm_count, st);
#endif
        return launch_tier<T1_G, T1_F, T1_V, T1_O, L0_F, L0_V>(C, sm_count, st);
    }
    return launch_tier<T2_G, T2_F, T2_V, T2_O, T1_F, T1_V>(C, sm_count, st);
}

}  // namespace cb2
