// cb2_epa_team.cuh — EPA3 (EPA3::process, epa3d.rs:27-65) with a TEAM OF W WARPS per 32 pairs.
//
// What the measurements of this round said (profiles/README.md, round 2):
//  * round 1's kernel (4 lanes of one warp per pair, cb2_epa.cu) is issue-bound: the lanes of a
//    group that have nothing to do in a phase (two of four during a Minkowski support, all but one
//    in the swap_remove walk) are predicated off but their warp still issues every instruction —
//    20.8 of 32 lanes active, 205 warp instructions per pair-iteration;
//  * one lane per pair (cb2_epa_lane.cuh) removes that waste (1.6x fewer warp instructions) but
//    shared memory then allows only 5 warps per SM and the kernel drowns in latency (IPC 0.69).
// This kernel keeps the lane-per-pair data layout — the polytopes of 32 pairs in shared memory as
// [index][lane], bank-conflict free for any per-lane index — and puts W warps on those 32 pairs:
// lane l of EVERY warp of the block works on pair l.  Work is split by WARP, not by lane:
//  * warp 0 evaluates the left shape's support point, warp 1 the right one;
//  * the per-face loops (closest face, visibility, Face::new, swap_remove moves) are striped over
//    the warps (face f belongs to warp f mod W);
//  * the serial part (the swap_remove order and the horizon) runs on warp 0 alone.
// A warp with nothing to do in a phase waits at the block barrier and costs NO issue slots, every
// instruction that is issued works on up to 32 pairs, and an SM holds 5 x W warps.  Every warp
// tracks the pair's scalar state itself (replicated state machine: all of them read the same
// published values and take the same decisions), so the only traffic between warps is a few words
// per pair per phase in shared memory, separated by six block barriers per round.
//
// The horizon (remove_or_add_edge, epa3d.rs:212-219) is computed per lane from a vertex-adjacency
// bit matrix kept in the spare word of each vertex record: edges of the examined faces, in
// examination order, whose reverse edge is not an edge of an examined face.  That equals the
// reference's list whenever no directed edge occurs twice; a duplicate is detected while the bits
// are set and the lane then replays remove_or_add_edge literally.
//
// The body compiles for the host (cb2_hostsim.h); tests/host_sim runs it phase by phase.
#pragma once
#include "cb2_epa_lane.cuh"

namespace cb2 {

template <int W, int FCAP, int VCAP, int ECAP>
struct team_store {
    float4 fnd[FCAP * 32];        // normal.xyz, distance
    float4 pv[VCAP * 32];         // SupportPoint.v; .w = the vertex's adjacency row (all zero between rounds)
    uint32_t fidx[FCAP * 32];     // v0 | v1<<8 | v2<<16
    uint16_t el[ECAP * 32];       // horizon edges a | b<<8, in the reference's Vec order
    uint16_t mv[ECAP * 32];       // swap_remove moves: dst | src<<8
    uint8_t ex[ECAP * 32];        // slots of the examined (visible) faces, in examination order
    float sup[6 * 32];            // support points of the two shapes
    uint32_t part[W * 2 * 32];    // per-warp partial results: closest face (distance, slot), then visibility (lo, hi)
    uint32_t ctl[4 * 32];         // fetch: entry, a, b, k   support: -, -, hang l, hang r   walk: packed, dump slot
};

#ifdef CB2_HOST_SIM
static bool g_sim_force_literal_horizon = false;  // tests/host_sim: exercise the rare non-manifold path
#endif
constexpr uint32_t TEAM_NO_ENTRY = 0xFFFFFFFFu;
constexpr uint32_t TEAM_NO_FACE = 0xFFFFFFFFu;

template <int W, int FCAP, int VCAP, int ECAP, int RF, int RV>
struct epa_team {
    static_assert(W >= 2 && W <= 8, "warp 0 and warp 1 evaluate one shape each");
    static_assert(FCAP <= 64, "the visibility mask is one 64-bit word");
    static_assert(VCAP <= 32, "a vertex's adjacency row is one word");
    static_assert(ECAP <= 255, "indices are bytes");
    static_assert(RF <= FCAP && RV <= VCAP, "a resumed polytope must fit");
    using store = team_store<W, FCAP, VCAP, ECAP>;
    using DLR = dump_layout<(RF > 0 ? RF : 1), (RV > 0 ? RV : 1)>;
    using DL = dump_layout<FCAP, VCAP>;

    struct regs {
        int state;
        uint32_t pair, it, apex;
        int nf, nv, len, ne;
        bool init4;
        int best;
        float4 fb;
        f3 sv, sa;
        bool live, add, conv;
        uint8_t result;
        shape sh;  // warp 0: the left shape, warp 1: the right one
        xform xf;
        // warp 0: the claimed next work-list entry (see cb2_epa_lane.cuh, PF)
        int nx_stage;
        uint32_t nx_k, nx_entry, nx_a, nx_b;
    };

    static CB2_DI void init(regs& R) {
        R.state = L_FETCH;
        R.pair = R.it = R.apex = 0;
        R.nf = R.nv = R.len = R.ne = 0;
        R.init4 = false;
        R.best = 0;
        R.fb = make_float4(0.f, 0.f, 0.f, 0.f);
        R.sv = R.sa = zero3();
        R.live = R.add = R.conv = false;
        R.result = ST_NONE;
        R.sh.kind = K_CUBOID; R.sh.h = zero3(); R.sh.hang = 0u;
        R.sh.poly.voff = 0; R.sh.poly.vcnt = 0; R.sh.poly.cell_off = 0; R.sh.poly.cfg = 0;
        R.xf = make_xform(make_float4(1.f, 1.f, 0.f, 0.f), make_float4(0.f, 0.f, 0.f, 0.f));
        R.nx_stage = 0;
        R.nx_k = R.nx_entry = R.nx_a = R.nx_b = 0;
    }

#define FND(i) S.fnd[(i) * 32 + lane]
#define PV(i) S.pv[(i) * 32 + lane]
#define ADJ(i) (reinterpret_cast<uint32_t*>(&S.pv[(i) * 32 + lane])[3])
#define FIDX(i) S.fidx[(i) * 32 + lane]
#define EL(i) S.el[(i) * 32 + lane]
#define MV(i) S.mv[(i) * 32 + lane]
#define EX(i) S.ex[(i) * 32 + lane]
#define SUP(i) S.sup[(i) * 32 + lane]
#define PART(i) S.part[(i) * 32 + lane]
#define CTL(i) S.ctl[(i) * 32 + lane]
#define PA(i) pa[(i) * 32 + lane]

    // ---- part 0 (warp 0): claim work for the lanes whose pair finished, publish it -------------------
    static CB2_DI void part0(const epa_args& C, store& S, float4* __restrict__ pa, regs& R, const int w,
                             const unsigned lane) {
        if (w != 0) return;
        const narrow_args& A = C.A;
        const uint32_t* const list = C.W.list[C.tier];
        const uint32_t count = C.W.EC->n_list[C.tier];
        uint32_t* const cursor = &C.W.EC->next[C.tier];
        if (R.nx_stage == 1 && R.state == L_RUN) {
            // second hop of the claimed entry (its first hop was issued a round ago)
            const uint32_t np = R.nx_entry & ~OVER_RESUME;
            if (A.pair_body) {
                const uint2 b = __ldg(reinterpret_cast<const uint2*>(A.pair_body) + np);
                R.nx_a = b.x;
                R.nx_b = b.y;
            } else {
                R.nx_a = __ldg(A.l_shape + np);
                R.nx_b = __ldg(A.r_shape + np);
                prefetch_l2(A.l_t + 2 * (size_t)np);
                prefetch_l2(A.r_t + 2 * (size_t)np);
            }
            if (!(RF > 0 && (R.nx_entry & OVER_RESUME))) prefetch_l2(C.W.simplex + (size_t)np * 8);
            R.nx_stage = 2;
        }
        const bool want = R.state == L_FETCH;
        uint32_t k = 0, entry = 0, a = 0, b = 0;
        bool have = false, resolved = false;
        {   // only the first pair of a lane is claimed here (the vote inside is taken by the whole warp)
            const bool cold = want && R.nx_stage == 0;
            k = epa_claim(cursor, cold, lane);
            if (cold && k < count) {
                entry = __ldg(list + k);
                have = true;
            }
        }
        if (want && R.nx_stage != 0) {
            have = R.nx_stage != 3;
            k = R.nx_k;
            entry = R.nx_entry;
            if (R.nx_stage == 2) {
                a = R.nx_a;
                b = R.nx_b;
                resolved = true;
            }
        }
        if (want) {
            if (!have) {
                CTL(0) = TEAM_NO_ENTRY;
            } else {
                const uint32_t np = entry & ~OVER_RESUME;
                if (!resolved) {
                    if (A.pair_body) {
                        const uint2 bb = __ldg(reinterpret_cast<const uint2*>(A.pair_body) + np);
                        a = bb.x;
                        b = bb.y;
                    } else {
                        a = __ldg(A.l_shape + np);
                        b = __ldg(A.r_shape + np);
                    }
                }
                CTL(0) = entry;
                CTL(1) = a;
                CTL(2) = b;
                CTL(3) = k;
                if (!(RF > 0 && (entry & OVER_RESUME))) {
                    // Polytope::new (epa3d.rs:41-45, 129-135): the 4 simplex points; the faces
                    // (3,2,1) (3,1,0) (3,0,2) (2,0,1) of Face::new (:201-208) are built by the
                    // `new faces` phase from this edge list (apex 3; 2 for the fourth)
                    const float4* sx = C.W.simplex + (size_t)np * 8;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float4 q = sx[j];
                        PV(j) = make_float4(q.x, q.y, q.z, 0.f);
                        PA(j) = sx[4 + j];
                    }
                    EL(0) = (uint16_t)(2u | (1u << 8));
                    EL(1) = (uint16_t)(1u | (0u << 8));
                    EL(2) = (uint16_t)(0u | (2u << 8));
                    EL(3) = (uint16_t)(0u | (1u << 8));
                }
            }
        }
        {   // claim the entry after this one; its list word is consumed a round from now
            const bool more = want && have;
            const uint32_t k2 = epa_claim(cursor, more, lane);
            if (more) {
                R.nx_k = k2;
                if (k2 < count) {
                    R.nx_entry = __ldg(list + k2);
                    R.nx_stage = 1;
                } else {
                    R.nx_stage = 3;
                }
            }
        }
    }

    // ---- part 1: adopt the new pair; Face::new_impl (epa3d.rs:189-199) over this warp's share of the
    //      edge list; this warp's share of closest_face_to_origin (:137-145) ---------------------------
    static CB2_DI void part1(const epa_args& C, store& S, float4* __restrict__ pa, regs& R, const int w,
                             const unsigned lane) {
        const narrow_args& A = C.A;
        if (R.state == L_FETCH) {
            const uint32_t e = CTL(0);
            if (e == TEAM_NO_ENTRY) {
                R.state = L_EXIT;
                R.nf = R.len = R.ne = 0;
            } else {
                R.pair = e & ~OVER_RESUME;
                if (w < 2) {  // this warp's side of the pair
                    const uint32_t x = w == 0 ? CTL(1) : CTL(2);
                    const shape_table T = table_of(A.T);
                    if (A.pair_body) {
                        R.sh = load_shape(T, __ldg(A.l_shape + x));
                        R.xf = load_xform(A.l_t, x);
                    } else {
                        R.sh = load_shape(T, x);
                        R.xf = load_xform(w == 0 ? A.l_t : A.r_t, R.pair);
                    }
                }
                if (RF > 0 && (e & OVER_RESUME)) {
                    // resume a polytope that outgrew the previous tier, at the same iteration
                    const float4* rec = C.W.dump[C.tier - 1] + (size_t)CTL(3) * DLR::SIZE;
                    const float4 h = rec[0];
                    R.nf = (int)__float_as_uint(h.x);
                    R.nv = (int)__float_as_uint(h.y);
                    R.it = __float_as_uint(h.z);
                    const uint32_t* fi = reinterpret_cast<const uint32_t*>(rec + DLR::FIDX);
#pragma unroll 2
                    for (int j = w; j < R.nf; j += W) {
                        FND(j) = rec[DLR::FND + j];
                        FIDX(j) = fi[j];
                    }
#pragma unroll 2
                    for (int j = w; j < R.nv; j += W) {
                        const float4 q = rec[DLR::PV + j];
                        PV(j) = make_float4(q.x, q.y, q.z, 0.f);
                        PA(j) = rec[DLR::PA + j];
                    }
                    R.len = R.nf;
                    R.ne = 0;
                    R.init4 = false;
                } else {
                    R.ne = 4;
                    R.len = 0;
                    R.nv = 4;
                    R.it = 1;
                    R.apex = 3;
                    R.init4 = true;
                }
                R.state = L_RUN;
            }
        }
        float bd = INFINITY;
        uint32_t bi = TEAM_NO_FACE;
        if (R.state == L_RUN) {
            // surviving faces first, then the new ones: slots ascend, so `<` keeps the first minimum
#pragma unroll 4
            for (int f = w; f < R.len; f += W) {
                float dd = FND(f).w;
                if (dd != dd) dd = INFINITY;  // a NaN never compares smaller
                if (dd < bd) {
                    bd = dd;
                    bi = (uint32_t)f;
                }
            }
#pragma unroll 2
            for (int k = w; k < R.ne; k += W) {
                const uint32_t e = EL(k);
                const uint32_t a = (R.init4 && k == 3) ? 2u : R.apex, b = e & 255u, c = e >> 8;
                const float4 qa = PV(a), qb = PV(b), qc = PV(c);
                const f3 va = mk3(qa.x, qa.y, qa.z);
                const f3 ab = mk3(qb.x, qb.y, qb.z) - va;
                const f3 ac = mk3(qc.x, qc.y, qc.z) - va;
                const f3 n = normalize(cross(ab, ac));
                const float dist = dot(n, va);
                FND(R.len + k) = make_float4(n.x, n.y, n.z, dist);
                FIDX(R.len + k) = a | (b << 8) | (c << 16);
                const float dd = dist != dist ? INFINITY : dist;
                if (dd < bd) {
                    bd = dd;
                    bi = (uint32_t)(R.len + k);
                }
            }
            R.nf = R.len + R.ne;
            R.ne = 0;
            R.init4 = false;
        }
        PART(2 * w) = __float_as_uint(bd);
        PART(2 * w + 1) = bi;
    }

    // ---- part 2: the closest face; SupportPoint::from_minkowski(normal), minkowski/mod.rs:39-60 -----
    static CB2_DI void part2(const epa_args& C, store& S, float4* __restrict__ pa, regs& R, const int w,
                             const unsigned lane) {
        (void)pa;
        R.result = ST_NONE;
        R.live = false;
        if (R.state == L_RUN) {
            if (R.nf == 0) {
                R.result = ST_PANIC_EPA_EMPTY;  // faces[0] on an empty Vec (epa3d.rs:138)
            } else {
                float bd = INFINITY;
                uint32_t bi = TEAM_NO_FACE;
#pragma unroll
                for (int q = 0; q < W; ++q) {
                    const float ob = __uint_as_float(PART(2 * q));
                    const uint32_t oi = PART(2 * q + 1);
                    if (ob < bd || (ob == bd && oi < bi)) {
                        bd = ob;
                        bi = oi;
                    }
                }
                // the reference starts from faces[0] and replaces on strictly smaller: a NaN there sticks
                const float d0 = FND(0).w;
                R.best = (d0 != d0 || bi == TEAM_NO_FACE) ? 0 : (int)bi;
                R.fb = FND(R.best);
                R.live = true;
            }
        }
        if (w < 2 && R.live) {
            const shape_table T = table_of(C.A.T);
            const f3 n = mk3(R.fb.x, R.fb.y, R.fb.z);
            const f3 p = support_point(T, R.sh, R.xf, w == 0 ? n : -n);
            SUP(3 * w + 0) = p.x;
            SUP(3 * w + 1) = p.y;
            SUP(3 * w + 2) = p.z;
            CTL(2 + w) = R.sh.hang;
        }
    }

    // ---- part 3: convergence (epa3d.rs:46-64); this warp's share of the visibility tests -------------
    static CB2_DI void part3(const epa_args& C, store& S, float4* __restrict__ pa, regs& R, const int w,
                             const unsigned lane) {
        (void)pa;
        R.add = R.conv = false;
        uint32_t vlo = 0u, vhi = 0u;
        if (R.live) {
            const f3 l = mk3(SUP(0), SUP(1), SUP(2)), r = mk3(SUP(3), SUP(4), SUP(5));
            R.sv = l - r;
            R.sa = l;
            const f3 n = mk3(R.fb.x, R.fb.y, R.fb.z);
            if ((CTL(2) | CTL(3)) != 0u) {
                R.result = ST_NONCONVERGED;  // a HalfEdge hill climb met a NaN: the reference never returns
            } else if (fsub(dot(R.sv, n), R.fb.w) < C.A.settings.epa_tolerance ||
                       R.it >= C.A.settings.max_iterations) {
                R.conv = true;
                R.result = ST_SOME;
            } else {
                R.add = true;
            }
        }
        if (R.add) {
            // Polytope::add (epa3d.rs:147-175), step 1: which faces can see the new point
            const f3 sv = R.sv;
#pragma unroll 4
            for (int f = w; f < R.nf; f += W) {
                const float4 fn = FND(f);
                const float4 q = PV(FIDX(f) & 255u);
                const bool v = dot(mk3(fn.x, fn.y, fn.z), sv - mk3(q.x, q.y, q.z)) > 0.0f;
                if (v) {
                    if (f < 32) vlo |= 1u << f;
                    else vhi |= 1u << (f - 32);
                }
            }
        }
        PART(2 * w) = vlo;
        PART(2 * w + 1) = vhi;
    }

    // ---- part 4 (warp 0): the swap_remove order and the horizon ---------------------------------------
    static CB2_DI void part4(const epa_args& C, store& S, float4* __restrict__ pa, regs& R, const int w,
                             const unsigned lane) {
        (void)pa;
        if (w != 0 || !R.add) return;
        uint32_t vlo = 0u, vhi = 0u;
#pragma unroll
        for (int q = 0; q < W; ++q) {
            vlo |= PART(2 * q);
            vhi |= PART(2 * q + 1);
        }
        unsigned long long vis = (unsigned long long)vlo | ((unsigned long long)vhi << 32);
        // (2) the swap_remove sweep as a two-pointer walk over the mask: `i` climbs from the front,
        //     `len` shrinks from the back; a visible face at slot i is examined, then the tail face
        //     that swap_remove drops into slot i is examined next if it is visible, else recorded as
        //     a move.
        int len = R.nf, nvis = 0, nmv = 0, ne = 0;
        bool overflow = false;
        if (vis) {
            int i = __ffsll((long long)vis) - 1;
            int cur = i;
            for (;;) {
                if (nvis >= ECAP) {
                    overflow = true;
                    break;
                }
                EX(nvis) = (uint8_t)cur;
                ++nvis;
                vis &= ~(1ull << cur);
                --len;
                if (i == len) break;
                if ((vis >> len) & 1ull) {
                    cur = len;
                    continue;
                }
                MV(nmv) = (uint16_t)((uint32_t)i | ((uint32_t)len << 8));  // nmv < nvis <= ECAP
                ++nmv;
                const unsigned long long rest = vis & (~0ull << (i + 1));
                if (!rest) break;
                i = __ffsll((long long)rest) - 1;
                if (i >= len) break;
                cur = i;
            }
        }
        // (3) the horizon.  Set the directed edges of the examined faces in the adjacency rows ...
        bool dup = false;
        if (!overflow) {
            for (int k = 0; k < nvis; ++k) {
                const uint32_t idx = FIDX(EX(k));
                const uint32_t v0 = idx & 255u, v1 = (idx >> 8) & 255u, v2 = (idx >> 16) & 255u;
                const uint32_t r0 = ADJ(v0);
                dup |= ((r0 >> v1) & 1u) != 0u;
                ADJ(v0) = r0 | (1u << v1);
                const uint32_t r1 = ADJ(v1);
                dup |= ((r1 >> v2) & 1u) != 0u;
                ADJ(v1) = r1 | (1u << v2);
                const uint32_t r2 = ADJ(v2);
                dup |= ((r2 >> v0) & 1u) != 0u;
                ADJ(v2) = r2 | (1u << v0);
            }
#ifdef CB2_HOST_SIM
            if (g_sim_force_literal_horizon) dup = true;
#endif
            if (!dup) {
                // ... an edge stays iff its reverse is not an edge of an examined face
                for (int k = 0; k < nvis; ++k) {
                    const uint32_t idx = FIDX(EX(k));
                    const uint32_t v0 = idx & 255u, v1 = (idx >> 8) & 255u, v2 = (idx >> 16) & 255u;
                    const uint32_t r0 = ADJ(v0), r1 = ADJ(v1), r2 = ADJ(v2);
                    if (!((r1 >> v0) & 1u)) {
                        if (ne < ECAP) EL(ne) = (uint16_t)(v0 | (v1 << 8));
                        ++ne;
                    }
                    if (!((r2 >> v1) & 1u)) {
                        if (ne < ECAP) EL(ne) = (uint16_t)(v1 | (v2 << 8));
                        ++ne;
                    }
                    if (!((r0 >> v2) & 1u)) {
                        if (ne < ECAP) EL(ne) = (uint16_t)(v2 | (v0 << 8));
                        ++ne;
                    }
                }
                if (ne > ECAP) overflow = true;
            } else {
                // a directed edge occurs twice (non-manifold polytope): replay remove_or_add_edge
                // (epa3d.rs:212-219) literally — drop the first stored reverse edge, else push
                for (int k = 0; k < nvis && !overflow; ++k) {
                    const uint32_t idx = FIDX(EX(k));
                    const uint32_t v0 = idx & 255u, v1 = (idx >> 8) & 255u, v2 = (idx >> 16) & 255u;
#pragma unroll 1
                    for (int q = 0; q < 3; ++q) {
                        const uint32_t ea = q == 0 ? v0 : (q == 1 ? v1 : v2);
                        const uint32_t eb = q == 0 ? v1 : (q == 1 ? v2 : v0);
                        const uint32_t rev = eb | (ea << 8);
                        int m = 0;
                        while (m < ne && (uint32_t)EL(m) != rev) ++m;
                        if (m < ne) {
#pragma unroll 1
                            for (; m + 1 < ne; ++m) EL(m) = EL(m + 1);
                            --ne;
                        } else if (ne >= ECAP) {
                            overflow = true;
                        } else {
                            EL(ne) = (uint16_t)(ea | (eb << 8));
                            ++ne;
                        }
                    }
                }
            }
            // ... and clear the rows again
            for (int k = 0; k < nvis; ++k) {
                const uint32_t idx = FIDX(EX(k));
                ADJ(idx & 255u) = 0u;
                ADJ((idx >> 8) & 255u) = 0u;
                ADJ((idx >> 16) & 255u) = 0u;
            }
        }
        const bool spill = overflow || R.nv >= VCAP || len + ne > FCAP;
        uint32_t h = 0;
        if (spill) {
            len = ne = nmv = 0;
            if (C.tier + 1 < (uint32_t)EPA_TIERS) h = atomicAdd(&C.W.EC->n_list[C.tier + 1], 1u);
        }
        CTL(0) = (uint32_t)len | ((uint32_t)ne << 8) | ((uint32_t)nmv << 16) | (spill ? 1u << 24 : 0u);
        CTL(1) = h;
    }

    // ---- part 5: apply the moves, push the vertex; finished pairs write their contact or their dump --
    static CB2_DI void part5(const epa_args& C, store& S, float4* __restrict__ pa, regs& R, const int w,
                             const unsigned lane) {
        const narrow_args& A = C.A;
        uint32_t h = 0;
        if (R.add) {
            const uint32_t c = CTL(0);
            if (c >> 24) {
                // the polytope itself is untouched: the next tier redoes this round
                R.result = ST_SCRATCH_OVERFLOW;
                h = CTL(1);
            } else {
                R.len = (int)(c & 255u);
                R.ne = (int)((c >> 8) & 255u);
                const int nmv = (int)((c >> 16) & 255u);
                // sources lie behind every destination, so the moves are independent
                for (int m = w; m < nmv; m += W) {
                    const uint32_t mvw = MV(m);
                    const int dst = (int)(mvw & 255u), src = (int)(mvw >> 8);
                    FND(dst) = FND(src);
                    FIDX(dst) = FIDX(src);
                }
                if (w == W - 1) {
                    PV(R.nv) = make_float4(R.sv.x, R.sv.y, R.sv.z, 0.f);
                    PA(R.nv) = make_float4(R.sa.x, R.sa.y, R.sa.z, 0.f);
                }
                R.apex = (uint32_t)R.nv;
                ++R.nv;
                ++R.it;
            }
        }
        if (R.result != ST_NONE) {
            const bool next_tier = C.tier + 1 < (uint32_t)EPA_TIERS;
            if (R.result == ST_SCRATCH_OVERFLOW && next_tier) {
                // re-queue for the next tier with the polytope as it stands
                const bool room = h < C.W.dump_cap[C.tier];
                if (room) {
                    float4* rec = C.W.dump[C.tier] + (size_t)h * DL::SIZE;
                    uint32_t* fi = reinterpret_cast<uint32_t*>(rec + DL::FIDX);
                    if (w == 0)
                        rec[0] = make_float4(__uint_as_float((uint32_t)R.nf), __uint_as_float((uint32_t)R.nv),
                                             __uint_as_float(R.it), 0.f);
#pragma unroll 2
                    for (int j = w; j < R.nf; j += W) {
                        rec[DL::FND + j] = FND(j);
                        fi[j] = FIDX(j);
                    }
#pragma unroll 2
                    for (int j = w; j < R.nv; j += W) {
                        rec[DL::PV + j] = PV(j);
                        rec[DL::PA + j] = PA(j);
                    }
                }
                if (w == 0) C.W.list[C.tier + 1][h] = R.pair | (room ? OVER_RESUME : 0u);
            } else if (w == W - 1) {
                contact_out co = zero_contact();
                if (R.conv) {
                    // contact(), epa3d.rs:84-114
                    const f3 n = mk3(R.fb.x, R.fb.y, R.fb.z);
                    const uint32_t idx = FIDX(R.best);
                    const uint32_t i0 = idx & 255u, i1 = (idx >> 8) & 255u, i2 = (idx >> 16) & 255u;
                    const float4 q0 = PV(i0), q1 = PV(i1), q2 = PV(i2);
                    float u, v, ww;
                    barycentric_vector(n * R.fb.w, mk3(q0.x, q0.y, q0.z), mk3(q1.x, q1.y, q1.z),
                                       mk3(q2.x, q2.y, q2.z), u, v, ww);
                    const float4 a0 = PA(i0), a1 = PA(i1), a2 = PA(i2);
                    co.normal = n;
                    co.depth = R.fb.w;
                    co.point = (mk3(a0.x, a0.y, a0.z) * u + mk3(a1.x, a1.y, a1.z) * v) +
                               mk3(a2.x, a2.y, a2.z) * ww;
                }
                store_contact(A.out_contact, R.pair, CB2_FULL_RESOLUTION, co);
                A.out_status[R.pair] = R.result;
            }
            R.state = L_FETCH;
            R.nf = R.len = R.ne = 0;
        }
    }
#undef FND
#undef PV
#undef ADJ
#undef FIDX
#undef EL
#undef MV
#undef EX
#undef SUP
#undef PART
#undef CTL
#undef PA
};

}  // namespace cb2
