// cb2_epa_lane.cuh — lane-per-pair EPA3 (EPA3::process, epa3d.rs:27-65): the first EPA tier of
// GJK3::intersection(FullResolution) (gjk/mod.rs:362-391).
//
// Round 1 gave every pair a group of 4 lanes (cb2_epa.cu).  ncu showed that kernel issue-bound
// with 20.8 of 32 lanes active and ~205 warp instructions per pair-iteration: the two support
// calls of a Minkowski point kept only 2 of 4 lanes busy, every phase ran to the warp maximum of
// eight groups, and the parallel horizon needed an adjacency matrix, three passes and six warp
// barriers per iteration.  Here ONE lane owns a pair and the polytope of the 32 pairs of a warp
// lives in shared memory as [index][lane]:
//  * every shared-memory access is bank-conflict free whatever index a lane uses (the lane picks
//    the bank, the index picks the row), so the serial data-dependent parts of Polytope::add
//    (swap_remove, remove_or_add_edge) are plain per-lane code — no ballots, shuffles or barriers;
//  * both support calls of a pair are useful work in the same lane;
//  * remove_or_add_edge (epa3d.rs:212-219) is replayed LITERALLY on a per-lane edge list, so
//    non-manifold polytopes need no second chance;
//  * the loop is one flat state machine per lane (fetch -> new faces -> closest face -> support ->
//    visibility -> horizon walk) with a full-warp vote per round: lanes refill from the device
//    queue the moment their pair converges, and a refilled lane builds its first four faces in
//    the same `new faces` phase the other lanes use for their horizon faces.
//  * visibility is evaluated once per face into a 64-bit register mask; the swap_remove sweep
//    then becomes a two-pointer walk over that mask which touches only the visible faces.
// A pair that outgrows the tier's caps leaves its polytope UNTOUCHED (the walk only writes the
// edge / move lists), dumps it in cb2_epa.cu's record format and the next tier resumes it at the
// same iteration.
//
// The body compiles for the host as well (cb2_hostsim.h): tests/host_sim runs it one lane at a
// time against the oracle.
#pragma once
#include "cb2_pair.cuh"

namespace cb2 {

struct epa_args {
    narrow_args A;
    narrow_ws W;
    uint32_t tier;      // 0 .. EPA_TIERS-1
    uint32_t sm_count;  // SMs of the device (the later tiers size their active grid by it)
    float4* supply;     // tier 0 of the group kernel: 32 prepared hits per warp (behind the sup_a scratch)
};

// polytope dump record, in float4 units: header (nf, nv, it) | fnd | pv | pa | fidx
constexpr uint32_t OVER_RESUME = 0x80000000u;  // over-list entry flag: a dump record exists
template <int F, int V>
struct dump_layout {
    static constexpr int FND = 1, PV = FND + F, PA = PV + V, FIDX = PA + V, SIZE = FIDX + (F + 3) / 4;
};

// the polytopes of one warp: element i of lane l at [i * 32 + l]
template <int FCAP, int VCAP, int ECAP>
struct lane_store {
    float4 fnd[FCAP * 32];     // normal.xyz, distance
    float4 pv[VCAP * 32];      // SupportPoint.v
    uint32_t fidx[FCAP * 32];  // v0 | v1<<8 | v2<<16
    uint16_t el[ECAP * 32];    // horizon edges a | b<<8, in the reference's Vec order
    uint16_t mv[ECAP * 32];    // swap_remove moves: dst | src<<8
};

// next entry of the tier's work list for every lane that wants one (one atomic per warp)
CB2_DI uint32_t epa_claim(uint32_t* cursor, bool want, unsigned lane) {
#ifdef CB2_HOST_SIM
    (void)lane;
    return want ? atomicAdd(cursor, 1u) : 0u;
#else
    const unsigned need = __ballot_sync(0xffffffffu, want);
    if (!need) return 0u;
    const int leader = __ffs((int)need) - 1;
    uint32_t base = 0;
    if ((int)lane == leader) base = atomicAdd(cursor, (uint32_t)__popc(need));
    base = __shfl_sync(0xffffffffu, base, leader);
    return base + (uint32_t)__popc(need & ((1u << lane) - 1u));
#endif
}

enum : int { L_FETCH = 0, L_RUN = 1, L_EXIT = 2 };

CB2_DI void prefetch_l2(const void* p) {
#ifndef CB2_HOST_SIM
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
#else
    (void)p;
#endif
}

// RF/RV: caps of the tier whose dumps this instantiation resumes (0 = none).
// PF: every lane claims its NEXT work-list entry when it starts a pair and resolves the entry's
// dependent loads (list -> shape indices) one round at a time while the current pair runs, so a
// refill costs one memory round trip (shapes, transforms, simplex) instead of three back to back —
// with one warp per block and 5-8 blocks per SM there are few other warps to hide them behind.
template <int FCAP, int VCAP, int ECAP, int RF, int RV, bool PF>
CB2_DI void epa_lane_run(const epa_args& C, lane_store<FCAP, VCAP, ECAP>& S, float4* __restrict__ pa_warp,
                         const unsigned lane) {
    static_assert(FCAP <= 64, "the visibility mask is one 64-bit word");
    static_assert(VCAP <= 255 && ECAP <= 255, "indices are bytes");
    static_assert(RF <= FCAP && RV <= VCAP, "a resumed polytope must fit");
    const narrow_args& A = C.A;
    float4* const fnd = S.fnd + lane;
    float4* const pv = S.pv + lane;
    uint32_t* const fidx = S.fidx + lane;
    uint16_t* const el = S.el + lane;
    uint16_t* const mv = S.mv + lane;
    float4* const pa = pa_warp + lane;
#define FND(i) fnd[(i) * 32]
#define PV(i) pv[(i) * 32]
#define FIDX(i) fidx[(i) * 32]
#define EL(i) el[(i) * 32]
#define MV(i) mv[(i) * 32]
#define PA(i) pa[(i) * 32]

    const uint32_t* const list = C.W.list[C.tier];
    const uint32_t count = C.W.EC->n_list[C.tier];
    uint32_t* over_list = nullptr;
    uint32_t* over_count = nullptr;
    if (C.tier + 1 < (uint32_t)EPA_TIERS) {
        over_list = C.W.list[C.tier + 1];
        over_count = &C.W.EC->n_list[C.tier + 1];
    }
    uint32_t* const cursor = &C.W.EC->next[C.tier];
    const uint32_t max_it = A.settings.max_iterations;
    const float epa_tol = A.settings.epa_tolerance;

    pair_in p;
    p.T = table_of(A.T);
    p.ls.kind = K_CUBOID; p.ls.h = zero3(); p.ls.hang = 0u;
    p.ls.poly.voff = 0; p.ls.poly.vcnt = 0; p.ls.poly.cell_off = 0; p.ls.poly.cfg = 0;
    p.rs = p.ls;
    p.lt = make_xform(make_float4(1.f, 1.f, 0.f, 0.f), make_float4(0.f, 0.f, 0.f, 0.f));
    p.rt = p.lt;

    int state = L_FETCH;
    uint32_t pair = 0, it = 0, apex = 0;
    int nf = 0, nv = 0, len = 0, ne = 0;
    bool init4 = false;
    // the claimed next entry: nx_stage 0 none, 1 list entry in flight, 2 indices in flight, 3 queue empty
    int nx_stage = 0;
    uint32_t nx_k = 0, nx_entry = 0, nx_a = 0, nx_b = 0;

    for (;;) {
        // ================= fetch (only the lanes whose pair finished) ====================================
        if (PF && nx_stage == 1 && state == L_RUN) {
            // second hop of the claimed entry (its first hop was issued a round ago)
            const uint32_t np = nx_entry & ~OVER_RESUME;
            if (A.pair_body) {
                const uint2 b = __ldg(reinterpret_cast<const uint2*>(A.pair_body) + np);
                nx_a = b.x;
                nx_b = b.y;
            } else {
                nx_a = __ldg(A.l_shape + np);
                nx_b = __ldg(A.r_shape + np);
                prefetch_l2(A.l_t + 2 * (size_t)np);
                prefetch_l2(A.r_t + 2 * (size_t)np);
            }
            if (!(RF > 0 && (nx_entry & OVER_RESUME))) prefetch_l2(C.W.simplex + (size_t)np * 8);
            nx_stage = 2;
        }
        {
            const bool want = state == L_FETCH;
            uint32_t k = 0, entry = 0;
            bool have = false;
            {  // (with PF: only the first pair of a lane; the vote inside is taken by every lane)
                const bool cold = want && (!PF || nx_stage == 0);
                k = epa_claim(cursor, cold, lane);
                if (cold && k < count) {
                    entry = __ldg(list + k);
                    have = true;
                }
            }
            if (PF && want && nx_stage != 0) {
                have = nx_stage != 3;
                k = nx_k;
                entry = nx_entry;
            }
            if (want) {
                if (!have) {
                    state = L_EXIT;
                    nf = len = ne = 0;
                } else {
                    pair = entry & ~OVER_RESUME;
                    uint32_t li, ri;
                    uint64_t lx, rx;
                    if (A.pair_body) {
                        if (PF && nx_stage == 2) {
                            lx = nx_a;
                            rx = nx_b;
                        } else {
                            const uint2 b = __ldg(reinterpret_cast<const uint2*>(A.pair_body) + pair);
                            lx = b.x;
                            rx = b.y;
                        }
                        li = __ldg(A.l_shape + lx);
                        ri = __ldg(A.l_shape + rx);
                    } else {
                        lx = rx = pair;
                        if (PF && nx_stage == 2) {
                            li = nx_a;
                            ri = nx_b;
                        } else {
                            li = __ldg(A.l_shape + pair);
                            ri = __ldg(A.r_shape + pair);
                        }
                    }
                    p.ls = load_shape(p.T, li);
                    p.rs = load_shape(p.T, ri);
                    p.lt = load_xform(A.l_t, lx);
                    p.rt = load_xform(A.r_t, rx);
                    if (RF > 0 && (entry & OVER_RESUME)) {
                        // resume a polytope that outgrew the previous tier, at the same iteration
                        using DL = dump_layout<(RF > 0 ? RF : 1), (RV > 0 ? RV : 1)>;
                        const float4* rec = C.W.dump[C.tier - 1] + (size_t)k * DL::SIZE;
                        const float4 h = rec[0];
                        nf = (int)__float_as_uint(h.x);
                        nv = (int)__float_as_uint(h.y);
                        it = __float_as_uint(h.z);
                        const uint32_t* fi = reinterpret_cast<const uint32_t*>(rec + DL::FIDX);
#pragma unroll 1
                        for (int j = 0; j < nf; ++j) {
                            FND(j) = rec[DL::FND + j];
                            FIDX(j) = fi[j];
                        }
#pragma unroll 1
                        for (int j = 0; j < nv; ++j) {
                            PV(j) = rec[DL::PV + j];
                            PA(j) = rec[DL::PA + j];
                        }
                        len = nf;
                        ne = 0;
                        init4 = false;
                    } else {
                        // Polytope::new (epa3d.rs:41-45, 129-135): the 4 simplex points; the faces
                        // (3,2,1) (3,1,0) (3,0,2) (2,0,1) of Face::new (:201-208) are built by the
                        // `new faces` phase below from this edge list (apex 3; 2 for the fourth)
                        const float4* sx = C.W.simplex + (size_t)pair * 8;
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            PV(j) = sx[j];
                            PA(j) = sx[4 + j];
                        }
                        EL(0) = (uint16_t)(2u | (1u << 8));
                        EL(1) = (uint16_t)(1u | (0u << 8));
                        EL(2) = (uint16_t)(0u | (2u << 8));
                        EL(3) = (uint16_t)(0u | (1u << 8));
                        ne = 4;
                        len = 0;
                        nv = 4;
                        it = 1;
                        apex = 3;
                        init4 = true;
                    }
                    state = L_RUN;
                }
            }
            if (PF) {
                // claim the entry after this one; its list word is consumed a round from now
                const bool more = want && have;
                const uint32_t k2 = epa_claim(cursor, more, lane);
                if (more) {
                    nx_k = k2;
                    if (k2 < count) {
                        nx_entry = __ldg(list + k2);
                        nx_stage = 1;
                    } else {
                        nx_stage = 3;
                    }
                }
            }
        }
        if (__all_sync(0xffffffffu, state == L_EXIT)) break;

        // ================= new faces: Face::new_impl (epa3d.rs:189-199) over the edge list ==============
        // (the horizon of the previous round's Polytope::add, or the four faces of a new pair)
        {
#pragma unroll 2
            for (int k = 0; k < ne; ++k) {
                const uint32_t e = EL(k);
                const uint32_t a = (init4 && k == 3) ? 2u : apex, b = e & 255u, c = e >> 8;
                const float4 qa = PV(a), qb = PV(b), qc = PV(c);
                const f3 va = mk3(qa.x, qa.y, qa.z);
                const f3 ab = mk3(qb.x, qb.y, qb.z) - va;
                const f3 ac = mk3(qc.x, qc.y, qc.z) - va;
                const f3 n = normalize(cross(ab, ac));
                const float dist = dot(n, va);
                FND(len + k) = make_float4(n.x, n.y, n.z, dist);
                FIDX(len + k) = a | (b << 8) | (c << 16);
            }
            nf = len + ne;
            ne = 0;
            init4 = false;
        }

        const bool run = state == L_RUN;
        uint8_t result = ST_NONE;
        // ================= closest_face_to_origin (epa3d.rs:137-145) ====================================
        // literal: start from faces[0], replace on strictly smaller (a NaN at index 0 sticks)
        int best = 0;
        float4 fb = make_float4(0.f, 0.f, 0.f, 0.f);
        if (run) {
            if (nf == 0) {
                result = ST_PANIC_EPA_EMPTY;  // faces[0] on an empty Vec (epa3d.rs:138)
            } else {
                float bd = FND(0).w;
#pragma unroll 4
                for (int f = 1; f < nf; ++f) {
                    const float dd = FND(f).w;
                    if (dd < bd) {
                        bd = dd;
                        best = f;
                    }
                }
                fb = FND(best);
            }
        }
        const bool live = run && result == ST_NONE;
        const f3 n = mk3(fb.x, fb.y, fb.z);

        // ================= SupportPoint::from_minkowski(n), minkowski/mod.rs:39-60 ======================
        f3 sv = zero3(), sa = zero3();
        if (live) minkowski<true>(p, n, sv, sa);
        bool add = false, conv = false;
        if (live) {
            if ((p.ls.hang | p.rs.hang) != 0u) {
                result = ST_NONCONVERGED;  // a HalfEdge hill climb met a NaN: the reference never returns
            } else if (fsub(dot(sv, n), fb.w) < epa_tol || it >= max_it) {
                conv = true;  // epa3d.rs:46-64
                result = ST_SOME;
            } else {
                add = true;
            }
        }

        // ================= Polytope::add (epa3d.rs:147-175) =============================================
        if (add) {
            // (1) visibility of every face, once
            uint32_t vlo = 0u, vhi = 0u;
            {
                const int n0 = nf < 32 ? nf : 32;
#pragma unroll 4
                for (int f = 0; f < n0; ++f) {
                    const float4 fn = FND(f);
                    const float4 q = PV(FIDX(f) & 255u);
                    if (dot(mk3(fn.x, fn.y, fn.z), sv - mk3(q.x, q.y, q.z)) > 0.0f) vlo |= 1u << f;
                }
#pragma unroll 4
                for (int f = 32; f < nf; ++f) {
                    const float4 fn = FND(f);
                    const float4 q = PV(FIDX(f) & 255u);
                    if (dot(mk3(fn.x, fn.y, fn.z), sv - mk3(q.x, q.y, q.z)) > 0.0f) vhi |= 1u << (f - 32);
                }
            }
            // (2) the swap_remove sweep as a two-pointer walk over the mask: `i` climbs from the
            //     front, `len` shrinks from the back; a visible face at slot i is examined, then the
            //     tail face that swap_remove drops into slot i is examined next if it is visible,
            //     else recorded as a move.  Examined faces toggle their three edges in the list.
            unsigned long long vis = (unsigned long long)vlo | ((unsigned long long)vhi << 32);
            int nmv = 0;
            bool overflow = false;
            len = nf;
            ne = 0;
            if (vis) {
                int i = __ffsll((long long)vis) - 1;
                int cur = i;
                for (;;) {
                    const uint32_t idx = FIDX(cur);
                    const uint32_t v0 = idx & 255u, v1 = (idx >> 8) & 255u, v2 = (idx >> 16) & 255u;
#pragma unroll
                    for (int q = 0; q < 3; ++q) {
                        // remove_or_add_edge (epa3d.rs:212-219): drop the first stored reverse
                        // edge, else push
                        const uint32_t ea = q == 0 ? v0 : (q == 1 ? v1 : v2);
                        const uint32_t eb = q == 0 ? v1 : (q == 1 ? v2 : v0);
                        const uint32_t rev = eb | (ea << 8);
                        int k = 0;
                        while (k < ne && (uint32_t)EL(k) != rev) ++k;
                        if (k < ne) {
#pragma unroll 1
                            for (int m = k; m + 1 < ne; ++m) EL(m) = EL(m + 1);
                            --ne;
                        } else if (ne >= ECAP) {
                            overflow = true;
                        } else {
                            EL(ne) = (uint16_t)(ea | (eb << 8));
                            ++ne;
                        }
                    }
                    if (overflow) break;
                    vis &= ~(1ull << cur);
                    --len;
                    if (i == len) break;
                    if ((vis >> len) & 1ull) {
                        cur = len;
                        continue;
                    }
                    if (nmv >= ECAP) {
                        overflow = true;
                        break;
                    }
                    MV(nmv) = (uint16_t)((uint32_t)i | ((uint32_t)len << 8));
                    ++nmv;
                    const unsigned long long rest = vis & (~0ull << (i + 1));
                    if (!rest) break;
                    i = __ffsll((long long)rest) - 1;
                    if (i >= len) break;
                    cur = i;
                }
            }
            if (overflow || nv >= VCAP || len + ne > FCAP) {
                // the polytope itself is untouched: the next tier redoes this round
                result = ST_SCRATCH_OVERFLOW;
            } else {
                // (3) apply the moves (sources lie behind every destination), push the vertex
                for (int m = 0; m < nmv; ++m) {
                    const uint32_t w = MV(m);
                    const int dst = (int)(w & 255u), src = (int)(w >> 8);
                    FND(dst) = FND(src);
                    FIDX(dst) = FIDX(src);
                }
                PV(nv) = make_float4(sv.x, sv.y, sv.z, 0.f);
                PA(nv) = make_float4(sa.x, sa.y, sa.z, 0.f);
                apex = (uint32_t)nv;
                ++nv;
                ++it;
            }
        }

        // ================= finished pairs ===============================================================
        if (result != ST_NONE) {
            if (result == ST_SCRATCH_OVERFLOW && over_list) {
                // re-queue for the next tier with the polytope as it stands
                using DL = dump_layout<FCAP, VCAP>;
                const uint32_t h = atomicAdd(over_count, 1u);
                const bool room = h < C.W.dump_cap[C.tier];
                if (room) {
                    float4* rec = C.W.dump[C.tier] + (size_t)h * DL::SIZE;
                    rec[0] = make_float4(__uint_as_float((uint32_t)nf), __uint_as_float((uint32_t)nv),
                                         __uint_as_float(it), 0.f);
                    uint32_t* fi = reinterpret_cast<uint32_t*>(rec + DL::FIDX);
#pragma unroll 4
                    for (int j = 0; j < nf; ++j) {
                        rec[DL::FND + j] = FND(j);
                        fi[j] = FIDX(j);
                    }
#pragma unroll 8
                    for (int j = 0; j < nv; ++j) {
                        rec[DL::PV + j] = PV(j);
                        rec[DL::PA + j] = PA(j);
                    }
                }
                over_list[h] = pair | (room ? OVER_RESUME : 0u);
            } else {
                contact_out co = zero_contact();
                if (conv) {
                    // contact(), epa3d.rs:84-114
                    const uint32_t idx = FIDX(best);
                    const uint32_t i0 = idx & 255u, i1 = (idx >> 8) & 255u, i2 = (idx >> 16) & 255u;
                    const float4 q0 = PV(i0), q1 = PV(i1), q2 = PV(i2);
                    float u, v, w;
                    barycentric_vector(n * fb.w, mk3(q0.x, q0.y, q0.z), mk3(q1.x, q1.y, q1.z),
                                       mk3(q2.x, q2.y, q2.z), u, v, w);
                    const float4 a0 = PA(i0), a1 = PA(i1), a2 = PA(i2);
                    co.normal = n;
                    co.depth = fb.w;
                    co.point = (mk3(a0.x, a0.y, a0.z) * u + mk3(a1.x, a1.y, a1.z) * v) +
                               mk3(a2.x, a2.y, a2.z) * w;
                }
                store_contact(A.out_contact, pair, CB2_FULL_RESOLUTION, co);
                A.out_status[pair] = result;
            }
            state = L_FETCH;
            nf = len = ne = 0;
        }
    }
#undef FND
#undef PV
#undef FIDX
#undef EL
#undef MV
#undef PA
}

}  // namespace cb2
