// cb2_poly.cu — sub-warp cooperative GJK3+EPA3 (GJK3::intersection, gjk/mod.rs:362-391) for pairs
// that involve a ConvexPolyhedron.
//
// Design (B200-first, not a translation of the reference's per-pair loop):
//  * A GROUP of G lanes (4/8/16/32, chosen per size class) owns one pair.  Everything scalar
//    (transforms, simplex logic) is computed redundantly by the G lanes — it is uniform inside
//    the group, so it costs one instruction stream — while the O(V) work is split across them:
//    the support_point vertex scan (polyhedron.rs:160-176), EPA's closest-face search,
//    visibility tests and new-face construction (epa3d.rs:137-175, 189-199).
//  * Both polyhedra of the pair are staged ONCE into shared memory (12-byte packed xyz) and
//    re-scanned ~20 times from there; lanes read consecutive vertices (bank-conflict free) and
//    the first-strict-max rule of the reference is kept by a (dot desc, index asc) butterfly.
//  * The kernel is a persistent state machine whose loop body is exactly ONE Minkowski support
//    evaluation (the dominant, state-independent cost) followed by a short state-specific step
//    (GJK seed / GJK iteration / EPA iteration).  Groups whose pair finished fetch the next pair
//    from a global atomic queue, so lanes never idle waiting for the slowest pair of a warp and
//    all groups of a warp execute the support scans convergently whatever state they are in.
//  * Pairs are bucketed by total vertex count first (k_classify/k_scatter), so each launch has
//    a shared-memory footprint and group width that fit its class.
//  * The EPA polytope lives in per-group shared memory, spilling to a per-group global scratch
//    only beyond 48 faces / 28 vertices / 32 horizon edges (rare).
#include "cb2_gjk.cuh"
#include "cb2_kernels.h"

namespace cb2 {

// ---- size classes ----------------------------------------------------------------------------------
// class 0: no polyhedron in the pair (thread-per-pair kernel); 1..5: total vertices <= 64, 128,
// 256, 512, 1024 (staged in shared memory); 6: larger (scanned straight from global/L2).
CB2_DI uint32_t class_of(uint32_t vt, bool any_poly) {
    if (!any_poly) return 0;
    if (vt <= 64) return 1;
    if (vt <= 128) return 2;
    if (vt <= 256) return 3;
    if (vt <= 512) return 4;
    if (vt <= 1024) return 5;
    return 6;
}

CB2_DI void pair_shape_indices(const narrow_args& A, uint64_t i, uint32_t& li, uint32_t& ri,
                               uint64_t& lx, uint64_t& rx) {
    if (A.pair_body) {
        const uint2 b = __ldg(reinterpret_cast<const uint2*>(A.pair_body) + i);
        lx = b.x;
        rx = b.y;
        li = __ldg(A.l_shape + lx);
        ri = __ldg(A.l_shape + rx);
    } else {
        lx = rx = i;
        li = __ldg(A.l_shape + i);
        ri = __ldg(A.r_shape + i);
    }
}

__global__ void __launch_bounds__(256) k_classify(narrow_args A, uint8_t* __restrict__ cls,
                                                 uint32_t* __restrict__ counts) {
    __shared__ uint32_t h[NUM_CLASSES];
    if (threadIdx.x < NUM_CLASSES) h[threadIdx.x] = 0;
    __syncthreads();
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < A.n) {
        uint32_t li, ri;
        uint64_t lx, rx;
        pair_shape_indices(A, i, li, ri, lx, rx);
        const dshape* sl = A.shapes + li;
        const dshape* sr = A.shapes + ri;
        const uint32_t kl = __ldg(&sl->kind), kr = __ldg(&sr->kind);
        const uint32_t vl = kl == K_POLY ? __ldg(&sl->vcnt) : 0u;
        const uint32_t vr = kr == K_POLY ? __ldg(&sr->vcnt) : 0u;
        const uint32_t c = class_of(vl + vr, kl == K_POLY || kr == K_POLY);
        cls[i] = (uint8_t)c;
        atomicAdd(&h[c], 1u);
    }
    __syncthreads();
    if (threadIdx.x < NUM_CLASSES && h[threadIdx.x]) atomicAdd(&counts[threadIdx.x], h[threadIdx.x]);
}

// perm is grouped by class: class c occupies [sum(counts[<c]), +counts[c]); order inside a class is
// arbitrary (results are written back by original pair index, so outputs are deterministic).
__global__ void __launch_bounds__(256) k_scatter(uint64_t n, const uint8_t* __restrict__ cls,
                                                const uint32_t* __restrict__ counts,
                                                uint32_t* __restrict__ cursor,
                                                uint32_t* __restrict__ perm) {
    __shared__ uint32_t h[NUM_CLASSES], base[NUM_CLASSES];
    if (threadIdx.x < NUM_CLASSES) h[threadIdx.x] = 0;
    __syncthreads();
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t c = 0, local = 0;
    if (i < n) {
        c = cls[i];
        local = atomicAdd(&h[c], 1u);
    }
    __syncthreads();
    if (threadIdx.x < NUM_CLASSES) {
        uint32_t off = 0;
        for (uint32_t k = 0; k < threadIdx.x; ++k) off += counts[k];
        base[threadIdx.x] = off + (h[threadIdx.x] ? atomicAdd(&cursor[threadIdx.x], h[threadIdx.x]) : 0u);
    }
    __syncthreads();
    if (i < n) perm[base[c] + local] = (uint32_t)i;
}

// ---- cooperative kernel ------------------------------------------------------------------------------
constexpr int PF_CAP = 48;   // faces kept in shared memory per group
constexpr int PV_CAP = 28;   // polytope vertices kept in shared memory per group
constexpr int PE_CAP = 32;   // horizon edges kept in shared memory per group
constexpr int GF_CAP = 256;  // hard caps (shared + global spill), as in cb2_narrow.cu
constexpr int GV_CAP = 104;
constexpr int GE_CAP = 192;
// global spill words per group: faces (5 words) + verts (6 words) + edges (u16 -> 1/2 word)
constexpr int SPILL_WORDS = (GF_CAP - PF_CAP) * 5 + (GV_CAP - PV_CAP) * 6 + (GE_CAP - PE_CAP) / 2;
// shared words per group excluding staged vertices: faces 5, verts 6, edges 1/2, vis mask 8
constexpr int POLY_WORDS = PF_CAP * 5 + PV_CAP * 6 + PE_CAP / 2 + GF_CAP / 32;

enum : int { S_FETCH = 0, S_SEED = 1, S_GJK = 2, S_EPA = 3, S_EXIT = 4 };

struct poly_store {
    float* s_fnd;       // PF_CAP x 4
    uint32_t* s_fidx;   // PF_CAP
    float* s_pv;        // PV_CAP x 3
    float* s_pa;        // PV_CAP x 3
    uint16_t* s_edge;   // PE_CAP
    uint32_t* s_vis;    // GF_CAP / 32
    float* g_fnd;       // spill
    uint32_t* g_fidx;
    float* g_pv;
    float* g_pa;
    uint16_t* g_edge;

    CB2_DI float4* fnd(int i) const {
        return reinterpret_cast<float4*>(i < PF_CAP ? s_fnd + 4 * i : g_fnd + 4 * (i - PF_CAP));
    }
    CB2_DI uint32_t* fidx(int i) const { return i < PF_CAP ? s_fidx + i : g_fidx + (i - PF_CAP); }
    CB2_DI float* pv(int i) const { return i < PV_CAP ? s_pv + 3 * i : g_pv + 3 * (i - PV_CAP); }
    CB2_DI float* pa(int i) const { return i < PV_CAP ? s_pa + 3 * i : g_pa + 3 * (i - PV_CAP); }
    CB2_DI uint16_t* edge(int i) const { return i < PE_CAP ? s_edge + i : g_edge + (i - PE_CAP); }
    CB2_DI f3 get_pv(int i) const {
        const float* p = pv(i);
        return mk3(p[0], p[1], p[2]);
    }
    CB2_DI f3 get_pa(int i) const {
        const float* p = pa(i);
        return mk3(p[0], p[1], p[2]);
    }
};

// Face::new_impl (epa3d.rs:189-199) into slot
CB2_DI void coop_face_new(const poly_store& P, int slot, uint32_t a, uint32_t b, uint32_t c) {
    const f3 va = P.get_pv(a);
    const f3 ab = P.get_pv(b) - va;
    const f3 ac = P.get_pv(c) - va;
    const f3 n = normalize(cross(ab, ac));
    const float dist = dot(n, va);
    *P.fnd(slot) = make_float4(n.x, n.y, n.z, dist);
    *P.fidx(slot) = a | (b << 8) | (c << 16);
}

struct coop_args {
    narrow_args A;
    const uint32_t* perm;      // pairs of this class: perm[begin .. begin+count)
    const uint32_t* counts;    // per-class counts (device)
    uint32_t* next;            // per-class work-queue cursors (device)
    uint32_t cls;
    uint32_t vcap;             // staged vertices per group (L + R), 0 => scan from global
    float* spill;              // SPILL_WORDS per group, global
};

template <int G>
CB2_DI void group_argmax(float& best, uint32_t& bi, unsigned mask) {
#pragma unroll
    for (int off = G / 2; off >= 1; off >>= 1) {
        const float ob = __shfl_xor_sync(mask, best, off);
        const uint32_t oi = __shfl_xor_sync(mask, bi, off);
        if (ob > best || (ob == best && oi < bi)) {
            best = ob;
            bi = oi;
        }
    }
}

template <int G, bool FULL>
__global__ void __launch_bounds__(32) k_coop_intersection(coop_args C) {
    extern __shared__ __align__(16) float smem[];
    constexpr int NG = 32 / G;  // groups per warp (block = one warp)
    const int lane = threadIdx.x;
    const int gl = lane & (G - 1);
    const int gw = lane / G;
    const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (gw * G));
    const narrow_args& A = C.A;

    // ---- shared memory carve-up: [staged vertices (vcap*3 floats)] [polytope] per group ----------
    const uint32_t vwords = (C.vcap * 3 + 3) & ~3u;
    const uint32_t gwords = vwords + POLY_WORDS;
    float* gs = smem + (size_t)gw * gwords;
    float* s_verts = gs;
    poly_store P;
    {
        float* p = gs + vwords;
        P.s_fnd = p;                                   p += PF_CAP * 4;
        P.s_fidx = reinterpret_cast<uint32_t*>(p);     p += PF_CAP;
        P.s_pv = p;                                    p += PV_CAP * 3;
        P.s_pa = p;                                    p += PV_CAP * 3;
        P.s_edge = reinterpret_cast<uint16_t*>(p);     p += PE_CAP / 2;
        P.s_vis = reinterpret_cast<uint32_t*>(p);
        const size_t group_global = (size_t)blockIdx.x * NG + gw;
        float* g = C.spill + group_global * SPILL_WORDS;
        P.g_fnd = g;                                   g += (GF_CAP - PF_CAP) * 4;
        P.g_fidx = reinterpret_cast<uint32_t*>(g);     g += (GF_CAP - PF_CAP);
        P.g_pv = g;                                    g += (GV_CAP - PV_CAP) * 3;
        P.g_pa = g;                                    g += (GV_CAP - PV_CAP) * 3;
        P.g_edge = reinterpret_cast<uint16_t*>(g);
    }

    uint32_t begin = 0;
    for (uint32_t k = 0; k < C.cls; ++k) begin += __ldg(C.counts + k);
    const uint32_t count = __ldg(C.counts + C.cls);
    const uint32_t max_it = A.settings.max_iterations;
    const float epa_tol = A.settings.epa_tolerance;

    // ---- per-group state (uniform across the group's lanes) ------------------------------------------
    int state = S_FETCH;
    uint64_t pair = 0;
    xform lt, rt;
    uint32_t lk = K_SPHERE, rk = K_SPHERE, lvn = 0, rvn = 0;
    f3 lh = zero3(), rh = zero3();
    const float4* lgv = nullptr;  // global vertex pointers (unstaged scan)
    const float4* rgv = nullptr;
    bool staged = false;
    simplex<FULL> sx;
    sx.n = 0;
    f3 d = zero3();
    uint32_t it = 0;
    int nf = 0, nv = 0, best_face = 0;
    // make_xform needs defined inputs even before the first fetch
    lt = make_xform(make_float4(1.f, 1.f, 0.f, 0.f), make_float4(0.f, 0.f, 0.f, 0.f));
    rt = lt;

    for (;;) {
        // ================= FETCH: grab the next pair of this class ===================================
        if (state == S_FETCH) {
            uint32_t q = 0;
            if (gl == 0) q = atomicAdd(C.next + C.cls, 1u);
            q = __shfl_sync(gmask, q, gw * G);
            if (q >= count) {
                state = S_EXIT;
            } else {
                pair = __ldg(C.perm + begin + q);
                uint32_t li, ri;
                uint64_t lx, rx;
                pair_shape_indices(A, pair, li, ri, lx, rx);
                const shape ls = load_shape(A.shapes, A.verts, li);
                const shape rs = load_shape(A.shapes, A.verts, ri);
                lt = make_xform(__ldg(A.l_t + 2 * lx), __ldg(A.l_t + 2 * lx + 1));
                rt = make_xform(__ldg(A.r_t + 2 * rx), __ldg(A.r_t + 2 * rx + 1));
                lk = ls.kind; rk = rs.kind;
                lh = ls.h; rh = rs.h;
                lvn = lk == K_POLY ? ls.vcnt : 0u;
                rvn = rk == K_POLY ? rs.vcnt : 0u;
                lgv = ls.verts; rgv = rs.verts;
                staged = (lvn + rvn) <= C.vcap;
                if (staged) {
                    // stage both vertex arrays: coalesced 16-byte loads, packed 12-byte stores
                    __syncwarp(gmask);  // previous pair's readers are done
                    for (uint32_t j = gl; j < lvn; j += G) {
                        const float4 v = __ldg(lgv + j);
                        s_verts[3 * j] = v.x; s_verts[3 * j + 1] = v.y; s_verts[3 * j + 2] = v.z;
                    }
                    float* sr = s_verts + 3 * lvn;
                    for (uint32_t j = gl; j < rvn; j += G) {
                        const float4 v = __ldg(rgv + j);
                        sr[3 * j] = v.x; sr[3 * j + 1] = v.y; sr[3 * j + 2] = v.z;
                    }
                    __syncwarp(gmask);
                }
                if (ulps_eq_zero(lt.scale) || ulps_eq_zero(rt.scale)) {
                    // inverse_transform_vector(..).unwrap() panics (util.rs:18, polyhedron.rs:394)
                    if (gl == 0) {
                        store_contact(A.out_contact, pair,
                                      FULL ? CB2_FULL_RESOLUTION : CB2_COLLISION_ONLY, zero_contact());
                        A.out_status[pair] = ST_PANIC_INV_SCALE;
                    }
                    // stay in FETCH: next loop iteration grabs another pair; run a dummy support now
                    lvn = rvn = 0; lk = rk = K_SPHERE;
                } else {
                    // seed direction, gjk/mod.rs:117-122
                    const f3 lp = transform_point(lt, zero3());
                    const f3 rp = transform_point(rt, zero3());
                    d = rp - lp;
                    if (ulps_eq_zero(d)) d = mk3(1.0f, 1.0f, 1.0f);
                    state = S_SEED;
                }
            }
        }
        if (__all_sync(0xffffffffu, state == S_EXIT)) break;

        // ================= SUPPORT: SupportPoint::from_minkowski(d), minkowski/mod.rs:39-60 ===========
        // (executed convergently by every group of the warp, whatever its state)
        f3 sup_v, sup_a;
        {
            const f3 dl = inverse_transform_vector(lt, d);
            const f3 dr = inverse_transform_vector(rt, -d);
            f3 pl = zero3(), pr = zero3();
            if (lk == K_CUBOID) pl = cuboid_local_support(lh, dl);
            else if (lk == K_SPHERE) pl = sphere_local_support(lh.x, dl);
            if (rk == K_CUBOID) pr = cuboid_local_support(rh, dr);
            else if (rk == K_SPHERE) pr = sphere_local_support(rh.x, dr);
            // polyhedron scans (polyhedron.rs:160-176): first strictly greatest dot
            float bl = -INFINITY, br = -INFINITY;
            uint32_t il = 0xFFFFFFFFu, ir = 0xFFFFFFFFu;
            if (staged) {
                const float* sl = s_verts;
                const float* sr = s_verts + 3 * lvn;
                const uint32_t nmax = lvn > rvn ? lvn : rvn;
                for (uint32_t j = gl; j < nmax; j += G) {
                    if (j < lvn) {
                        const float t = dot(mk3(sl[3 * j], sl[3 * j + 1], sl[3 * j + 2]), dl);
                        if (t > bl) { bl = t; il = j; }
                    }
                    if (j < rvn) {
                        const float t = dot(mk3(sr[3 * j], sr[3 * j + 1], sr[3 * j + 2]), dr);
                        if (t > br) { br = t; ir = j; }
                    }
                }
            } else {
                for (uint32_t j = gl; j < lvn; j += G) {
                    const float4 v = __ldg(lgv + j);
                    const float t = dot(mk3(v.x, v.y, v.z), dl);
                    if (t > bl) { bl = t; il = j; }
                }
                for (uint32_t j = gl; j < rvn; j += G) {
                    const float4 v = __ldg(rgv + j);
                    const float t = dot(mk3(v.x, v.y, v.z), dr);
                    if (t > br) { br = t; ir = j; }
                }
            }
            group_argmax<G>(bl, il, 0xffffffffu);
            group_argmax<G>(br, ir, 0xffffffffu);
            if (lk == K_POLY && il != 0xFFFFFFFFu) {
                if (staged) pl = mk3(s_verts[3 * il], s_verts[3 * il + 1], s_verts[3 * il + 2]);
                else { const float4 v = __ldg(lgv + il); pl = mk3(v.x, v.y, v.z); }
            }
            if (rk == K_POLY && ir != 0xFFFFFFFFu) {
                const float* sr = s_verts + 3 * lvn;
                if (staged) pr = mk3(sr[3 * ir], sr[3 * ir + 1], sr[3 * ir + 2]);
                else { const float4 v = __ldg(rgv + ir); pr = mk3(v.x, v.y, v.z); }
            }
            const f3 wl = transform_point(lt, pl);
            const f3 wr = transform_point(rt, pr);
            sup_v = wl - wr;
            sup_a = wl;
        }

        // ================= state-specific step ==========================================================
        bool finished = false;
        uint8_t result = ST_NONE;
        contact_out co = zero_contact();

        if (state == S_SEED) {  // gjk/mod.rs:123-129
            if (dot(sup_v, d) <= 0.0f) {
                finished = true;
            } else {
                sx.n = 0;
                sx.push(sup_v, sup_a);
                d = -d;
                it = 0;
                if (max_it == 0) finished = true;  // `for _ in 0..0` => None
                else state = S_GJK;
            }
        } else if (state == S_GJK) {  // gjk/mod.rs:130-143
            if (dot(sup_v, d) <= 0.0f) {
                finished = true;
            } else {
                sx.push(sup_v, sup_a);
                if (reduce_to_closest_feature(sx, d)) {
                    if (!FULL) {
                        finished = true;
                        result = ST_SOME;
                    } else {
                        // EPA3::process prologue: Polytope::new (epa3d.rs:41-45, 129-135, 201-208)
                        __syncwarp(gmask);
                        if (gl < 4) {
                            float* pv = P.pv(gl);
                            const f3 v = gl == 0 ? sx.v[0] : (gl == 1 ? sx.v[1] : (gl == 2 ? sx.v[2] : sx.v[3]));
                            pv[0] = v.x; pv[1] = v.y; pv[2] = v.z;
                            if constexpr (FULL) {
                                float* pa = P.pa(gl);
                                const f3 a = gl == 0 ? sx.a[0] : (gl == 1 ? sx.a[1] : (gl == 2 ? sx.a[2] : sx.a[3]));
                                pa[0] = a.x; pa[1] = a.y; pa[2] = a.z;
                            }
                        }
                        __syncwarp(gmask);
                        if (gl == 0) coop_face_new(P, 0, 3, 2, 1);  // Face::new, epa3d.rs:201-208
                        if (gl == 1) coop_face_new(P, 1, 3, 1, 0);
                        if (gl == 2) coop_face_new(P, 2, 3, 0, 2);
                        if (gl == 3) coop_face_new(P, 3, 2, 0, 1);
                        __syncwarp(gmask);
                        nv = 4;
                        nf = 4;
                        it = 1;
                        state = S_EPA;
                    }
                } else {
                    ++it;
                    if (it >= max_it) finished = true;  // loop exhausted => None
                }
            }
        } else if (state == S_EPA) {  // epa3d.rs:46-64 with d == face.normal of best_face
            const float4 fb = *P.fnd(best_face);
            const f3 n = mk3(fb.x, fb.y, fb.z);
            const float dd = dot(sup_v, n);
            if (fsub(dd, fb.w) < epa_tol || it >= max_it) {
                // contact(), epa3d.rs:84-114
                const uint32_t idx = *P.fidx(best_face);
                const uint32_t i0 = idx & 255u, i1 = (idx >> 8) & 255u, i2 = (idx >> 16) & 255u;
                float u, v, w;
                barycentric_vector(n * fb.w, P.get_pv(i0), P.get_pv(i1), P.get_pv(i2), u, v, w);
                co.normal = n;
                co.depth = fb.w;
                co.point = (P.get_pa(i0) * u + P.get_pa(i1) * v) + P.get_pa(i2) * w;
                finished = true;
                result = ST_SOME;
            } else {
                // ---- Polytope::add (epa3d.rs:147-175) ------------------------------------------------
                // (1) visibility of every face, in parallel, into a bit mask in shared memory
                for (int wq = gl; wq < GF_CAP / 32; wq += G) P.s_vis[wq] = 0u;
                __syncwarp(gmask);
                for (int f0 = 0; f0 < nf; f0 += G) {
                    const int f = f0 + gl;
                    bool vis = false;
                    if (f < nf) {
                        const float4 fn = *P.fnd(f);
                        const uint32_t idx = *P.fidx(f);
                        vis = dot(mk3(fn.x, fn.y, fn.z), sup_v - P.get_pv(idx & 255u)) > 0.0f;
                    }
                    if (vis) atomicOr(&P.s_vis[f >> 5], 1u << (f & 31));
                }
                __syncwarp(gmask);
                // (2) the reference's sequential swap_remove sweep + edge toggling, replayed in
                //     order by the group's leader lane (the leaders of the warp's groups run it
                //     side by side); visible faces are found with ffs on the bit mask
                int ne = 0;
                int n_faces = nf;
                int overflow = 0;
                if (gl == 0) {
                    int i = 0;
                    while (i < n_faces) {
                        int wi = i >> 5;
                        uint32_t wbits = P.s_vis[wi] & (0xFFFFFFFFu << (i & 31));
                        while (wbits == 0u && (wi + 1) * 32 < n_faces) {
                            ++wi;
                            wbits = P.s_vis[wi];
                        }
                        if (wbits == 0u) break;
                        i = wi * 32 + (__ffs(wbits) - 1);
                        if (i >= n_faces) break;
                        const uint32_t idx = *P.fidx(i);
                        --n_faces;
                        // swap_remove: the last face moves into slot i with its visibility bit
                        const uint32_t last_vis = (P.s_vis[n_faces >> 5] >> (n_faces & 31)) & 1u;
                        P.s_vis[n_faces >> 5] &= ~(1u << (n_faces & 31));
                        if (i != n_faces) {
                            *P.fnd(i) = *P.fnd(n_faces);
                            *P.fidx(i) = *P.fidx(n_faces);
                            if (last_vis) P.s_vis[i >> 5] |= (1u << (i & 31));
                            else P.s_vis[i >> 5] &= ~(1u << (i & 31));
                        }
                        const uint32_t v0 = idx & 255u, v1 = (idx >> 8) & 255u, v2 = (idx >> 16) & 255u;
#pragma unroll
                        for (int e = 0; e < 3; ++e) {  // remove_or_add_edge, epa3d.rs:212-219
                            const uint32_t ea = e == 0 ? v0 : (e == 1 ? v1 : v2);
                            const uint32_t eb = e == 0 ? v1 : (e == 1 ? v2 : v0);
                            const uint16_t rev = (uint16_t)(eb | (ea << 8));
                            int k = 0;
                            while (k < ne && *P.edge(k) != rev) ++k;
                            if (k < ne) {
                                for (int m = k; m + 1 < ne; ++m) *P.edge(m) = *P.edge(m + 1);
                                --ne;
                            } else if (ne >= GE_CAP) {
                                overflow = 1;
                            } else {
                                *P.edge(ne) = (uint16_t)(ea | (eb << 8));
                                ++ne;
                            }
                        }
                    }
                }
                __syncwarp(gmask);
                ne = __shfl_sync(gmask, ne, gw * G);
                n_faces = __shfl_sync(gmask, n_faces, gw * G);
                overflow = __shfl_sync(gmask, overflow, gw * G);
                if (overflow || nv >= GV_CAP || n_faces + ne > GF_CAP) {
                    finished = true;
                    result = ST_SCRATCH_OVERFLOW;
                } else {
                    // (3) append the vertex, build the new faces in parallel (one edge per lane)
                    const uint32_t nn = (uint32_t)nv;
                    if (gl == 0) {
                        float* pv = P.pv(nv);
                        float* pa = P.pa(nv);
                        pv[0] = sup_v.x; pv[1] = sup_v.y; pv[2] = sup_v.z;
                        pa[0] = sup_a.x; pa[1] = sup_a.y; pa[2] = sup_a.z;
                    }
                    ++nv;
                    __syncwarp(gmask);
                    for (int k = gl; k < ne; k += G) {
                        const uint32_t e = *P.edge(k);
                        coop_face_new(P, n_faces + k, nn, e & 255u, e >> 8);
                    }
                    nf = n_faces + ne;
                    ++it;
                    __syncwarp(gmask);
                    if (nf == 0) {  // faces[0] on an empty Vec panics (epa3d.rs:138)
                        finished = true;
                        result = ST_PANIC_EPA_EMPTY;
                    }
                }
            }
        }

        // EPA: choose the face / direction of the next support call (closest_face_to_origin,
        // epa3d.rs:137-145: first strictly smallest distance; a NaN at index 0 sticks)
        if (state == S_EPA && !finished) {
            float bd = INFINITY;
            uint32_t bi = 0xFFFFFFFFu;
            for (int f = gl; f < nf; f += G) {
                float dd = P.fnd(f)->w;
                if (dd != dd) dd = INFINITY;  // NaN never compares smaller
                if (dd < bd || (dd == bd && (uint32_t)f < bi)) { bd = dd; bi = (uint32_t)f; }
            }
#pragma unroll
            for (int off = G / 2; off >= 1; off >>= 1) {
                const float ob = __shfl_xor_sync(gmask, bd, off);
                const uint32_t oi = __shfl_xor_sync(gmask, bi, off);
                if (ob < bd || (ob == bd && oi < bi)) { bd = ob; bi = oi; }
            }
            const float d0 = P.fnd(0)->w;
            best_face = (d0 != d0) ? 0 : (int)bi;
            const float4 fb = *P.fnd(best_face);
            d = mk3(fb.x, fb.y, fb.z);
        }

        if (finished) {
            if (gl == 0) {
                store_contact(A.out_contact, pair, FULL ? CB2_FULL_RESOLUTION : CB2_COLLISION_ONLY,
                              co);
                A.out_status[pair] = result;
            }
            state = S_FETCH;
        }
    }
}

// ---- host side: classify, scatter, one launch per class ------------------------------------------------
struct class_cfg {
    int G;
    uint32_t vcap;
};
static const class_cfg k_class_cfg[NUM_CLASSES] = {
    {0, 0}, {4, 64}, {4, 128}, {8, 256}, {8, 512}, {16, 1024}, {32, 0}};

template <int G>
static cudaError_t launch_coop(const coop_args& C, bool full, int sm_count, cudaStream_t st,
                               int* blocks_out) {
    const int ng = 32 / G;
    const uint32_t vwords = (C.vcap * 3 + 3) & ~3u;
    const size_t smem = (size_t)ng * (vwords + POLY_WORDS) * sizeof(float);
    auto kern = full ? k_coop_intersection<G, true> : k_coop_intersection<G, false>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 32, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 24) per_sm = 24;
    const int blocks = per_sm * sm_count;
    *blocks_out = blocks;
    kern<<<blocks, 32, smem, st>>>(C);
    return cudaGetLastError();
}

size_t coop_spill_bytes(int sm_count) {
    // upper bound on resident groups: 24 blocks/SM x 8 groups
    return (size_t)sm_count * 24 * 8 * SPILL_WORDS * sizeof(float);
}

cudaError_t launch_intersection_classified(const narrow_args& A, bool full, const narrow_ws& W,
                                           int sm_count, cudaStream_t st, uint64_t* launches) {
    if (A.n == 0) return cudaSuccess;
    cudaError_t e = cudaMemsetAsync(W.counts, 0, sizeof(uint32_t) * 3 * NUM_CLASSES, st);
    if (e != cudaSuccess) return e;
    uint32_t* counts = W.counts;
    uint32_t* cursor = W.counts + NUM_CLASSES;
    uint32_t* next = W.counts + 2 * NUM_CLASSES;
    const unsigned nb = (unsigned)((A.n + 255) / 256);
    k_classify<<<nb, 256, 0, st>>>(A, W.cls, counts);
    k_scatter<<<nb, 256, 0, st>>>(A.n, W.cls, counts, cursor, W.perm);
    *launches += 2;
    // class 0: thread-per-pair kernel over perm[0 .. counts[0])
    narrow_args P = A;
    P.perm = W.perm;
    P.dev_count = counts;
    e = launch_intersection(P, full, st);
    ++*launches;
    if (e != cudaSuccess) return e;
    for (uint32_t c = 1; c < NUM_CLASSES; ++c) {
        coop_args C;
        C.A = A;
        C.perm = W.perm;
        C.counts = counts;
        C.next = next;
        C.cls = c;
        C.vcap = k_class_cfg[c].vcap;
        C.spill = W.spill;
        int blocks = 0;
        switch (k_class_cfg[c].G) {
            case 4: e = launch_coop<4>(C, full, sm_count, st, &blocks); break;
            case 8: e = launch_coop<8>(C, full, sm_count, st, &blocks); break;
            case 16: e = launch_coop<16>(C, full, sm_count, st, &blocks); break;
            default: e = launch_coop<32>(C, full, sm_count, st, &blocks); break;
        }
        ++*launches;
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

}  // namespace cb2
