// cb2_poly.cu — sub-warp cooperative GJK3 + EPA3 (GJK3::intersection, gjk/mod.rs:362-391) for
// pairs that involve a ConvexPolyhedron.
//
// Design (B200-first, not a translation of the reference's per-pair loop):
//  * A GROUP of G lanes (8/16/32, chosen per size class) owns one pair.  Scalar work (transforms,
//    simplex logic) is uniform inside the group; the O(V) and O(F) work is split across its lanes:
//    the support_point vertex scan (polyhedron.rs:160-176) and, in EPA, the visibility tests, the
//    horizon extraction, new-face construction and the closest-face search (epa3d.rs:137-175).
//  * Both polyhedra of a pair are staged ONCE per kernel into shared memory as structure-of-arrays
//    planes (x[], y[], z[], padded to 4 with NaN — a NaN dot never wins `dot > max`), so a lane
//    reads 4 vertices with three conflict-free 128-bit loads.  The first-strict-max rule of the
//    reference is kept by (chunk max, first chunk) per lane and a (dot desc, index asc) butterfly.
//  * Two persistent kernels per size class instead of one state machine: k_pgjk runs
//    GJK::intersect (gjk/mod.rs:101-146) and hands the terminal simplex of each hit to k_pepa,
//    which runs EPA3::process (epa3d.rs:27-65).  Every group of a warp is then in the same phase,
//    so the only divergence left is pair turnover; finished groups pull the next pair from a
//    device-side queue, so lanes never idle behind the slowest pair of a warp.
//  * Polytope::add (epa3d.rs:147-175) is reproduced in ORDER without replaying it serially: face
//    visibility becomes a register bit mask (ballot), the swap_remove sweep is a two-pointer walk
//    over that mask (which faces are examined in which order, which tail face lands in which
//    hole), and the horizon list is "edges of visible faces, in examination order, whose reverse
//    edge is not an edge of a visible face" — evaluated in parallel through a vertex-adjacency bit
//    matrix in shared memory.  That equals remove_or_add_edge's list (epa3d.rs:212-219) whenever
//    no directed edge occurs twice; the (non-manifold) exception is detected by the same atomicOr
//    and the pair is handed to the serial second-chance kernel (cb2_narrow.cu).
//  * The polytope lives in per-group shared memory with a small cap (64 faces / 36 vertices covers
//    ~99 % of pairs); pairs that outgrow it are re-run by the same kernel instantiated with the
//    manifold worst case (208 faces / 104 vertices = 2V-4 at the 100-iteration cap).
//  * Pairs are counting-sorted by (size class, vertex-count sub-bucket) first, so each launch has
//    the shared-memory footprint of its class and the groups of a warp scan similar lengths.
#include "cb2_gjk.cuh"
#include "cb2_kernels.h"

namespace cb2 {

// ---- size classes ----------------------------------------------------------------------------------
// class 0: no polyhedron in the pair (thread-per-pair kernel, cb2_narrow.cu)
// class 1..9: padded vertex total pad4(Vl)+pad4(Vr) <= cap (staged in shared memory)
// class 10: larger (scanned straight from global / L2)
__constant__ uint32_t c_class_cap[POLY_NC] = {0, 64, 128, 192, 256, 320, 384, 512, 768, 1024, 0xFFFFFFFFu};
static const uint32_t h_class_cap[POLY_NC] = {0, 64, 128, 192, 256, 320, 384, 512, 768, 1024, 0xFFFFFFFFu};

CB2_DI uint32_t pad4(uint32_t v) { return (v + 3u) & ~3u; }

CB2_DI uint32_t bucket_of(uint32_t lvn, uint32_t rvn, bool any_poly) {
    if (!any_poly) return 0;
    const uint32_t lp = pad4(lvn), rp = pad4(rvn), vt = lp + rp;
    uint32_t c = 1;
    while (c < POLY_NC - 1 && vt > c_class_cap[c]) ++c;
    // sub-bucket: quantised (Vl, Vr) so neighbouring queue entries scan similar lengths
    const uint32_t cap = c == POLY_NC - 1 ? 2048u : c_class_cap[c];
    uint32_t ql = (4u * lp) / (cap + 1u), qr = (4u * rp) / (cap + 1u);
    ql = ql > 3u ? 3u : ql;
    qr = qr > 3u ? 3u : qr;
    return c * POLY_SUB + ql * 4u + qr;
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

__global__ void __launch_bounds__(256) k_classify(narrow_args A, uint8_t* __restrict__ key,
                                                 poly_counters* __restrict__ PC) {
    __shared__ uint32_t h[POLY_NB];
    for (int k = threadIdx.x; k < POLY_NB; k += blockDim.x) h[k] = 0;
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
        const uint32_t b = bucket_of(vl, vr, kl == K_POLY || kr == K_POLY);
        key[i] = (uint8_t)b;
        atomicAdd(&h[b], 1u);
    }
    __syncthreads();
    for (int k = threadIdx.x; k < POLY_NB; k += blockDim.x)
        if (h[k]) atomicAdd(&PC->hist[k], h[k]);
}

__global__ void __launch_bounds__(32) k_bucket_scan(poly_counters* __restrict__ PC) {
    if (threadIdx.x != 0) return;
    uint32_t off = 0;
    for (int b = 0; b < POLY_NB; ++b) {
        if (b % POLY_SUB == 0) PC->class_begin[b / POLY_SUB] = off;
        PC->start[b] = off;
        off += PC->hist[b];
    }
    PC->class_begin[POLY_NC] = off;
}

// perm is grouped by bucket; order inside a bucket is arbitrary (results are written back by
// original pair index, so outputs are deterministic).
__global__ void __launch_bounds__(256) k_scatter(uint64_t n, const uint8_t* __restrict__ key,
                                                poly_counters* __restrict__ PC,
                                                uint32_t* __restrict__ perm) {
    __shared__ uint32_t h[POLY_NB], base[POLY_NB];
    for (int k = threadIdx.x; k < POLY_NB; k += blockDim.x) h[k] = 0;
    __syncthreads();
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t b = 0, local = 0;
    if (i < n) {
        b = key[i];
        local = atomicAdd(&h[b], 1u);
    }
    __syncthreads();
    for (int k = threadIdx.x; k < POLY_NB; k += blockDim.x)
        base[k] = h[k] ? PC->start[k] + atomicAdd(&PC->cursor[k], h[k]) : 0u;
    __syncthreads();
    if (i < n) perm[base[b] + local] = (uint32_t)i;
}

// ---- per-group pair context -------------------------------------------------------------------------
struct pair_ctx {
    xform lt, rt;
    uint32_t lk, rk;      // shape kinds
    f3 lh, rh;            // cuboid half dims / sphere radius
    uint32_t lvn, rvn;    // polyhedron vertex counts (0 for other kinds)
    uint32_t lvp, rvp;    // padded to 4
    const float4* lgv;    // global vertex arrays (x,y,z,0)
    const float4* rgv;
    uint64_t pair;        // original pair index
};

CB2_DI void idle_ctx(pair_ctx& c) {
    c.lt = make_xform(make_float4(1.f, 1.f, 0.f, 0.f), make_float4(0.f, 0.f, 0.f, 0.f));
    c.rt = c.lt;
    c.lk = c.rk = K_CUBOID;
    c.lh = c.rh = zero3();
    c.lvn = c.rvn = c.lvp = c.rvp = 0;
    c.lgv = c.rgv = nullptr;
    c.pair = 0;
}

// load shapes + transforms of `pair`; returns false when inverse_transform_vector would panic
CB2_DI bool load_ctx(const narrow_args& A, uint64_t pair, pair_ctx& c) {
    uint32_t li, ri;
    uint64_t lx, rx;
    pair_shape_indices(A, pair, li, ri, lx, rx);
    const shape ls = load_shape(A.shapes, A.verts, li);
    const shape rs = load_shape(A.shapes, A.verts, ri);
    c.lt = make_xform(__ldg(A.l_t + 2 * lx), __ldg(A.l_t + 2 * lx + 1));
    c.rt = make_xform(__ldg(A.r_t + 2 * rx), __ldg(A.r_t + 2 * rx + 1));
    c.lk = ls.kind;
    c.rk = rs.kind;
    c.lh = ls.h;
    c.rh = rs.h;
    c.lvn = c.lk == K_POLY ? ls.vcnt : 0u;
    c.rvn = c.rk == K_POLY ? rs.vcnt : 0u;
    c.lvp = pad4(c.lvn);
    c.rvp = pad4(c.rvn);
    c.lgv = ls.verts;
    c.rgv = rs.verts;
    c.pair = pair;
    return !(ulps_eq_zero(c.lt.scale) || ulps_eq_zero(c.rt.scale));
}

// stage one polyhedron into SoA planes x[vp] y[vp] z[vp]; pad slots are NaN
template <int G>
CB2_DI void stage_soa(float* __restrict__ s, const float4* __restrict__ gv, uint32_t vn, uint32_t vp,
                      int gl) {
    const float qnan = __int_as_float(0x7fc00000);
    for (uint32_t j = gl; j < vp; j += G) {
        float4 v = make_float4(qnan, qnan, qnan, 0.f);
        if (j < vn) v = __ldg(gv + j);
        s[j] = v.x;
        s[vp + j] = v.y;
        s[2 * vp + j] = v.z;
    }
}

CB2_DI float dot3(float x, float y, float z, f3 d) {
    return fadd(fadd(fmul(x, d.x), fmul(y, d.y)), fmul(z, d.z));
}

// polyhedron.rs:160-176 over staged planes: this lane's (best dot, lowest index reaching it)
template <int G>
CB2_DI void scan_soa(const float* __restrict__ s, uint32_t vp, f3 d, int gl, float& best,
                     uint32_t& bidx) {
    const float4* xs = reinterpret_cast<const float4*>(s);
    const float4* ys = reinterpret_cast<const float4*>(s + vp);
    const float4* zs = reinterpret_cast<const float4*>(s + 2 * vp);
    const uint32_t nc = vp >> 2;
    best = -INFINITY;
    uint32_t bc = 0xFFFFFFFFu;
#pragma unroll 2
    for (uint32_t c = gl; c < nc; c += G) {
        const float4 X = xs[c], Y = ys[c], Z = zs[c];
        const float d0 = dot3(X.x, Y.x, Z.x, d), d1 = dot3(X.y, Y.y, Z.y, d);
        const float d2 = dot3(X.z, Y.z, Z.z, d), d3 = dot3(X.w, Y.w, Z.w, d);
        const float m = fmaxf(fmaxf(d0, d1), fmaxf(d2, d3));  // fmaxf drops NaN operands
        if (m > best) {
            best = m;
            bc = c;
        }
    }
    bidx = 0xFFFFFFFFu;
    if (bc != 0xFFFFFFFFu) {  // which slot of the winning chunk: first one equal to the max
        const float4 X = xs[bc], Y = ys[bc], Z = zs[bc];
        const float d0 = dot3(X.x, Y.x, Z.x, d), d1 = dot3(X.y, Y.y, Z.y, d);
        const float d2 = dot3(X.z, Y.z, Z.z, d);
        const uint32_t slot = d0 == best ? 0u : (d1 == best ? 1u : (d2 == best ? 2u : 3u));
        bidx = 4u * bc + slot;
    }
}

// the same scan straight from global memory (class 10: too large to stage)
template <int G>
CB2_DI void scan_global(const float4* __restrict__ gv, uint32_t vn, f3 d, int gl, float& best,
                        uint32_t& bidx) {
    best = -INFINITY;
    bidx = 0xFFFFFFFFu;
    for (uint32_t j = gl; j < vn; j += G) {
        const float4 v = __ldg(gv + j);
        const float t = dot3(v.x, v.y, v.z, d);
        if (t > best) {
            best = t;
            bidx = j;
        }
    }
}

// (dot desc, index asc) butterfly over the G lanes of every group, two keys at once
template <int G>
CB2_DI void group_argmax2(float& b0, uint32_t& i0, float& b1, uint32_t& i1) {
#pragma unroll
    for (int off = G / 2; off >= 1; off >>= 1) {
        const float ob0 = __shfl_xor_sync(0xffffffffu, b0, off);
        const uint32_t oi0 = __shfl_xor_sync(0xffffffffu, i0, off);
        const float ob1 = __shfl_xor_sync(0xffffffffu, b1, off);
        const uint32_t oi1 = __shfl_xor_sync(0xffffffffu, i1, off);
        if (ob0 > b0 || (ob0 == b0 && oi0 < i0)) {
            b0 = ob0;
            i0 = oi0;
        }
        if (ob1 > b1 || (ob1 == b1 && oi1 < i1)) {
            b1 = ob1;
            i1 = oi1;
        }
    }
}

// SupportPoint::from_minkowski(d), minkowski/mod.rs:39-60.  Executed convergently by the whole
// warp (idle groups run it on an empty context).
template <int G, bool STAGED>
CB2_DI void minkowski_support(const pair_ctx& c, const float* __restrict__ s_verts, f3 d, int gl,
                              f3& sup_v, f3& sup_a) {
    const f3 dl = inverse_transform_vector(c.lt, d);
    const f3 dr = inverse_transform_vector(c.rt, -d);
    f3 pl = zero3(), pr = zero3();
    if (c.lk == K_CUBOID) pl = cuboid_local_support(c.lh, dl);
    else if (c.lk == K_SPHERE) pl = sphere_local_support(c.lh.x, dl);
    if (c.rk == K_CUBOID) pr = cuboid_local_support(c.rh, dr);
    else if (c.rk == K_SPHERE) pr = sphere_local_support(c.rh.x, dr);
    float bl, br;
    uint32_t il, ir;
    const float* sl = s_verts;
    const float* sr = s_verts + 3 * c.lvp;
    if (STAGED) {
        scan_soa<G>(sl, c.lvp, dl, gl, bl, il);
        scan_soa<G>(sr, c.rvp, dr, gl, br, ir);
    } else {
        scan_global<G>(c.lgv, c.lvn, dl, gl, bl, il);
        scan_global<G>(c.rgv, c.rvn, dr, gl, br, ir);
    }
    group_argmax2<G>(bl, il, br, ir);
    if (c.lk == K_POLY && il != 0xFFFFFFFFu) {
        if (STAGED) pl = mk3(sl[il], sl[c.lvp + il], sl[2 * c.lvp + il]);
        else { const float4 v = __ldg(c.lgv + il); pl = mk3(v.x, v.y, v.z); }
    }
    if (c.rk == K_POLY && ir != 0xFFFFFFFFu) {
        if (STAGED) pr = mk3(sr[ir], sr[c.rvp + ir], sr[2 * c.rvp + ir]);
        else { const float4 v = __ldg(c.rgv + ir); pr = mk3(v.x, v.y, v.z); }
    }
    const f3 wl = transform_point(c.lt, pl);
    const f3 wr = transform_point(c.rt, pr);
    sup_v = wl - wr;
    sup_a = wl;
}

struct coop_args {
    narrow_args A;
    const uint32_t* perm;     // pairs sorted by bucket
    poly_counters* PC;
    uint32_t* hits;           // queue positions of FullResolution hits, per class range
    float4* simplex;          // 6 x float4 per queue position: v[0..3], sup_a[0..3]
    uint32_t* over;           // tier-2 list (queue positions whose polytope outgrew tier 1)
    uint32_t cls;
    uint32_t vcap;            // staged floats per group = 3 * vcap
    uint32_t tier;            // EPA: 1 = class hit list, 2 = overflow list
};

enum : int { S_FETCH = 0, S_SEED = 1, S_RUN = 2, S_EXIT = 3 };

// ---- k_pgjk: GJK::intersect (gjk/mod.rs:101-146) ----------------------------------------------------
template <int G, bool STAGED, bool FULL>
__global__ void __launch_bounds__(32) k_pgjk(coop_args C) {
    extern __shared__ __align__(16) float smem[];
    const int lane = threadIdx.x;
    const int gl = lane & (G - 1);
    const int gw = lane / G;
    const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (gw * G));
    const narrow_args& A = C.A;
    float* s_verts = smem + (size_t)gw * (3 * C.vcap);

    const uint32_t begin = C.PC->class_begin[C.cls];
    const uint32_t count = C.PC->class_begin[C.cls + 1] - begin;
    const uint32_t max_it = A.settings.max_iterations;

    int state = S_FETCH;
    pair_ctx c;
    idle_ctx(c);
    uint32_t q = 0;
    simplex<FULL> sx;
    sx.n = 0;
    f3 d = zero3();
    uint32_t it = 0;

    for (;;) {
        if (state == S_FETCH) {
            uint32_t k = 0;
            if (gl == 0) k = atomicAdd(&C.PC->gjk_next[C.cls], 1u);
            k = __shfl_sync(gmask, k, gw * G);
            if (k >= count) {
                state = S_EXIT;
                idle_ctx(c);
            } else {
                q = begin + k;
                const bool ok = load_ctx(A, __ldg(C.perm + q), c);
                if (STAGED) {
                    __syncwarp(gmask);  // previous pair's readers are done
                    stage_soa<G>(s_verts, c.lgv, c.lvn, c.lvp, gl);
                    stage_soa<G>(s_verts + 3 * c.lvp, c.rgv, c.rvn, c.rvp, gl);
                    __syncwarp(gmask);
                }
                if (!ok) {
                    // inverse_transform_vector(..).unwrap() panics (util.rs:18, polyhedron.rs:394)
                    if (gl == 0) {
                        store_contact(A.out_contact, c.pair,
                                      FULL ? CB2_FULL_RESOLUTION : CB2_COLLISION_ONLY, zero_contact());
                        A.out_status[c.pair] = ST_PANIC_INV_SCALE;
                    }
                    idle_ctx(c);  // stay in FETCH; this round's support runs on an empty context
                } else {
                    // seed direction, gjk/mod.rs:117-122
                    const f3 lp = transform_point(c.lt, zero3());
                    const f3 rp = transform_point(c.rt, zero3());
                    d = rp - lp;
                    if (ulps_eq_zero(d)) d = mk3(1.0f, 1.0f, 1.0f);
                    state = S_SEED;
                }
            }
        }
        if (__all_sync(0xffffffffu, state == S_EXIT)) break;

        f3 sup_v, sup_a;
        minkowski_support<G, STAGED>(c, s_verts, d, gl, sup_v, sup_a);

        bool finished = false;
        uint8_t result = ST_NONE;
        if (state == S_SEED) {  // gjk/mod.rs:123-129
            if (dot(sup_v, d) <= 0.0f) {
                finished = true;
            } else {
                sx.n = 0;
                sx.push(sup_v, sup_a);
                d = -d;
                it = 0;
                if (max_it == 0) finished = true;  // `for _ in 0..0` => None
                else state = S_RUN;
            }
        } else if (state == S_RUN) {  // gjk/mod.rs:130-143
            if (dot(sup_v, d) <= 0.0f) {
                finished = true;
            } else {
                sx.push(sup_v, sup_a);
                if (reduce_to_closest_feature(sx, d)) {
                    finished = true;
                    result = ST_SOME;
                } else {
                    ++it;
                    if (it >= max_it) finished = true;  // loop exhausted => None
                }
            }
        }
        if (finished) {
            if (gl == 0) {
                if (FULL && result == ST_SOME) {
                    // hand the terminal simplex to k_pepa (EPA3::process, epa3d.rs:27-45)
                    if constexpr (FULL) {
                        float4* o = C.simplex + (size_t)q * 6;
                        o[0] = make_float4(sx.v[0].x, sx.v[0].y, sx.v[0].z, sx.v[1].x);
                        o[1] = make_float4(sx.v[1].y, sx.v[1].z, sx.v[2].x, sx.v[2].y);
                        o[2] = make_float4(sx.v[2].z, sx.v[3].x, sx.v[3].y, sx.v[3].z);
                        o[3] = make_float4(sx.a[0].x, sx.a[0].y, sx.a[0].z, sx.a[1].x);
                        o[4] = make_float4(sx.a[1].y, sx.a[1].z, sx.a[2].x, sx.a[2].y);
                        o[5] = make_float4(sx.a[2].z, sx.a[3].x, sx.a[3].y, sx.a[3].z);
                    }
                    const uint32_t h = atomicAdd(&C.PC->hit_count[C.cls], 1u);
                    C.hits[begin + h] = q;
                } else {
                    store_contact(A.out_contact, c.pair,
                                  FULL ? CB2_FULL_RESOLUTION : CB2_COLLISION_ONLY, zero_contact());
                    A.out_status[c.pair] = result;
                }
            }
            state = S_FETCH;
        }
    }
}

// ---- k_pepa: EPA3::process (epa3d.rs:27-65) ----------------------------------------------------------
// per-group polytope scratch in shared memory
template <int FCAP, int VCAP, int OCAP>
struct poly_smem {
    static constexpr int AW = (VCAP + 31) / 32;  // adjacency words per vertex row
    float4 fnd[FCAP];          // normal.xyz, distance
    uint32_t fidx[FCAP];       // v0 | v1<<8 | v2<<16
    float pv[VCAP * 3];        // SupportPoint.v
    float pa[VCAP * 3];        // SupportPoint.sup_a
    uint32_t adj[VCAP * AW];   // directed-edge bit matrix (all zero between iterations)
    uint16_t elist[OCAP];      // horizon edges a | b<<8 in reference order
    uint8_t order[OCAP];       // visible faces in the order Polytope::add examines them
    uint8_t mv_dst[OCAP];      // swap_remove: tail face mv_src lands in hole mv_dst
    uint8_t mv_src[OCAP];
};

template <class PS>
CB2_DI f3 ps_pv(const PS& P, uint32_t i) { return mk3(P.pv[3 * i], P.pv[3 * i + 1], P.pv[3 * i + 2]); }
template <class PS>
CB2_DI f3 ps_pa(const PS& P, uint32_t i) { return mk3(P.pa[3 * i], P.pa[3 * i + 1], P.pa[3 * i + 2]); }

// Face::new_impl (epa3d.rs:189-199) into slot
template <class PS>
CB2_DI void ps_face_new(PS& P, int slot, uint32_t a, uint32_t b, uint32_t c) {
    const f3 va = ps_pv(P, a);
    const f3 ab = ps_pv(P, b) - va;
    const f3 ac = ps_pv(P, c) - va;
    const f3 n = normalize(cross(ab, ac));
    const float dist = dot(n, va);
    P.fnd[slot] = make_float4(n.x, n.y, n.z, dist);
    P.fidx[slot] = a | (b << 8) | (c << 16);
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

template <int G, bool STAGED, int FCAP, int VCAP, int OCAP>
__global__ void __launch_bounds__(32) k_pepa(coop_args C) {
    using PS = poly_smem<FCAP, VCAP, OCAP>;
    constexpr int FW = (FCAP + 31) / 32;
    constexpr int AW = PS::AW;
    extern __shared__ __align__(16) float smem[];
    const int lane = threadIdx.x;
    const int gl = lane & (G - 1);
    const int gw = lane / G;
    const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (gw * G));
    const narrow_args& A = C.A;
    constexpr int NG = 32 / G;
    // layout: [NG polytope blocks][NG staged vertex blocks]
    PS& P = reinterpret_cast<PS*>(smem)[gw];
    float* s_verts = smem + (sizeof(PS) / sizeof(float)) * NG + (size_t)gw * (3 * C.vcap);

    uint32_t begin, count;
    uint32_t* cursor;
    const uint32_t* list;
    if (C.tier == 1) {
        begin = C.PC->class_begin[C.cls];
        count = C.PC->hit_count[C.cls];
        cursor = &C.PC->epa_next[C.cls];
        list = C.hits + begin;
    } else {
        count = C.PC->over_count;
        cursor = &C.PC->over_next;
        list = C.over;
    }
    const uint32_t max_it = A.settings.max_iterations;
    const float epa_tol = A.settings.epa_tolerance;

    // the adjacency matrix is kept all-zero between iterations
    for (int k = gl; k < VCAP * AW; k += G) P.adj[k] = 0u;
    __syncwarp(gmask);

    int state = S_FETCH;
    pair_ctx c;
    idle_ctx(c);
    uint32_t q = 0;
    f3 d = zero3();
    uint32_t it = 0;
    int nf = 0, nv = 0, best_face = 0;

    for (;;) {
        bool finished = false;
        uint8_t result = ST_NONE;
        contact_out co = zero_contact();

        if (state == S_FETCH) {
            uint32_t k = 0;
            if (gl == 0) k = atomicAdd(cursor, 1u);
            k = __shfl_sync(gmask, k, gw * G);
            if (k >= count) {
                state = S_EXIT;
                idle_ctx(c);
            } else {
                q = list[k];
                load_ctx(A, __ldg(C.perm + q), c);  // scale was checked by k_pgjk
                __syncwarp(gmask);  // previous pair's readers are done
                if (STAGED) {
                    stage_soa<G>(s_verts, c.lgv, c.lvn, c.lvp, gl);
                    stage_soa<G>(s_verts + 3 * c.lvp, c.rgv, c.rvn, c.rvp, gl);
                }
                // Polytope::new (epa3d.rs:41-45, 129-135): the 4 simplex points ...
                if (gl < 6) {
                    const float4 r = C.simplex[(size_t)q * 6 + gl];
                    float* dst = (gl < 3 ? P.pv : P.pa) + 4 * (gl < 3 ? gl : gl - 3);
                    dst[0] = r.x; dst[1] = r.y; dst[2] = r.z; dst[3] = r.w;
                }
                __syncwarp(gmask);
                // ... and Face::new (3,2,1) (3,1,0) (3,0,2) (2,0,1), epa3d.rs:201-208
                if (gl == 0) ps_face_new(P, 0, 3, 2, 1);
                if (gl == 1) ps_face_new(P, 1, 3, 1, 0);
                if (gl == 2) ps_face_new(P, 2, 3, 0, 2);
                if (gl == 3) ps_face_new(P, 3, 2, 0, 1);
                __syncwarp(gmask);
                nv = 4;
                nf = 4;
                it = 1;
                state = S_SEED;  // pick the first closest face below
            }
        }
        if (__all_sync(0xffffffffu, state == S_EXIT)) break;

        // closest_face_to_origin (epa3d.rs:137-145): first strictly smallest distance; a NaN at
        // index 0 sticks.  Its normal is the next support direction.
        {
            float bd = INFINITY;
            uint32_t bi = 0xFFFFFFFFu;
            for (int f = gl; f < nf; f += G) {
                float dd = P.fnd[f].w;
                if (dd != dd) dd = INFINITY;  // NaN never compares smaller
                if (dd < bd || (dd == bd && (uint32_t)f < bi)) { bd = dd; bi = (uint32_t)f; }
            }
#pragma unroll
            for (int off = G / 2; off >= 1; off >>= 1) {
                const float ob = __shfl_xor_sync(0xffffffffu, bd, off);
                const uint32_t oi = __shfl_xor_sync(0xffffffffu, bi, off);
                if (ob < bd || (ob == bd && oi < bi)) { bd = ob; bi = oi; }
            }
            if (state != S_EXIT) {
                const float d0 = P.fnd[0].w;
                best_face = (d0 != d0) ? 0 : (int)bi;
                const float4 fb = P.fnd[best_face];
                d = mk3(fb.x, fb.y, fb.z);
                state = S_RUN;
            }
        }

        f3 sup_v, sup_a;
        minkowski_support<G, STAGED>(c, s_verts, d, gl, sup_v, sup_a);

        if (state == S_RUN) {  // epa3d.rs:46-64 with d == faces[best_face].normal
            const float4 fb = P.fnd[best_face];
            const f3 n = mk3(fb.x, fb.y, fb.z);
            const float dd = dot(sup_v, n);
            if (fsub(dd, fb.w) < epa_tol || it >= max_it) {
                // contact(), epa3d.rs:84-114
                const uint32_t idx = P.fidx[best_face];
                const uint32_t i0 = idx & 255u, i1 = (idx >> 8) & 255u, i2 = (idx >> 16) & 255u;
                float u, v, w;
                barycentric_vector(n * fb.w, ps_pv(P, i0), ps_pv(P, i1), ps_pv(P, i2), u, v, w);
                co.normal = n;
                co.depth = fb.w;
                co.point = (ps_pa(P, i0) * u + ps_pa(P, i1) * v) + ps_pa(P, i2) * w;
                finished = true;
                result = ST_SOME;
            } else {
                // ---- Polytope::add (epa3d.rs:147-175) ------------------------------------------------
                // (1) visibility of every face -> bit mask (uniform across the group)
                bitmask<FW> vis;
                vis.clear();
                for (int f0 = 0; f0 < nf; f0 += G) {
                    const int f = f0 + gl;
                    bool v = false;
                    if (f < nf) {
                        const float4 fn = P.fnd[f];
                        const uint32_t idx = P.fidx[f];
                        v = dot(mk3(fn.x, fn.y, fn.z), sup_v - ps_pv(P, idx & 255u)) > 0.0f;
                    }
                    const uint32_t bits = (__ballot_sync(gmask, v) >> (gw * G)) &
                                          (G == 32 ? 0xFFFFFFFFu : ((1u << G) - 1u));
                    vis.or_bits(f0, bits);
                }
                // (2) the swap_remove sweep as a two-pointer walk over the mask: `i` climbs from the
                //     front, `len` shrinks from the back; a visible face at slot i is examined, then
                //     the tail face dropped into slot i is examined next, and so on.
                int len = nf, nvis = 0, nmv = 0;
                bool overflow = false;
                {
                    int i = vis.next(0);
                    int cur = i;
                    while (i >= 0 && i < len) {
                        if (nvis >= OCAP) { overflow = true; break; }
                        if (gl == 0) P.order[nvis] = (uint8_t)cur;
                        ++nvis;
                        vis.reset(cur);
                        --len;
                        if (i == len) break;
                        if (vis.test(len)) {
                            cur = len;
                            continue;
                        }
                        if (gl == 0) {
                            P.mv_dst[nmv] = (uint8_t)i;
                            P.mv_src[nmv] = (uint8_t)len;
                        }
                        ++nmv;
                        i = vis.next(i + 1);
                        cur = i;
                    }
                }
                __syncwarp(gmask);
                // (3) horizon = edges of the examined faces, in examination order, whose reverse edge
                //     is not an edge of an examined face (remove_or_add_edge, epa3d.rs:212-219)
                const int T = 3 * nvis;
                bool dup = false;
                if (!overflow) {
                    for (int t = gl; t < T; t += G) {
                        const uint32_t idx = P.fidx[P.order[t / 3]];
                        const int e = t % 3;
                        const uint32_t ea = (idx >> (8 * e)) & 255u;
                        const uint32_t eb = (idx >> (8 * (e == 2 ? 0 : e + 1))) & 255u;
                        const uint32_t bit = 1u << (eb & 31);
                        const uint32_t old = atomicOr(&P.adj[ea * AW + (eb >> 5)], bit);
                        dup |= (old & bit) != 0u;
                    }
                }
                __syncwarp(gmask);
                int ne = 0;
                if (!overflow) {
                    for (int t0 = 0; t0 < T; t0 += G) {
                        const int t = t0 + gl;
                        bool keep = false;
                        uint32_t ea = 0, eb = 0;
                        if (t < T) {
                            const uint32_t idx = P.fidx[P.order[t / 3]];
                            const int e = t % 3;
                            ea = (idx >> (8 * e)) & 255u;
                            eb = (idx >> (8 * (e == 2 ? 0 : e + 1))) & 255u;
                            keep = ((P.adj[eb * AW + (ea >> 5)] >> (ea & 31)) & 1u) == 0u;
                        }
                        const uint32_t bits = (__ballot_sync(gmask, keep) >> (gw * G)) &
                                              (G == 32 ? 0xFFFFFFFFu : ((1u << G) - 1u));
                        if (keep) {
                            const int pos = ne + __popc(bits & ((1u << gl) - 1u));
                            if (pos < OCAP) P.elist[pos] = (uint16_t)(ea | (eb << 8));
                        }
                        ne += __popc(bits);
                    }
                }
                dup = __any_sync(gmask, dup);
                __syncwarp(gmask);
                // (4) clear the adjacency bits again; apply the swap_remove moves
                for (int t = gl; t < T; t += G) {
                    const uint32_t idx = P.fidx[P.order[t / 3]];
                    const uint32_t ea = (idx >> (8 * (t % 3))) & 255u;
#pragma unroll
                    for (int k = 0; k < AW; ++k) P.adj[ea * AW + k] = 0u;
                }
                __syncwarp(gmask);
                if (overflow || dup || ne > OCAP || nv >= VCAP || len + ne > FCAP) {
                    finished = true;
                    result = ST_SCRATCH_OVERFLOW;
                } else {
                    for (int m = gl; m < nmv; m += G) {
                        const int dst = P.mv_dst[m], src = P.mv_src[m];
                        P.fnd[dst] = P.fnd[src];
                        P.fidx[dst] = P.fidx[src];
                    }
                    // (5) append the vertex, build the new faces (one horizon edge per lane)
                    const uint32_t nn = (uint32_t)nv;
                    if (gl == 0) {
                        P.pv[3 * nn] = sup_v.x; P.pv[3 * nn + 1] = sup_v.y; P.pv[3 * nn + 2] = sup_v.z;
                        P.pa[3 * nn] = sup_a.x; P.pa[3 * nn + 1] = sup_a.y; P.pa[3 * nn + 2] = sup_a.z;
                    }
                    ++nv;
                    __syncwarp(gmask);
                    for (int k = gl; k < ne; k += G) {
                        const uint32_t e = P.elist[k];
                        ps_face_new(P, len + k, nn, e & 255u, e >> 8);
                    }
                    nf = len + ne;
                    ++it;
                    __syncwarp(gmask);
                    if (nf == 0) {  // faces[0] on an empty Vec panics (epa3d.rs:138)
                        finished = true;
                        result = ST_PANIC_EPA_EMPTY;
                    }
                }
            }
        }

        if (finished) {
            if (gl == 0) {
                if (result == ST_SCRATCH_OVERFLOW && C.tier == 1) {
                    const uint32_t h = atomicAdd(&C.PC->over_count, 1u);
                    C.over[h] = q;
                } else {
                    store_contact(A.out_contact, c.pair, CB2_FULL_RESOLUTION, co);
                    A.out_status[c.pair] = result;
                }
            }
            state = S_FETCH;
            nf = 0;
        }
    }
}

// ---- host side: classify, sort, two launches per class ---------------------------------------------------
struct class_cfg {
    int G;
    bool staged;
};
static const class_cfg k_class_cfg[POLY_NC] = {{0, false}, {8, true}, {8, true}, {8, true},
                                               {8, true},  {8, true}, {8, true}, {8, true},
                                               {16, true}, {16, true}, {32, false}};

constexpr int T1_F = 64, T1_V = 36, T1_O = 32;    // tier 1 polytope caps
constexpr int T2_F = 208, T2_V = 104, T2_O = 96;  // tier 2: manifold worst case at 100 iterations

template <class K>
static cudaError_t launch_persistent(K kern, const coop_args& C, size_t smem, int sm_count,
                                     cudaStream_t st) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 32, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    kern<<<per_sm * sm_count, 32, smem, st>>>(C);
    return cudaGetLastError();
}

template <int G, bool STAGED>
static cudaError_t launch_gjk(const coop_args& C, bool full, int sm_count, cudaStream_t st) {
    const size_t smem = (size_t)(32 / G) * 3 * C.vcap * sizeof(float);
    if (full) return launch_persistent(k_pgjk<G, STAGED, true>, C, smem, sm_count, st);
    return launch_persistent(k_pgjk<G, STAGED, false>, C, smem, sm_count, st);
}
template <int G, bool STAGED, int F, int V, int O>
static cudaError_t launch_epa(const coop_args& C, int sm_count, cudaStream_t st) {
    const size_t smem = (size_t)(32 / G) * (sizeof(poly_smem<F, V, O>) + 3 * C.vcap * sizeof(float));
    return launch_persistent(k_pepa<G, STAGED, F, V, O>, C, smem, sm_count, st);
}

size_t poly_ws_bytes_per_pair() { return 1 + 4 + 4 + 4 + 6 * sizeof(float4); }

cudaError_t launch_intersection_classified(const narrow_args& A, bool full, const narrow_ws& W,
                                           int sm_count, uint32_t max_vcnt, cudaStream_t st,
                                           uint64_t* launches) {
    if (A.n == 0) return cudaSuccess;
    cudaError_t e = cudaMemsetAsync(W.PC, 0, sizeof(poly_counters), st);
    if (e != cudaSuccess) return e;
    const unsigned nb = (unsigned)((A.n + 255) / 256);
    k_classify<<<nb, 256, 0, st>>>(A, W.key, W.PC);
    k_bucket_scan<<<1, 32, 0, st>>>(W.PC);
    k_scatter<<<nb, 256, 0, st>>>(A.n, W.key, W.PC, W.perm);
    *launches += 3;
    // class 0: thread-per-pair kernel over perm[0 .. class_begin[1])
    narrow_args P0 = A;
    P0.perm = W.perm;
    P0.dev_count = &W.PC->class_begin[1];
    e = launch_intersection(P0, full, st);
    ++*launches;
    if (e != cudaSuccess) return e;

    coop_args C;
    C.A = A;
    C.perm = W.perm;
    C.PC = W.PC;
    C.hits = W.hits;
    C.simplex = W.simplex;
    C.over = W.over;
    C.tier = 1;
    // classes whose smallest member needs more vertices than any two table shapes have are empty
    const uint32_t vt_max = 2u * ((max_vcnt + 3u) & ~3u);
    for (int phase = 0; phase < (full ? 2 : 1); ++phase) {
        for (uint32_t c = 1; c < POLY_NC; ++c) {
            if (h_class_cap[c - 1] >= vt_max && c > 1) continue;
            C.cls = c;
            C.vcap = k_class_cfg[c].staged ? h_class_cap[c] : 0u;
            if (phase == 0) {
                switch (k_class_cfg[c].G) {
                    case 8: e = launch_gjk<8, true>(C, full, sm_count, st); break;
                    case 16: e = launch_gjk<16, true>(C, full, sm_count, st); break;
                    default: e = launch_gjk<32, false>(C, full, sm_count, st); break;
                }
            } else {
                switch (k_class_cfg[c].G) {
                    case 8: e = launch_epa<8, true, T1_F, T1_V, T1_O>(C, sm_count, st); break;
                    case 16: e = launch_epa<16, true, T1_F, T1_V, T1_O>(C, sm_count, st); break;
                    default: e = launch_epa<32, false, T1_F, T1_V, T1_O>(C, sm_count, st); break;
                }
            }
            ++*launches;
            if (e != cudaSuccess) return e;
        }
    }
    if (full) {
        // tier 2: pairs whose polytope outgrew the tier-1 scratch, any class, vertices from global
        C.tier = 2;
        C.cls = 0;
        C.vcap = 0;
        e = launch_epa<32, false, T2_F, T2_V, T2_O>(C, sm_count, st);
        ++*launches;
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

}  // namespace cb2
