// cb2_narrow.cu — thread-per-pair GJK3 / EPA3 kernels (Sphere / Cuboid pairs; polyhedra are
// accepted here with a serial vertex scan, the fast path for them is cb2_poly.cu).
//
// Replaces, batched: GJK3::intersection (gjk/mod.rs:362-391 -> intersect :101-146 +
// EPA3::process epa3d.rs:27-65), GJK3::distance (:271-321) and
// GJK3::intersection_time_of_impact (:163-256).
//
// Layout: transforms are 32-byte records read as two 128-bit loads per thread (a warp reads
// 1 KiB contiguous); shape-table entries are 32 bytes, two 128-bit __ldg loads.  The simplex
// is held in registers.  The EPA polytope (<= 104 vertices, <= 256 faces) is per-thread local
// memory: the hardware interleaves local memory per lane, so a warp walking the face list in
// lockstep reads 128-byte coalesced lines out of L1.
#include <cstdlib>

#include "cb2_pair.cuh"

namespace cb2 {

// ---- EPA3 (epa3d.rs) over a polytope store ------------------------------------------------------
// local_store: per-thread local memory (the hardware interleaves it per lane, so a warp walking
// the face list in lockstep reads coalesced 128-byte lines out of L1).  Sized for the manifold
// bound (2V-4 = 202 faces at the 100-iteration cap) plus slack.  About 5 pairs per million
// (curved shapes) grow a NON-manifold polytope past that (observed up to 1050 faces); those
// return ST_SCRATCH_OVERFLOW here and are redone by k_retry_overflow over a global_store.
constexpr int EPA_MAX_VERTS = 104;   // 4 + (CB2_MAX_ITERATIONS - 1) adds = 103
constexpr int EPA_MAX_FACES = 256;
constexpr int EPA_MAX_EDGES = 192;

struct local_store {
    float4 fnd_[EPA_MAX_FACES];     // normal.xyz, distance
    uint32_t fidx_[EPA_MAX_FACES];  // v0 | v1<<8 | v2<<16
    f3 pv_[EPA_MAX_VERTS];          // SupportPoint.v
    f3 pa_[EPA_MAX_VERTS];          // SupportPoint.sup_a
    uint16_t edges_[EPA_MAX_EDGES]; // a | b<<8, horizon list in reference order
    CB2_DI float4& fnd(int i) { return fnd_[i]; }
    CB2_DI uint32_t& fidx(int i) { return fidx_[i]; }
    CB2_DI f3& pv(int i) { return pv_[i]; }
    CB2_DI f3& pa(int i) { return pa_[i]; }
    CB2_DI uint16_t& edge(int i) { return edges_[i]; }
    CB2_DI int fcap() const { return EPA_MAX_FACES; }
    CB2_DI int vcap() const { return EPA_MAX_VERTS; }
    CB2_DI int ecap() const { return EPA_MAX_EDGES; }
};
// Face::new_impl, epa3d.rs:189-199
template <class S>
CB2_DI void face_new(S& P, int slot, uint32_t a, uint32_t b, uint32_t c) {
    const f3 va = P.pv(a);
    const f3 ab = P.pv(b) - va;
    const f3 ac = P.pv(c) - va;
    const f3 n = normalize(cross(ab, ac));
    const float dist = dot(n, va);
    P.fnd(slot) = make_float4(n.x, n.y, n.z, dist);
    P.fidx(slot) = a | (b << 8) | (c << 16);
}

// EPA3::process, epa3d.rs:27-65 (+ Polytope::add :147-175, contact :84-114)
template <class S>
CB2_DI uint8_t epa_process(S& P, const pair_in& p, const simplex<true>& s, float tol,
                           uint32_t max_iterations, contact_out& out) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        P.pv(i) = s.v[i];
        P.pa(i) = s.a[i];
    }
    int nv = 4;
    face_new(P, 0, 3, 2, 1);  // Face::new, epa3d.rs:201-208
    face_new(P, 1, 3, 1, 0);
    face_new(P, 2, 3, 0, 2);
    face_new(P, 3, 2, 0, 1);
    int nf = 4;
    uint32_t it = 1;
    for (;;) {
        if (nf == 0) return ST_PANIC_EPA_EMPTY;
        // closest_face_to_origin: first strictly smallest distance (:137-145)
        int best = 0;
        float best_d = P.fnd(0).w;
        for (int f = 1; f < nf; ++f) {
            const float d = P.fnd(f).w;
            if (d < best_d) {
                best_d = d;
                best = f;
            }
        }
        const float4 fb = P.fnd(best);
        const f3 n = mk3(fb.x, fb.y, fb.z);
        f3 sv, sa;
        minkowski<true>(p, n, sv, sa);
        const float d = dot(sv, n);
        if (fsub(d, fb.w) < tol || it >= max_iterations) {
            // contact(): barycentric of normal*distance on the face, blend sup_a (:84-114)
            const uint32_t idx = P.fidx(best);
            const uint32_t i0 = idx & 255u, i1 = (idx >> 8) & 255u, i2 = (idx >> 16) & 255u;
            float u, v, w;
            barycentric_vector(n * fb.w, P.pv(i0), P.pv(i1), P.pv(i2), u, v, w);
            out.normal = n;
            out.depth = fb.w;
            out.point = (P.pa(i0) * u + P.pa(i1) * v) + P.pa(i2) * w;
            out.toi = 0.0f;
            return ST_SOME;
        }
        // Polytope::add: remove visible faces (swap_remove, no i++ on removal), toggle edges
        int ne = 0;
        int f = 0;
        while (f < nf) {
            const float4 fn = P.fnd(f);
            const uint32_t idx = P.fidx(f);
            const float vis = dot(mk3(fn.x, fn.y, fn.z), sv - P.pv(idx & 255u));
            if (vis > 0.0f) {
                --nf;
                P.fnd(f) = P.fnd(nf);
                P.fidx(f) = P.fidx(nf);
                const uint32_t v0 = idx & 255u, v1 = (idx >> 8) & 255u, v2 = (idx >> 16) & 255u;
#pragma unroll
                for (int e = 0; e < 3; ++e) {
                    // remove_or_add_edge (:212-219): drop the first stored reverse edge, else push
                    const uint32_t ea = e == 0 ? v0 : (e == 1 ? v1 : v2);
                    const uint32_t eb = e == 0 ? v1 : (e == 1 ? v2 : v0);
                    const uint16_t rev = (uint16_t)(eb | (ea << 8));
                    int k = 0;
                    while (k < ne && P.edge(k) != rev) ++k;
                    if (k < ne) {
                        for (int m = k; m + 1 < ne; ++m) P.edge(m) = P.edge(m + 1);
                        --ne;
                    } else {
                        if (ne >= P.ecap()) return ST_SCRATCH_OVERFLOW;
                        P.edge(ne++) = (uint16_t)(ea | (eb << 8));
                    }
                }
            } else {
                ++f;
            }
        }
        if (nv >= P.vcap() || nf + ne > P.fcap()) return ST_SCRATCH_OVERFLOW;
        const uint32_t nn = (uint32_t)nv;
        P.pv(nn) = sv;
        P.pa(nn) = sa;
        ++nv;
        for (int k = 0; k < ne; ++k) {
            const uint32_t e = P.edge(k);
            face_new(P, nf + k, nn, e & 255u, e >> 8);
        }
        nf += ne;
        ++it;
    }
}

__device__ __noinline__ uint8_t epa_process_local(const pair_in& p, const simplex<true>& s,
                                                  float tol, uint32_t max_iterations,
                                                  contact_out& out) {
    local_store P;
    return epa_process(P, p, s, tol, max_iterations, out);
}

// ---- kernels -----------------------------------------------------------------------------------
template <bool FULL>
__global__ void __launch_bounds__(128) k_intersection(narrow_args A) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= A.n) return;
    pair_in p;
    contact_out c = zero_contact();
    uint8_t st = ST_NONE;
    if (const uint8_t bad = load_pair(A, i, A.l_t, A.r_t, p)) {
        st = bad;
    } else {
        simplex<FULL> s;
        if (gjk_intersect(p, A.settings.max_iterations, s)) {
            if constexpr (FULL) {
                st = epa_process_local(p, reinterpret_cast<const simplex<true>&>(s),
                                       A.settings.epa_tolerance, A.settings.max_iterations, c);
                if (st != ST_SOME) c = zero_contact();
            } else {
                st = ST_SOME;
            }
        }
        if (p.ls.hang | p.rs.hang) {  // the reference's hill climb never returned
            st = ST_NONCONVERGED;
            c = zero_contact();
        }
    }
    store_contact(A.out_contact, i, FULL ? CB2_FULL_RESOLUTION : CB2_COLLISION_ONLY, c);
    A.out_status[i] = st;
}

// GJK::intersect (gjk/mod.rs:101-146), thread per pair, as the first stage of the GJK -> EPA
// pipeline: misses (and CollisionOnly hits) are final here; FullResolution hits hand their
// terminal simplex to the cooperative EPA kernel (cb2_epa.cu) through a compacted work list.
// Pairs with a Sphere go straight to the largest polytope tier (curved shapes run EPA to the
// iteration cap: ~137 faces).
#ifndef CB2_GJK_FULL_MINB
#define CB2_GJK_FULL_MINB 4
#endif
template <bool FULL>
__global__ void __launch_bounds__(128, FULL ? CB2_GJK_FULL_MINB : 4) k_gjk_hits(narrow_args A, narrow_ws W) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool hit = false, curved = false;
    if (i < A.n) {
        pair_in p;
        uint8_t st = ST_NONE;
        if (const uint8_t bad = load_pair(A, i, A.l_t, A.r_t, p)) {
            st = bad;
        } else {
            simplex<FULL> s;
            const bool inside = gjk_intersect(p, A.settings.max_iterations, s);
            if (p.ls.hang | p.rs.hang) {  // the reference's hill climb never returned
                st = ST_NONCONVERGED;
            } else if (inside) {
                st = ST_SOME;
                if constexpr (FULL) {
                    hit = true;
                    curved = is_curved(p.ls.kind) || is_curved(p.rs.kind);
                    float4* o = W.simplex + i * 8;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        o[k] = make_float4(s.v[k].x, s.v[k].y, s.v[k].z, 0.f);
                        o[4 + k] = make_float4(s.a[k].x, s.a[k].y, s.a[k].z, 0.f);
                    }
                }
            }
        }
        if (!hit) {
            store_contact(A.out_contact, i, FULL ? CB2_FULL_RESOLUTION : CB2_COLLISION_ONLY,
                          zero_contact());
            A.out_status[i] = st;
        }
    }
    if constexpr (FULL) {
        // warp-aggregated append to the first / last tier's work list
        const unsigned lane = threadIdx.x & 31u;
        const unsigned m1 = __ballot_sync(0xffffffffu, hit && !curved);
        const unsigned m3 = __ballot_sync(0xffffffffu, hit && curved);
        uint32_t b1 = 0, b3 = 0;
        if (lane == 0) {
            if (m1) b1 = atomicAdd(&W.EC->n_list[0], __popc(m1));
            if (m3) b3 = atomicAdd(&W.EC->n_list[EPA_TIERS - 1], __popc(m3));
        }
        b1 = __shfl_sync(0xffffffffu, b1, 0);
        b3 = __shfl_sync(0xffffffffu, b3, 0);
        const unsigned below = (1u << lane) - 1u;
        if (hit && !curved) W.list[0][b1 + __popc(m1 & below)] = (uint32_t)i;
        if (hit && curved) W.list[EPA_TIERS - 1][b3 + __popc(m3 & below)] = (uint32_t)i;
    }
}

// GJK::intersect on its own (gjk/mod.rs:101-146): Some(simplex) -> the tetrahedron's four
// SupportPoints {v, sup_a, sup_b} (minkowski/mod.rs:16-23) in the reference's Vec order
__global__ void __launch_bounds__(128) k_gjk_intersect(narrow_args A) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= A.n) return;
    pair_in p;
    simplex<true, true> s;
    s.n = 0;
    uint8_t st = ST_NONE;
    if (const uint8_t bad = load_pair(A, i, A.l_t, A.r_t, p)) {
        st = bad;
    } else {
        const bool inside = gjk_intersect(p, A.settings.max_iterations, s);
        if (p.ls.hang | p.rs.hang) st = ST_NONCONVERGED;  // the reference's hill climb never returned
        else if (inside) st = ST_SOME;
    }
    float* o = reinterpret_cast<float*>(A.out_simplex + 4 * i);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const bool have = st == ST_SOME;
        const f3 v = have ? s.v[k] : zero3(), a = have ? s.a[k] : zero3(), b = have ? s.b[k] : zero3();
        o[9 * k + 0] = v.x; o[9 * k + 1] = v.y; o[9 * k + 2] = v.z;
        o[9 * k + 3] = a.x; o[9 * k + 4] = a.y; o[9 * k + 5] = a.z;
        o[9 * k + 6] = b.x; o[9 * k + 7] = b.y; o[9 * k + 8] = b.z;
    }
    A.out_status[i] = st;
}

// GJK::get_contact_manifold (gjk/mod.rs:326-344) = EPA3::process on a caller-held simplex: this
// kernel stands where k_gjk_hits stands in GJK::intersection — it queues every simplex of four
// points for the EPA tiers (fewer points: None, epa3d.rs:38-40)
__global__ void __launch_bounds__(128) k_manifold_seed(narrow_args A, narrow_ws W) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool hit = false, curved = false;
    if (i < A.n) {
        pair_in p;
        uint8_t st = ST_NONE;
        if (const uint8_t bad = load_pair(A, i, A.l_t, A.r_t, p)) {
            st = bad;  // EPA's first support call is where the reference would panic
        } else if (!A.in_simplex_len || __ldg(A.in_simplex_len + i) >= 4u) {
            hit = true;
            curved = is_curved(p.ls.kind) || is_curved(p.rs.kind);
            const float* s = reinterpret_cast<const float*>(A.in_simplex + 4 * i);
            float4* o = W.simplex + i * 8;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                o[k] = make_float4(__ldg(s + 9 * k), __ldg(s + 9 * k + 1), __ldg(s + 9 * k + 2), 0.f);
                o[4 + k] = make_float4(__ldg(s + 9 * k + 3), __ldg(s + 9 * k + 4), __ldg(s + 9 * k + 5), 0.f);
            }
        }
        if (!hit) {
            store_contact(A.out_contact, i, CB2_FULL_RESOLUTION, zero_contact());
            A.out_status[i] = st;
        }
    }
    const unsigned lane = threadIdx.x & 31u;
    const unsigned m1 = __ballot_sync(0xffffffffu, hit && !curved);
    const unsigned m3 = __ballot_sync(0xffffffffu, hit && curved);
    uint32_t b1 = 0, b3 = 0;
    if (lane == 0) {
        if (m1) b1 = atomicAdd(&W.EC->n_list[0], __popc(m1));
        if (m3) b3 = atomicAdd(&W.EC->n_list[EPA_TIERS - 1], __popc(m3));
    }
    b1 = __shfl_sync(0xffffffffu, b1, 0);
    b3 = __shfl_sync(0xffffffffu, b3, 0);
    const unsigned below = (1u << lane) - 1u;
    if (hit && !curved) W.list[0][b1 + __popc(m1 & below)] = (uint32_t)i;
    if (hit && curved) W.list[EPA_TIERS - 1][b3 + __popc(m3 & below)] = (uint32_t)i;
}

// GJK::distance, gjk/mod.rs:271-321
__global__ void __launch_bounds__(128) k_distance(narrow_args A) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= A.n) return;
    pair_in p;
    uint8_t st = ST_NONE;
    float result = 0.0f;
    if (const uint8_t bad = load_pair(A, i, A.l_t, A.r_t, p)) {
        st = bad;
    } else {
        const f3 d0 = seed_direction(p.lt, p.rt);
        simplex<false> s;
        s.n = 0;
        f3 v, a;
        minkowski<false>(p, d0, v, a);
        s.push(v, a);
        minkowski<false>(p, -d0, v, a);
        s.push(v, a);
        for (uint32_t it = 0; it < A.settings.max_iterations; ++it) {
            uint8_t panic = 0;
            f3 d = closest_point_to_origin(s, panic);
            if (panic) {
                st = panic;
                break;
            }
            if (ulps_eq_zero(d)) break;  // None
            d = -d;
            minkowski<false>(p, d, v, a);
            const float dp = dot(v, d);
            const float dz = dot(s.v[0], d);
            if (fsub(dp, dz) < A.settings.distance_tolerance) {
                result = magnitude(d);
                st = ST_SOME;
                break;
            }
            s.push(v, a);
        }
        if (p.ls.hang | p.rs.hang) {
            st = ST_NONCONVERGED;
            result = 0.0f;
        }
    }
    A.out_distance[i] = result;
    A.out_status[i] = st;
}

// GJK::intersection_time_of_impact, gjk/mod.rs:163-256 (+ iter_cap -> ST_NONCONVERGED)
__global__ void __launch_bounds__(128) k_time_of_impact(narrow_args A) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= A.n) return;
    pair_in p;
    contact_out c = zero_contact();
    uint8_t st = ST_NONE;
    if (const uint8_t bad = load_pair(A, i, A.l_t, A.r_t, p)) {
        st = bad;
    } else {
        uint64_t lx = i, rx = i;
        const xform lt1 = load_xform(A.l_t1, lx);
        const xform rt1 = load_xform(A.r_t1, rx);
        const f3 left_lin_vel = transform_point(lt1, zero3()) - transform_point(p.lt, zero3());
        const f3 right_lin_vel = transform_point(rt1, zero3()) - transform_point(p.rt, zero3());
        const f3 ray = right_lin_vel - left_lin_vel;
        float lambda = 0.0f;
        f3 normal = zero3();
        f3 ray_origin = zero3();
        simplex<false> s;
        s.n = 0;
        f3 v, pv, a;
        minkowski<false>(p, -ray, v, a);
        uint32_t it = 0;
        bool done = false;
        while (magnitude2(v) > A.settings.continuous_tolerance) {
            if (it >= A.iter_cap) {
                st = ST_NONCONVERGED;
                done = true;
                break;
            }
            ++it;
            minkowski<false>(p, -v, pv, a);
            const float vp = dot(v, pv);
            const float vr = dot(v, ray);
            if (vp > fmul(lambda, vr)) {
                if (vr > 0.0f) {
                    lambda = fdiv(vp, vr);
                    if (lambda > 1.0f) {
                        done = true;
                        break;
                    }
                    ray_origin = ray * lambda;
                    s.n = 0;
                    normal = -v;
                } else {
                    done = true;
                    break;
                }
            }
            s.push(pv - ray_origin, a);
            uint8_t panic = 0;
            v = closest_point_to_origin(s, panic);
            if (panic) {
                st = panic;
                done = true;
                break;
            }
        }
        if (!done && magnitude2(v) <= A.settings.continuous_tolerance) {
            // translation_interpolate (traits.rs:233-246): disp lerp, rot/scale from END
            xform t = rt1;
            t.disp = lerp(p.rt.disp, rt1.disp, lambda);
            c.normal = -normalize(normal);
            c.depth = magnitude(v);
            c.point = transform_point(t, ray_origin);
            c.toi = lambda;
            st = ST_SOME;
        }
        if (p.ls.hang | p.rs.hang) {
            st = ST_NONCONVERGED;
            c = zero_contact();
        }
    }
    store_contact(A.out_contact, i, CB2_FULL_RESOLUTION, c);
    A.out_status[i] = st;
}

// ---- persistent thread-per-pair kernels with lane refill --------------------------------------------
// GJK::distance runs 2..100 support evaluations per pair (Sphere-Cuboid pairs mostly to the cap) and
// the time-of-impact loop 1..iter_cap; in a one-thread-per-pair grid a warp is as slow as its slowest
// lane (~25 % lane utilisation on cfg 2; one never-converging pair in 500 holds its warp for 1024
// rounds on cfg 5).  Here every loop round is exactly ONE Minkowski support per lane plus the
// state's pre/post step, and a lane whose pair finished pulls the next pair from a device queue
// (warp-aggregated atomic), so the lanes stay busy whatever their pairs' trip counts.
CB2_DI bool refill_lane(const narrow_args& A, bool want, uint64_t& i) {
    const unsigned need = __ballot_sync(0xffffffffu, want);
    if (!need) return false;
    const unsigned lane = threadIdx.x & 31u;
    uint32_t base = 0;
    if (lane == (unsigned)(__ffs(need) - 1)) base = atomicAdd(A.queue, (uint32_t)__popc(need));
    base = __shfl_sync(0xffffffffu, base, __ffs(need) - 1);
    i = (uint64_t)base + __popc(need & ((1u << lane) - 1u));
    return want;
}

// GJK::intersect (gjk/mod.rs:101-146) as the first stage of the GJK -> EPA pipeline, lane-refill form
// of k_gjk_hits: misses exit after 1-2 supports, hits after ~4, so a refilled lane is worth ~2x.
#ifndef CB2_GJK_MINB  // resident blocks per SM the GJK stage is compiled for (4: 128 registers)
#define CB2_GJK_MINB 4
#endif
template <bool FULL>
__global__ void __launch_bounds__(128, CB2_GJK_MINB) k_gjk_hits_p(narrow_args A, narrow_ws W) {
    const uint32_t max_it = A.settings.max_iterations;
    bool active = false, done = false;
    uint64_t i = 0;
    pair_in p;
    p.T = table_of(A.T);
    simplex<FULL> s;
    s.n = 0;
    f3 d = zero3();
    int state = 0;
    uint32_t it = 0;
    for (;;) {
        uint64_t ni;
        if (refill_lane(A, !active && !done, ni)) {
            if (ni >= A.n) {
                done = true;
            } else {
                i = ni;
                if (const uint8_t bad = load_pair(A, i, A.l_t, A.r_t, p)) {
                    store_contact(A.out_contact, i, FULL ? CB2_FULL_RESOLUTION : CB2_COLLISION_ONLY,
                                  zero_contact());
                    A.out_status[i] = bad;
                } else {
                    d = seed_direction(p.lt, p.rt);
                    state = 0;
                    active = true;
                }
            }
        }
        if (__all_sync(0xffffffffu, done && !active)) break;
        // ---- one Minkowski support, then the state's step ---------------------------------------------
        f3 v = zero3(), a = zero3();
        if (active) minkowski<FULL, true>(p, d, v, a);
        bool finished = false, inside = false;
        if (active) {
            if (dot(v, d) <= 0.0f) {
                finished = true;  // gjk/mod.rs:124, 132: no intersection
            } else if (state == 0) {
                s.n = 0;
                s.push(v, a);
                d = -d;
                it = 0;
                state = 1;
                finished = max_it == 0;  // `for _ in 0..0` => None
            } else {
                s.push(v, a);
                if (reduce_to_closest_feature(s, d)) {
                    finished = true;
                    inside = true;
                } else {
                    ++it;
                    finished = it >= max_it;  // loop exhausted => None
                }
            }
        }
        bool hit = false, curved = false;
        if (finished) {
            uint8_t st = inside ? ST_SOME : ST_NONE;
            if (p.ls.hang | p.rs.hang) st = ST_NONCONVERGED;  // the reference's hill climb never returned
            if (FULL && st == ST_SOME) {
                if constexpr (FULL) {
                    hit = true;
                    curved = is_curved(p.ls.kind) || is_curved(p.rs.kind);
                    float4* o = W.simplex + i * 8;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        o[k] = make_float4(s.v[k].x, s.v[k].y, s.v[k].z, 0.f);
                        o[4 + k] = make_float4(s.a[k].x, s.a[k].y, s.a[k].z, 0.f);
                    }
                }
            } else {
                store_contact(A.out_contact, i, FULL ? CB2_FULL_RESOLUTION : CB2_COLLISION_ONLY,
                              zero_contact());
                A.out_status[i] = st;
            }
            active = false;
        }
        if constexpr (FULL) {
            // warp-aggregated append of this round's hits to the first / last tier's work list
            const unsigned lane = threadIdx.x & 31u;
            const unsigned m1 = __ballot_sync(0xffffffffu, hit && !curved);
            const unsigned m3 = __ballot_sync(0xffffffffu, hit && curved);
            if (m1 | m3) {
                uint32_t b1 = 0, b3 = 0;
                if (lane == 0) {
                    if (m1) b1 = atomicAdd(&W.EC->n_list[0], __popc(m1));
                    if (m3) b3 = atomicAdd(&W.EC->n_list[EPA_TIERS - 1], __popc(m3));
                }
                b1 = __shfl_sync(0xffffffffu, b1, 0);
                b3 = __shfl_sync(0xffffffffu, b3, 0);
                const unsigned below = (1u << lane) - 1u;
                if (hit && !curved) W.list[0][b1 + __popc(m1 & below)] = (uint32_t)i;
                if (hit && curved) W.list[EPA_TIERS - 1][b3 + __popc(m3 & below)] = (uint32_t)i;
            }
        }
    }
}

// GJK::distance, gjk/mod.rs:271-321
__global__ void __launch_bounds__(128, 4) k_distance_p(narrow_args A) {
    const uint32_t max_it = A.settings.max_iterations;
    const float tol = A.settings.distance_tolerance;
    bool active = false, done = false;
    uint64_t i = 0;
    pair_in p;
    p.T = table_of(A.T);
    simplex<false> s;
    s.n = 0;
    f3 d0 = zero3();
    int state = 0;
    uint32_t it = 0;
    for (;;) {
        // ---- refill ---------------------------------------------------------------------------------
        uint64_t ni;
        if (refill_lane(A, !active && !done, ni)) {
            if (ni >= A.n) {
                done = true;
            } else {
                i = ni;
                if (const uint8_t bad = load_pair(A, i, A.l_t, A.r_t, p)) {
                    A.out_distance[i] = 0.0f;
                    A.out_status[i] = bad;
                } else {
                    d0 = seed_direction(p.lt, p.rt);
                    s.n = 0;
                    state = 0;
                    active = true;
                }
            }
        }
        if (__all_sync(0xffffffffu, done && !active)) break;
        // ---- pre: the direction of this round's support ---------------------------------------------
        bool go = active;
        uint8_t st = ST_NONE;
        float result = 0.0f;
        f3 dir = zero3();
        if (active) {
            if (state == 0) {
                dir = d0;
            } else if (state == 1) {
                dir = -d0;
            } else if (it >= max_it) {
                go = false;  // loop exhausted => None
            } else {
                uint8_t panic = 0;
                const f3 d = closest_point_to_origin(s, panic);
                if (panic) {
                    st = panic;
                    go = false;
                } else if (ulps_eq_zero(d)) {
                    go = false;  // None
                } else {
                    dir = -d;
                }
            }
        }
        // ---- one Minkowski support ---------------------------------------------------------------------
        f3 v = zero3(), a;
        if (go) minkowski<false>(p, dir, v, a);
        // ---- post ---------------------------------------------------------------------------------------
        bool finished = active && !go;
        if (go) {
            if (state < 2) {
                s.push(v, a);
                ++state;
                it = 0;
            } else {
                const float dp = dot(v, dir);
                const float dz = dot(s.v[0], dir);
                if (fsub(dp, dz) < tol) {
                    result = magnitude(dir);
                    st = ST_SOME;
                    finished = true;
                } else {
                    s.push(v, a);
                    ++it;
                }
            }
        }
        if (finished) {
            if (p.ls.hang | p.rs.hang) {
                st = ST_NONCONVERGED;
                result = 0.0f;
            }
            A.out_distance[i] = result;
            A.out_status[i] = st;
            active = false;
        }
    }
}


// ---- staged refill ------------------------------------------------------------------------------------
// A lane whose pair finished used to fetch and prepare its next pair on the spot (shape and transform
// loads, two quaternion inversions, for time of impact the end poses and the ray): ~400 instructions
// executed with the two or three lanes that happened to finish in that round — a quarter of
// k_time_of_impact_p's instructions at 2.6 of 32 lanes (ncu, cfg 5).  Now the WARP prepares 32 pairs
// at once into its slice of shared memory (every lane one pair: coalesced loads, full lanes) and a
// finished lane merely pops a prepared entry ([word][entry] layout: lanes that pop different entries
// hit different banks).  The queue hands out blocks of 32 pairs per warp.
constexpr int STG_PAIR = 42;  // i, status, ls (8), rs (8), lt (12), rt (12)
constexpr int STG_TOI = STG_PAIR + 11;  // + ray (3), end rotation (4), end scale, end displacement (3)
struct stage_io {
    uint32_t* p;  // word k of this entry at p[k * 32]
    CB2_DI void put(int k, uint32_t v) const { p[k * 32] = v; }
    CB2_DI void putf(int k, float v) const { p[k * 32] = __float_as_uint(v); }
    CB2_DI uint32_t get(int k) const { return p[k * 32]; }
    CB2_DI float getf(int k) const { return __uint_as_float(p[k * 32]); }
    CB2_DI void put_shape(int k, const shape& s) const {
        put(k, s.kind); putf(k + 1, s.h.x); putf(k + 2, s.h.y); putf(k + 3, s.h.z);
        put(k + 4, s.poly.voff); put(k + 5, s.poly.vcnt); put(k + 6, s.poly.cell_off); put(k + 7, s.poly.cfg);
    }
    CB2_DI shape get_shape(int k) const {
        shape s;
        s.kind = get(k); s.h = mk3(getf(k + 1), getf(k + 2), getf(k + 3));
        s.poly.voff = get(k + 4); s.poly.vcnt = get(k + 5); s.poly.cell_off = get(k + 6); s.poly.cfg = get(k + 7);
        s.hang = 0u;
        return s;
    }
    CB2_DI void put_xform(int k, const xform& t) const {
        putf(k, t.q.s); putf(k + 1, t.q.x); putf(k + 2, t.q.y); putf(k + 3, t.q.z);
        putf(k + 4, t.qi.s); putf(k + 5, t.qi.x); putf(k + 6, t.qi.y); putf(k + 7, t.qi.z);
        putf(k + 8, t.disp.x); putf(k + 9, t.disp.y); putf(k + 10, t.disp.z); putf(k + 11, t.scale);
    }
    CB2_DI xform get_xform(int k) const {
        xform t;
        t.q.s = getf(k); t.q.x = getf(k + 1); t.q.y = getf(k + 2); t.q.z = getf(k + 3);
        t.qi.s = getf(k + 4); t.qi.x = getf(k + 5); t.qi.y = getf(k + 6); t.qi.z = getf(k + 7);
        t.disp = mk3(getf(k + 8), getf(k + 9), getf(k + 10));
        t.scale = getf(k + 11);
        t.unit = (t.scale == 1.0f);
        return t;
    }
};
// prepare pair i (start poses lt / rt) into an entry; TOI: also the end poses' part
template <bool TOI>
CB2_NOINLINE_DEVICE void stage_prepare(const narrow_args& A, uint64_t i, uint32_t* entry) {
    const stage_io io{entry};
    pair_in p;
    const uint8_t bad = load_pair(A, i, A.l_t, A.r_t, p);
    io.put(0, (uint32_t)i);
    io.put(1, (uint32_t)bad);
    if (bad) return;
    io.put_shape(2, p.ls);
    io.put_shape(10, p.rs);
    io.put_xform(18, p.lt);
    io.put_xform(30, p.rt);
    if (TOI) {
        const xform lt1 = load_xform(A.l_t1, i);
        const xform rt1 = load_xform(A.r_t1, i);
        const f3 left_lin_vel = transform_point(lt1, zero3()) - transform_point(p.lt, zero3());
        const f3 right_lin_vel = transform_point(rt1, zero3()) - transform_point(p.rt, zero3());
        const f3 ray = right_lin_vel - left_lin_vel;
        io.putf(STG_PAIR, ray.x); io.putf(STG_PAIR + 1, ray.y); io.putf(STG_PAIR + 2, ray.z);
        io.putf(STG_PAIR + 3, rt1.q.s); io.putf(STG_PAIR + 4, rt1.q.x); io.putf(STG_PAIR + 5, rt1.q.y);
        io.putf(STG_PAIR + 6, rt1.q.z); io.putf(STG_PAIR + 7, rt1.scale);
        io.putf(STG_PAIR + 8, rt1.disp.x); io.putf(STG_PAIR + 9, rt1.disp.y); io.putf(STG_PAIR + 10, rt1.disp.z);
    }
}
// The warp's supply of prepared pairs.  next(): true for a lane that wanted a pair and got one
// (`entry` then points at it); lanes that wanted one and got none stay idle until the next round,
// or are `done` when the queue is exhausted.  Must be called by the whole warp.
template <bool TOI, int NW>
struct stage_supply {
    uint32_t* buf;       // this warp's NW x 32 words
    uint32_t head, count;
    bool exhausted;
    CB2_DI void init(uint32_t* b) { buf = b; head = count = 0u; exhausted = false; }
    CB2_DI bool next(const narrow_args& A, bool want, bool& done, stage_io& entry) {
        const unsigned need = __ballot_sync(0xffffffffu, want);
        if (!need) return false;
        const unsigned lane = threadIdx.x & 31u;
        if (head == count && !exhausted) {
            uint32_t base = 0;
            if (lane == 0) base = atomicAdd(A.queue, 32u);
            base = __shfl_sync(0xffffffffu, base, 0);
            const uint64_t left = (uint64_t)base < A.n ? A.n - base : 0ull;
            count = left < 32ull ? (uint32_t)left : 32u;
            head = 0u;
            exhausted = count == 0u;
            __syncwarp();  // every pop of the previous fill has been read
            if (lane < count) stage_prepare<TOI>(A, (uint64_t)base + lane, buf + lane);
            __syncwarp();
        }
        if (head == count) {  // (only when exhausted)
            if (want) done = true;
            return false;
        }
        const uint32_t e = head + (uint32_t)__popc(need & ((1u << lane) - 1u));
        const bool got = want && e < count;
        entry.p = buf + (got ? e : 0u);
        const uint32_t served = (uint32_t)__popc(need);
        head = head + served < count ? head + served : count;
        return got;
    }
};

// GJK::intersection_time_of_impact, gjk/mod.rs:163-256 (+ iter_cap -> ST_NONCONVERGED)
__global__ void __launch_bounds__(128, 4) k_time_of_impact_p(narrow_args A) {
    const float tol = A.settings.continuous_tolerance;
    bool active = false, done = false;
    uint64_t i = 0;
    pair_in p;
    p.T = table_of(A.T);
    simplex<false> s;
    s.n = 0;
    // right.end's rotation/scale and displacement (translation_interpolate, traits.rs:233-246)
    quat r1q = {1.f, 0.f, 0.f, 0.f};
    float r1scale = 1.f;
    f3 r1disp = zero3();
    f3 ray = zero3(), normal = zero3(), ray_origin = zero3(), v = zero3();
    float lambda = 0.0f;
    int state = 0;
    uint32_t it = 0;
    __shared__ uint32_t stage[4][STG_TOI * 32];
    stage_supply<true, STG_TOI> supply;
    supply.init(stage[threadIdx.x >> 5]);
    for (;;) {
        stage_io e{nullptr};
        if (supply.next(A, !active && !done, done, e)) {
            i = e.get(0);
            if (const uint8_t bad = (uint8_t)e.get(1)) {
                store_contact(A.out_contact, i, CB2_FULL_RESOLUTION, zero_contact());
                A.out_status[i] = bad;
            } else {
                p.ls = e.get_shape(2);
                p.rs = e.get_shape(10);
                p.lt = e.get_xform(18);
                p.rt = e.get_xform(30);
                ray = mk3(e.getf(STG_PAIR), e.getf(STG_PAIR + 1), e.getf(STG_PAIR + 2));
                r1q.s = e.getf(STG_PAIR + 3); r1q.x = e.getf(STG_PAIR + 4); r1q.y = e.getf(STG_PAIR + 5);
                r1q.z = e.getf(STG_PAIR + 6);
                r1scale = e.getf(STG_PAIR + 7);
                r1disp = mk3(e.getf(STG_PAIR + 8), e.getf(STG_PAIR + 9), e.getf(STG_PAIR + 10));
                lambda = 0.0f;
                normal = zero3();
                ray_origin = zero3();
                s.n = 0;
                it = 0;
                state = 0;
                active = true;
            }
        }
        if (__all_sync(0xffffffffu, done && !active)) break;
        // ---- pre ---------------------------------------------------------------------------------------
        bool go = active;
        uint8_t st = ST_NONE;
        contact_out c = zero_contact();
        f3 dir = zero3();
        if (active) {
            if (state == 0) {
                dir = -ray;
            } else {
                const float m2 = magnitude2(v);
                if (!(m2 > tol)) {  // the while loop ends
                    go = false;
                    if (m2 <= tol) {
                        xform t;
                        t.q = r1q;
                        t.scale = r1scale;
                        t.unit = (r1scale == 1.0f);
                        t.disp = lerp(p.rt.disp, r1disp, lambda);
                        c.normal = -normalize(normal);
                        c.depth = magnitude(v);
                        c.point = transform_point(t, ray_origin);
                        c.toi = lambda;
                        st = ST_SOME;
                    }
                } else if (it >= A.iter_cap) {
                    st = ST_NONCONVERGED;
                    go = false;
                } else {
                    ++it;
                    dir = -v;
                }
            }
        }
        // ---- one Minkowski support ---------------------------------------------------------------------
        f3 pv = zero3(), a;
        if (go) minkowski<false>(p, dir, pv, a);
        // ---- post ---------------------------------------------------------------------------------------
        bool finished = active && !go;
        if (go) {
            if (state == 0) {
                v = pv;
                state = 1;
            } else {
                const float vp = dot(v, pv);
                const float vr = dot(v, ray);
                if (vp > fmul(lambda, vr)) {
                    if (vr > 0.0f) {
                        lambda = fdiv(vp, vr);
                        if (lambda > 1.0f) {
                            finished = true;  // None
                        } else {
                            ray_origin = ray * lambda;
                            s.n = 0;
                            normal = -v;
                        }
                    } else {
                        finished = true;  // None
                    }
                }
                if (!finished) {
                    s.push(pv - ray_origin, a);
                    uint8_t panic = 0;
                    v = closest_point_to_origin(s, panic);
                    if (panic) {
                        st = panic;
                        finished = true;
                    }
                }
            }
        }
        if (finished) {
            if (p.ls.hang | p.rs.hang) st = ST_NONCONVERGED;
            if (st != ST_SOME) c = zero_contact();
            store_contact(A.out_contact, i, CB2_FULL_RESOLUTION, c);
            A.out_status[i] = st;
            active = false;
        }
    }
}

// ---- second chance for ST_SCRATCH_OVERFLOW pairs ---------------------------------------------------
// Collect the (few per million) pairs whose polytope outgrew the fast kernels' scratch and redo
// them, one pair per block leader, over a large global scratch.  Same arithmetic, same order.
__global__ void __launch_bounds__(256) k_collect_overflow(const uint8_t* __restrict__ status,
                                                         uint64_t n, uint32_t* __restrict__ list,
                                                         uint32_t* __restrict__ count) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (status[i] == ST_SCRATCH_OVERFLOW) {
        const uint32_t k = atomicAdd(count, 1u);
        if (k < RETRY_LIST_CAP) list[k] = (uint32_t)i;
    }
}

// Warp-cooperative EPA3::process with the reference's SERIAL edge-list semantics, for the pairs the
// tier kernels hand back (non-manifold polytopes: a directed edge occurs twice, so the parallel
// horizon of k_epa does not apply; or polytopes past 208 faces).  One warp per pair, the polytope in
// shared memory:
//   closest face, visibility of every face, Face::new of the new faces   -> all 32 lanes
//   the swap_remove sweep with remove_or_add_edge (epa3d.rs:147-175, 212-219) -> lane 0, literally,
//     but only over the VISIBLE faces (the visibility bit mask lets it jump over the others: during
//     the sweep every position past the cursor still holds its original face)
// A serial single-thread version over global memory (epa_process<global_store>) took 2.3 ms for one
// 100-iteration / 246-face pair — 12 % of a 4 M-pair step on the GPU that drew that pair; it stays
// as the fallback for polytopes that outgrow the shared-memory store.
constexpr int COOP_FACES = 1536;
constexpr int COOP_EDGES = 2048;
struct coop_store {   // shared memory: what nearly every second-chance pair fits
    static constexpr int FCAP = COOP_FACES, ECAP = COOP_EDGES, VCAP = EPA_MAX_VERTS;
    float4 fnd[COOP_FACES];
    uint32_t fidx[COOP_FACES];
    float4 pv[EPA_MAX_VERTS];
    float4 pa[EPA_MAX_VERTS];
    uint16_t edges[COOP_EDGES];
    uint32_t vmask[COOP_FACES / 32];
};
// ... and the same arrays in global memory for the polytopes that do not (the 32 lanes stride over
// the faces, so the accesses coalesce): RETRY_FACES faces, RETRY_EDGES horizon edges, RETRY_VERTS
// vertices — the byte-sized vertex indices of a face record allow 255, i.e. max_iterations <= 250
struct coop_gstore {
    static constexpr int FCAP = RETRY_FACES, ECAP = RETRY_EDGES, VCAP = RETRY_VERTS;
    float4* fnd;
    uint32_t* fidx;
    float4* pv;
    float4* pa;
    uint16_t* edges;
    uint32_t* vmask;
};
template <class ST>
CB2_DI f3 cs_pv(const ST& P, uint32_t i) {
    const float4 v = P.pv[i];
    return mk3(v.x, v.y, v.z);
}
template <class ST>
CB2_DI void cs_face_new(ST& P, int slot, uint32_t a, uint32_t b, uint32_t c) {
    const f3 va = cs_pv(P, a);
    const f3 ab = cs_pv(P, b) - va;
    const f3 ac = cs_pv(P, c) - va;
    const f3 n = normalize(cross(ab, ac));
    P.fnd[slot] = make_float4(n.x, n.y, n.z, dot(n, va));
    P.fidx[slot] = a | (b << 8) | (c << 16);
}
// every lane runs this with identical arguments; control flow is warp-uniform
template <class ST>
__device__ __noinline__ uint8_t epa_process_coop(ST& P, const pair_in& p, const simplex<true>& s,
                                                 float tol, uint32_t max_iterations, contact_out& out) {
    const int lane = threadIdx.x & 31;
    if (lane < 4) {
        P.pv[lane] = make_float4(s.v[lane].x, s.v[lane].y, s.v[lane].z, 0.f);
        P.pa[lane] = make_float4(s.a[lane].x, s.a[lane].y, s.a[lane].z, 0.f);
    }
    __syncwarp();
    if (lane < 4) {  // Face::new (3,2,1) (3,1,0) (3,0,2) (2,0,1), epa3d.rs:201-208
        const uint32_t tri = lane == 0 ? 0x010203u : (lane == 1 ? 0x000103u : (lane == 2 ? 0x020003u : 0x010002u));
        cs_face_new(P, lane, tri & 255u, (tri >> 8) & 255u, tri >> 16);
    }
    __syncwarp();
    int nv = 4, nf = 4;
    uint32_t it = 1;
    for (;;) {
        if (nf == 0) return ST_PANIC_EPA_EMPTY;
        // closest_face_to_origin (:137-145): first strictly smallest distance; a NaN at index 0 sticks
        float bd = INFINITY;
        uint32_t bi = 0xFFFFFFFFu;
        for (int f = lane; f < nf; f += 32) {
            const float d = P.fnd[f].w;
            if (d < bd) {
                bd = d;
                bi = (uint32_t)f;
            }
        }
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) {
            const float ob = __shfl_xor_sync(0xffffffffu, bd, off);
            const uint32_t oi = __shfl_xor_sync(0xffffffffu, bi, off);
            if (ob < bd || (ob == bd && oi < bi)) {
                bd = ob;
                bi = oi;
            }
        }
        const float d0 = P.fnd[0].w;
        const int best = (d0 != d0 || bi == 0xFFFFFFFFu) ? 0 : (int)bi;
        const float4 fb = P.fnd[best];
        const f3 n = mk3(fb.x, fb.y, fb.z);
        f3 sv, sa;
        minkowski<true>(p, n, sv, sa);
        if (fsub(dot(sv, n), fb.w) < tol || it >= max_iterations) {
            const uint32_t idx = P.fidx[best];
            const uint32_t i0 = idx & 255u, i1 = (idx >> 8) & 255u, i2 = (idx >> 16) & 255u;
            float u, v, w;
            barycentric_vector(n * fb.w, cs_pv(P, i0), cs_pv(P, i1), cs_pv(P, i2), u, v, w);
            const float4 a0 = P.pa[i0], a1 = P.pa[i1], a2 = P.pa[i2];
            out.normal = n;
            out.depth = fb.w;
            out.point = (mk3(a0.x, a0.y, a0.z) * u + mk3(a1.x, a1.y, a1.z) * v) + mk3(a2.x, a2.y, a2.z) * w;
            out.toi = 0.0f;
            return ST_SOME;
        }
        // visibility of every face -> bit mask
        unsigned any_visible = 0u;
        for (int f0 = 0; f0 < nf; f0 += 32) {
            const int f = f0 + lane;
            bool vis = false;
            if (f < nf) {
                const float4 fn = P.fnd[f];
                vis = dot(mk3(fn.x, fn.y, fn.z), sv - cs_pv(P, P.fidx[f] & 255u)) > 0.0f;
            }
            const unsigned bits = __ballot_sync(0xffffffffu, vis);
            any_visible |= bits;
            if (lane == 0) P.vmask[f0 >> 5] = bits;
        }
        if (!any_visible) {
            // Fixed point: with no visible face Polytope::add removes nothing and adds nothing (the
            // new vertex is referenced by no face), so every later iteration finds the same closest
            // face, the same support point and fails the same tolerance test until `it` reaches
            // max_iterations (epa3d.rs:46-64) — and then returns the contact of this very face.
            // (NaN normals end here: no comparison with a NaN is ever true.)
            max_iterations = it;
            continue;
        }
        __syncwarp();
        // Polytope::add's sweep.  Every lane walks it with the same scalars (uniform control flow);
        // the linear search of remove_or_add_edge (:212-219) is the part spread over the lanes.
        int ne = 0;
        int overflow = 0;
        {
            const int nf0 = nf;
            auto next_visible = [&](int from) {  // first visible ORIGINAL position in [from, nf), else nf
                while (from < nf) {
                    const uint32_t wbits = P.vmask[from >> 5] >> (from & 31);
                    if (wbits) {
                        const int q = from + __ffs((int)wbits) - 1;
                        return q < nf ? q : nf;
                    }
                    from = (from | 31) + 1;
                }
                return nf;
            };
            int f = next_visible(0);
            int orig = f;
            while (f < nf && !overflow) {
                if (orig < nf0 && ((P.vmask[orig >> 5] >> (orig & 31)) & 1u)) {
                    const uint32_t idx = P.fidx[f];
                    --nf;
                    __syncwarp();
                    if (lane == 0) {
                        P.fnd[f] = P.fnd[nf];  // swap_remove: the tail face lands in the hole
                        P.fidx[f] = P.fidx[nf];
                    }
                    orig = nf;  // positions past f were never written: the tail is original
                    const uint32_t v0 = idx & 255u, v1 = (idx >> 8) & 255u, v2 = (idx >> 16) & 255u;
#pragma unroll
                    for (int e = 0; e < 3; ++e) {
                        // remove_or_add_edge: drop the FIRST stored reverse edge, else push
                        const uint32_t ea = e == 0 ? v0 : (e == 1 ? v1 : v2);
                        const uint32_t eb = e == 0 ? v1 : (e == 1 ? v2 : v0);
                        const uint16_t rev = (uint16_t)(eb | (ea << 8));
                        int k = ne;
                        for (int k0 = 0; k0 < ne; k0 += 32) {
                            const int q = k0 + lane;
                            const unsigned hit = __ballot_sync(0xffffffffu, q < ne && P.edges[q] == rev);
                            if (hit) {
                                k = k0 + __ffs((int)hit) - 1;
                                break;
                            }
                        }
                        if (k < ne) {
                            // erase position k, keeping the order: read, then write one slot down
                            for (int m0 = k + 1; m0 < ne; m0 += 32) {
                                const int m = m0 + lane;
                                uint16_t val = 0;
                                if (m < ne) val = P.edges[m];
                                __syncwarp();
                                if (m < ne) P.edges[m - 1] = val;
                                __syncwarp();
                            }
                            --ne;
                        } else if (ne >= ST::ECAP) {
                            overflow = 1;
                        } else {
                            if (lane == 0) P.edges[ne] = (uint16_t)(ea | (eb << 8));
                            ++ne;
                        }
                        __syncwarp();
                    }
                } else {
                    f = next_visible(f + 1);
                    orig = f;
                }
            }
        }
        __syncwarp();
        if (overflow || nv >= ST::VCAP || nf + ne > ST::FCAP) return ST_SCRATCH_OVERFLOW;
        const uint32_t nn = (uint32_t)nv;
        if (lane == 0) {
            P.pv[nn] = make_float4(sv.x, sv.y, sv.z, 0.f);
            P.pa[nn] = make_float4(sa.x, sa.y, sa.z, 0.f);
        }
        ++nv;
        __syncwarp();
        for (int k = lane; k < ne; k += 32) {
            const uint32_t e = P.edges[k];
            cs_face_new(P, nf + k, nn, e & 255u, e >> 8);
        }
        nf += ne;
        ++it;
        __syncwarp();
    }
}

__global__ void __launch_bounds__(32) k_retry_overflow(narrow_args A, const uint32_t* __restrict__ list,
                                                      const uint32_t* __restrict__ count,
                                                      uint32_t* __restrict__ cursor,
                                                      char* __restrict__ scratch) {
    __shared__ coop_store CS;
    const int lane = threadIdx.x;
    uint32_t total = *count;
    if (total > RETRY_LIST_CAP) total = RETRY_LIST_CAP;
    coop_gstore P;
    char* base = scratch + (size_t)blockIdx.x * RETRY_BYTES_PER_BLOCK;
    P.fnd = reinterpret_cast<float4*>(base);
    P.pv = P.fnd + RETRY_FACES;
    P.pa = P.pv + RETRY_VERTS;
    P.fidx = reinterpret_cast<uint32_t*>(P.pa + RETRY_VERTS);
    P.vmask = P.fidx + RETRY_FACES;
    P.edges = reinterpret_cast<uint16_t*>(P.vmask + RETRY_FACES / 32);
    for (;;) {
        uint32_t k = 0;
        if (lane == 0) k = atomicAdd(cursor, 1u);
        k = __shfl_sync(0xffffffffu, k, 0);
        if (k >= total) return;
        const uint64_t i = list[k];
        pair_in p;
        contact_out c = zero_contact();
        uint8_t st = ST_NONE;
        // every lane loads the pair and runs GJK::intersect redundantly (uniform: same data, same path)
        if (const uint8_t bad = load_pair(A, i, A.l_t, A.r_t, p)) {
            st = bad;
        } else {
            simplex<true> s;
            bool inside;
            if (A.seed_simplex) {  // get_contact_manifold: the caller's simplex, not a new GJK run
                const float4* sx = A.seed_simplex + i * 8;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float4 v = sx[k], a = sx[4 + k];
                    s.set(k, mk3(v.x, v.y, v.z), mk3(a.x, a.y, a.z));
                }
                s.n = 4;
                inside = true;
            } else {
                inside = gjk_intersect(p, A.settings.max_iterations, s);
            }
            if (inside) {
                st = epa_process_coop(CS, p, s, A.settings.epa_tolerance, A.settings.max_iterations, c);
                if (st == ST_SCRATCH_OVERFLOW)  // past the shared-memory store: the same, over global scratch
                    st = epa_process_coop(P, p, s, A.settings.epa_tolerance, A.settings.max_iterations, c);
                if (st != ST_SOME) c = zero_contact();
            }
            if (p.ls.hang | p.rs.hang) {
                st = ST_NONCONVERGED;
                c = zero_contact();
            }
        }
        if (lane == 0) {
            store_contact(A.out_contact, i, CB2_FULL_RESOLUTION, c);
            A.out_status[i] = st;
        }
        __syncwarp();
    }
}

// ---- world AABBs: shape.compute_bound().transform_volume(&t) ------------------------------------
// ComputeBound: sphere.rs:41-51, cuboid.rs:82-92, polyhedron.rs:113-115 (fold from Aabb3::zero(),
// so the bound always contains the origin); Aabb3::transform: aabb3.rs:39-50,90-100.
CB2_DI float rmin(float l, float r) { return (l > r) ? r : l; }  // aabb/mod.rs:21-26
CB2_DI float rmax(float l, float r) { return (l < r) ? r : l; }  // aabb/mod.rs:28-33

__global__ void __launch_bounds__(256) k_world_aabbs(aabb_args A) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= A.n) return;
    const shape s = load_shape(table_of(A.T), __ldg(A.shape + i));
    const xform t = load_xform(A.t, i);
    f3 mn, mx;
    // Aabb3::new(p1, p2) orders each component with the reference's min/max (aabb/mod.rs:21-33)
    auto ordered = [&](f3 a, f3 b) {
        mn = mk3(rmin(a.x, b.x), rmin(a.y, b.y), rmin(a.z, b.z));
        mx = mk3(rmax(a.x, b.x), rmax(a.y, b.y), rmax(a.z, b.z));
    };
    if (s.kind == K_SPHERE) {
        const float r = s.h.x;
        ordered(mk3(-r, -r, -r), mk3(r, r, r));
    } else if (s.kind == K_CUBOID) {
        ordered(-s.h, s.h);
    } else if (s.kind == K_PARTICLE) {  // Aabb3::zero(), primitive3.rs:120
        mn = zero3();
        mx = zero3();
    } else if (s.kind == K_QUAD) {  // quad.rs:80-90
        ordered(mk3(-s.h.x, -s.h.y, 0.0f), mk3(s.h.x, s.h.y, 0.0f));
    } else if (s.kind == K_CYLINDER) {  // cylinder.rs:81-91
        ordered(mk3(-s.h.y, -s.h.x, -s.h.y), mk3(s.h.y, s.h.x, s.h.y));
    } else if (s.kind == K_CAPSULE) {  // capsule.rs:64-74
        ordered(mk3(-s.h.y, fsub(-s.h.x, s.h.y), -s.h.y), mk3(s.h.y, fadd(s.h.x, s.h.y), s.h.y));
    } else {
        mn = zero3();
        mx = zero3();
        for (uint32_t k = 0; k < s.poly.vcnt; ++k) {
            const float4 v = __ldg(A.T.verts + s.poly.voff + k);
            // grow(): Aabb::new(min(self.min,p), max(self.max,p)) then Aabb3::new re-sorts
            const f3 a = mk3(rmin(mn.x, v.x), rmin(mn.y, v.y), rmin(mn.z, v.z));
            const f3 b = mk3(rmax(mx.x, v.x), rmax(mx.y, v.y), rmax(mx.z, v.z));
            mn = mk3(rmin(a.x, b.x), rmin(a.y, b.y), rmin(a.z, b.z));
            mx = mk3(rmax(a.x, b.x), rmax(a.y, b.y), rmax(a.z, b.z));
        }
    }
    f3 omn, omx;
#pragma unroll
    for (int c = 0; c < 8; ++c) {  // to_corners order, aabb3.rs:39-50
        const f3 corner = mk3((c & 1) ? mx.x : mn.x, (c & 2) ? mx.y : mn.y, (c & 4) ? mx.z : mn.z);
        const f3 q = transform_point(t, corner);
        if (c == 0) {
            omn = q;
            omx = q;
        } else {
            const f3 a = mk3(rmin(omn.x, q.x), rmin(omn.y, q.y), rmin(omn.z, q.z));
            const f3 b = mk3(rmax(omx.x, q.x), rmax(omx.y, q.y), rmax(omx.z, q.z));
            omn = mk3(rmin(a.x, b.x), rmin(a.y, b.y), rmin(a.z, b.z));
            omx = mk3(rmax(a.x, b.x), rmax(a.y, b.y), rmax(a.z, b.z));
        }
    }
    float* o = reinterpret_cast<float*>(A.out + i);
    o[0] = omn.x; o[1] = omn.y; o[2] = omn.z;
    o[3] = omx.x; o[4] = omx.y; o[5] = omx.z;
}

// ---- launchers -------------------------------------------------------------------------------------
static inline unsigned grid_for(uint64_t n, unsigned block) { return (unsigned)((n + block - 1) / block); }

cudaError_t launch_intersection(const narrow_args& A, bool full, cudaStream_t st) {
    if (A.n == 0) return cudaSuccess;
    if (full) k_intersection<true><<<grid_for(A.n, 128), 128, 0, st>>>(A);
    else k_intersection<false><<<grid_for(A.n, 128), 128, 0, st>>>(A);
    return cudaGetLastError();
}
template <class K>
static cudaError_t launch_refill2(K kern, const narrow_args& A, const narrow_ws& W, int sm_count,
                                  cudaStream_t st) {
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 128, 0) != cudaSuccess || per_sm < 1)
        per_sm = 1;
    cudaError_t e = cudaMemsetAsync(A.queue, 0, sizeof(uint32_t), st);
    if (e != cudaSuccess) return e;
    uint64_t blocks = (uint64_t)per_sm * sm_count;
    const uint64_t needed = (A.n + 127) / 128;
    if (blocks > needed) blocks = needed;
    kern<<<(unsigned)blocks, 128, 0, st>>>(A, W);
    return cudaGetLastError();
}
static bool no_refill();
cudaError_t launch_gjk_hits(const narrow_args& A, bool full, const narrow_ws& W, int sm_count,
                            cudaStream_t st) {
    if (A.n == 0) return cudaSuccess;
    if (no_refill() || !A.queue) {
        if (full) k_gjk_hits<true><<<grid_for(A.n, 128), 128, 0, st>>>(A, W);
        else k_gjk_hits<false><<<grid_for(A.n, 128), 128, 0, st>>>(A, W);
        return cudaGetLastError();
    }
    if (full) return launch_refill2(k_gjk_hits_p<true>, A, W, sm_count, st);
    return launch_refill2(k_gjk_hits_p<false>, A, W, sm_count, st);
}
cudaError_t launch_gjk_intersect(const narrow_args& A, cudaStream_t st) {
    if (A.n == 0) return cudaSuccess;
    k_gjk_intersect<<<grid_for(A.n, 128), 128, 0, st>>>(A);
    return cudaGetLastError();
}
cudaError_t launch_manifold_seed(const narrow_args& A, const narrow_ws& W, cudaStream_t st) {
    if (A.n == 0) return cudaSuccess;
    k_manifold_seed<<<grid_for(A.n, 128), 128, 0, st>>>(A, W);
    return cudaGetLastError();
}
size_t retry_scratch_bytes() { return (size_t)RETRY_BLOCKS * RETRY_BYTES_PER_BLOCK; }

cudaError_t launch_overflow_retry(const narrow_args& A, const retry_ws& R, cudaStream_t st) {
    if (A.n == 0) return cudaSuccess;
    cudaError_t e = cudaMemsetAsync(R.count, 0, 2 * sizeof(uint32_t), st);
    if (e != cudaSuccess) return e;
    k_collect_overflow<<<grid_for(A.n, 256), 256, 0, st>>>(A.out_status, A.n, R.list, R.count);
    k_retry_overflow<<<RETRY_BLOCKS, 32, 0, st>>>(A, R.list, R.count, R.count + 1, R.scratch);
    return cudaGetLastError();
}

// persistent launch: as many blocks as fit, the pairs come from the device queue A.queue (zeroed here)
template <class K>
static cudaError_t launch_refill(K kern, const narrow_args& A, int sm_count, cudaStream_t st) {
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 128, 0) != cudaSuccess || per_sm < 1)
        per_sm = 1;
    cudaError_t e = cudaMemsetAsync(A.queue, 0, sizeof(uint32_t), st);
    if (e != cudaSuccess) return e;
    uint64_t blocks = (uint64_t)per_sm * sm_count;
    const uint64_t needed = (A.n + 127) / 128;
    if (blocks > needed) blocks = needed;
    kern<<<(unsigned)blocks, 128, 0, st>>>(A);
    return cudaGetLastError();
}
static bool no_refill() {
    static const bool v = [] { const char* e = std::getenv("CB2_NO_REFILL"); return e && e[0] == '1'; }();
    return v;
}
cudaError_t launch_distance(const narrow_args& A, int sm_count, cudaStream_t st) {
    if (A.n == 0) return cudaSuccess;
    if (no_refill() || !A.queue || A.n > 0xFFFFFF00ull) {
        k_distance<<<grid_for(A.n, 128), 128, 0, st>>>(A);
        return cudaGetLastError();
    }
    return launch_refill(k_distance_p, A, sm_count, st);
}
cudaError_t launch_time_of_impact(const narrow_args& A, int sm_count, cudaStream_t st) {
    if (A.n == 0) return cudaSuccess;
    // (the staged refill claims 32 pairs per warp and overshoots the 32-bit queue by at most that per warp)
    if (no_refill() || !A.queue || A.n > 0xFFF00000ull) {
        k_time_of_impact<<<grid_for(A.n, 128), 128, 0, st>>>(A);
        return cudaGetLastError();
    }
    return launch_refill(k_time_of_impact_p, A, sm_count, st);
}
cudaError_t launch_world_aabbs(const aabb_args& A, cudaStream_t st) {
    if (A.n == 0) return cudaSuccess;
    k_world_aabbs<<<grid_for(A.n, 256), 256, 0, st>>>(A);
    return cudaGetLastError();
}

}  // namespace cb2
