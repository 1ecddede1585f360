// cb2_complex.cu — composite shapes: GJK3::intersection_complex / distance_complex /
// intersection_complex_time_of_impact (gjk/mod.rs:412-588).
//
// The reference loops over (left part x right part), concatenates the model transform with each
// part's local transform (Transform::concat), calls the single-primitive routine and reduces:
// first hit (CollisionOnly), max penetration depth (FullResolution), min distance, min time of
// impact.  Here the double loop is EXPANDED into one flat batch of sub-pairs that the ordinary
// narrow-phase kernels process (k_cx_expand writes their shapes and concatenated transforms),
// and k_cx_reduce replays the reference's loop order over the sub-pair results — including where
// the reference would have returned early or panicked, so later sub-pairs cannot leak into the
// answer.
#include <cub/device/device_scan.cuh>

#include "cb2_gjk.cuh"
#include "cb2_kernels.h"

namespace cb2 {

// Quaternion * Quaternion, cgmath 0.18 quaternion.rs (left-to-right evaluation)
CB2_DI quat quat_mul(const quat& a, const quat& b) {
    quat r;
    r.s = fsub(fsub(fsub(fmul(a.s, b.s), fmul(a.x, b.x)), fmul(a.y, b.y)), fmul(a.z, b.z));
    r.x = fsub(fadd(fadd(fmul(a.s, b.x), fmul(a.x, b.s)), fmul(a.y, b.z)), fmul(a.z, b.y));
    r.y = fsub(fadd(fadd(fmul(a.s, b.y), fmul(a.y, b.s)), fmul(a.z, b.x)), fmul(a.x, b.z));
    r.z = fsub(fadd(fadd(fmul(a.s, b.z), fmul(a.z, b.s)), fmul(a.x, b.y)), fmul(a.y, b.x));
    return r;
}

struct raw_xf {
    float scale;
    quat q;
    f3 d;
};
CB2_DI raw_xf load_raw(const float4* __restrict__ t, uint64_t i) {
    const float4 a = __ldg(t + 2 * i), b = __ldg(t + 2 * i + 1);
    raw_xf r;
    r.scale = a.x;
    r.q.s = a.y; r.q.x = a.z; r.q.y = a.w; r.q.z = b.x;
    r.d = mk3(b.y, b.z, b.w);
    return r;
}
// Decomposed::concat (cgmath 0.18 transform.rs): {s*os, rot*orot, rot.rotate(odisp*s) + disp}
CB2_DI void store_concat(float4* __restrict__ out, uint64_t i, const raw_xf& a, const raw_xf& b) {
    const float sc = fmul(a.scale, b.scale);
    const quat q = quat_mul(a.q, b.q);
    const f3 d = rotate(a.q, b.d * a.scale) + a.d;
    out[2 * i] = make_float4(sc, q.s, q.x, q.y);
    out[2 * i + 1] = make_float4(q.z, d.x, d.y, d.z);
}

// ---- the 2-D instantiation: Decomposed<Vector2, Basis2> (cb2_xform2: scale, c0.xy, c1.xy | disp.xy) ----
struct raw_xf2 {
    float scale, c0x, c0y, c1x, c1y, dx, dy;
};
CB2_DI raw_xf2 load_raw2(const float4* __restrict__ t, uint64_t i) {
    const float4 a = __ldg(t + 2 * i), b = __ldg(t + 2 * i + 1);
    raw_xf2 r;
    r.scale = a.x; r.c0x = a.y; r.c0y = a.z; r.c1x = a.w;
    r.c1y = b.x; r.dx = b.y; r.dy = b.z;
    return r;
}
// Matrix2 * Vector2 = (row0 . v, row1 . v) with dot(a, b) = a.x*b.x + a.y*b.y
CB2_DI void mat2_mul(const raw_xf2& m, float vx, float vy, float& ox, float& oy) {
    ox = fadd(fmul(m.c0x, vx), fmul(m.c1x, vy));
    oy = fadd(fmul(m.c0y, vx), fmul(m.c1y, vy));
}
// concat: {s*os, rot*orot (Matrix2 * Matrix2, column by column), rot.rotate(odisp*s) + disp}
CB2_DI void store_concat2(float4* __restrict__ out, uint64_t i, const raw_xf2& a, const raw_xf2& b) {
    float c0x, c0y, c1x, c1y, dx, dy;
    mat2_mul(a, b.c0x, b.c0y, c0x, c0y);
    mat2_mul(a, b.c1x, b.c1y, c1x, c1y);
    mat2_mul(a, fmul(b.dx, a.scale), fmul(b.dy, a.scale), dx, dy);
    out[2 * i] = make_float4(fmul(a.scale, b.scale), c0x, c0y, c1x);
    out[2 * i + 1] = make_float4(c1y, fadd(dx, a.dx), fadd(dy, a.dy), 0.0f);
}

__global__ void __launch_bounds__(256) k_cx_count(complex_args X) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= X.n) return;
    const uint32_t lc = __ldg(X.l_comp + i), rc = __ldg(X.r_comp + i);
    const uint64_t nl = __ldg(X.comp_off + lc + 1) - __ldg(X.comp_off + lc);
    const uint64_t nr = __ldg(X.comp_off + rc + 1) - __ldg(X.comp_off + rc);
    X.sub_count[i] = nl * nr;
}

template <bool D2>
__global__ void __launch_bounds__(128) k_cx_expand(complex_args X) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= X.n) return;
    const uint32_t lc = __ldg(X.l_comp + i), rc = __ldg(X.r_comp + i);
    const uint32_t lo = __ldg(X.comp_off + lc), le = __ldg(X.comp_off + lc + 1);
    const uint32_t ro = __ldg(X.comp_off + rc), re = __ldg(X.comp_off + rc + 1);
    uint64_t s = X.sub_off[i];
    if (D2) {
        const raw_xf2 lt0 = load_raw2(X.l_t0, i), rt0 = load_raw2(X.r_t0, i);
        raw_xf2 lt1 = lt0, rt1 = rt0;
        if (X.l_t1) {
            lt1 = load_raw2(X.l_t1, i);
            rt1 = load_raw2(X.r_t1, i);
        }
        for (uint32_t a = lo; a < le; ++a) {
            const raw_xf2 la = load_raw2(X.comp_local, a);
            for (uint32_t b = ro; b < re; ++b, ++s) {
                const raw_xf2 rb = load_raw2(X.comp_local, b);
                X.sub_l_shape[s] = __ldg(X.comp_shape + a);
                X.sub_r_shape[s] = __ldg(X.comp_shape + b);
                store_concat2(X.sub_l_t0, s, lt0, la);
                store_concat2(X.sub_r_t0, s, rt0, rb);
                if (X.l_t1) {
                    store_concat2(X.sub_l_t1, s, lt1, la);
                    store_concat2(X.sub_r_t1, s, rt1, rb);
                }
            }
        }
        return;
    }
    const raw_xf lt0 = load_raw(X.l_t0, i), rt0 = load_raw(X.r_t0, i);
    raw_xf lt1 = lt0, rt1 = rt0;
    if (X.l_t1) {
        lt1 = load_raw(X.l_t1, i);
        rt1 = load_raw(X.r_t1, i);
    }
    for (uint32_t a = lo; a < le; ++a) {      // same loop order as the reference
        const raw_xf la = load_raw(X.comp_local, a);
        for (uint32_t b = ro; b < re; ++b, ++s) {
            const raw_xf rb = load_raw(X.comp_local, b);
            X.sub_l_shape[s] = __ldg(X.comp_shape + a);
            X.sub_r_shape[s] = __ldg(X.comp_shape + b);
            store_concat(X.sub_l_t0, s, lt0, la);
            store_concat(X.sub_r_t0, s, rt0, rb);
            if (X.l_t1) {
                store_concat(X.sub_l_t1, s, lt1, la);
                store_concat(X.sub_r_t1, s, rt1, rb);
            }
        }
    }
}

CB2_DI int partial_cmp_f(float a, float b) {  // -1 Less, 0 Equal, 1 Greater, 2 None
    if (a < b) return -1;
    if (a > b) return 1;
    if (a == b) return 0;
    return 2;
}

template <bool D2>
__global__ void __launch_bounds__(128) k_cx_reduce(complex_args X, uint32_t op, uint32_t strategy) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= X.n) return;
    const uint64_t s0 = X.sub_off[i], s1 = s0 + X.sub_count[i];
    uint8_t st = ST_NONE;
    if (op == 1) {  // distance_complex, gjk/mod.rs:477-516
        bool have = false;
        float m = 0.0f;
        for (uint64_t s = s0; s < s1; ++s) {
            const uint8_t ss = X.sub_status[s];
            if (ss != ST_SOME) {  // None => None (colliding); panic => panic; either way stop
                st = ss;
                have = false;
                break;
            }
            const float d = X.sub_distance[s];
            m = have ? fminf(d, m) : d;  // f32::min
            have = true;
            st = ST_SOME;
        }
        X.out_distance[i] = have ? m : 0.0f;
        X.out_status[i] = st;
        return;
    }
    const bool toi = op == 2;
    bool have = false;
    contact_out best = zero_contact();
    uint32_t best_strategy = strategy;
    for (uint64_t s = s0; s < s1; ++s) {
        const uint8_t ss = X.sub_status[s];
        if (ss == ST_NONE) continue;
        if (ss != ST_SOME) {  // the reference panics (or never returns) right here
            st = ss;
            have = false;
            break;
        }
        contact_out cur;
        if (D2) {  // cb2_contact2: strategy, normal.xy, depth, point.xy, toi
            const float* c = reinterpret_cast<const float*>(X.sub_contact) + 7 * s;
            cur.normal = mk3(c[1], c[2], 0.0f);
            cur.depth = c[3];
            cur.point = mk3(c[4], c[5], 0.0f);
            cur.toi = c[6];
        } else {
            const float* c = reinterpret_cast<const float*>(X.sub_contact) + 9 * s;
            cur.normal = mk3(c[1], c[2], c[3]);
            cur.depth = c[4];
            cur.point = mk3(c[5], c[6], c[7]);
            cur.toi = c[8];
        }
        if (strategy == CB2_COLLISION_ONLY) {  // return Some(contact) at the first hit
            best = cur;
            have = true;
            st = ST_SOME;
            break;
        }
        if (!have) {
            best = cur;
            have = true;
            st = ST_SOME;
        } else if (!toi) {
            // Iterator::max_by(partial_cmp(..).unwrap()): Greater keeps the current one
            const int o = partial_cmp_f(best.depth, cur.depth);
            if (o == 2) {
                st = ST_PANIC_CMP_NAN;
                have = false;
                break;
            }
            if (o != 1) best = cur;
        } else {
            // Iterator::min_by(partial_cmp(..).unwrap_or(Equal)): only Greater replaces it
            if (partial_cmp_f(best.toi, cur.toi) == 1) best = cur;
        }
    }
    if (!have) best = zero_contact();
    if (D2) {
        float* o = reinterpret_cast<float*>(X.out_contact) + 7 * i;
        o[0] = __uint_as_float(best_strategy);
        o[1] = best.normal.x;
        o[2] = best.normal.y;
        o[3] = best.depth;
        o[4] = best.point.x;
        o[5] = best.point.y;
        o[6] = best.toi;
    } else {
        store_contact(reinterpret_cast<cb2_contact*>(X.out_contact), i, best_strategy, best);
    }
    X.out_status[i] = st;
}

// exclusive scan of the per-pair sub-pair counts; call with tmp == nullptr to size tmp
cudaError_t cx_scan(void* tmp, size_t* tmp_bytes, const uint64_t* in, uint64_t* out, uint64_t n,
                    cudaStream_t st) {
    return cub::DeviceScan::ExclusiveSum(tmp, *tmp_bytes, in, out, (int64_t)n, st);
}
cudaError_t launch_cx_count(const complex_args& X, cudaStream_t st) {
    k_cx_count<<<(unsigned)((X.n + 255) / 256), 256, 0, st>>>(X);
    return cudaGetLastError();
}
cudaError_t launch_cx_expand(const complex_args& X, cudaStream_t st) {
    if (X.dim2) k_cx_expand<true><<<(unsigned)((X.n + 127) / 128), 128, 0, st>>>(X);
    else k_cx_expand<false><<<(unsigned)((X.n + 127) / 128), 128, 0, st>>>(X);
    return cudaGetLastError();
}
cudaError_t launch_cx_reduce(const complex_args& X, uint32_t op, uint32_t strategy, cudaStream_t st) {
    if (X.dim2) k_cx_reduce<true><<<(unsigned)((X.n + 127) / 128), 128, 0, st>>>(X, op, strategy);
    else k_cx_reduce<false><<<(unsigned)((X.n + 127) / 128), 128, 0, st>>>(X, op, strategy);
    return cudaGetLastError();
}

}  // namespace cb2
