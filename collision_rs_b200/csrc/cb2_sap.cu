// cb2_sap.cu — SweepAndPrune3::find_collider_pairs on the device
// (algorithm/broad_phase/sweep_prune.rs:76-136).
//
// Reference semantics (finite extents): stable sort by (min[axis], max[axis]) under
// partial_cmp (so -0.0 == +0.0); every pair (a<b in sorted positions) whose boxes overlap
// STRICTLY on all three axes (aabb3.rs:246-253); Vec order = b ascending, then a ascending;
// sweep_axis <- 0 afterwards (Variance3 accumulates nothing, sweep_prune.rs:254-277).
// The reference's `active.retain(max_a >= min_b)` prune is implied by the strict test.
//
// Device pipeline:
//   k_sap_keys    64-bit radix keys ord(min[axis]) : ord(max[axis]) (−0 canonicalised)
//   radix sort    stable LSD sort of (key, index)            -> perm
//   k_sap_gather  sorted SoA: U=(min_u,max_u) V=(min_v,max_v) S=(min_s,max_s); upper(a) by
//                 binary search = first b with min_s[b] >= max_s[a] (monotone early exit)
//   k_sap_sweep   (count, then emit) block = 256 consecutive a's; the window of b's they
//                 share is staged tile by tile in shared memory and broadcast-read
//   exclusive scan of the per-a counts; emit gives (a asc, b asc); one stable radix sort
//   of the pairs by b yields the reference Vec order (b asc, a asc).
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "cb2_kernels.h"

namespace cb2 {

struct sap_workspace {
    void* buf = nullptr;
    size_t cap = 0;
    uint64_t* h_total = nullptr;  // pinned
};

sap_workspace* sap_workspace_create() {
    sap_workspace* w = new sap_workspace();
    if (cudaMallocHost(&w->h_total, 2 * sizeof(uint64_t)) != cudaSuccess) {
        delete w;
        return nullptr;
    }
    return w;
}
void sap_workspace_destroy(sap_workspace* w) {
    if (!w) return;
    if (w->buf) cudaFree(w->buf);
    if (w->h_total) cudaFreeHost(w->h_total);
    delete w;
}

__device__ __forceinline__ uint32_t ord_f32(float x) {
    x = __fadd_rn(x, 0.0f);  // -0.0 -> +0.0: partial_cmp treats them as Equal
    const uint32_t b = __float_as_uint(x);
    return b ^ ((b >> 31) ? 0xFFFFFFFFu : 0x80000000u);
}

__global__ void __launch_bounds__(256) k_sap_keys(const float* __restrict__ aabbs, uint64_t n,
                                                 uint32_t axis, uint64_t* __restrict__ keys,
                                                 uint32_t* __restrict__ idx,
                                                 uint32_t* __restrict__ nan_flag) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float mn = __ldg(aabbs + 6 * i + axis);
    const float mx = __ldg(aabbs + 6 * i + 3 + axis);
    if (mn != mn || mx != mx) atomicOr(nan_flag, 1u);
    keys[i] = ((uint64_t)ord_f32(mn) << 32) | (uint64_t)ord_f32(mx);
    idx[i] = (uint32_t)i;
}

__global__ void __launch_bounds__(256) k_sap_gather(const float* __restrict__ aabbs, uint64_t n,
                                                   uint32_t axis, const uint32_t* __restrict__ perm,
                                                   float2* __restrict__ S, float2* __restrict__ U,
                                                   float2* __restrict__ V) {
    const uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const float* b = aabbs + 6ull * __ldg(perm + k);
    const uint32_t u = (axis + 1) % 3, v = (axis + 2) % 3;
    S[k] = make_float2(__ldg(b + axis), __ldg(b + 3 + axis));
    U[k] = make_float2(__ldg(b + u), __ldg(b + 3 + u));
    V[k] = make_float2(__ldg(b + v), __ldg(b + 3 + v));
}

// upper[a] = first b in (a, n] with min_s[b] >= max_s[a]  (min_s is sorted ascending)
__global__ void __launch_bounds__(256) k_sap_upper(const float2* __restrict__ S, uint64_t n,
                                                  uint32_t* __restrict__ upper) {
    const uint64_t a = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= n) return;
    const float key = __ldg(&S[a].y);
    uint64_t lo = a + 1, hi = n;
    while (lo < hi) {
        const uint64_t mid = (lo + hi) >> 1;
        if (__ldg(&S[mid].x) < key) lo = mid + 1;
        else hi = mid;
    }
    upper[a] = (uint32_t)lo;
}

constexpr int SAP_BLOCK = 256;
constexpr int SAP_TILE = 1024;

template <bool EMIT>
__global__ void __launch_bounds__(SAP_BLOCK) k_sap_sweep(
    const float2* __restrict__ S, const float2* __restrict__ U, const float2* __restrict__ V,
    const uint32_t* __restrict__ upper, uint64_t n, uint64_t* __restrict__ counts,
    const uint64_t* __restrict__ offsets, uint32_t* __restrict__ out_a,
    uint32_t* __restrict__ out_b, uint64_t cap) {
    __shared__ float2 sU[SAP_TILE];
    __shared__ float2 sV[SAP_TILE];
    __shared__ float sMaxS[SAP_TILE];
    __shared__ uint32_t s_end;
    const uint64_t a0 = (uint64_t)blockIdx.x * SAP_BLOCK;
    const uint64_t a = a0 + threadIdx.x;
    const bool live = a < n;
    float2 ms = make_float2(0.f, 0.f), mu = ms, mv = ms;
    uint32_t up = 0;
    if (live) {
        ms = __ldg(S + a);
        mu = __ldg(U + a);
        mv = __ldg(V + a);
        up = __ldg(upper + a);
    }
    if (threadIdx.x == 0) s_end = 0;
    __syncthreads();
    atomicMax(&s_end, up);
    __syncthreads();
    const uint64_t b_end = s_end;
    uint32_t cnt = 0;
    uint64_t w = 0;
    if (EMIT && live) w = offsets[a];
    for (uint64_t bt = a0 + 1; bt < b_end; bt += SAP_TILE) {
        __syncthreads();
        for (int k = threadIdx.x; k < SAP_TILE; k += SAP_BLOCK) {
            const uint64_t b = bt + k;
            if (b < n) {
                sU[k] = __ldg(U + b);
                sV[k] = __ldg(V + b);
                sMaxS[k] = __ldg(&S[b].y);
            }
        }
        __syncthreads();
        if (live) {
            const uint64_t lo = (a + 1 > bt) ? a + 1 : bt;
            const uint64_t hi = (up < bt + SAP_TILE) ? (uint64_t)up : bt + SAP_TILE;
            for (uint64_t b = lo; b < hi; ++b) {
                const int k = (int)(b - bt);
                const float2 bu = sU[k];
                // Aabb3::intersects (aabb3.rs:246-253), strict. a1.s > b0.s holds for b < upper.
                if (mu.y > bu.x && mu.x < bu.y) {
                    const float2 bv = sV[k];
                    if (mv.y > bv.x && mv.x < bv.y && ms.x < sMaxS[k]) {
                        if (EMIT) {
                            if (w < cap) {
                                out_a[w] = (uint32_t)a;
                                out_b[w] = (uint32_t)b;
                            }
                            ++w;
                        }
                        ++cnt;
                    }
                }
            }
        }
    }
    if (!EMIT && live) counts[a] = cnt;
}

__global__ void __launch_bounds__(256) k_sap_total(const uint64_t* __restrict__ counts,
                                                  const uint64_t* __restrict__ offsets, uint64_t n,
                                                  const uint32_t* __restrict__ nan_flag,
                                                  uint64_t* __restrict__ h_total) {
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        h_total[0] = offsets[n - 1] + counts[n - 1];
        h_total[1] = *nan_flag;
    }
}

__global__ void __launch_bounds__(256) k_sap_interleave(const uint32_t* __restrict__ a,
                                                       const uint32_t* __restrict__ b, uint64_t np,
                                                       uint32_t* __restrict__ pairs) {
    const uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= np) return;
    reinterpret_cast<uint2*>(pairs)[p] = make_uint2(a[p], b[p]);
}

__global__ void __launch_bounds__(256) k_iota(uint32_t* __restrict__ p, uint64_t n) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = (uint32_t)i;
}

static inline size_t align_up(size_t x) { return (x + 255) & ~(size_t)255; }
static inline unsigned nblk(uint64_t n, unsigned b) { return (unsigned)((n + b - 1) / b); }

#define SAP_CK(x)                                   \
    do {                                            \
        cudaError_t e_ = (x);                       \
        if (e_ != cudaSuccess) {                    \
            *err = cudaGetErrorString(e_);          \
            return CB2_ERR_CUDA;                    \
        }                                           \
    } while (0)

int sap_find_pairs(sap_workspace* w, const cb2_aabb3* aabbs, uint64_t n, uint32_t axis,
                   uint32_t* perm_out, uint32_t* pairs_out, uint64_t cap, uint64_t* n_pairs,
                   cudaStream_t st, uint64_t* launches, const char** err) {
    *n_pairs = 0;
    if (n == 0) return CB2_OK;
    if (n <= 1) {  // sweep_prune.rs:83-85: nothing sorted
        k_iota<<<nblk(n, 256), 256, 0, st>>>(perm_out, n);
        ++*launches;
        SAP_CK(cudaGetLastError());
        return CB2_OK;
    }
    if (n > 0xFFFFFFFFull) {
        *err = "more than 2^32-1 boxes";
        return CB2_ERR_ARG;
    }
    const float* boxes = reinterpret_cast<const float*>(aabbs);
    // ---- workspace carve-up (phase 1) ------------------------------------------------------
    size_t sort_tmp = 0, scan_tmp = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, sort_tmp, (uint64_t*)nullptr, (uint64_t*)nullptr,
                                    (uint32_t*)nullptr, (uint32_t*)nullptr, (int64_t)n, 0, 64, st);
    cub::DeviceScan::ExclusiveSum(nullptr, scan_tmp, (uint64_t*)nullptr, (uint64_t*)nullptr,
                                  (int64_t)n, st);
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += align_up(bytes); return o; };
    const size_t o_keys0 = take(8 * n), o_keys1 = take(8 * n), o_idx0 = take(4 * n);
    const size_t o_S = take(8 * n), o_U = take(8 * n), o_V = take(8 * n);
    const size_t o_upper = take(4 * n), o_counts = take(8 * n), o_offs = take(8 * n);
    const size_t o_flag = take(256);
    const size_t o_tmp = take(sort_tmp > scan_tmp ? sort_tmp : scan_tmp);
    const size_t phase1 = off;
    if (w->cap < phase1) {
        if (w->buf) cudaFree(w->buf);
        w->buf = nullptr;
        w->cap = 0;
        SAP_CK(cudaMalloc(&w->buf, phase1 + (phase1 >> 2)));
        w->cap = phase1 + (phase1 >> 2);
    }
    char* base = static_cast<char*>(w->buf);
    uint64_t* keys0 = (uint64_t*)(base + o_keys0);
    uint64_t* keys1 = (uint64_t*)(base + o_keys1);
    uint32_t* idx0 = (uint32_t*)(base + o_idx0);
    float2* S = (float2*)(base + o_S);
    float2* U = (float2*)(base + o_U);
    float2* V = (float2*)(base + o_V);
    uint32_t* upper = (uint32_t*)(base + o_upper);
    uint64_t* counts = (uint64_t*)(base + o_counts);
    uint64_t* offs = (uint64_t*)(base + o_offs);
    uint32_t* flag = (uint32_t*)(base + o_flag);
    void* tmp = base + o_tmp;

    SAP_CK(cudaMemsetAsync(flag, 0, 4, st));
    k_sap_keys<<<nblk(n, 256), 256, 0, st>>>(boxes, n, axis, keys0, idx0, flag);
    size_t tb = sort_tmp;
    SAP_CK(cub::DeviceRadixSort::SortPairs(tmp, tb, keys0, keys1, idx0, perm_out, (int64_t)n, 0, 64,
                                           st));
    k_sap_gather<<<nblk(n, 256), 256, 0, st>>>(boxes, n, axis, perm_out, S, U, V);
    k_sap_upper<<<nblk(n, 256), 256, 0, st>>>(S, n, upper);
    k_sap_sweep<false><<<nblk(n, SAP_BLOCK), SAP_BLOCK, 0, st>>>(S, U, V, upper, n, counts, nullptr,
                                                                 nullptr, nullptr, 0);
    // exclusive scan of the per-a counts (64-bit: a dense scene can exceed 2^32 pairs)
    tb = scan_tmp;
    SAP_CK(cub::DeviceScan::ExclusiveSum(tmp, tb, counts, offs, (int64_t)n, st));
    k_sap_total<<<1, 32, 0, st>>>(counts, offs, n, flag, w->h_total);
    *launches += 8;  // keys, sort (>=1), gather, upper, sweep, scan (>=1), total + memset
    SAP_CK(cudaGetLastError());
    SAP_CK(cudaStreamSynchronize(st));
    if (w->h_total[1]) {
        *err = "NaN extent on the sweep axis";
        return CB2_ERR_NAN;
    }
    const uint64_t np = w->h_total[0];
    *n_pairs = np;
    if (np == 0) return CB2_OK;
    if (np > cap) return CB2_ERR_NOSPC;
    // ---- phase 2: emit (a asc, b asc), stable sort by b, interleave -------------------------
    // (phase-1 arrays keys0/keys1/idx0 are dead now; reuse the workspace after `offs`)
    int bits = 1;
    while (((uint64_t)1 << bits) < n) ++bits;
    size_t sort2_tmp = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, sort2_tmp, (uint32_t*)nullptr, (uint32_t*)nullptr,
                                    (uint32_t*)nullptr, (uint32_t*)nullptr, (int64_t)np, 0, bits,
                                    st);
    size_t off2 = phase1;
    auto take2 = [&](size_t bytes) { size_t o = off2; off2 += align_up(bytes); return o; };
    const size_t o_a0 = take2(4 * np), o_b0 = take2(4 * np), o_a1 = take2(4 * np),
                 o_b1 = take2(4 * np), o_tmp2 = take2(sort2_tmp);
    if (w->cap < off2) {
        // grow, preserving phase-1 contents
        void* nb = nullptr;
        SAP_CK(cudaMalloc(&nb, off2 + (off2 >> 2)));
        SAP_CK(cudaMemcpyAsync(nb, w->buf, phase1, cudaMemcpyDeviceToDevice, st));
        SAP_CK(cudaStreamSynchronize(st));
        cudaFree(w->buf);
        w->buf = nb;
        w->cap = off2 + (off2 >> 2);
        base = static_cast<char*>(w->buf);
        S = (float2*)(base + o_S);
        U = (float2*)(base + o_U);
        V = (float2*)(base + o_V);
        upper = (uint32_t*)(base + o_upper);
        counts = (uint64_t*)(base + o_counts);
        offs = (uint64_t*)(base + o_offs);
    }
    uint32_t* a0 = (uint32_t*)(base + o_a0);
    uint32_t* b0 = (uint32_t*)(base + o_b0);
    uint32_t* a1 = (uint32_t*)(base + o_a1);
    uint32_t* b1 = (uint32_t*)(base + o_b1);
    k_sap_sweep<true><<<nblk(n, SAP_BLOCK), SAP_BLOCK, 0, st>>>(S, U, V, upper, n, nullptr, offs, a0,
                                                                b0, np);
    size_t tb2 = sort2_tmp;
    SAP_CK(cub::DeviceRadixSort::SortPairs(base + o_tmp2, tb2, b0, b1, a0, a1, (int64_t)np, 0, bits,
                                           st));
    k_sap_interleave<<<nblk(np, 256), 256, 0, st>>>(a1, b1, np, pairs_out);
    *launches += 3;
    SAP_CK(cudaGetLastError());
    return CB2_OK;
}

}  // namespace cb2
