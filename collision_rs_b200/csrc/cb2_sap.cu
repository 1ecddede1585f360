// cb2_sap.cu — SweepAndPrune3::find_collider_pairs on the device
// (algorithm/broad_phase/sweep_prune.rs:76-136).
//
// Reference semantics (finite extents): stable sort by (min[axis], max[axis]) under
// partial_cmp (so -0.0 == +0.0); every pair (a<b in sorted positions) whose boxes overlap
// STRICTLY on all three axes (aabb3.rs:246-253); Vec order = b ascending, then a ascending;
// sweep_axis <- 0 afterwards (Variance3 accumulates nothing, sweep_prune.rs:254-277).
// The reference's `active.retain(max_a >= min_b)` prune is implied by the strict test.
//
// Device pipeline (B200-first: the reference's O(N * active-window) sweep — 1e10 box tests for
// 1 M boxes — is replaced by a sweep INSIDE the cells of a 2-D grid over the two other axes, which
// finds the identical pair set with ~30 tests per box):
//   k_sap_keys     64-bit radix keys ord(min[axis]) : ord(max[axis]) (−0 canonicalised)
//   radix sort     stable LSD sort of (key, index)                       -> perm (sorted positions)
//   k_sap_gather   sorted SoA: S=(min_s,max_s) U=(min_u,max_u) V=(min_v,max_v)
//   k_grid_stats   origin / extent of the (u,v) domain, log2 histogram of box extents
//   k_grid_params  cell size = 2^ceil(log2) of the 99.9 % extent quantile (most boxes touch <= 4
//                  cells), grid clamped to 2048 x 2048
//   k_grid_emit    one (cell, position) entry per cell a box touches (<= 9; boxes touching more are
//                  "large" and go to k_sap_large): 4 box-major slots plus a spill tail
//   radix sort     stable sort of the entries by cell        -> per cell: positions ascending
//   k_grid_pack    boxes in entry order, so the sweep reads memory sequentially
//   k_grid_sweep   one thread per entry (cell C, a): walk forward through C's entries b while
//                  min_s[b] < max_s[a]; strict Aabb3 test (aabb3.rs:246-253); a pair is reported only
//                  in its canonical cell (max of the two boxes' first cells), so exactly once
//   k_sap_large    the rare large boxes against every box of their sweep window
//   radix sort     of the packed keys (b << bits | a), then k_sap_unpack: the (a, b) u32 pair list
//                  in the reference's Vec order (b ascending, then a ascending)
// Monotone cell mapping (floor of a monotone f32 expression) guarantees that two boxes that overlap
// strictly on u and v share their canonical cell, so no pair is missed.  The old tiled sweep is kept
// behind CB2_SAP_LEGACY=1 for A/B tests.
#include <cstdlib>

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
    if (cudaMallocHost(&w->h_total, 4 * sizeof(uint64_t)) != cudaSuccess) {
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
                                                 uint32_t* __restrict__ nan_flag, bool nan_sorts_last) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float mn = __ldg(aabbs + 6 * i + axis);
    const float mx = __ldg(aabbs + 6 * i + 3 + axis);
    uint32_t hi = ord_f32(mn);
    if (mn != mn || mx != mx) {
        // SweepAndPrune: the reference's sort order is unspecified for NaN keys -> loud error.
        // BruteForce never sorts (brute_force.rs:20-39): a NaN box just intersects nothing; here it
        // sorts behind every other box, where the window walks and the strict test leave it alone.
        if (nan_sorts_last) hi = mn != mn ? 0xFFFFFFFFu : hi;
        else atomicOr(nan_flag, 1u);
    }
    keys[i] = ((uint64_t)hi << 32) | (uint64_t)ord_f32(mx);
    idx[i] = (uint32_t)i;
}

// The sort key of the reference is (min, max) on the sweep axis.  Two boxes rarely share a min, so the
// radix sort runs on the 32-bit min alone (4 passes instead of 8) and this kernel repairs the runs of
// equal mins: the first position of a run orders it by (max, original order) with a stable insertion
// sort.  A run longer than FIX_RUN_MAX (quantised coordinates) raises `need_full`: the host then
// repeats the call with the 64-bit keys.
constexpr uint32_t FIX_RUN_MAX = 64;
__global__ void __launch_bounds__(256) k_sap_keys32(const float* __restrict__ aabbs, uint64_t n, uint32_t axis,
                                                   uint32_t* __restrict__ keys, uint32_t* __restrict__ idx,
                                                   uint32_t* __restrict__ nan_flag, bool nan_sorts_last) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float mn = __ldg(aabbs + 6 * i + axis);
    const float mx = __ldg(aabbs + 6 * i + 3 + axis);
    uint32_t k = ord_f32(mn);
    if (mn != mn || mx != mx) {
        if (nan_sorts_last) k = mn != mn ? 0xFFFFFFFFu : k;
        else atomicOr(nan_flag, 1u);
    }
    keys[i] = k;
    idx[i] = (uint32_t)i;
}
__global__ void __launch_bounds__(256) k_sap_fix_runs(const float* __restrict__ aabbs, uint64_t n, uint32_t axis,
                                                     const uint32_t* __restrict__ keys,
                                                     uint32_t* __restrict__ perm, uint32_t* __restrict__ need_full) {
    const uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k + 1 >= n) return;
    const uint32_t key = __ldg(keys + k);
    if (__ldg(keys + k + 1) != key || (k > 0 && __ldg(keys + k - 1) == key)) return;  // not the head of a run
    uint64_t e = k + 2;
    while (e < n && e - k <= FIX_RUN_MAX && __ldg(keys + e) == key) ++e;
    const uint32_t len = (uint32_t)(e - k);
    if (len > FIX_RUN_MAX) {
        atomicOr(need_full, 1u);
        return;
    }
    uint32_t id[FIX_RUN_MAX], mx[FIX_RUN_MAX];
    for (uint32_t i = 0; i < len; ++i) {
        const uint32_t p = perm[k + i];
        const uint32_t m = ord_f32(__ldg(aabbs + 6ull * p + 3 + axis));
        uint32_t j = i;
        while (j > 0 && mx[j - 1] > m) {  // strict: equal (min, max) keep their original order
            mx[j] = mx[j - 1];
            id[j] = id[j - 1];
            --j;
        }
        mx[j] = m;
        id[j] = p;
    }
    for (uint32_t i = 0; i < len; ++i) perm[k + i] = id[i];
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

// ---- grid-accelerated sweep -------------------------------------------------------------------------
constexpr int GRID_MAX = 2048;        // cells per axis at most
constexpr int LARGE_CELLS = 9;        // a box touching more cells than this is "large"
constexpr int SLOT_CELLS = 4;         // entry array = n * SLOT_CELLS + n/4 records
constexpr int EXT_BUCKETS = 64;       // log2 histogram of extents: bucket = exponent + 32, clamped

struct grid_params {
    // written by k_grid_stats (ordered-uint encodings so atomicMin/Max work on floats)
    uint32_t min_u, min_v, max_u, max_v;
    uint32_t hist_u[EXT_BUCKETS], hist_v[EXT_BUCKETS];
    // written by k_grid_params
    float org_u, org_v, inv_u, inv_v;
    int nu, nv, bits;
    // counters
    uint32_t n_large;
    unsigned long long n_pairs;
};

__device__ __forceinline__ uint32_t ford(float x) {
    const uint32_t b = __float_as_uint(x);
    return b ^ ((b >> 31) ? 0xFFFFFFFFu : 0x80000000u);
}
__device__ __forceinline__ float funord(uint32_t k) {
    return __uint_as_float(k ^ ((k >> 31) ? 0x80000000u : 0xFFFFFFFFu));
}
__device__ __forceinline__ int ext_bucket(float e) {
    if (!(e > 0.0f)) return 0;
    int ex;
    frexpf(e, &ex);  // e = m * 2^ex, m in [0.5, 1)  =>  e <= 2^ex
    ex += 32;
    return ex < 0 ? 0 : (ex >= EXT_BUCKETS ? EXT_BUCKETS - 1 : ex);
}

__global__ void __launch_bounds__(256) k_grid_stats(const float2* __restrict__ U,
                                                   const float2* __restrict__ V, uint64_t n,
                                                   grid_params* __restrict__ G) {
    __shared__ uint32_t hu[EXT_BUCKETS], hv[EXT_BUCKETS], mn[2], mx[2];
    if (threadIdx.x < EXT_BUCKETS) hu[threadIdx.x] = hv[threadIdx.x] = 0;
    if (threadIdx.x < 2) { mn[threadIdx.x] = 0xFFFFFFFFu; mx[threadIdx.x] = 0u; }
    __syncthreads();
    for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n;
         k += (uint64_t)gridDim.x * blockDim.x) {
        const float2 u = __ldg(U + k), v = __ldg(V + k);
        // only finite coordinates shape the grid (the cell mapping clamps the others)
        if (isfinite(u.x)) atomicMin(&mn[0], ford(u.x));
        if (isfinite(v.x)) atomicMin(&mn[1], ford(v.x));
        if (isfinite(u.y)) atomicMax(&mx[0], ford(u.y));
        if (isfinite(v.y)) atomicMax(&mx[1], ford(v.y));
        atomicAdd(&hu[ext_bucket(u.y - u.x)], 1u);
        atomicAdd(&hv[ext_bucket(v.y - v.x)], 1u);
    }
    __syncthreads();
    if (threadIdx.x < EXT_BUCKETS) {
        if (hu[threadIdx.x]) atomicAdd(&G->hist_u[threadIdx.x], hu[threadIdx.x]);
        if (hv[threadIdx.x]) atomicAdd(&G->hist_v[threadIdx.x], hv[threadIdx.x]);
    }
    if (threadIdx.x == 0) {
        atomicMin(&G->min_u, mn[0]); atomicMin(&G->min_v, mn[1]);
        atomicMax(&G->max_u, mx[0]); atomicMax(&G->max_v, mx[1]);
    }
}

__device__ void grid_axis(const uint32_t* hist, uint64_t n, float lo, float hi, float& org,
                          float& inv, int& cells) {
    // cell = 2^e with e the bucket that covers 99.9 % of the extents (e <= 2^bucket)
    const uint64_t want = n - n / 1000;
    uint64_t acc = 0;
    int b = 0;
    for (; b < EXT_BUCKETS; ++b) {
        acc += hist[b];
        if (acc >= want) break;
    }
    if (b >= EXT_BUCKETS) b = EXT_BUCKETS - 1;
    float cell = ldexpf(1.0f, b - 32);
    float span = hi - lo;
    if (!(span > 0.0f) || !isfinite(span)) span = 0.0f;
    if (!(cell > 0.0f)) cell = 1.0f;
    if (span / cell > (float)(GRID_MAX - 1)) cell = span / (float)(GRID_MAX - 1);
    org = isfinite(lo) ? lo : 0.0f;
    inv = 1.0f / cell;
    cells = (int)(span * inv) + 1;
    if (cells < 1) cells = 1;
    if (cells > GRID_MAX) cells = GRID_MAX;
}

__global__ void k_grid_params(grid_params* __restrict__ G, uint64_t n) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    float lo_u = funord(G->min_u), hi_u = funord(G->max_u);
    float lo_v = funord(G->min_v), hi_v = funord(G->max_v);
    if (G->min_u == 0xFFFFFFFFu || G->max_u == 0u) lo_u = hi_u = 0.0f;  // no finite coordinate
    if (G->min_v == 0xFFFFFFFFu || G->max_v == 0u) lo_v = hi_v = 0.0f;
    grid_axis(G->hist_u, n, lo_u, hi_u, G->org_u, G->inv_u, G->nu);
    grid_axis(G->hist_v, n, lo_v, hi_v, G->org_v, G->inv_v, G->nv);
    int bits = 1;
    while ((1 << bits) < G->nu * G->nv + 1) ++bits;  // +1: the sentinel cell of unused slots
    G->bits = bits;
}

// monotone non-decreasing in x (fsub, fmul, floor, clamp are): overlapping intervals share a cell
__device__ __forceinline__ int cell_coord(float x, float org, float inv, int cells) {
    const float t = floorf(__fmul_rn(__fsub_rn(x, org), inv));
    if (!(t > 0.0f)) return 0;  // also NaN (rejected elsewhere) and -inf
    return t >= (float)(cells - 1) ? cells - 1 : (int)t;
}

struct cell_range {
    int u0, u1, v0, v1;
};
__device__ __forceinline__ cell_range box_cells(const grid_params* __restrict__ G, float2 u, float2 v) {
    cell_range r;
    r.u0 = cell_coord(u.x, G->org_u, G->inv_u, G->nu);
    r.u1 = cell_coord(u.y, G->org_u, G->inv_u, G->nu);
    r.v0 = cell_coord(v.x, G->org_v, G->inv_v, G->nv);
    r.v1 = cell_coord(v.y, G->org_v, G->inv_v, G->nv);
    return r;
}

// How a box takes part in the grid.  The reference tests every candidate with Aabb3::intersects
// literally (aabb3.rs:246-253), so
//  * a box with a NaN extent on a grid axis never intersects anything (every comparison is false):
//    it stays in the permutation but is neither gridded nor swept;
//  * a box whose extents are inverted on a grid axis (min > max) can still pass the strict test, but
//    its cell range is empty or reversed: it goes the way of the large boxes (literal test against
//    every box);
//  * everything else is gridded when it touches at most LARGE_CELLS cells.
// (Inverted extents on the SWEEP axis need nothing special: the window walk and the strict test
// are literal there; a NaN on the sweep axis is rejected with CB2_ERR_NAN — the reference's sort
// order is unspecified for it.)
enum : int { BOX_GRID = 0, BOX_LARGE = 1, BOX_DEAD = 2 };
__device__ __forceinline__ int box_class(const grid_params* __restrict__ G, float2 u, float2 v, cell_range& r,
                                         int& cnt) {
    cnt = 0;
    if (u.x != u.x || u.y != u.y || v.x != v.x || v.y != v.y) return BOX_DEAD;
    if (u.y < u.x || v.y < v.x) return BOX_LARGE;
    r = box_cells(G, u, v);
    cnt = (r.u1 - r.u0 + 1) * (r.v1 - r.v0 + 1);
    return cnt > LARGE_CELLS ? BOX_LARGE : BOX_GRID;
}

// A sharded call (this GPU reports the pairs whose later box sits in [b_lo, b_hi)) only needs the
// boxes that can take part in such a pair: the shard's own boxes, and the earlier ones whose sweep
// interval reaches the shard's first box (min_s ascends with the position, so max_s[k] > min_s[b_lo]
// is necessary for k < b_lo).  Everything else stays out of the grid: entries, cell sort, packing and
// sweep shrink with the shard.
__device__ __forceinline__ bool in_shard_reach(const float2* __restrict__ S, uint64_t k, uint32_t b_lo,
                                               uint32_t b_hi, uint64_t n) {
    if (b_lo == 0u && b_hi >= n) return true;
    if (k >= b_hi) return false;
    if (k >= b_lo) return true;
    return __ldg(&S[k].y) > __ldg(&S[b_lo].x);
}

// Entries (cell, position), compacted in POSITION order through an exclusive scan of the per-box
// cell counts: the stable sort by cell then leaves every cell's entries ascending by position.
// The entry array holds n * SLOT_CELLS + n/4 records (the cell size covers the 99.9 % extent
// quantile, so nearly every box touches <= 2 x 2 cells); a box that does not fit is treated as large.
__global__ void __launch_bounds__(256) k_grid_count(const float2* __restrict__ S, const float2* __restrict__ U,
                                                   const float2* __restrict__ V, uint64_t n,
                                                   const grid_params* __restrict__ G,
                                                   uint32_t* __restrict__ counts, uint32_t b_lo, uint32_t b_hi) {
    const uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    cell_range r;
    int cnt;
    const int cls = box_class(G, __ldg(U + k), __ldg(V + k), r, cnt);
    counts[k] = (cls == BOX_GRID && in_shard_reach(S, k, b_lo, b_hi, n)) ? (uint32_t)cnt : 0u;
}

__global__ void __launch_bounds__(256) k_grid_emit(const float2* __restrict__ S, const float2* __restrict__ U,
                                                  const float2* __restrict__ V, uint64_t n,
                                                  grid_params* __restrict__ G,
                                                  const uint32_t* __restrict__ offs, uint64_t n_entries,
                                                  uint32_t* __restrict__ ecell,
                                                  uint32_t* __restrict__ epos,
                                                  uint32_t* __restrict__ large_list, uint32_t b_lo,
                                                  uint32_t b_hi) {
    const uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    cell_range r;
    int cnt;
    const int cls = box_class(G, __ldg(U + k), __ldg(V + k), r, cnt);
    if (cls == BOX_DEAD) return;
    // (a large box is listed even outside the shard's reach: k_sap_large filters its pairs, and its
    //  backward pass must recognise which partners are large themselves)
    if (cls == BOX_GRID && !in_shard_reach(S, k, b_lo, b_hi, n)) return;
    const uint64_t o = offs[k];
    if (cls == BOX_LARGE || o + (uint64_t)cnt > n_entries) {
        large_list[atomicAdd(&G->n_large, 1u)] = (uint32_t)k;
        return;
    }
    uint32_t* c = ecell + o;
    uint32_t* p = epos + o;
    for (int iu = r.u0; iu <= r.u1; ++iu)
        for (int iv = r.v0; iv <= r.v1; ++iv) {
            *c++ = (uint32_t)(iu * G->nv + iv);
            *p++ = (uint32_t)k;
        }
}

// after the sort by cell: the boxes of each entry, in entry order, so the sweep reads them
// sequentially instead of gathering
__global__ void __launch_bounds__(256) k_grid_pack(const float2* __restrict__ S,
                                                  const float2* __restrict__ U,
                                                  const float2* __restrict__ V,
                                                  const uint32_t* __restrict__ ecell,
                                                  const uint32_t* __restrict__ epos, uint64_t n_entries,
                                                  const grid_params* __restrict__ G,
                                                  float2* __restrict__ eS, float2* __restrict__ eU,
                                                  float2* __restrict__ eV) {
    const uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_entries) return;
    if (__ldg(ecell + e) >= (uint32_t)(G->nu * G->nv)) return;  // sentinel slots sort to the end
    const uint32_t b = __ldg(epos + e);
    eS[e] = __ldg(S + b);
    eU[e] = __ldg(U + b);
    eV[e] = __ldg(V + b);
}

// SweepAndPrune: key = b << pos_bits | a over sorted positions a < b; sorting it gives the
// reference's Vec order (b asc, then a asc).  BruteForce (remap = the sort permutation): the pair in
// ORIGINAL indices lo < hi, key = lo << pos_bits | hi, i.e. (left asc, right asc) (brute_force.rs:28-35)
__device__ __forceinline__ unsigned long long pair_key(uint32_t a, uint32_t b, int pos_bits,
                                                       const uint32_t* __restrict__ remap) {
    if (remap) {
        const uint32_t x = __ldg(remap + a), y = __ldg(remap + b);
        const uint32_t lo = x < y ? x : y, hi = x < y ? y : x;
        return ((unsigned long long)lo << pos_bits) | hi;
    }
    return ((unsigned long long)b << pos_bits) | a;
}
__device__ __forceinline__ void emit_pair(grid_params* __restrict__ G, unsigned long long* __restrict__ keys,
                                          uint64_t cap, int pos_bits, uint32_t a, uint32_t b,
                                          const uint32_t* __restrict__ remap, uint32_t* __restrict__ seg_cnt) {
    const unsigned long long w = atomicAdd(&G->n_pairs, 1ull);
    const unsigned long long key = pair_key(a, b, pos_bits, remap);
    if (w < cap) keys[w] = key;
    if (seg_cnt) atomicAdd(seg_cnt + (uint32_t)(key >> pos_bits), 1u);
}

// ---- ordering the pair list without a global sort ---------------------------------------------------
// The reference's Vec is ordered by the later box b, then by a (BruteForce: by the lower original
// index, then the higher).  The high part of a packed key is therefore a SEGMENT id below n, and a
// segment holds a handful of pairs (4 on average for cfg 4).  A 40-bit radix sort of all keys moved
// 10x the pair list through HBM; instead the emitters count pairs per segment, an exclusive scan turns
// the counts into segment offsets, one pass scatters every key into its segment, and every segment is
// sorted on its own: by one thread when it is short, by one block (bitonic network in shared memory)
// when it is long.  A segment beyond SEG_BLOCK_MAX pairs sends the whole call down the radix-sort path.
constexpr uint32_t SEG_THREAD_MAX = 32;
constexpr uint32_t SEG_BLOCK_MAX = 4096;

__global__ void __launch_bounds__(256) k_seg_max(const uint32_t* __restrict__ cnt, uint64_t n,
                                                uint32_t* __restrict__ out_max) {
    uint32_t m = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t c = __ldg(cnt + i);
        m = c > m ? c : m;
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        const uint32_t o = __shfl_xor_sync(0xffffffffu, m, off);
        m = o > m ? o : m;
    }
    if ((threadIdx.x & 31u) == 0 && m) atomicMax(out_max, m);
}

__global__ void __launch_bounds__(256) k_seg_scatter(const unsigned long long* __restrict__ keys, uint64_t np,
                                                    int pos_bits, const uint64_t* __restrict__ off,
                                                    uint32_t* __restrict__ fill,
                                                    unsigned long long* __restrict__ out) {
    const uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= np) return;
    const unsigned long long k = keys[p];
    const uint32_t seg = (uint32_t)(k >> pos_bits);
    out[__ldg(off + seg) + atomicAdd(fill + seg, 1u)] = k;
}

// short segments: one thread each; long ones are queued for k_seg_sort_long
__global__ void __launch_bounds__(128) k_seg_sort(const unsigned long long* __restrict__ keys,
                                                 const uint64_t* __restrict__ off, uint64_t n, int pos_bits,
                                                 bool high_first, uint32_t* __restrict__ pairs,
                                                 uint32_t* __restrict__ long_list, uint32_t* __restrict__ n_long) {
    const uint64_t seg = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (seg >= n) return;
    const uint64_t b = __ldg(off + seg), e = __ldg(off + seg + 1);
    const uint32_t len = (uint32_t)(e - b);
    if (len == 0) return;
    if (len > SEG_THREAD_MAX) {
        long_list[atomicAdd(n_long, 1u)] = (uint32_t)seg;
        return;
    }
    const uint32_t mask = (uint32_t)((1ull << pos_bits) - 1ull);
    uint32_t lo[SEG_THREAD_MAX];
    for (uint32_t i = 0; i < len; ++i) {  // insertion sort on the low parts (all distinct)
        const uint32_t v = (uint32_t)keys[b + i] & mask;
        uint32_t j = i;
        while (j > 0 && lo[j - 1] > v) {
            lo[j] = lo[j - 1];
            --j;
        }
        lo[j] = v;
    }
    uint2* o = reinterpret_cast<uint2*>(pairs) + b;
    const uint32_t high = (uint32_t)seg;
    for (uint32_t i = 0; i < len; ++i) o[i] = high_first ? make_uint2(high, lo[i]) : make_uint2(lo[i], high);
}

__global__ void __launch_bounds__(256) k_seg_sort_long(const unsigned long long* __restrict__ keys,
                                                      const uint64_t* __restrict__ off, int pos_bits,
                                                      bool high_first, uint32_t* __restrict__ pairs,
                                                      const uint32_t* __restrict__ long_list,
                                                      const uint32_t* __restrict__ n_long) {
    __shared__ uint32_t lo[SEG_BLOCK_MAX];
    const uint32_t mask = (uint32_t)((1ull << pos_bits) - 1ull);
    const uint32_t total = *n_long;
    for (uint32_t li = blockIdx.x; li < total; li += gridDim.x) {
        const uint32_t seg = long_list[li];
        const uint64_t b = off[seg];
        const uint32_t len = (uint32_t)(off[seg + 1] - b);  // <= SEG_BLOCK_MAX (the host checked)
        uint32_t m = 1;
        while (m < len) m <<= 1;
        for (uint32_t i = threadIdx.x; i < m; i += blockDim.x)
            lo[i] = i < len ? ((uint32_t)keys[b + i] & mask) : 0xFFFFFFFFu;
        __syncthreads();
        for (uint32_t k = 2; k <= m; k <<= 1)
            for (uint32_t j = k >> 1; j > 0; j >>= 1) {
                for (uint32_t i = threadIdx.x; i < m; i += blockDim.x) {
                    const uint32_t x = i ^ j;
                    if (x > i) {
                        const uint32_t u = lo[i], v = lo[x];
                        const bool up = (i & k) == 0;
                        if ((u > v) == up) {
                            lo[i] = v;
                            lo[x] = u;
                        }
                    }
                }
                __syncthreads();
            }
        uint2* o = reinterpret_cast<uint2*>(pairs) + b;
        for (uint32_t i = threadIdx.x; i < len; i += blockDim.x)
            o[i] = high_first ? make_uint2(seg, lo[i]) : make_uint2(lo[i], seg);
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256) k_sap_unpack(const unsigned long long* __restrict__ keys,
                                                   uint64_t np, int pos_bits, bool high_first,
                                                   uint32_t* __restrict__ pairs) {
    const uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= np) return;
    const unsigned long long k = keys[p];
    const uint32_t low = (uint32_t)(k & ((1ull << pos_bits) - 1ull)), high = (uint32_t)(k >> pos_bits);
    reinterpret_cast<uint2*>(pairs)[p] = high_first ? make_uint2(high, low) : make_uint2(low, high);
}

#ifndef CB2_SWEEP_L   // lanes per entry of k_grid_sweep.  Measured at 1 M boxes (the whole call, ordered):
#define CB2_SWEEP_L 1 // 1: 0.688 ms, 2: 0.717, 4: 0.820, 8: 1.023 — most windows are short, more lanes per
#endif                // entry only multiply the idle trips; a thread per entry stays
template <int L>
__global__ void __launch_bounds__(256) k_grid_sweep(const float2* __restrict__ eS,
                                                   const float2* __restrict__ eU,
                                                   const float2* __restrict__ eV,
                                                   const uint32_t* __restrict__ ecell,
                                                   const uint32_t* __restrict__ epos, uint64_t n_entries,
                                                   grid_params* __restrict__ G,
                                                   unsigned long long* __restrict__ keys, uint64_t cap,
                                                   int pos_bits, const uint32_t* __restrict__ remap,
                                                   uint32_t b_lo, uint32_t b_hi,
                                                   uint32_t* __restrict__ seg_cnt) {
    // b_lo/b_hi: this GPU's shard of the pair list — only pairs whose LATER box (the `b` the
    // reference's Vec is ordered by) sits at a sorted position in [b_lo, b_hi) are reported
    const bool sharded = b_lo != 0u || b_hi != 0xFFFFFFFFu;
    // L lanes share one entry and take its window L candidates at a time (an experiment against the
    // 10 of 32 lanes one thread per entry keeps busy; L = 1 is the product path, see CB2_SWEEP_L)
    const uint64_t e = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) / L;
    const unsigned sub = threadIdx.x & (L - 1);
    const float org_u = G->org_u, inv_u = G->inv_u, org_v = G->org_v, inv_v = G->inv_v;
    const int nu = G->nu, nv = G->nv;
    // no early return: every lane takes part in the warp-wide reservation below
    uint32_t cell = 0xFFFFFFFFu;
    if (e < n_entries) cell = __ldg(ecell + e);
    const bool valid = cell < (uint32_t)(nu * nv);  // not a sentinel slot
    float2 as = make_float2(0.f, 0.f), au = as, av = as;
    if (valid) {
        as = __ldg(eS + e);
        au = __ldg(eU + e);
        av = __ldg(eV + e);
    }
    const int au0 = cell_coord(au.x, org_u, inv_u, nu);
    const int av0 = cell_coord(av.x, org_v, inv_v, nv);
    const int cu = (int)(cell / (uint32_t)nv), cv = (int)(cell % (uint32_t)nv);
    // One pass over the (short) window.  The warp stays convergent (a lane whose window is closed
    // idles), the hits of a trip are ranked by a ballot and staged in the warp's slice of shared
    // memory; a full slice — and the remainder at the end — is flushed with ONE reservation on the
    // global counter and coalesced stores.  (Round 1 walked every window twice: count, reserve per
    // warp, write.)
    constexpr int STAGE = 96;  // keys per warp slice; a trip adds at most 32
    __shared__ unsigned long long stage[8][STAGE];
    const unsigned lane = threadIdx.x & 31u;
    unsigned long long* mystage = stage[threadIdx.x >> 5];
    uint32_t fill = 0;  // warp-uniform
    auto flush = [&]() {
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(&G->n_pairs, (unsigned long long)fill);
        base = __shfl_sync(0xffffffffu, base, 0);
        __syncwarp();
        for (uint32_t j = lane; j < fill; j += 32u)
            if (base + j < cap) keys[base + j] = mystage[j];
        __syncwarp();
        fill = 0;
    };
    const uint32_t a = valid ? __ldg(epos + e) : 0u;
    uint64_t f = e + 1 + sub;
    bool open = valid;  // uniform over the L lanes of an entry
    const unsigned lane_bit = 1u << lane;
    const unsigned last_of_sub = 1u << ((lane & ~(unsigned)(L - 1)) + (L - 1));
    while (__any_sync(0xffffffffu, open)) {
        bool hit = false, inwin = false;
        uint32_t bp = 0;
        if (open && f < n_entries && __ldg(ecell + f) == cell) {
            const float2 bs = __ldg(eS + f);  // entries of a cell ascend by position, so by min_s
            inwin = bs.x < as.y;              // ... and a closed window stays closed for every later f
            if (inwin && as.x < bs.y) {
                // Aabb3::intersects (aabb3.rs:246-253), strict on all three axes
                const float2 bu = __ldg(eU + f);
                const float2 bv = __ldg(eV + f);
                if (au.y > bu.x && au.x < bu.y && av.y > bv.x && av.x < bv.y) {
                    // report only in the canonical cell of the pair
                    const int bu0 = cell_coord(bu.x, org_u, inv_u, nu);
                    const int bv0 = cell_coord(bv.x, org_v, inv_v, nv);
                    if ((au0 > bu0 ? au0 : bu0) == cu && (av0 > bv0 ? av0 : bv0) == cv) {
                        bp = __ldg(epos + f);
                        hit = !sharded || (bp >= b_lo && bp < b_hi);
                    }
                }
            }
        }
        f += L;
        // the window goes on iff the entry's LAST lane was still inside it
        const unsigned win = __ballot_sync(0xffffffffu, inwin);
        open = open && (win & last_of_sub) != 0u;
        const unsigned m = __ballot_sync(0xffffffffu, hit);
        if (m) {
            if (hit) {
                const unsigned long long key = pair_key(a, bp, pos_bits, remap);
                mystage[fill + __popc(m & (lane_bit - 1u))] = key;
                if (seg_cnt) atomicAdd(seg_cnt + (uint32_t)(key >> pos_bits), 1u);
            }
            fill += __popc(m);
            if (fill > STAGE - 32) flush();
        }
    }
    if (fill) flush();
}

// large boxes (not in the grid): forward against every b of their window, backward against every
// non-large a whose window reaches them
__global__ void __launch_bounds__(256) k_sap_large(const float2* __restrict__ S,
                                                  const float2* __restrict__ U,
                                                  const float2* __restrict__ V, uint64_t n,
                                                  const uint32_t* __restrict__ large_list,
                                                  const uint32_t* __restrict__ offs, uint64_t n_entries,
                                                  grid_params* __restrict__ G,
                                                  unsigned long long* __restrict__ keys, uint64_t cap,
                                                  int pos_bits, const uint32_t* __restrict__ remap,
                                                  uint32_t b_lo, uint32_t b_hi, uint32_t* __restrict__ seg_cnt) {
    const uint32_t nl = G->n_large;
    for (uint32_t li = 0; li < nl; ++li) {
        const uint32_t p = __ldg(large_list + li);
        const float2 ps = __ldg(S + p), pu = __ldg(U + p), pv = __ldg(V + p);
        for (uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n;
             q += (uint64_t)gridDim.x * blockDim.x) {
            if (q == p) continue;
            const float2 qs = __ldg(S + q), qu = __ldg(U + q), qv = __ldg(V + q);
            if (!(ps.y > qs.x && ps.x < qs.y && pu.y > qu.x && pu.x < qu.y && pv.y > qv.x && pv.x < qv.y))
                continue;
            const uint32_t later = q > p ? (uint32_t)q : p;
            if (later < b_lo || later >= b_hi) continue;  // another GPU's shard
            if (q > p) {
                emit_pair(G, keys, cap, pos_bits, p, (uint32_t)q, remap, seg_cnt);
            } else {
                // (q, p) with q < p: if q is large too, q's own forward pass reports it
                cell_range r;
                int cnt;
                if (box_class(G, qu, qv, r, cnt) == BOX_GRID &&
                    (uint64_t)__ldg(offs + q) + (uint64_t)cnt <= n_entries)
                    emit_pair(G, keys, cap, pos_bits, (uint32_t)q, p, remap, seg_cnt);
            }
        }
    }
}

// entries the grid really holds (the exclusive scan's total, clamped to the array): read back by the
// host so that the cell sort, the packing and the sweep run over them and not over the capacity
__global__ void k_grid_entries(const uint32_t* __restrict__ counts, const uint32_t* __restrict__ offs, uint64_t n,
                               uint64_t capacity, uint64_t* __restrict__ h_out) {
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        const uint64_t t = (uint64_t)offs[n - 1] + counts[n - 1];
        *h_out = t < capacity ? t : capacity;
    }
}

__global__ void k_grid_total(const grid_params* __restrict__ G, const uint32_t* __restrict__ nan_flag,
                             uint64_t* __restrict__ h_total) {
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        h_total[0] = G->n_pairs;
        // NaN flag (bit 0) | "a run of equal mins was too long for the 32-bit sort" (bit 1) | longest segment
        h_total[1] = (uint64_t)(nan_flag[0] & 1u) | ((uint64_t)(nan_flag[3] & 1u) << 1) | ((uint64_t)nan_flag[1] << 32);
    }
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

static int sap_find_pairs_legacy(sap_workspace* w, const cb2_aabb3* aabbs, uint64_t n, uint32_t axis,
                                 uint32_t* perm_out, uint32_t* pairs_out, uint64_t cap,
                                 uint64_t* n_pairs, cudaStream_t st, uint64_t* launches,
                                 const char** err) {
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
    k_sap_keys<<<nblk(n, 256), 256, 0, st>>>(boxes, n, axis, keys0, idx0, flag, false);
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


static int ensure_cap(sap_workspace* w, size_t bytes, const char** err) {
    if (w->cap >= bytes) return CB2_OK;
    if (w->buf) cudaFree(w->buf);
    w->buf = nullptr;
    w->cap = 0;
    SAP_CK(cudaMalloc(&w->buf, bytes + (bytes >> 2)));
    w->cap = bytes + (bytes >> 2);
    return CB2_OK;
}

int sap_find_pairs(sap_workspace* w, const cb2_aabb3* aabbs, uint64_t n, uint32_t axis,
                   uint32_t* perm_out, uint32_t* pairs_out, uint64_t cap, uint64_t* n_pairs,
                   cudaStream_t st, uint64_t* launches, const char** err, bool brute_force,
                   uint32_t b_lo, uint32_t b_hi, bool unordered) {
    {
        const char* e = std::getenv("CB2_SAP_LEGACY");
        if (e && e[0] == '1' && !brute_force && b_lo == 0u && b_hi == 0xFFFFFFFFu && !unordered)
            return sap_find_pairs_legacy(w, aabbs, n, axis, perm_out, pairs_out, cap, n_pairs, st,
                                         launches, err);
    }
    *n_pairs = 0;
    if (n == 0) return CB2_OK;
    if (n <= 1) {  // sweep_prune.rs:83-85: nothing sorted (brute_force.rs:22-24: no pairs)
        if (!brute_force) {
            k_iota<<<nblk(n, 256), 256, 0, st>>>(perm_out, n);
            ++*launches;
            SAP_CK(cudaGetLastError());
        }
        return CB2_OK;
    }
    if (n > 0x7FFFFFFFull / (SLOT_CELLS + 1)) {
        *err = "too many boxes for one call";
        return CB2_ERR_ARG;
    }
    if (cap && (reinterpret_cast<uintptr_t>(pairs_out) & 7u)) {
        *err = "pairs_out must be 8-byte aligned";
        return CB2_ERR_ARG;
    }
    const float* boxes = reinterpret_cast<const float*>(aabbs);
    const uint64_t ne = n * SLOT_CELLS + n / 4 + 1024;
    const int cell_bits = 23;  // GRID_MAX^2 cells + the sentinel
    int pos_bits = 1;
    while (((uint64_t)1 << pos_bits) < n) ++pos_bits;
    // internal pair buffer: the caller's cap, but start no larger than 16 pairs per box
    uint64_t kcap = cap < 16 * n + 1024 ? cap : 16 * n + 1024;
    static const bool keys64_env = std::getenv("CB2_SAP_KEYS64") && std::getenv("CB2_SAP_KEYS64")[0] == '1';
    bool keys64 = keys64_env;  // false: sort on the 32-bit min and repair the runs of equal mins
    for (int attempt = 0; attempt < 3; ++attempt) {
        size_t sort1 = 0, sort1b = 0, sort2 = 0, sort3 = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, sort1, (uint64_t*)nullptr, (uint64_t*)nullptr,
                                        (uint32_t*)nullptr, (uint32_t*)nullptr, (int64_t)n, 0, 64, st);
        cub::DeviceRadixSort::SortPairs(nullptr, sort1b, (uint32_t*)nullptr, (uint32_t*)nullptr,
                                        (uint32_t*)nullptr, (uint32_t*)nullptr, (int64_t)n, 0, 32, st);
        if (sort1b > sort1) sort1 = sort1b;
        cub::DeviceRadixSort::SortPairs(nullptr, sort2, (uint32_t*)nullptr, (uint32_t*)nullptr,
                                        (uint32_t*)nullptr, (uint32_t*)nullptr, (int64_t)ne, 0,
                                        cell_bits, st);
        size_t scan_tmp = 0, scan2_tmp = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, scan_tmp, (uint32_t*)nullptr, (uint32_t*)nullptr,
                                      (int64_t)n, st);
        cub::DeviceScan::ExclusiveScan(nullptr, scan2_tmp, (uint32_t*)nullptr, (uint64_t*)nullptr, cub::Sum(),
                                       (uint64_t)0, (int64_t)(n + 1), st);
        if (scan2_tmp > scan_tmp) scan_tmp = scan2_tmp;
        cub::DeviceRadixSort::SortKeys(nullptr, sort3, (unsigned long long*)nullptr,
                                       (unsigned long long*)nullptr, (int64_t)(kcap ? kcap : 1), 0,
                                       2 * pos_bits, st);
        size_t off = 0;
        auto take = [&](size_t bytes) { size_t o = off; off += align_up(bytes); return o; };
        const size_t o_keys0 = take(8 * n), o_keys1 = take(8 * n), o_idx0 = take(4 * n);
        const size_t o_S = take(8 * n), o_U = take(8 * n), o_V = take(8 * n);
        const size_t o_c0 = take(4 * ne), o_c1 = take(4 * ne), o_p0 = take(4 * ne), o_p1 = take(4 * ne);
        const size_t o_large = take(4 * n), o_G = take(sizeof(grid_params)), o_flag = take(256);
        const size_t o_cnt = take(4 * n), o_offs = take(4 * n), o_perm = take(4 * n);
        const size_t o_eS = take(8 * ne), o_eU = take(8 * ne), o_eV = take(8 * ne);
        const size_t o_pk = take(8 * (kcap ? kcap : 1)), o_pk2 = take(8 * (kcap ? kcap : 1));
        const size_t o_seg = take(4 * (n + 1)), o_fill = take(4 * (n + 1)), o_soff = take(8 * (n + 2));
        const size_t o_long = take(4 * n);
        size_t tmpsz = sort1 > sort2 ? sort1 : sort2;
        tmpsz = tmpsz > sort3 ? tmpsz : sort3;
        tmpsz = tmpsz > scan_tmp ? tmpsz : scan_tmp;
        const size_t o_tmp = take(tmpsz);
        int rc = ensure_cap(w, off, err);
        if (rc) return rc;
        char* base = static_cast<char*>(w->buf);
        uint64_t* keys0 = (uint64_t*)(base + o_keys0);
        uint64_t* keys1 = (uint64_t*)(base + o_keys1);
        uint32_t* idx0 = (uint32_t*)(base + o_idx0);
        float2* S = (float2*)(base + o_S);
        float2* U = (float2*)(base + o_U);
        float2* V = (float2*)(base + o_V);
        uint32_t* c0 = (uint32_t*)(base + o_c0);
        uint32_t* c1 = (uint32_t*)(base + o_c1);
        uint32_t* p0 = (uint32_t*)(base + o_p0);
        uint32_t* p1 = (uint32_t*)(base + o_p1);
        uint32_t* large = (uint32_t*)(base + o_large);
        grid_params* G = (grid_params*)(base + o_G);
        uint32_t* flag = (uint32_t*)(base + o_flag);
        uint32_t* cnts = (uint32_t*)(base + o_cnt);
        uint32_t* offs = (uint32_t*)(base + o_offs);
        if (brute_force) perm_out = (uint32_t*)(base + o_perm);  // BruteForce keeps the caller's order
        const uint32_t* remap = brute_force ? perm_out : nullptr;
        float2* eS = (float2*)(base + o_eS);
        float2* eU = (float2*)(base + o_eU);
        float2* eV = (float2*)(base + o_eV);
        unsigned long long* pk = (unsigned long long*)(base + o_pk);
        unsigned long long* pk2 = (unsigned long long*)(base + o_pk2);
        void* tmp = base + o_tmp;
        uint32_t* seg_cnt = (uint32_t*)(base + o_seg);
        uint32_t* seg_fill = (uint32_t*)(base + o_fill);
        uint64_t* seg_off = (uint64_t*)(base + o_soff);
        uint32_t* long_list = (uint32_t*)(base + o_long);
        // the pair list is ordered segment by segment unless the caller wants the bare set
        static const bool seg_env = !(std::getenv("CB2_SAP_RADIX_PAIRS") && std::getenv("CB2_SAP_RADIX_PAIRS")[0] == '1');
        const bool by_segment = seg_env && !(unordered && !brute_force);
        if (by_segment) {
            SAP_CK(cudaMemsetAsync(seg_cnt, 0, 4 * (n + 1), st));
            SAP_CK(cudaMemsetAsync(seg_fill, 0, 4 * (n + 1), st));
        }
        uint32_t* seg_arg = by_segment ? seg_cnt : nullptr;

        SAP_CK(cudaMemsetAsync(flag, 0, 256, st));
        SAP_CK(cudaMemsetAsync(G, 0, sizeof(grid_params), st));
        SAP_CK(cudaMemsetAsync(G, 0xFF, 2 * sizeof(uint32_t), st));  // min_u, min_v = +max
        SAP_CK(cudaMemsetAsync(c0, 0xFF, 4 * ne, st));  // unused entry records = sentinel cells
        size_t tb = sort1;
        if (keys64) {
            k_sap_keys<<<nblk(n, 256), 256, 0, st>>>(boxes, n, axis, keys0, idx0, flag, brute_force);
            SAP_CK(cub::DeviceRadixSort::SortPairs(tmp, tb, keys0, keys1, idx0, perm_out, (int64_t)n, 0,
                                                   64, st));
        } else {
            uint32_t* k32a = reinterpret_cast<uint32_t*>(keys0);
            uint32_t* k32b = reinterpret_cast<uint32_t*>(keys1);
            k_sap_keys32<<<nblk(n, 256), 256, 0, st>>>(boxes, n, axis, k32a, idx0, flag, brute_force);
            SAP_CK(cub::DeviceRadixSort::SortPairs(tmp, tb, k32a, k32b, idx0, perm_out, (int64_t)n, 0, 32, st));
            k_sap_fix_runs<<<nblk(n, 256), 256, 0, st>>>(boxes, n, axis, k32b, perm_out, flag + 3);
            ++*launches;
        }
        k_sap_gather<<<nblk(n, 256), 256, 0, st>>>(boxes, n, axis, perm_out, S, U, V);
        k_grid_stats<<<592, 256, 0, st>>>(U, V, n, G);
        k_grid_params<<<1, 32, 0, st>>>(G, n);
        k_grid_count<<<nblk(n, 256), 256, 0, st>>>(S, U, V, n, G, cnts, b_lo, b_hi);
        tb = scan_tmp;
        SAP_CK(cub::DeviceScan::ExclusiveSum(tmp, tb, cnts, offs, (int64_t)n, st));
        k_grid_entries<<<1, 32, 0, st>>>(cnts, offs, n, ne, w->h_total + 2);
        k_grid_emit<<<nblk(n, 256), 256, 0, st>>>(S, U, V, n, G, offs, ne, c0, p0, large, b_lo, b_hi);
        ++*launches;
        // one short host wait buys the entry count: the array is sized for 4.25 entries per box, the
        // grid of cfg 4 holds 1.6 (and a shard a fraction of that)
        SAP_CK(cudaStreamSynchronize(st));
        const uint64_t nu = w->h_total[2] ? w->h_total[2] : 1;
        tb = sort2;
        SAP_CK(cub::DeviceRadixSort::SortPairs(tmp, tb, c0, c1, p0, p1, (int64_t)nu, 0, cell_bits, st));
        k_grid_pack<<<nblk(nu, 256), 256, 0, st>>>(S, U, V, c1, p1, nu, G, eS, eU, eV);
        k_grid_sweep<CB2_SWEEP_L><<<nblk(nu * CB2_SWEEP_L, 256), 256, 0, st>>>(eS, eU, eV, c1, p1, nu, G, pk, kcap,
                                                                               pos_bits, remap, b_lo, b_hi, seg_arg);
        k_sap_large<<<592, 256, 0, st>>>(S, U, V, n, large, offs, ne, G, pk, kcap, pos_bits, remap, b_lo,
                                         b_hi, seg_arg);
        if (by_segment) {
            k_seg_max<<<296, 256, 0, st>>>(seg_cnt, n, flag + 1);
            ++*launches;
        }
        k_grid_total<<<1, 32, 0, st>>>(G, flag, w->h_total);
        *launches += 13;
        SAP_CK(cudaGetLastError());
        SAP_CK(cudaStreamSynchronize(st));
        if (w->h_total[1] & 1ull) {
            *err = "NaN extent on the sweep axis";
            return CB2_ERR_NAN;
        }
        if (w->h_total[1] & 2ull) {  // many boxes share a min (quantised input): full (min, max) keys
            keys64 = true;
            continue;
        }
        const uint32_t longest = (uint32_t)(w->h_total[1] >> 32);
        const uint64_t np = w->h_total[0];
        *n_pairs = np;
        if (np == 0) return CB2_OK;
        if (np > cap) return CB2_ERR_NOSPC;
        if (np > kcap) {  // denser than 16 pairs per box: run again with the exact room
            kcap = np;
            continue;
        }
        if (by_segment && longest <= SEG_BLOCK_MAX) {
            // counts -> offsets, keys -> their segments, every segment sorted on its own
            tb = scan_tmp;
            SAP_CK(cub::DeviceScan::ExclusiveScan(tmp, tb, seg_cnt, seg_off, cub::Sum(), (uint64_t)0,
                                                  (int64_t)(n + 1), st));
            k_seg_scatter<<<nblk(np, 256), 256, 0, st>>>(pk, np, pos_bits, seg_off, seg_fill, pk2);
            SAP_CK(cudaMemsetAsync(flag + 2, 0, 4, st));
            k_seg_sort<<<nblk(n, 128), 128, 0, st>>>(pk2, seg_off, n, pos_bits, brute_force, pairs_out, long_list,
                                                      flag + 2);
            *launches += 3;
            if (longest > SEG_THREAD_MAX) {
                k_seg_sort_long<<<592, 256, 0, st>>>(pk2, seg_off, pos_bits, brute_force, pairs_out, long_list,
                                                      flag + 2);
                ++*launches;
            }
            SAP_CK(cudaGetLastError());
            return CB2_OK;
        }
        if (unordered && !brute_force) {
            // the caller wants the pair SET (e.g. to feed the narrow phase): no final sort
            k_sap_unpack<<<nblk(np, 256), 256, 0, st>>>(pk, np, pos_bits, false, pairs_out);
            *launches += 1;
            return CB2_OK;
        }
        // sort the packed keys (b << pos_bits | a): the reference's Vec order, then unpack
        tb = sort3;
        SAP_CK(cub::DeviceRadixSort::SortKeys(tmp, tb, pk, pk2, (int64_t)np, 0, 2 * pos_bits, st));
        k_sap_unpack<<<nblk(np, 256), 256, 0, st>>>(pk2, np, pos_bits, brute_force, pairs_out);
        *launches += 2;
        return CB2_OK;
    }
    *err = "pair count changed between passes";
    return CB2_ERR_CUDA;
}

}  // namespace cb2
