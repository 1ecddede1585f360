// cb2_kernels.h — internal launch interface between the C ABI (cb2_api.cu) and the kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/collision_b200.h"

namespace cb2 {

struct dshape;

// the device-resident shape table (mirrors cb2::shape_table in cb2_gjk.cuh)
struct shape_table_args {
    const dshape* shapes;     // 32-byte entries
    const float4* verts;      // polyhedron vertices (x,y,z,ring offset) + HalfEdge ring records
    const uint4* cells;       // support-cell records (cb2_cells.cuh)
    const uint16_t* ext;      // support-cell extension pool
};

// All pointers are device pointers.
struct narrow_args {
    shape_table_args T;
    const uint32_t* l_shape;  // per pair; per BODY when pair_body != nullptr
    const uint32_t* r_shape;
    const float4* l_t;        // cb2_xform records as 2 x float4 (start pose for time of impact)
    const float4* r_t;
    const float4* l_t1;       // end poses (time of impact only)
    const float4* r_t1;
    const uint32_t* pair_body;  // optional 2 x n body indices (then l_t == r_t == body_t)
    const uint32_t* perm;       // optional: process pairs perm[0 .. *dev_count) (classified batches)
    const uint32_t* dev_count;
    uint64_t n;
    uint32_t n_shapes;          // entries of the shape table: every shape index is checked against it
    uint64_t n_bodies;          // entries of the body table when pair_body is set and the caller wants the
                                // body indices checked as well (0: trusted, the device-buffer contract)
    uint32_t* bad_index;        // optional device flag, set when a pair's shape index is out of range
    uint32_t* queue;            // device work-queue cursor of the persistent (lane-refill) kernels
    cb2_settings settings;
    uint32_t iter_cap;
    cb2_contact* out_contact;
    float* out_distance;
    uint8_t* out_status;
    // GJK::intersect / get_contact_manifold exported on their own (gjk/mod.rs:101-146, 326-344)
    cb2_support_point* out_simplex;        // 4 records per pair
    const cb2_support_point* in_simplex;   // 4 records per pair
    const uint32_t* in_simplex_len;        // optional: points actually held (EPA3 returns None below 4)
    const float4* seed_simplex;            // set for get_contact_manifold: the second-chance kernel then
                                           // starts from ws.simplex instead of running GJK again
};

struct aabb_args {
    shape_table_args T;
    const uint32_t* shape;
    const float4* t;
    uint64_t n;
    cb2_aabb3* out;
};

cudaError_t launch_intersection(const narrow_args& A, bool full, cudaStream_t st);

// ---- second chance for pairs whose EPA polytope outgrew the fast kernels' scratch ----------------
constexpr int RETRY_FACES = 65536;
constexpr int RETRY_EDGES = 16384;
constexpr int RETRY_VERTS = 256;
constexpr int RETRY_BLOCKS = 64;
constexpr uint32_t RETRY_LIST_CAP = 1u << 16;
constexpr size_t RETRY_BYTES_PER_BLOCK =  // fnd | pv | pa | fidx | vmask | edges
    (size_t)RETRY_FACES * 16 + 2 * (size_t)RETRY_VERTS * 16 + (size_t)RETRY_FACES * 4 + RETRY_FACES / 8 +
    (size_t)RETRY_EDGES * 2 + 256;
struct retry_ws {
    uint32_t* list;   // RETRY_LIST_CAP
    uint32_t* count;  // [0] = overflow count, [1] = work cursor
    char* scratch;    // retry_scratch_bytes()
};
size_t retry_scratch_bytes();
cudaError_t launch_overflow_retry(const narrow_args& A, const retry_ws& R, cudaStream_t st);

// ---- GJK -> EPA pipeline (cb2_narrow.cu: k_gjk_hits, cb2_epa.cu: k_epa) ------------------------
constexpr int EPA_TIERS = 3;
struct epa_counters {             // device-side bookkeeping, zeroed per batch
    uint32_t n_list[EPA_TIERS];   // entries queued for each tier ([0] = FullResolution hits; the
                                  // last tier also receives curved pairs straight from k_gjk_hits)
    uint32_t next[EPA_TIERS];     // work-queue cursors
};
struct narrow_ws {                // per-stream device workspace, owned by the ctx
    uint32_t* list[EPA_TIERS];    // n words each: pair indices (| OVER_RESUME when a dump exists)
    float4* simplex;              // 8 x float4 per pair: GJK terminal simplex v[0..3], sup_a[0..3]
    float4* pa;                   // SupportPoint.sup_a scratch, per resident EPA group
    float4* dump[EPA_TIERS - 1];  // polytope dumps of tier k, resumed by tier k+1
    uint32_t dump_cap[EPA_TIERS - 1];
    epa_counters* EC;
};
size_t epa_dump_record_bytes(int tier);
size_t epa_pa_scratch_bytes(int sm_count);
cudaError_t launch_gjk_hits(const narrow_args& A, bool full, const narrow_ws& W, int sm_count,
                            cudaStream_t st);
cudaError_t launch_gjk_intersect(const narrow_args& A, cudaStream_t st);
cudaError_t launch_manifold_seed(const narrow_args& A, const narrow_ws& W, cudaStream_t st);
cudaError_t launch_epa_tier(int tier, const narrow_args& A, const narrow_ws& W, int sm_count,
                            cudaStream_t st);

// ---- composite shapes (cb2_complex.cu) -----------------------------------------------------------
struct complex_args {
    bool dim2;  // the 2-D instantiation: cb2_xform2 transforms (same 32-byte stride), cb2_contact2 records
    // composite table: parts [comp_off[c], comp_off[c+1]) = (comp_shape, comp_local)
    const uint32_t* comp_off;
    const uint32_t* comp_shape;
    const float4* comp_local;
    // composite pairs
    const uint32_t* l_comp;
    const uint32_t* r_comp;
    const float4* l_t0;
    const float4* r_t0;
    const float4* l_t1;  // time of impact only (else nullptr)
    const float4* r_t1;
    uint64_t n;
    // expansion
    uint64_t* sub_count;  // n
    uint64_t* sub_off;    // n (exclusive scan of sub_count)
    uint32_t* sub_l_shape;
    uint32_t* sub_r_shape;
    float4* sub_l_t0;
    float4* sub_r_t0;
    float4* sub_l_t1;
    float4* sub_r_t1;
    // sub-pair results
    const void* sub_contact;     // cb2_contact (36 B) or cb2_contact2 (28 B) records
    const float* sub_distance;
    const uint8_t* sub_status;
    // composite results
    void* out_contact;
    float* out_distance;
    uint8_t* out_status;
};
cudaError_t cx_scan(void* tmp, size_t* tmp_bytes, const uint64_t* in, uint64_t* out, uint64_t n,
                    cudaStream_t st);
cudaError_t launch_cx_count(const complex_args& X, cudaStream_t st);
cudaError_t launch_cx_expand(const complex_args& X, cudaStream_t st);
cudaError_t launch_cx_reduce(const complex_args& X, uint32_t op, uint32_t strategy, cudaStream_t st);

// ---- support cells (cb2_cells.cu) ---------------------------------------------------------------
uint32_t cells_resolution(uint32_t vertex_count);
cudaError_t build_support_cells(const dshape* d_shapes, const float4* d_verts, uint32_t n_shapes,
                                uint32_t max_N, uint4* d_cells, uint16_t* d_ext, uint32_t ext_cap,
                                uint32_t* d_ext_cursor, cudaStream_t st);
cudaError_t launch_distance(const narrow_args& A, int sm_count, cudaStream_t st);
cudaError_t launch_time_of_impact(const narrow_args& A, int sm_count, cudaStream_t st);
cudaError_t launch_world_aabbs(const aabb_args& A, cudaStream_t st);

// ---- broad phase (cb2_sap.cu) ---------------------------------------------------------------
struct sap_workspace;  // opaque, owned by the ctx
sap_workspace* sap_workspace_create();
void sap_workspace_destroy(sap_workspace* w);
// Device-resident SweepAndPrune3::find_collider_pairs.  Returns 0 / cb2_error; *n_pairs is the
// true pair count (pairs beyond cap are not written).  launches += kernels launched.
int sap_find_pairs(sap_workspace* w, const cb2_aabb3* aabbs, uint64_t n, uint32_t axis,
                   uint32_t* perm_out, uint32_t* pairs_out, uint64_t cap, uint64_t* n_pairs,
                   cudaStream_t st, uint64_t* launches, const char** err, bool brute_force = false,
                   uint32_t b_lo = 0u, uint32_t b_hi = 0xFFFFFFFFu,  // [b_lo, b_hi): this GPU's shard
                   bool unordered = false);  // the pair SET in emission order: skips the final pair sort

// ---- 2-D narrow phase (cb2_narrow2d.cu): GJK2 / EPA2, one thread per pair -------------------------
struct narrow2_args {           // all device pointers
    const cb2_shape2* shapes;   // 32-byte entries
    const float* verts;         // x,y pairs
    const uint32_t* l_shape;
    const uint32_t* r_shape;
    const cb2_xform2* l_t;      // start poses for time of impact
    const cb2_xform2* r_t;
    const cb2_xform2* l_t1;     // end poses (time of impact only)
    const cb2_xform2* r_t1;
    const uint32_t* pair_body;  // optional 2 x n body indices (intersection only): l_shape / l_t per body
    uint64_t n;
    cb2_settings settings;
    uint32_t iter_cap;
    uint32_t strategy;
    cb2_contact2* out_contact;
    float* out_distance;
    uint8_t* out_status;
    // GJK -> EPA2 hand-off workspace (FullResolution intersection only; owned by the ctx)
    uint32_t* counters;  // 8 words
    uint32_t* hits;      // 3 n words: pair per hit | small-tier queue | large-tier queue
    void* simplex;       // n x 48 bytes
    bool long_tails;     // the shape table mixes kinds whose distance loops differ by 10-50x in length
                         // (circles against anything else run to the iteration cap; long polygon
                         // walks): use the lane-refill kernels.  Uniform tables finish sooner without.
};
// op: 0 intersection (k_gjk2_intersection [+ the two k_epa2 tiers]), 1 distance, 2 time of impact
cudaError_t launch_narrow2(int op, const narrow2_args& A, int sm_count, cudaStream_t st);
cudaError_t launch_lift_aabb2(const cb2_aabb2* in, uint64_t n, cb2_aabb3* out, cudaStream_t st);
cudaError_t launch_world_aabbs2(const cb2_shape2* shapes, const float* verts, const uint32_t* shape,
                                const cb2_xform2* t, uint64_t n, cb2_aabb2* out, cudaStream_t st);

}  // namespace cb2
