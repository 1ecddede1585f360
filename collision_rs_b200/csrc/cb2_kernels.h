// cb2_kernels.h — internal launch interface between the C ABI (cb2_api.cu) and the kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/collision_b200.h"

namespace cb2 {

struct dshape;

// All pointers are device pointers.
struct narrow_args {
    const dshape* shapes;     // device shape table (32-byte entries)
    const float4* verts;      // polyhedron vertices, (x,y,z,0)
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
    cb2_settings settings;
    uint32_t iter_cap;
    cb2_contact* out_contact;
    float* out_distance;
    uint8_t* out_status;
};

struct aabb_args {
    const dshape* shapes;
    const float4* verts;
    const uint32_t* shape;
    const float4* t;
    uint64_t n;
    cb2_aabb3* out;
};

cudaError_t launch_intersection(const narrow_args& A, bool full, cudaStream_t st);

// ---- second chance for pairs whose EPA polytope outgrew the fast kernels' scratch ----------------
constexpr int RETRY_FACES = 16384;
constexpr int RETRY_EDGES = 4096;
constexpr int RETRY_BLOCKS = 64;
constexpr uint32_t RETRY_LIST_CAP = 1u << 16;
constexpr size_t RETRY_BYTES_PER_BLOCK =
    (size_t)RETRY_FACES * 20 + 2 * 104 * 12 + (size_t)RETRY_EDGES * 2 + 64;
struct retry_ws {
    uint32_t* list;   // RETRY_LIST_CAP
    uint32_t* count;  // [0] = overflow count, [1] = work cursor
    char* scratch;    // retry_scratch_bytes()
};
size_t retry_scratch_bytes();
cudaError_t launch_overflow_retry(const narrow_args& A, const retry_ws& R, cudaStream_t st);

// ---- classified (polyhedron-aware) intersection, cb2_poly.cu ----------------------------------
constexpr int POLY_NC = 11;                  // size classes (0 = no polyhedron in the pair)
constexpr int POLY_SUB = 16;                 // (Vl, Vr) sub-buckets per class
constexpr int POLY_NB = POLY_NC * POLY_SUB;  // counting-sort buckets (fits the u8 key)
struct poly_counters {                       // device-side bookkeeping, zeroed per batch
    uint32_t hist[POLY_NB];
    uint32_t start[POLY_NB];
    uint32_t cursor[POLY_NB];
    uint32_t class_begin[POLY_NC + 1];
    uint32_t gjk_next[POLY_NC];              // work-queue cursors
    uint32_t hit_count[POLY_NC];
    uint32_t epa_next[POLY_NC];
    uint32_t over_count;
    uint32_t over_next;
};
struct narrow_ws {       // per-stream device workspace, owned by the ctx
    uint8_t* key;        // n bytes: bucket of each pair
    uint32_t* perm;      // n words: pairs sorted by bucket
    uint32_t* hits;      // n words: queue positions of FullResolution hits, per class range
    uint32_t* over;      // n words: queue positions whose polytope outgrew the tier-1 scratch
    float4* simplex;     // 6 x float4 per queue position (GJK terminal simplex -> EPA)
    poly_counters* PC;
};
cudaError_t launch_intersection_classified(const narrow_args& A, bool full, const narrow_ws& W,
                                           int sm_count, uint32_t max_vcnt, cudaStream_t st,
                                           uint64_t* launches);
cudaError_t launch_distance(const narrow_args& A, cudaStream_t st);
cudaError_t launch_time_of_impact(const narrow_args& A, cudaStream_t st);
cudaError_t launch_world_aabbs(const aabb_args& A, cudaStream_t st);

// ---- broad phase (cb2_sap.cu) ---------------------------------------------------------------
struct sap_workspace;  // opaque, owned by the ctx
sap_workspace* sap_workspace_create();
void sap_workspace_destroy(sap_workspace* w);
// Device-resident SweepAndPrune3::find_collider_pairs.  Returns 0 / cb2_error; *n_pairs is the
// true pair count (pairs beyond cap are not written).  launches += kernels launched.
int sap_find_pairs(sap_workspace* w, const cb2_aabb3* aabbs, uint64_t n, uint32_t axis,
                   uint32_t* perm_out, uint32_t* pairs_out, uint64_t cap, uint64_t* n_pairs,
                   cudaStream_t st, uint64_t* launches, const char** err);

}  // namespace cb2
