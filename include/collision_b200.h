/*
 * collision_b200.h — C ABI of the B200-native batched collision pipeline.
 *
 * This is the drop-in boundary for ONE hot path of rustgd/collision-rs 0.20.1:
 * the 3-D narrow phase GJK3 (+EPA3) and the SweepAndPrune3 broad phase.  The
 * reference has no FFI of its own (it is safe, generic Rust); each entry point
 * below replaces one generic method specialised to
 *     S = f32, P = Point3<f32>, T = Decomposed<Vector3<f32>, Quaternion<f32>>,
 *     shapes = Primitive3<f32>: Particle, Quad, Sphere, Cuboid, Cube, Cylinder, Capsule,
 *     ConvexPolyhedron (VertexOnly and HalfEdge mode)
 * and batched (one record per pair instead of one call per pair).
 * Citations are relative to the reference tree (src/...).
 *
 * Everything is plain C: POD structs, raw pointers, sizes.  No torch types.
 * All functions return 0 on success and a negative cb2_error on failure;
 * cb2_last_error(ctx) gives a human readable message.  There is NO CPU
 * fallback: without a CUDA device cb2_create fails with CB2_ERR_NO_DEVICE.
 */
#ifndef COLLISION_B200_H
#define COLLISION_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CB2_ABI_VERSION 1u

/* ---- error codes (function return values) -------------------------------- */
typedef enum cb2_error {
    CB2_OK = 0,
    CB2_ERR_NO_DEVICE = -1,    /* no CUDA device / cudaSetDevice failed          */
    CB2_ERR_CUDA = -2,         /* a CUDA runtime call failed (see last_error)    */
    CB2_ERR_ARG = -3,          /* null pointer, bad enum, index out of range     */
    CB2_ERR_NOSPC = -4,        /* pairs_out too small; *n_pairs = required count */
    CB2_ERR_UNSUPPORTED = -5,  /* e.g. max_iterations above CB2_MAX_ITERATIONS   */
    CB2_ERR_NAN = -6           /* SAP input contains NaN (reference order is
                                  unspecified for NaN keys, sweep_prune.rs:87-97) */
} cb2_error;

/* ---- per-pair status byte -------------------------------------------------
 * The reference returns Option<..> or panics.  A batch cannot panic per pair,
 * so every narrow-phase output record carries one status byte:              */
typedef enum cb2_status {
    CB2_NONE = 0,               /* Option::None                                  */
    CB2_SOME = 1,               /* Option::Some(record)                          */
    CB2_PANIC_UNWRAP_PLANE = 2, /* Plane::from_points(..).unwrap() on a degenerate
                                   face, gjk/simplex/simplex3d.rs:146            */
    CB2_PANIC_ASSERT_RANGE = 3, /* assert!(in_range(u) && ..), simplex3d.rs:150  */
    CB2_PANIC_INV_SCALE = 4,    /* inverse_transform_vector(..).unwrap() with
                                   scale ~ 0, primitive/util.rs:18, sphere.rs:36,
                                   polyhedron.rs:394                             */
    CB2_NONCONVERGED = 5,       /* the reference would never return: the time-of-impact
                                   loop (gjk/mod.rs:204, no cap) hit iter_cap, or a HalfEdge
                                   polyhedron's hill climb met a NaN direction
                                   (polyhedron.rs:184-206 loops while best_dot != previous) */
    CB2_PANIC_EPA_EMPTY = 6,    /* Polytope::closest_face_to_origin indexes
                                   faces[0] of an empty Vec, epa/epa3d.rs:138    */
    CB2_SCRATCH_OVERFLOW = 7,   /* device EPA scratch exhausted: a non-manifold horizon grew the
                                   polytope past 65 536 faces or 16 384 horizon edges (a Particle
                                   inside a Cylinder reaches 533 000 faces and takes the reference
                                   100 s for ONE pair).  Result NOT computed — loud, never a silent
                                   wrong answer; INTEGRATION.md lists it as a deliberate limit  */
    CB2_PANIC_CMP_NAN = 8,      /* intersection_complex: max_by(partial_cmp(..).unwrap())
                                   met a NaN penetration depth, gjk/mod.rs:455-459 */
    /* 9 and 10 are the 2-D path's CB2_PANIC_EPA2_EDGE / CB2_PANIC_INV_ROT below */
    CB2_BAD_INDEX = 11          /* a shape index outside the uploaded table: a Rust slice index
                                   would panic.  The *_dev entry points report it per pair (nothing
                                   is read through the index); the host-buffer entry points fail the
                                   whole call with CB2_ERR_ARG                                   */
} cb2_status;

/* ---- CollisionStrategy (contact.rs:16-22); discriminants in declaration order */
typedef enum cb2_strategy {
    CB2_FULL_RESOLUTION = 0,
    CB2_COLLISION_ONLY = 1
} cb2_strategy;

/* ---- shape kinds: every Primitive3 variant (primitive3.rs:20-40) --------------- */
typedef enum cb2_shape_kind {
    CB2_SPHERE = 0,     /* p[0] = radius                        (sphere.rs:14-24)   */
    CB2_CUBOID = 1,     /* p[0..3] = dim (full extents)         (cuboid.rs:18-42)   */
    CB2_POLYHEDRON = 2, /* vertices [vtx_off, vtx_off+vtx_cnt), VertexOnly mode
                           (polyhedron.rs:100-122)                                   */
    CB2_PARTICLE = 3,   /* Particle3: no parameters             (primitive3.rs:164) */
    CB2_QUAD = 4,       /* p[0..2] = dim x, y (in the z = 0 plane)  (quad.rs:24-57)  */
    CB2_CYLINDER = 5,   /* p[0] = half_height, p[1] = radius (axis y) (cylinder.rs:21-36) */
    CB2_CAPSULE = 6,    /* p[0] = half_height, p[1] = radius (axis y) (capsule.rs:22-37)  */
    CB2_CUBE = 7,       /* p[0] = dim: Cuboid::new(dim, dim, dim)   (cuboid.rs:144-158)  */
    CB2_POLYHEDRON_HE = 8 /* ConvexPolyhedron::new_with_faces (polyhedron.rs:124-139), HalfEdge
                           mode: support_point is the reference's hill climb over the half-edge
                           rings (:179-209), which can return a different vertex than the scan on
                           ties.  Vertices [vtx_off, +vtx_cnt); faces [f_off, +f_cnt) of the face
                           table passed to cb2_shapes_upload_faces, with f_off / f_cnt stored as
                           the uint32 BIT PATTERNS of p[0] / p[1].  Needs cb2_shapes_upload_faces. */
} cb2_shape_kind;

/* One entry of the shape table (24 B). */
typedef struct cb2_shape {
    uint32_t kind;      /* cb2_shape_kind */
    float p[3];
    uint32_t vtx_off;   /* first vertex in the vertex table (polyhedra only) */
    uint32_t vtx_cnt;
} cb2_shape;

/* cgmath Decomposed<Vector3<f32>, Quaternion<f32>> flattened (32 B):
 * transform_point(p) = rot * (p * scale) + disp. Quaternion is (s; x,y,z).   */
typedef struct cb2_xform {
    float scale;
    float qs, qx, qy, qz;
    float dx, dy, dz;
} cb2_xform;

/* Contact<Point3<f32>> (contact.rs:30-46), 36 B. */
typedef struct cb2_contact {
    uint32_t strategy;  /* cb2_strategy */
    float normal[3];
    float penetration_depth;
    float contact_point[3];
    float time_of_impact;
} cb2_contact;

/* GJK::new_with_settings arguments (gjk/mod.rs:71-84). Defaults: gjk/mod.rs:22-24,
 * epa/mod.rs:15-16 => {1e-6, 1e-6, 1e-5, 100}.                                */
typedef struct cb2_settings {
    float distance_tolerance;
    float continuous_tolerance;
    float epa_tolerance;
    uint32_t max_iterations;
} cb2_settings;

/* Largest max_iterations the 3-D entry points accept.  GJK::new_with_settings (gjk/mod.rs:71-84)
 * takes any u32; here an EPA polytope names its vertices with bytes (<= 255 vertices = 4 + 250 adds).
 * Up to 100 iterations (the crate's default MAX_ITERATIONS) every pair stays in the fast shared-
 * memory tiers; beyond that the long pairs finish in the second-chance kernel over global scratch
 * (65 536 faces, 16 384 horizon edges) — same results, slower.  The 2-D entry points keep the cap of
 * 100 (their polygon ring is sized for it).  Larger values: CB2_ERR_UNSUPPORTED.                  */
#define CB2_MAX_ITERATIONS 250u
#define CB2_MAX_ITERATIONS_2D 100u

/* Aabb3<f32> (volume/aabb/aabb3.rs:20-25): min.xyz, max.xyz (24 B). */
typedef struct cb2_aabb3 {
    float min[3];
    float max[3];
} cb2_aabb3;

/* Aabb2<f32> (volume/aabb/aabb2.rs): min.xy, max.xy (16 B). */
typedef struct cb2_aabb2 {
    float min[2];
    float max[2];
} cb2_aabb2;

/* SupportPoint<Point3<f32>> (algorithm/minkowski/mod.rs:16-23): a point of the Minkowski difference
 * and the two shape points it came from (36 B). */
typedef struct cb2_support_point {
    float v[3];
    float sup_a[3];
    float sup_b[3];
} cb2_support_point;

typedef struct cb2_ctx cb2_ctx;

/* ---- context ---------------------------------------------------------------- */
uint32_t cb2_abi_version(void);
void cb2_default_settings(cb2_settings* out);         /* GJK::new(), gjk/mod.rs:60-68 */
int cb2_create(int device, cb2_ctx** out);            /* one ctx per GPU / host thread; its *_dev
                                                          calls must be ordered on one stream */
void cb2_destroy(cb2_ctx* ctx);
const char* cb2_last_error(const cb2_ctx* ctx);       /* ctx may be NULL: create errors */
int cb2_device_sm_count(const cb2_ctx* ctx);
/* Number of this library's kernels launched through ctx since creation. */
uint64_t cb2_kernel_launches(const cb2_ctx* ctx);

/* Stage timing of the narrow phase (measurement aid for bench.py's roofline: CUDA events on the
 * launching stream around the stages of the LAST FullResolution cb2_gjk3_intersection*_dev call).
 * out4 = milliseconds of { GJK (k_gjk_hits), EPA tier 0 (k_epa), EPA higher tiers (k_epa x2),
 * second-chance (k_collect_overflow + k_retry_overflow) }.  cb2_last_stage_ms synchronises.     */
int cb2_set_stage_timing(cb2_ctx* ctx, int on);
int cb2_last_stage_ms(cb2_ctx* ctx, float* out4);

/* ---- shape table -------------------------------------------------------------
 * Replaces constructing Sphere::new / Cuboid::new / ConvexPolyhedron::new values
 * (sphere.rs:20-24, cuboid.rs:29-42, polyhedron.rs:100-122).  verts_xyz is
 * n_verts x 3 floats.  Host pointers; copied to the device.                      */
int cb2_shapes_upload(cb2_ctx* ctx, const cb2_shape* shapes, uint32_t n_shapes,
                      const float* verts_xyz, uint64_t n_verts);
/* Same, with a face table for CB2_POLYHEDRON_HE shapes: faces = n_faces x 3 vertex indices, local
 * to the owning shape's vertex range, in the order given to ConvexPolyhedron::new_with_faces (the
 * half-edge numbering, hence the hill climb's tie behaviour, depends on it).  The half-edge
 * structure is built on the host exactly as build_half_edges does (polyhedron.rs:282-380).
 * CB2_ERR_ARG where the reference would panic or misbehave at construction: a degenerate face
 * (Plane::from_points(..).unwrap(), :306-311), an index out of range, a vertex ring that does not
 * close (not a closed 2-manifold), a HalfEdge shape without faces.                              */
int cb2_shapes_upload_faces(cb2_ctx* ctx, const cb2_shape* shapes, uint32_t n_shapes,
                            const float* verts_xyz, uint64_t n_verts, const uint32_t* faces,
                            uint64_t n_faces);

/* ---- narrow phase, HOST buffers (H2D, kernels, D2H inside the call) -----------
 * out / status hold n records; out[i] is valid only where status[i] == CB2_SOME.
 * The batch streams through the device in chunks (copy-in, kernels and copy-out of different chunks
 * overlap; pinned host buffers make the copies truly asynchronous).  A shape index past the table
 * (a Rust slice index would panic) fails the whole call with CB2_ERR_ARG; it is detected chunk by
 * chunk, so the contents of out / status after a FAILED call are unspecified.
 *
 * cb2_gjk3_intersection  replaces GJK3::intersection      (gjk/mod.rs:362-391)
 * cb2_gjk3_distance      replaces GJK3::distance          (gjk/mod.rs:271-321)
 * cb2_gjk3_time_of_impact replaces GJK3::intersection_time_of_impact (gjk/mod.rs:163-256);
 *   l_t0..l_t1 is the Range<&TL> start..end, likewise r_t0..r_t1; iter_cap bounds
 *   the reference's uncapped while loop (0 => default 1024).                     */
int cb2_gjk3_intersection(cb2_ctx* ctx, uint32_t strategy, const cb2_settings* settings,
                          const uint32_t* l_shape, const uint32_t* r_shape,
                          const cb2_xform* l_t, const cb2_xform* r_t, uint64_t n,
                          cb2_contact* out, uint8_t* status);
int cb2_gjk3_distance(cb2_ctx* ctx, const cb2_settings* settings,
                      const uint32_t* l_shape, const uint32_t* r_shape,
                      const cb2_xform* l_t, const cb2_xform* r_t, uint64_t n,
                      float* out, uint8_t* status);
int cb2_gjk3_time_of_impact(cb2_ctx* ctx, const cb2_settings* settings,
                            const uint32_t* l_shape, const uint32_t* r_shape,
                            const cb2_xform* l_t0, const cb2_xform* l_t1,
                            const cb2_xform* r_t0, const cb2_xform* r_t1, uint64_t n,
                            uint32_t iter_cap, cb2_contact* out, uint8_t* status);

/* GJK::intersect (gjk/mod.rs:101-146) on its own: status CB2_SOME = Some(simplex), and
 * simplex_out[4i .. 4i+3] is the tetrahedron in the reference's Vec order (zeros otherwise).
 * GJK::get_contact_manifold (gjk/mod.rs:326-344) = EPA3::process (epa/epa3d.rs:27-65) on such a
 * simplex: simplex[4i .. 4i+3] with simplex_len[i] points actually held (NULL: four everywhere);
 * fewer than four -> CB2_NONE (epa3d.rs:38-40).  cb2_gjk3_intersect followed by
 * cb2_gjk3_get_contact_manifold is bit-identical to cb2_gjk3_intersection(FULL_RESOLUTION).      */
int cb2_gjk3_intersect(cb2_ctx* ctx, const cb2_settings* settings,
                       const uint32_t* l_shape, const uint32_t* r_shape,
                       const cb2_xform* l_t, const cb2_xform* r_t, uint64_t n,
                       cb2_support_point* simplex_out /* 4 x n */, uint8_t* status);
int cb2_gjk3_get_contact_manifold(cb2_ctx* ctx, const cb2_settings* settings,
                                  const cb2_support_point* simplex /* 4 x n */,
                                  const uint32_t* simplex_len /* n, may be NULL */,
                                  const uint32_t* l_shape, const uint32_t* r_shape,
                                  const cb2_xform* l_t, const cb2_xform* r_t, uint64_t n,
                                  cb2_contact* out, uint8_t* status);

/* ---- narrow phase, DEVICE buffers (all pointers are device pointers; the work
 * is enqueued on `stream` (a cudaStream_t, may be NULL) and NOT synchronised).
 * Every SHAPE index is checked on the device against the uploaded table: a pair that names a shape
 * outside it gets status CB2_BAD_INDEX and nothing is read through the index (the host-buffer entry
 * points fail the whole call with CB2_ERR_ARG instead).  BODY indices (pair_body) are trusted: the
 * entry points are not told how many bodies there are.
 * All *_dev calls on one context share its workspaces: issue them on ONE stream at a time (or order
 * the streams with events); two calls in flight on different streams of the same context race.
 * With pair_body != NULL the transforms are per BODY and pair i uses bodies
 * pair_body[2i], pair_body[2i+1] (shape = body_shape[..], xform = body_t[..]):
 * this is how SweepAndPrune3 output feeds the narrow phase without a host trip. */
int cb2_gjk3_intersection_dev(cb2_ctx* ctx, uint32_t strategy, const cb2_settings* settings,
                              const uint32_t* l_shape, const uint32_t* r_shape,
                              const cb2_xform* l_t, const cb2_xform* r_t, uint64_t n,
                              cb2_contact* out, uint8_t* status, void* stream);
int cb2_gjk3_intersect_dev(cb2_ctx* ctx, const cb2_settings* settings,
                           const uint32_t* l_shape, const uint32_t* r_shape,
                           const cb2_xform* l_t, const cb2_xform* r_t, uint64_t n,
                           cb2_support_point* simplex_out, uint8_t* status, void* stream);
int cb2_gjk3_get_contact_manifold_dev(cb2_ctx* ctx, const cb2_settings* settings,
                                      const cb2_support_point* simplex, const uint32_t* simplex_len,
                                      const uint32_t* l_shape, const uint32_t* r_shape,
                                      const cb2_xform* l_t, const cb2_xform* r_t, uint64_t n,
                                      cb2_contact* out, uint8_t* status, void* stream);
int cb2_gjk3_distance_dev(cb2_ctx* ctx, const cb2_settings* settings,
                          const uint32_t* l_shape, const uint32_t* r_shape,
                          const cb2_xform* l_t, const cb2_xform* r_t, uint64_t n,
                          float* out, uint8_t* status, void* stream);
int cb2_gjk3_time_of_impact_dev(cb2_ctx* ctx, const cb2_settings* settings,
                                const uint32_t* l_shape, const uint32_t* r_shape,
                                const cb2_xform* l_t0, const cb2_xform* l_t1,
                                const cb2_xform* r_t0, const cb2_xform* r_t1, uint64_t n,
                                uint32_t iter_cap, cb2_contact* out, uint8_t* status,
                                void* stream);
/* The same with HOST buffers: the body table (n_bodies shape indices and transforms) is uploaded
 * once per call and only the pairs' body indices are streamed (8 bytes per pair instead of the 72
 * of cb2_gjk3_intersection) — how a caller that iterates a broad phase's pair list over its bodies
 * holds the data (`intersection(&shape[a], &pose[a], &shape[b], &pose[b])`, gjk/mod.rs:362-391).
 * Body and shape indices are checked (CB2_ERR_ARG for the call).                               */
int cb2_gjk3_intersection_bodies(cb2_ctx* ctx, uint32_t strategy, const cb2_settings* settings,
                                 const uint32_t* body_shape, const cb2_xform* body_t,
                                 uint64_t n_bodies, const uint32_t* pair_body /* 2 x n, interleaved */,
                                 uint64_t n, cb2_contact* out, uint8_t* status);
int cb2_gjk3_intersection_bodies_dev(cb2_ctx* ctx, uint32_t strategy,
                                     const cb2_settings* settings,
                                     const uint32_t* body_shape, const cb2_xform* body_t,
                                     const uint32_t* pair_body /* 2 x n, interleaved */,
                                     uint64_t n, cb2_contact* out, uint8_t* status,
                                     void* stream);

/* ---- composite shapes ----------------------------------------------------------
 * A composite is the reference's `&[(Primitive, local transform)]`: parts
 * [part_off[c], part_off[c+1]) of (part_shape[k] = index into the shape table, part_local[k]).
 * cb2_composites_upload replaces building those slices; host pointers, copied to the device.
 *
 * cb2_gjk3_intersection_complex      replaces GJK3::intersection_complex      (gjk/mod.rs:412-461)
 * cb2_gjk3_distance_complex          replaces GJK3::distance_complex          (gjk/mod.rs:477-516)
 * cb2_gjk3_time_of_impact_complex    replaces GJK3::intersection_complex_time_of_impact (:538-588)
 * l_comp / r_comp index the composite table; l_t / r_t are the model-to-world transforms.  Each
 * part pair runs as one sub-pair of the ordinary kernels with Transform::concat applied; the
 * reduction replays the reference's loop order (first hit for CollisionOnly; max depth — last
 * among equals, Iterator::max_by; min distance; min time of impact — first among equals), incl.
 * its early returns and panics.  Host buffers.                                                 */
/* The composite table refers to the shape table it was uploaded against: uploading a new shape
 * table drops it (upload the composites again afterwards).                                        */
int cb2_composites_upload(cb2_ctx* ctx, const uint32_t* part_off /* n_comp + 1 */,
                          const uint32_t* part_shape, const cb2_xform* part_local,
                          uint32_t n_comp);
int cb2_gjk3_intersection_complex(cb2_ctx* ctx, uint32_t strategy, const cb2_settings* settings,
                                  const uint32_t* l_comp, const uint32_t* r_comp,
                                  const cb2_xform* l_t, const cb2_xform* r_t, uint64_t n,
                                  cb2_contact* out, uint8_t* status);
int cb2_gjk3_distance_complex(cb2_ctx* ctx, const cb2_settings* settings,
                              const uint32_t* l_comp, const uint32_t* r_comp,
                              const cb2_xform* l_t, const cb2_xform* r_t, uint64_t n,
                              float* out, uint8_t* status);
int cb2_gjk3_time_of_impact_complex(cb2_ctx* ctx, uint32_t strategy, const cb2_settings* settings,
                                    const uint32_t* l_comp, const uint32_t* r_comp,
                                    const cb2_xform* l_t0, const cb2_xform* l_t1,
                                    const cb2_xform* r_t0, const cb2_xform* r_t1, uint64_t n,
                                    uint32_t iter_cap, cb2_contact* out, uint8_t* status);

/* ---- broad phase ---------------------------------------------------------------
 * cb2_sap3_find_collider_pairs replaces SweepAndPrune3::find_collider_pairs
 * (algorithm/broad_phase/sweep_prune.rs:76-136) for finite extents:
 *   perm_out[k]  = original index of the box that the reference's stable sort
 *                  (sweep_prune.rs:87-97) leaves at position k (the caller applies
 *                  it to its &mut [A]); identity when n <= 1.
 *   pairs_out    = (a, b) index pairs INTO THE SORTED ORDER, a < b, in the exact
 *                  order of the reference's Vec (b ascending, then a ascending).
 *   *axis_inout  = sweep axis used (in) / axis for the next call (out; always 0
 *                  when n >= 2 — Variance3 accumulates nothing, sweep_prune.rs:254-277).
 * If cap < number of pairs: returns CB2_ERR_NOSPC with *n_pairs = required.      */
int cb2_sap3_find_collider_pairs(cb2_ctx* ctx, const cb2_aabb3* aabbs, uint64_t n,
                                 uint32_t* axis_inout, uint32_t* perm_out,
                                 uint32_t* pairs_out /* 2 x cap */, uint64_t cap,
                                 uint64_t* n_pairs);
/* Device-resident variant: aabbs/perm_out/pairs_out are device pointers, the
 * pair count is returned through the host pointer n_pairs (this call synchronises
 * `stream` once, to learn the count).                                            */
int cb2_sap3_find_collider_pairs_dev(cb2_ctx* ctx, const cb2_aabb3* aabbs, uint64_t n,
                                     uint32_t axis, uint32_t* perm_out,
                                     uint32_t* pairs_out, uint64_t cap,
                                     uint64_t* n_pairs, void* stream);

/* The same pairs as a SET: pairs_out holds exactly the reference's pairs (indices into the sorted
 * order, a < b) but in unspecified order.  For callers that hand the pairs straight to the narrow
 * phase (cb2_gjk3_intersection_bodies_dev), where order does not matter; skips the final sort.   */
int cb2_sap3_find_collider_pairs_unordered_dev(cb2_ctx* ctx, const cb2_aabb3* aabbs, uint64_t n,
                                               uint32_t axis, uint32_t* perm_out, uint32_t* pairs_out,
                                               uint64_t cap, uint64_t* n_pairs, void* stream);

/* Multi-GPU broad phase (one process per GPU): the same call restricted to shard
 * `shard` of `n_shards`.  Every GPU holds the whole AABB table and computes the whole (identical)
 * permutation; pairs are split by the sorted position of their LATER box into n_shards equal
 * position ranges (the first n % n_shards ranges hold one more), so pairs_out of shards
 * 0 .. n_shards-1 concatenated in order is exactly the reference's Vec and each shard can feed its
 * own GPU's narrow phase with no exchange.  *n_pairs / CB2_ERR_NOSPC refer to this shard only.    */
int cb2_sap3_find_collider_pairs_shard_dev(cb2_ctx* ctx, const cb2_aabb3* aabbs, uint64_t n,
                                           uint32_t axis, uint32_t shard, uint32_t n_shards,
                                           uint32_t* perm_out, uint32_t* pairs_out, uint64_t cap,
                                           uint64_t* n_pairs, void* stream);

/* SweepAndPrune2::find_collider_pairs (sweep_prune.rs:18,76-136; Variance2 :171-224): the 2-D alias
 * of the same algorithm (the one the reference's own tests exercise, :317-361).  Same contract as
 * cb2_sap3_find_collider_pairs with axis in {0, 1}.  Host buffers.                              */
int cb2_sap2_find_collider_pairs(cb2_ctx* ctx, const cb2_aabb2* aabbs, uint64_t n,
                                 uint32_t* axis_inout, uint32_t* perm_out,
                                 uint32_t* pairs_out /* 2 x cap */, uint64_t cap, uint64_t* n_pairs);

/* BruteForce::find_collider_pairs (algorithm/broad_phase/brute_force.rs:20-39): every (left, right),
 * left < right in the CALLER's indices, whose boxes overlap strictly; order = left ascending, then
 * right ascending.  Same machinery as the sweep (the O(n^2) double loop is not reproduced).
 * A box with a NaN extent intersects nothing, as in the reference (no error: BruteForce never sorts). */
int cb2_brute_force_find_collider_pairs(cb2_ctx* ctx, const cb2_aabb3* aabbs, uint64_t n,
                                        uint32_t* pairs_out /* 2 x cap */, uint64_t cap,
                                        uint64_t* n_pairs);
int cb2_brute_force_find_collider_pairs_dev(cb2_ctx* ctx, const cb2_aabb3* aabbs, uint64_t n,
                                            uint32_t* pairs_out, uint64_t cap, uint64_t* n_pairs,
                                            void* stream);

/* ---- several GPUs of one box behind one handle ---------------------------------------------------
 * "cb2_create(devices, n_dev) with the context owning the devices".  A cb2_multi holds
 * one cb2_ctx per device.  Its host-buffer calls split the batch into contiguous slices, one per
 * device, and run every device's own pipeline from its own host thread (pinned to the CPUs of the
 * GPU's NUMA node); results land in the caller's arrays exactly where a single-device call would
 * put them.  Pairs are independent, so no data is exchanged between devices; the shape table is
 * uploaded to each of them.  cb2_multi_sap3_find_collider_pairs gives every device one shard of
 * the pair list (cb2_sap3_find_collider_pairs_shard_dev) and concatenates the shards, which is the
 * reference's Vec.  cb2_multi_ctx(m, i) is device i's context for the *_dev entry points.        */
typedef struct cb2_multi cb2_multi;
int cb2_create_multi(const int* devices, int n_devices, cb2_multi** out);
void cb2_destroy_multi(cb2_multi* m);
int cb2_multi_device_count(const cb2_multi* m);
cb2_ctx* cb2_multi_ctx(cb2_multi* m, int i);
const char* cb2_multi_last_error(const cb2_multi* m);
int cb2_multi_shapes_upload(cb2_multi* m, const cb2_shape* shapes, uint32_t n_shapes,
                            const float* verts_xyz, uint64_t n_verts);
int cb2_multi_gjk3_intersection(cb2_multi* m, uint32_t strategy, const cb2_settings* settings,
                                const uint32_t* l_shape, const uint32_t* r_shape,
                                const cb2_xform* l_t, const cb2_xform* r_t, uint64_t n,
                                cb2_contact* out, uint8_t* status);
int cb2_multi_gjk3_distance(cb2_multi* m, const cb2_settings* settings,
                            const uint32_t* l_shape, const uint32_t* r_shape,
                            const cb2_xform* l_t, const cb2_xform* r_t, uint64_t n,
                            float* out, uint8_t* status);
int cb2_multi_gjk3_time_of_impact(cb2_multi* m, const cb2_settings* settings,
                                  const uint32_t* l_shape, const uint32_t* r_shape,
                                  const cb2_xform* l_t0, const cb2_xform* l_t1,
                                  const cb2_xform* r_t0, const cb2_xform* r_t1, uint64_t n,
                                  uint32_t iter_cap, cb2_contact* out, uint8_t* status);
int cb2_multi_sap3_find_collider_pairs(cb2_multi* m, const cb2_aabb3* aabbs, uint64_t n,
                                       uint32_t* axis_inout, uint32_t* perm_out,
                                       uint32_t* pairs_out /* 2 x cap */, uint64_t cap,
                                       uint64_t* n_pairs);

/* ---- world bounds (the step before the broad phase) -----------------------------
 * out[i] = shape.compute_bound().transform_volume(&t[i])
 * (ComputeBound: sphere.rs:41-51, cuboid.rs:82-92, polyhedron.rs:113-115,403-410;
 *  Aabb3::transform: volume/aabb/aabb3.rs:90-100).                               */
int cb2_world_aabbs(cb2_ctx* ctx, const uint32_t* shape, const cb2_xform* t, uint64_t n,
                    cb2_aabb3* out);
int cb2_world_aabbs_dev(cb2_ctx* ctx, const uint32_t* shape, const cb2_xform* t, uint64_t n,
                        cb2_aabb3* out, void* stream);

/* ==== 2-D narrow phase: GJK2 (+EPA2) — SURVEY §8 row N4 ============================================
 * The same generic GJK (gjk/mod.rs:101-391) instantiated with SimplexProcessor2
 * (gjk/simplex/simplex2d.rs:17-97) and EPA2 (epa/epa2d.rs:21-157), specialised to
 *     S = f32, P = Point2<f32>, T = Decomposed<Vector2<f32>, Basis2<f32>>,
 *     shapes = Primitive2<f32>: Particle, Line, Circle, Rectangle, Square, ConvexPolygon.
 * Status bytes as above; two panic sites exist only here:                                         */
#define CB2_PANIC_EPA2_EDGE 9u    /* closest_edge: assert_ulps_ne!(inf, edge.distance) — no edge had
                                     a distance below infinity (NaN simplex), epa/epa2d.rs:154      */
#define CB2_PANIC_INV_ROT 10u     /* Basis2::invert = Matrix2::invert().unwrap() on a singular (or
                                     non-finite-determinant-zero) rotation matrix (cgmath 0.18
                                     rotation.rs; called by Decomposed::inverse_transform_vector)   */

/* Primitive2 variants in declaration order (primitive/primitive2.rs:20-33). */
typedef enum cb2_shape2_kind {
    CB2_PARTICLE2 = 0, /* support = transform_point(origin)              (primitive2.rs:96)       */
    CB2_LINE2 = 1,     /* p[0..2] = origin.xy, p[2..4] = dest.xy          (primitive/line.rs:7-26)  */
    CB2_CIRCLE = 2,    /* p[0] = radius                                   (circle.rs:14-41)         */
    CB2_RECTANGLE = 3, /* p[0..2] = dim (full extents)                    (rectangle.rs:18-77)      */
    CB2_SQUARE = 4,    /* p[0] = dim: Rectangle::new(dim, dim)            (rectangle.rs:122-160)    */
    CB2_POLYGON = 5    /* vertices [vtx_off, +vtx_cnt) of the 2-D vertex table: get_max_point scan
                          below 10 vertices, the neighbour walk from 10 up (polygon.rs:26-43,122-190) */
} cb2_shape2_kind;

/* One entry of the 2-D shape table (32 B). */
typedef struct cb2_shape2 {
    uint32_t kind; /* cb2_shape2_kind */
    float p[4];
    uint32_t vtx_off;
    uint32_t vtx_cnt;
    uint32_t reserved; /* 0 */
} cb2_shape2;

/* cgmath Decomposed<Vector2<f32>, Basis2<f32>> flattened (32 B).  Basis2 wraps a column-major
 * Matrix2: Rotation2::from_angle(t) = [c0 = (cos t, sin t), c1 = (-sin t, cos t)];
 * transform_point(p) = mat * (p * scale) + disp.                                                 */
typedef struct cb2_xform2 {
    float scale;
    float c0x, c0y, c1x, c1y;
    float dx, dy;
    float reserved; /* 0 */
} cb2_xform2;

/* Contact<Point2<f32>> (contact.rs:30-46), 28 B. */
typedef struct cb2_contact2 {
    uint32_t strategy; /* cb2_strategy */
    float normal[2];
    float penetration_depth;
    float contact_point[2];
    float time_of_impact;
} cb2_contact2;

/* Replaces constructing Circle::new / Rectangle::new / Square::new / ConvexPolygon::new /
 * Line2::new values.  verts_xy is n_verts x 2 floats.  Host pointers; copied to the device.
 * The 2-D table is independent of the 3-D one.                                                    */
int cb2_shapes2_upload(cb2_ctx* ctx, const cb2_shape2* shapes, uint32_t n_shapes,
                       const float* verts_xy, uint64_t n_verts);

/* cb2_gjk2_intersection    replaces GJK2::intersection (gjk/mod.rs:362-391 -> EPA2::process)
 * cb2_gjk2_distance        replaces GJK2::distance     (gjk/mod.rs:271-321)
 * cb2_gjk2_time_of_impact  replaces GJK2::intersection_time_of_impact (gjk/mod.rs:163-256)
 * Host buffers; same record/status contract as the 3-D entry points.                              */
int cb2_gjk2_intersection(cb2_ctx* ctx, uint32_t strategy, const cb2_settings* settings,
                          const uint32_t* l_shape, const uint32_t* r_shape,
                          const cb2_xform2* l_t, const cb2_xform2* r_t, uint64_t n,
                          cb2_contact2* out, uint8_t* status);
int cb2_gjk2_distance(cb2_ctx* ctx, const cb2_settings* settings,
                      const uint32_t* l_shape, const uint32_t* r_shape,
                      const cb2_xform2* l_t, const cb2_xform2* r_t, uint64_t n,
                      float* out, uint8_t* status);
int cb2_gjk2_time_of_impact(cb2_ctx* ctx, const cb2_settings* settings,
                            const uint32_t* l_shape, const uint32_t* r_shape,
                            const cb2_xform2* l_t0, const cb2_xform2* l_t1,
                            const cb2_xform2* r_t0, const cb2_xform2* r_t1, uint64_t n,
                            uint32_t iter_cap, cb2_contact2* out, uint8_t* status);
/* Device-pointer variants: enqueued on `stream`, not synchronised. */
int cb2_gjk2_intersection_dev(cb2_ctx* ctx, uint32_t strategy, const cb2_settings* settings,
                              const uint32_t* l_shape, const uint32_t* r_shape,
                              const cb2_xform2* l_t, const cb2_xform2* r_t, uint64_t n,
                              cb2_contact2* out, uint8_t* status, void* stream);
int cb2_gjk2_distance_dev(cb2_ctx* ctx, const cb2_settings* settings,
                          const uint32_t* l_shape, const uint32_t* r_shape,
                          const cb2_xform2* l_t, const cb2_xform2* r_t, uint64_t n,
                          float* out, uint8_t* status, void* stream);
int cb2_gjk2_time_of_impact_dev(cb2_ctx* ctx, const cb2_settings* settings,
                                const uint32_t* l_shape, const uint32_t* r_shape,
                                const cb2_xform2* l_t0, const cb2_xform2* l_t1,
                                const cb2_xform2* r_t0, const cb2_xform2* r_t1, uint64_t n,
                                uint32_t iter_cap, cb2_contact2* out, uint8_t* status,
                                void* stream);

/* 2-D composite shapes: GJK2::intersection_complex / distance_complex /
 * intersection_complex_time_of_impact (the same generic callers, gjk/mod.rs:412-588, over
 * `&[(Primitive2, Decomposed<Vector2, Basis2>)]`).  Same table layout, loop-order reduction, early
 * returns and panics as the 3-D entry points above; the table indexes the 2-D shape table.        */
int cb2_composites2_upload(cb2_ctx* ctx, const uint32_t* part_off /* n_comp + 1 */,
                           const uint32_t* part_shape, const cb2_xform2* part_local, uint32_t n_comp);
int cb2_gjk2_intersection_complex(cb2_ctx* ctx, uint32_t strategy, const cb2_settings* settings,
                                  const uint32_t* l_comp, const uint32_t* r_comp,
                                  const cb2_xform2* l_t, const cb2_xform2* r_t, uint64_t n,
                                  cb2_contact2* out, uint8_t* status);
int cb2_gjk2_distance_complex(cb2_ctx* ctx, const cb2_settings* settings,
                              const uint32_t* l_comp, const uint32_t* r_comp,
                              const cb2_xform2* l_t, const cb2_xform2* r_t, uint64_t n,
                              float* out, uint8_t* status);
int cb2_gjk2_time_of_impact_complex(cb2_ctx* ctx, uint32_t strategy, const cb2_settings* settings,
                                    const uint32_t* l_comp, const uint32_t* r_comp,
                                    const cb2_xform2* l_t0, const cb2_xform2* l_t1,
                                    const cb2_xform2* r_t0, const cb2_xform2* r_t1, uint64_t n,
                                    uint32_t iter_cap, cb2_contact2* out, uint8_t* status);

/* 2-D world bounds, the step before SweepAndPrune2: out[i] = shape.compute_bound().transform_volume(&t[i])
 * (ComputeBound<Aabb2> for Primitive2: primitive2.rs:68-80; Aabb2::transform: volume/aabb/aabb2.rs:78-88). */
int cb2_world_aabbs2(cb2_ctx* ctx, const uint32_t* shape, const cb2_xform2* t, uint64_t n, cb2_aabb2* out);
int cb2_world_aabbs2_dev(cb2_ctx* ctx, const uint32_t* shape, const cb2_xform2* t, uint64_t n,
                         cb2_aabb2* out, void* stream);

/* The 2-D chain on the device, mirroring cb2_sap3_find_collider_pairs_dev -> cb2_gjk3_intersection_bodies_dev:
 * SweepAndPrune2 on device boxes (same contract as the 3-D call, axis in {0, 1}), then GJK2 on the
 * pairs, which index one body table (shape = body_shape[..], transform = body_t[..]).              */
int cb2_sap2_find_collider_pairs_dev(cb2_ctx* ctx, const cb2_aabb2* aabbs, uint64_t n, uint32_t axis,
                                     uint32_t* perm_out, uint32_t* pairs_out, uint64_t cap,
                                     uint64_t* n_pairs, void* stream);
int cb2_gjk2_intersection_bodies_dev(cb2_ctx* ctx, uint32_t strategy, const cb2_settings* settings,
                                     const uint32_t* body_shape, const cb2_xform2* body_t,
                                     const uint32_t* pair_body /* 2 x n, interleaved */, uint64_t n,
                                     cb2_contact2* out, uint8_t* status, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* COLLISION_B200_H */
