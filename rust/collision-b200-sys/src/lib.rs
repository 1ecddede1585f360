//! Raw `extern "C"` declarations of `include/collision_b200.h` (ABI version 1).
//!
//! UNCOMPILED SKETCH: the build image of this repository has no Rust toolchain, so neither this
//! crate nor `collision-b200` has ever been through `cargo check`.  The structs are `#[repr(C)]`
//! mirrors of the header; `tests/test_abi_cpu.py` checks that every function and struct of the
//! header is declared here and that the field counts give the header's sizes, and the `const _`
//! assertions below make `rustc` refuse to build if a layout drifts.
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int, c_void};

pub const CB2_ABI_VERSION: u32 = 1;
pub const CB2_MAX_ITERATIONS: u32 = 250;
pub const CB2_MAX_ITERATIONS_2D: u32 = 100;

// cb2_error
pub const CB2_OK: c_int = 0;
pub const CB2_ERR_NO_DEVICE: c_int = -1;
pub const CB2_ERR_CUDA: c_int = -2;
pub const CB2_ERR_ARG: c_int = -3;
pub const CB2_ERR_NOSPC: c_int = -4;
pub const CB2_ERR_UNSUPPORTED: c_int = -5;
pub const CB2_ERR_NAN: c_int = -6;

// cb2_status (one byte per pair)
pub const CB2_NONE: u8 = 0;
pub const CB2_SOME: u8 = 1;
pub const CB2_PANIC_UNWRAP_PLANE: u8 = 2;
pub const CB2_PANIC_ASSERT_RANGE: u8 = 3;
pub const CB2_PANIC_INV_SCALE: u8 = 4;
pub const CB2_NONCONVERGED: u8 = 5;
pub const CB2_PANIC_EPA_EMPTY: u8 = 6;
pub const CB2_SCRATCH_OVERFLOW: u8 = 7;
pub const CB2_PANIC_CMP_NAN: u8 = 8;
pub const CB2_PANIC_EPA2_EDGE: u8 = 9;
pub const CB2_PANIC_INV_ROT: u8 = 10;
pub const CB2_BAD_INDEX: u8 = 11;
// cb2_shape2_kind: Primitive2 variants in declaration order
pub const CB2_PARTICLE2: u32 = 0;
pub const CB2_LINE2: u32 = 1;
pub const CB2_CIRCLE: u32 = 2;
pub const CB2_RECTANGLE: u32 = 3;
pub const CB2_SQUARE: u32 = 4;
pub const CB2_POLYGON: u32 = 5;

// cb2_strategy: contact.rs:16-22 discriminants in declaration order
pub const CB2_FULL_RESOLUTION: u32 = 0;
pub const CB2_COLLISION_ONLY: u32 = 1;

// cb2_shape_kind
pub const CB2_SPHERE: u32 = 0;
pub const CB2_CUBOID: u32 = 1;
pub const CB2_POLYHEDRON: u32 = 2;
pub const CB2_PARTICLE: u32 = 3;
pub const CB2_QUAD: u32 = 4;      // p = [dim.x, dim.y, _]
pub const CB2_CYLINDER: u32 = 5;  // p = [half_height, radius, _]
pub const CB2_CAPSULE: u32 = 6;   // p = [half_height, radius, _]
pub const CB2_CUBE: u32 = 7;      // p = [dim, _, _]
pub const CB2_POLYHEDRON_HE: u32 = 8; // new_with_faces; p[0], p[1] = f32::from_bits(face_off / face_cnt)

#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct cb2_shape {
    pub kind: u32,
    pub p: [f32; 3],
    pub vtx_off: u32,
    pub vtx_cnt: u32,
}

/// cgmath `Decomposed<Vector3<f32>, Quaternion<f32>>` flattened (32 B).
#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct cb2_xform {
    pub scale: f32,
    pub qs: f32,
    pub qx: f32,
    pub qy: f32,
    pub qz: f32,
    pub dx: f32,
    pub dy: f32,
    pub dz: f32,
}

/// `Contact<Point3<f32>>` (contact.rs:30-46), 36 B.
#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct cb2_contact {
    pub strategy: u32,
    pub normal: [f32; 3],
    pub penetration_depth: f32,
    pub contact_point: [f32; 3],
    pub time_of_impact: f32,
}

#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct cb2_settings {
    pub distance_tolerance: f32,
    pub continuous_tolerance: f32,
    pub epa_tolerance: f32,
    pub max_iterations: u32,
}

#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct cb2_aabb3 {
    pub min: [f32; 3],
    pub max: [f32; 3],
}

#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct cb2_aabb2 {
    pub min: [f32; 2],
    pub max: [f32; 2],
}

/// `SupportPoint<Point3<f32>>` (algorithm/minkowski/mod.rs:16-23), 36 B.
#[repr(C)]
#[derive(Clone, Copy, Debug, Default, PartialEq)]
pub struct cb2_support_point {
    pub v: [f32; 3],
    pub sup_a: [f32; 3],
    pub sup_b: [f32; 3],
}

#[repr(C)]
pub struct cb2_ctx {
    _private: [u8; 0],
}

/// Several GPUs of one box behind one handle (`cb2_create_multi`).
#[repr(C)]
pub struct cb2_multi {
    _private: [u8; 0],
}

// layouts the C header promises (include/collision_b200.h); a drift fails the build
const _: () = assert!(std::mem::size_of::<cb2_shape>() == 24);
const _: () = assert!(std::mem::size_of::<cb2_xform>() == 32);
const _: () = assert!(std::mem::size_of::<cb2_contact>() == 36);
const _: () = assert!(std::mem::size_of::<cb2_settings>() == 16);
const _: () = assert!(std::mem::size_of::<cb2_aabb3>() == 24);
const _: () = assert!(std::mem::size_of::<cb2_aabb2>() == 16);
const _: () = assert!(std::mem::size_of::<cb2_support_point>() == 36);
const _: () = assert!(std::mem::size_of::<cb2_shape2>() == 32);
const _: () = assert!(std::mem::size_of::<cb2_xform2>() == 32);
const _: () = assert!(std::mem::size_of::<cb2_contact2>() == 28);

/// One entry of the 2-D shape table (32 B); `kind` = Primitive2 variant in declaration order
/// (0 Particle, 1 Line, 2 Circle, 3 Rectangle, 4 Square, 5 ConvexPolygon).
#[repr(C)]
#[derive(Clone, Copy, Debug, Default, PartialEq)]
pub struct cb2_shape2 {
    pub kind: u32,
    pub p: [f32; 4],
    pub vtx_off: u32,
    pub vtx_cnt: u32,
    pub reserved: u32,
}

/// `Decomposed<Vector2<f32>, Basis2<f32>>` flattened (32 B): scale, the column-major Matrix2
/// behind `Basis2` (`AsRef<Matrix2<f32>>`: c0.x, c0.y, c1.x, c1.y), displacement.
#[repr(C)]
#[derive(Clone, Copy, Debug, Default, PartialEq)]
pub struct cb2_xform2 {
    pub scale: f32,
    pub c0x: f32,
    pub c0y: f32,
    pub c1x: f32,
    pub c1y: f32,
    pub dx: f32,
    pub dy: f32,
    pub reserved: f32,
}

/// `Contact<Point2<f32>>` (28 B).
#[repr(C)]
#[derive(Clone, Copy, Debug, Default, PartialEq)]
pub struct cb2_contact2 {
    pub strategy: u32,
    pub normal: [f32; 2],
    pub penetration_depth: f32,
    pub contact_point: [f32; 2],
    pub time_of_impact: f32,
}

extern "C" {
    pub fn cb2_abi_version() -> u32;
    pub fn cb2_default_settings(out: *mut cb2_settings);
    pub fn cb2_create(device: c_int, out: *mut *mut cb2_ctx) -> c_int;
    pub fn cb2_destroy(ctx: *mut cb2_ctx);
    pub fn cb2_last_error(ctx: *const cb2_ctx) -> *const c_char;
    pub fn cb2_device_sm_count(ctx: *const cb2_ctx) -> c_int;
    pub fn cb2_kernel_launches(ctx: *const cb2_ctx) -> u64;
    pub fn cb2_set_stage_timing(ctx: *mut cb2_ctx, on: c_int) -> c_int;
    pub fn cb2_last_stage_ms(ctx: *mut cb2_ctx, out4: *mut f32) -> c_int;

    pub fn cb2_shapes_upload(ctx: *mut cb2_ctx, shapes: *const cb2_shape, n_shapes: u32,
                             verts_xyz: *const f32, n_verts: u64) -> c_int;

    pub fn cb2_shapes_upload_faces(ctx: *mut cb2_ctx, shapes: *const cb2_shape, n_shapes: u32,
                                   verts_xyz: *const f32, n_verts: u64, faces: *const u32,
                                   n_faces: u64) -> c_int;

    pub fn cb2_gjk3_intersection(ctx: *mut cb2_ctx, strategy: u32, settings: *const cb2_settings,
                                 l_shape: *const u32, r_shape: *const u32,
                                 l_t: *const cb2_xform, r_t: *const cb2_xform, n: u64,
                                 out: *mut cb2_contact, status: *mut u8) -> c_int;
    pub fn cb2_gjk3_distance(ctx: *mut cb2_ctx, settings: *const cb2_settings,
                             l_shape: *const u32, r_shape: *const u32,
                             l_t: *const cb2_xform, r_t: *const cb2_xform, n: u64,
                             out: *mut f32, status: *mut u8) -> c_int;
    pub fn cb2_gjk3_time_of_impact(ctx: *mut cb2_ctx, settings: *const cb2_settings,
                                   l_shape: *const u32, r_shape: *const u32,
                                   l_t0: *const cb2_xform, l_t1: *const cb2_xform,
                                   r_t0: *const cb2_xform, r_t1: *const cb2_xform, n: u64,
                                   iter_cap: u32, out: *mut cb2_contact, status: *mut u8) -> c_int;

    pub fn cb2_gjk3_intersection_dev(ctx: *mut cb2_ctx, strategy: u32, settings: *const cb2_settings,
                                     l_shape: *const u32, r_shape: *const u32,
                                     l_t: *const cb2_xform, r_t: *const cb2_xform, n: u64,
                                     out: *mut cb2_contact, status: *mut u8, stream: *mut c_void) -> c_int;
    pub fn cb2_gjk3_distance_dev(ctx: *mut cb2_ctx, settings: *const cb2_settings,
                                 l_shape: *const u32, r_shape: *const u32,
                                 l_t: *const cb2_xform, r_t: *const cb2_xform, n: u64,
                                 out: *mut f32, status: *mut u8, stream: *mut c_void) -> c_int;
    pub fn cb2_gjk3_time_of_impact_dev(ctx: *mut cb2_ctx, settings: *const cb2_settings,
                                       l_shape: *const u32, r_shape: *const u32,
                                       l_t0: *const cb2_xform, l_t1: *const cb2_xform,
                                       r_t0: *const cb2_xform, r_t1: *const cb2_xform, n: u64,
                                       iter_cap: u32, out: *mut cb2_contact, status: *mut u8,
                                       stream: *mut c_void) -> c_int;
    pub fn cb2_gjk3_intersection_bodies(ctx: *mut cb2_ctx, strategy: u32, settings: *const cb2_settings,
                                        body_shape: *const u32, body_t: *const cb2_xform, n_bodies: u64,
                                        pair_body: *const u32, n: u64, out: *mut cb2_contact,
                                        status: *mut u8) -> c_int;
    pub fn cb2_gjk3_intersection_bodies_dev(ctx: *mut cb2_ctx, strategy: u32,
                                            settings: *const cb2_settings,
                                            body_shape: *const u32, body_t: *const cb2_xform,
                                            pair_body: *const u32, n: u64,
                                            out: *mut cb2_contact, status: *mut u8,
                                            stream: *mut c_void) -> c_int;

    pub fn cb2_gjk3_intersect(ctx: *mut cb2_ctx, settings: *const cb2_settings,
                              l_shape: *const u32, r_shape: *const u32,
                              l_t: *const cb2_xform, r_t: *const cb2_xform, n: u64,
                              simplex_out: *mut cb2_support_point, status: *mut u8) -> c_int;
    pub fn cb2_gjk3_get_contact_manifold(ctx: *mut cb2_ctx, settings: *const cb2_settings,
                                         simplex: *const cb2_support_point, simplex_len: *const u32,
                                         l_shape: *const u32, r_shape: *const u32,
                                         l_t: *const cb2_xform, r_t: *const cb2_xform, n: u64,
                                         out: *mut cb2_contact, status: *mut u8) -> c_int;
    pub fn cb2_gjk3_intersect_dev(ctx: *mut cb2_ctx, settings: *const cb2_settings,
                                  l_shape: *const u32, r_shape: *const u32,
                                  l_t: *const cb2_xform, r_t: *const cb2_xform, n: u64,
                                  simplex_out: *mut cb2_support_point, status: *mut u8,
                                  stream: *mut c_void) -> c_int;
    pub fn cb2_gjk3_get_contact_manifold_dev(ctx: *mut cb2_ctx, settings: *const cb2_settings,
                                             simplex: *const cb2_support_point, simplex_len: *const u32,
                                             l_shape: *const u32, r_shape: *const u32,
                                             l_t: *const cb2_xform, r_t: *const cb2_xform, n: u64,
                                             out: *mut cb2_contact, status: *mut u8,
                                             stream: *mut c_void) -> c_int;

    // ---- several GPUs of one box behind one handle ----
    pub fn cb2_create_multi(devices: *const c_int, n_devices: c_int, out: *mut *mut cb2_multi) -> c_int;
    pub fn cb2_destroy_multi(m: *mut cb2_multi);
    pub fn cb2_multi_device_count(m: *const cb2_multi) -> c_int;
    pub fn cb2_multi_ctx(m: *mut cb2_multi, i: c_int) -> *mut cb2_ctx;
    pub fn cb2_multi_last_error(m: *const cb2_multi) -> *const c_char;
    pub fn cb2_multi_shapes_upload(m: *mut cb2_multi, shapes: *const cb2_shape, n_shapes: u32,
                                   verts_xyz: *const f32, n_verts: u64) -> c_int;
    pub fn cb2_multi_gjk3_intersection(m: *mut cb2_multi, strategy: u32, settings: *const cb2_settings,
                                       l_shape: *const u32, r_shape: *const u32,
                                       l_t: *const cb2_xform, r_t: *const cb2_xform, n: u64,
                                       out: *mut cb2_contact, status: *mut u8) -> c_int;
    pub fn cb2_multi_gjk3_distance(m: *mut cb2_multi, settings: *const cb2_settings,
                                   l_shape: *const u32, r_shape: *const u32,
                                   l_t: *const cb2_xform, r_t: *const cb2_xform, n: u64,
                                   out: *mut f32, status: *mut u8) -> c_int;
    pub fn cb2_multi_gjk3_time_of_impact(m: *mut cb2_multi, settings: *const cb2_settings,
                                         l_shape: *const u32, r_shape: *const u32,
                                         l_t0: *const cb2_xform, l_t1: *const cb2_xform,
                                         r_t0: *const cb2_xform, r_t1: *const cb2_xform, n: u64,
                                         iter_cap: u32, out: *mut cb2_contact, status: *mut u8) -> c_int;
    pub fn cb2_multi_sap3_find_collider_pairs(m: *mut cb2_multi, aabbs: *const cb2_aabb3, n: u64,
                                              axis_inout: *mut u32, perm_out: *mut u32,
                                              pairs_out: *mut u32, cap: u64, n_pairs: *mut u64) -> c_int;

    pub fn cb2_composites_upload(ctx: *mut cb2_ctx, part_off: *const u32, part_shape: *const u32,
                                 part_local: *const cb2_xform, n_comp: u32) -> c_int;
    pub fn cb2_gjk3_intersection_complex(ctx: *mut cb2_ctx, strategy: u32, settings: *const cb2_settings,
                                         l_comp: *const u32, r_comp: *const u32,
                                         l_t: *const cb2_xform, r_t: *const cb2_xform, n: u64,
                                         out: *mut cb2_contact, status: *mut u8) -> c_int;
    pub fn cb2_gjk3_distance_complex(ctx: *mut cb2_ctx, settings: *const cb2_settings,
                                     l_comp: *const u32, r_comp: *const u32,
                                     l_t: *const cb2_xform, r_t: *const cb2_xform, n: u64,
                                     out: *mut f32, status: *mut u8) -> c_int;
    pub fn cb2_gjk3_time_of_impact_complex(ctx: *mut cb2_ctx, strategy: u32, settings: *const cb2_settings,
                                           l_comp: *const u32, r_comp: *const u32,
                                           l_t0: *const cb2_xform, l_t1: *const cb2_xform,
                                           r_t0: *const cb2_xform, r_t1: *const cb2_xform, n: u64,
                                           iter_cap: u32, out: *mut cb2_contact, status: *mut u8) -> c_int;

    pub fn cb2_sap3_find_collider_pairs(ctx: *mut cb2_ctx, aabbs: *const cb2_aabb3, n: u64,
                                        axis_inout: *mut u32, perm_out: *mut u32,
                                        pairs_out: *mut u32, cap: u64, n_pairs: *mut u64) -> c_int;
    pub fn cb2_sap3_find_collider_pairs_dev(ctx: *mut cb2_ctx, aabbs: *const cb2_aabb3, n: u64,
                                            axis: u32, perm_out: *mut u32, pairs_out: *mut u32,
                                            cap: u64, n_pairs: *mut u64, stream: *mut c_void) -> c_int;

    pub fn cb2_sap2_find_collider_pairs(ctx: *mut cb2_ctx, aabbs: *const cb2_aabb2, n: u64,
                                        axis_inout: *mut u32, perm_out: *mut u32,
                                        pairs_out: *mut u32, cap: u64, n_pairs: *mut u64) -> c_int;
    pub fn cb2_brute_force_find_collider_pairs(ctx: *mut cb2_ctx, aabbs: *const cb2_aabb3, n: u64,
                                               pairs_out: *mut u32, cap: u64, n_pairs: *mut u64) -> c_int;
    pub fn cb2_brute_force_find_collider_pairs_dev(ctx: *mut cb2_ctx, aabbs: *const cb2_aabb3, n: u64,
                                                   pairs_out: *mut u32, cap: u64, n_pairs: *mut u64,
                                                   stream: *mut c_void) -> c_int;

    pub fn cb2_world_aabbs(ctx: *mut cb2_ctx, shape: *const u32, t: *const cb2_xform, n: u64,
                           out: *mut cb2_aabb3) -> c_int;
    pub fn cb2_world_aabbs_dev(ctx: *mut cb2_ctx, shape: *const u32, t: *const cb2_xform, n: u64,
                               out: *mut cb2_aabb3, stream: *mut c_void) -> c_int;
    pub fn cb2_sap3_find_collider_pairs_shard_dev(ctx: *mut cb2_ctx, aabbs: *const cb2_aabb3, n: u64,
                                                  axis: u32, shard: u32, n_shards: u32,
                                                  perm_out: *mut u32, pairs_out: *mut u32, cap: u64,
                                                  n_pairs: *mut u64, stream: *mut c_void) -> c_int;

    // ---- 2-D narrow phase: GJK2 / EPA2 over Primitive2 and Decomposed<Vector2, Basis2> ----
    pub fn cb2_shapes2_upload(ctx: *mut cb2_ctx, shapes: *const cb2_shape2, n_shapes: u32,
                              verts_xy: *const f32, n_verts: u64) -> c_int;
    pub fn cb2_gjk2_intersection(ctx: *mut cb2_ctx, strategy: u32, settings: *const cb2_settings,
                                 l_shape: *const u32, r_shape: *const u32,
                                 l_t: *const cb2_xform2, r_t: *const cb2_xform2, n: u64,
                                 out: *mut cb2_contact2, status: *mut u8) -> c_int;
    pub fn cb2_gjk2_distance(ctx: *mut cb2_ctx, settings: *const cb2_settings,
                             l_shape: *const u32, r_shape: *const u32,
                             l_t: *const cb2_xform2, r_t: *const cb2_xform2, n: u64,
                             out: *mut f32, status: *mut u8) -> c_int;
    pub fn cb2_gjk2_time_of_impact(ctx: *mut cb2_ctx, settings: *const cb2_settings,
                                   l_shape: *const u32, r_shape: *const u32,
                                   l_t0: *const cb2_xform2, l_t1: *const cb2_xform2,
                                   r_t0: *const cb2_xform2, r_t1: *const cb2_xform2, n: u64,
                                   iter_cap: u32, out: *mut cb2_contact2, status: *mut u8) -> c_int;
    pub fn cb2_gjk2_intersection_dev(ctx: *mut cb2_ctx, strategy: u32, settings: *const cb2_settings,
                                     l_shape: *const u32, r_shape: *const u32,
                                     l_t: *const cb2_xform2, r_t: *const cb2_xform2, n: u64,
                                     out: *mut cb2_contact2, status: *mut u8, stream: *mut c_void) -> c_int;
    pub fn cb2_gjk2_distance_dev(ctx: *mut cb2_ctx, settings: *const cb2_settings,
                                 l_shape: *const u32, r_shape: *const u32,
                                 l_t: *const cb2_xform2, r_t: *const cb2_xform2, n: u64,
                                 out: *mut f32, status: *mut u8, stream: *mut c_void) -> c_int;
    pub fn cb2_gjk2_time_of_impact_dev(ctx: *mut cb2_ctx, settings: *const cb2_settings,
                                       l_shape: *const u32, r_shape: *const u32,
                                       l_t0: *const cb2_xform2, l_t1: *const cb2_xform2,
                                       r_t0: *const cb2_xform2, r_t1: *const cb2_xform2, n: u64,
                                       iter_cap: u32, out: *mut cb2_contact2, status: *mut u8,
                                       stream: *mut c_void) -> c_int;
    pub fn cb2_composites2_upload(ctx: *mut cb2_ctx, part_off: *const u32, part_shape: *const u32,
                                  part_local: *const cb2_xform2, n_comp: u32) -> c_int;
    pub fn cb2_gjk2_intersection_complex(ctx: *mut cb2_ctx, strategy: u32, settings: *const cb2_settings,
                                         l_comp: *const u32, r_comp: *const u32,
                                         l_t: *const cb2_xform2, r_t: *const cb2_xform2, n: u64,
                                         out: *mut cb2_contact2, status: *mut u8) -> c_int;
    pub fn cb2_gjk2_distance_complex(ctx: *mut cb2_ctx, settings: *const cb2_settings,
                                     l_comp: *const u32, r_comp: *const u32,
                                     l_t: *const cb2_xform2, r_t: *const cb2_xform2, n: u64,
                                     out: *mut f32, status: *mut u8) -> c_int;
    pub fn cb2_gjk2_time_of_impact_complex(ctx: *mut cb2_ctx, strategy: u32, settings: *const cb2_settings,
                                           l_comp: *const u32, r_comp: *const u32,
                                           l_t0: *const cb2_xform2, l_t1: *const cb2_xform2,
                                           r_t0: *const cb2_xform2, r_t1: *const cb2_xform2, n: u64,
                                           iter_cap: u32, out: *mut cb2_contact2, status: *mut u8) -> c_int;
    pub fn cb2_sap3_find_collider_pairs_unordered_dev(ctx: *mut cb2_ctx, aabbs: *const cb2_aabb3, n: u64,
                                                      axis: u32, perm_out: *mut u32, pairs_out: *mut u32,
                                                      cap: u64, n_pairs: *mut u64, stream: *mut c_void) -> c_int;
    pub fn cb2_world_aabbs2(ctx: *mut cb2_ctx, shape: *const u32, t: *const cb2_xform2, n: u64,
                            out: *mut cb2_aabb2) -> c_int;
    pub fn cb2_world_aabbs2_dev(ctx: *mut cb2_ctx, shape: *const u32, t: *const cb2_xform2, n: u64,
                                out: *mut cb2_aabb2, stream: *mut c_void) -> c_int;
    pub fn cb2_sap2_find_collider_pairs_dev(ctx: *mut cb2_ctx, aabbs: *const cb2_aabb2, n: u64, axis: u32,
                                            perm_out: *mut u32, pairs_out: *mut u32, cap: u64,
                                            n_pairs: *mut u64, stream: *mut c_void) -> c_int;
    pub fn cb2_gjk2_intersection_bodies_dev(ctx: *mut cb2_ctx, strategy: u32, settings: *const cb2_settings,
                                            body_shape: *const u32, body_t: *const cb2_xform2,
                                            pair_body: *const u32, n: u64, out: *mut cb2_contact2,
                                            status: *mut u8, stream: *mut c_void) -> c_int;
}
