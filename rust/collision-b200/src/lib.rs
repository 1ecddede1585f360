//! Safe wrapper that keeps the crate's own API shapes on top of `collision-b200-sys`.
//!
//! UNCOMPILED SKETCH (no Rust toolchain in the build image: never been through `cargo check`); it
//! is the binding a maintainer of collision-rs would add, see INTEGRATION.md.  Mirrors:
//!   * `GJK3::{intersection, distance, intersection_time_of_impact}` (gjk/mod.rs:163-391) as
//!     batch methods plus batch-of-1 methods with the original signatures,
//!   * `CollisionStrategy`, `Contact<Point3<f32>>` (contact.rs:16-82),
//!   * `GJK3::{intersect, get_contact_manifold}` (gjk/mod.rs:101-146, 326-344) with
//!     `Vec<SupportPoint>`-shaped simplices, the `*_complex` callers (gjk/mod.rs:412-588) over
//!     composite tables, `ConvexPolyhedron::new_with_faces` shapes, `ComputeBound` world boxes,
//!   * `Multi`: several GPUs behind one handle (`cb2_create_multi`),
//!   * `SweepAndPrune3::find_collider_pairs` (sweep_prune.rs:76-136): permutes the caller's
//!     slice in place, returns `Vec<(usize, usize)>` in the reference order, resets the sweep
//!     axis to 0 exactly as the reference's no-op `Variance3` does.
//! Reference panics (status 2,3,4,6) are re-raised with `panic!` so behaviour stays drop-in;
//! status 5 (time of impact never converges: the reference would loop forever) is an `Err`.
use cgmath::{Decomposed, Point3, Quaternion, Vector3};
use collision::{CollisionStrategy, Contact};
use collision_b200_sys as sys;
use std::ffi::CStr;

pub type Transform3 = Decomposed<Vector3<f32>, Quaternion<f32>>;

#[derive(Debug)]
pub struct Error(pub i32, pub String);

/// Index of a shape uploaded with [`Context::upload_shapes`].
#[derive(Clone, Copy, Debug, PartialEq, Eq)]
pub struct ShapeId(pub u32);

/// `Primitive3<f32>` (primitive3.rs:20-40).
pub enum Shape<'a> {
    Particle,
    Quad { dim_x: f32, dim_y: f32 },
    Sphere { radius: f32 },
    Cuboid { dim: Vector3<f32> },
    Cube { dim: f32 },
    Cylinder { half_height: f32, radius: f32 },
    Capsule { half_height: f32, radius: f32 },
    /// `ConvexPolyhedron::new(vertices)` — VertexOnly mode (polyhedron.rs:100-122).  The crate
    /// keeps the vertices private, so the caller passes the same slice it built the shape from.
    ConvexPolyhedron { vertices: &'a [Point3<f32>] },
    /// `ConvexPolyhedron::new_with_faces(vertices, faces)` — HalfEdge mode (polyhedron.rs:124-139):
    /// support points come from the reference's hill climb over the half-edge rings.
    ConvexPolyhedronWithFaces { vertices: &'a [Point3<f32>], faces: &'a [(usize, usize, usize)] },
}

/// `SupportPoint<Point3<f32>>` (minkowski/mod.rs:16-23); the crate keeps its own type private to
/// the algorithm module, so the wrapper hands out this mirror.
#[derive(Clone, Copy, Debug, PartialEq)]
pub struct SupportPoint {
    pub v: Vector3<f32>,
    pub sup_a: Point3<f32>,
    pub sup_b: Point3<f32>,
}

pub struct Context {
    raw: *mut sys::cb2_ctx,
}
// one submitting thread at a time per ctx (SURVEY §8b "Threading")
unsafe impl Send for Context {}

fn xf(t: &Transform3) -> sys::cb2_xform {
    sys::cb2_xform { scale: t.scale, qs: t.rot.s, qx: t.rot.v.x, qy: t.rot.v.y, qz: t.rot.v.z,
                     dx: t.disp.x, dy: t.disp.y, dz: t.disp.z }
}

fn contact(c: &sys::cb2_contact) -> Contact<Point3<f32>> {
    let strategy = if c.strategy == sys::CB2_COLLISION_ONLY { CollisionStrategy::CollisionOnly }
                   else { CollisionStrategy::FullResolution };
    let mut out = Contact::new_with_point(
        strategy,
        Vector3::new(c.normal[0], c.normal[1], c.normal[2]),
        c.penetration_depth,
        Point3::new(c.contact_point[0], c.contact_point[1], c.contact_point[2]));
    out.time_of_impact = c.time_of_impact;
    out
}

fn reraise(status: u8) -> ! {
    match status {
        sys::CB2_PANIC_UNWRAP_PLANE => panic!("called `Option::unwrap()` on a `None` value (simplex3d.rs:146)"),
        sys::CB2_PANIC_ASSERT_RANGE => panic!("assertion failed: in_range(u) && in_range(v) && in_range(w) (simplex3d.rs:150)"),
        sys::CB2_PANIC_INV_SCALE => panic!("called `Option::unwrap()` on a `None` value (primitive/util.rs:18)"),
        sys::CB2_PANIC_EPA_EMPTY => panic!("index out of bounds: the len is 0 but the index is 0 (epa3d.rs:138)"),
        sys::CB2_PANIC_EPA2_EDGE => panic!("assert_ulps_ne!(S::infinity(), edge.distance) (epa2d.rs:154)"),
        sys::CB2_PANIC_INV_ROT => panic!("called `Option::unwrap()` on a `None` value (Basis2::invert)"),
        s => panic!("collision-b200: device status {}", s),
    }
}

impl Context {
    pub fn new(device: i32) -> Result<Self, Error> {
        let mut raw = std::ptr::null_mut();
        let rc = unsafe { sys::cb2_create(device, &mut raw) };
        if rc != 0 {
            let msg = unsafe { CStr::from_ptr(sys::cb2_last_error(std::ptr::null())) };
            return Err(Error(rc, msg.to_string_lossy().into_owned()));
        }
        Ok(Context { raw })
    }

    fn check(&self, rc: i32) -> Result<(), Error> {
        if rc == 0 { return Ok(()); }
        let msg = unsafe { CStr::from_ptr(sys::cb2_last_error(self.raw)) };
        Err(Error(rc, msg.to_string_lossy().into_owned()))
    }

    /// Replaces constructing the primitives; returns ids in input order.
    pub fn upload_shapes(&mut self, shapes: &[Shape]) -> Result<Vec<ShapeId>, Error> {
        let mut recs = Vec::with_capacity(shapes.len());
        let mut verts: Vec<f32> = Vec::new();
        let mut faces: Vec<u32> = Vec::new();
        for s in shapes {
            recs.push(match s {
                Shape::ConvexPolyhedronWithFaces { vertices, faces: fs } => {
                    let off = (verts.len() / 3) as u32;
                    for v in vertices.iter() { verts.extend_from_slice(&[v.x, v.y, v.z]); }
                    let f_off = (faces.len() / 3) as u32;
                    for f in fs.iter() { faces.extend_from_slice(&[f.0 as u32, f.1 as u32, f.2 as u32]); }
                    // the face range rides in p[0], p[1] as bit patterns (include/collision_b200.h)
                    sys::cb2_shape { kind: sys::CB2_POLYHEDRON_HE,
                                     p: [f32::from_bits(f_off), f32::from_bits(fs.len() as u32), 0.0],
                                     vtx_off: off, vtx_cnt: vertices.len() as u32 }
                }
                Shape::Particle => sys::cb2_shape { kind: sys::CB2_PARTICLE, p: [0.0; 3], vtx_off: 0, vtx_cnt: 0 },
                Shape::Quad { dim_x, dim_y } => sys::cb2_shape { kind: sys::CB2_QUAD, p: [*dim_x, *dim_y, 0.0], vtx_off: 0, vtx_cnt: 0 },
                Shape::Cube { dim } => sys::cb2_shape { kind: sys::CB2_CUBE, p: [*dim, 0.0, 0.0], vtx_off: 0, vtx_cnt: 0 },
                Shape::Cylinder { half_height, radius } => sys::cb2_shape { kind: sys::CB2_CYLINDER, p: [*half_height, *radius, 0.0], vtx_off: 0, vtx_cnt: 0 },
                Shape::Capsule { half_height, radius } => sys::cb2_shape { kind: sys::CB2_CAPSULE, p: [*half_height, *radius, 0.0], vtx_off: 0, vtx_cnt: 0 },
                Shape::Sphere { radius } => sys::cb2_shape { kind: sys::CB2_SPHERE, p: [*radius, 0.0, 0.0], vtx_off: 0, vtx_cnt: 0 },
                Shape::Cuboid { dim } => sys::cb2_shape { kind: sys::CB2_CUBOID, p: [dim.x, dim.y, dim.z], vtx_off: 0, vtx_cnt: 0 },
                Shape::ConvexPolyhedron { vertices } => {
                    let off = (verts.len() / 3) as u32;
                    for v in vertices.iter() { verts.extend_from_slice(&[v.x, v.y, v.z]); }
                    sys::cb2_shape { kind: sys::CB2_POLYHEDRON, p: [0.0; 3], vtx_off: off, vtx_cnt: vertices.len() as u32 }
                }
            });
        }
        let rc = if faces.is_empty() {
            unsafe { sys::cb2_shapes_upload(self.raw, recs.as_ptr(), recs.len() as u32,
                                            verts.as_ptr(), (verts.len() / 3) as u64) }
        } else {
            unsafe { sys::cb2_shapes_upload_faces(self.raw, recs.as_ptr(), recs.len() as u32, verts.as_ptr(),
                                                  (verts.len() / 3) as u64, faces.as_ptr(), (faces.len() / 3) as u64) }
        };
        self.check(rc)?;
        Ok((0..recs.len() as u32).map(ShapeId).collect())
    }

    /// `shape.compute_bound().transform_volume(&t)` per body (ComputeBound: sphere.rs:41-51,
    /// cuboid.rs:82-92, polyhedron.rs:403-410; Aabb3::transform: aabb3.rs:90-100): the boxes
    /// `SweepAndPrune3` sorts.
    pub fn world_aabbs(&mut self, shape: &[ShapeId], transform: &[Transform3]) -> Result<Vec<collision::Aabb3<f32>>, Error> {
        let n = shape.len();
        assert!(transform.len() == n);
        let s: Vec<u32> = shape.iter().map(|s| s.0).collect();
        let t: Vec<sys::cb2_xform> = transform.iter().map(xf).collect();
        let mut out = vec![sys::cb2_aabb3::default(); n];
        let rc = unsafe { sys::cb2_world_aabbs(self.raw, s.as_ptr(), t.as_ptr(), n as u64, out.as_mut_ptr()) };
        self.check(rc)?;
        Ok(out.iter().map(|b| collision::Aabb3::new(Point3::new(b.min[0], b.min[1], b.min[2]),
                                                    Point3::new(b.max[0], b.max[1], b.max[2]))).collect())
    }

    /// Composite shapes `&[(Primitive3, Transform3)]` (gjk/mod.rs:412-588): composite c owns the
    /// parts `parts[c]`, each an uploaded shape with its local transform.  Returns composite ids.
    pub fn upload_composites(&mut self, parts: &[&[(ShapeId, Transform3)]]) -> Result<Vec<ShapeId>, Error> {
        let mut off = vec![0u32];
        let mut shape = Vec::new();
        let mut local = Vec::new();
        for c in parts {
            for (s, t) in c.iter() { shape.push(s.0); local.push(xf(t)); }
            off.push(shape.len() as u32);
        }
        let rc = unsafe { sys::cb2_composites_upload(self.raw, off.as_ptr(), shape.as_ptr(), local.as_ptr(), parts.len() as u32) };
        self.check(rc)?;
        Ok((0..parts.len() as u32).map(ShapeId).collect())
    }
}

impl Drop for Context {
    fn drop(&mut self) { unsafe { sys::cb2_destroy(self.raw) } }
}

/// `GJK3<f32>` (gjk/mod.rs:44-50) over a device context.
pub struct GJK3<'c> {
    ctx: &'c mut Context,
    settings: sys::cb2_settings,
}

impl<'c> GJK3<'c> {
    /// `GJK::new()` (gjk/mod.rs:60-68).
    pub fn new(ctx: &'c mut Context) -> Self {
        let mut s = sys::cb2_settings { distance_tolerance: 0.0, continuous_tolerance: 0.0, epa_tolerance: 0.0, max_iterations: 0 };
        unsafe { sys::cb2_default_settings(&mut s) };
        GJK3 { ctx, settings: s }
    }

    /// `GJK::new_with_settings` (gjk/mod.rs:71-84).
    pub fn new_with_settings(ctx: &'c mut Context, distance_tolerance: f32, continuous_tolerance: f32,
                             epa_tolerance: f32, max_iterations: u32) -> Self {
        GJK3 { ctx, settings: sys::cb2_settings { distance_tolerance, continuous_tolerance, epa_tolerance, max_iterations } }
    }

    /// Batched `GJK3::intersection` (gjk/mod.rs:362-391): element i is the reference's result
    /// for `(left[i], left_transform[i], right[i], right_transform[i])`.
    pub fn intersection_batch(&mut self, strategy: &CollisionStrategy, left: &[ShapeId], left_transform: &[Transform3],
                              right: &[ShapeId], right_transform: &[Transform3])
                              -> Result<Vec<Option<Contact<Point3<f32>>>>, Error> {
        let n = left.len();
        assert!(right.len() == n && left_transform.len() == n && right_transform.len() == n);
        let l: Vec<u32> = left.iter().map(|s| s.0).collect();
        let r: Vec<u32> = right.iter().map(|s| s.0).collect();
        let lt: Vec<sys::cb2_xform> = left_transform.iter().map(xf).collect();
        let rt: Vec<sys::cb2_xform> = right_transform.iter().map(xf).collect();
        let mut out = vec![sys::cb2_contact::default(); n];
        let mut status = vec![0u8; n];
        let strat = match strategy { CollisionStrategy::FullResolution => sys::CB2_FULL_RESOLUTION, CollisionStrategy::CollisionOnly => sys::CB2_COLLISION_ONLY };
        let rc = unsafe { sys::cb2_gjk3_intersection(self.ctx.raw, strat, &self.settings, l.as_ptr(), r.as_ptr(),
                                                     lt.as_ptr(), rt.as_ptr(), n as u64, out.as_mut_ptr(), status.as_mut_ptr()) };
        self.ctx.check(rc)?;
        Ok(out.iter().zip(status.iter()).map(|(c, s)| match *s {
            sys::CB2_NONE => None,
            sys::CB2_SOME => Some(contact(c)),
            other => reraise(other),
        }).collect())
    }

    /// `GJK3::intersection` for every `(a, b)` of a broad phase's pair list over one body table:
    /// element i is the reference's `intersection(strategy, &shape[a], &pose[a], &shape[b], &pose[b])`.
    /// Only the index pairs are streamed to the device per call (8 bytes per pair).
    pub fn intersection_bodies(&mut self, strategy: &CollisionStrategy, body_shape: &[ShapeId], body_pose: &[Transform3],
                               pairs: &[(usize, usize)]) -> Result<Vec<Option<Contact<Point3<f32>>>>, Error> {
        assert!(body_shape.len() == body_pose.len());
        let sh: Vec<u32> = body_shape.iter().map(|s| s.0).collect();
        let t: Vec<sys::cb2_xform> = body_pose.iter().map(xf).collect();
        let pb: Vec<u32> = pairs.iter().flat_map(|&(a, b)| [a as u32, b as u32]).collect();
        let n = pairs.len();
        let mut out = vec![sys::cb2_contact::default(); n];
        let mut status = vec![0u8; n];
        let strat = match strategy { CollisionStrategy::FullResolution => sys::CB2_FULL_RESOLUTION, CollisionStrategy::CollisionOnly => sys::CB2_COLLISION_ONLY };
        let rc = unsafe { sys::cb2_gjk3_intersection_bodies(self.ctx.raw, strat, &self.settings, sh.as_ptr(), t.as_ptr(),
                                                            sh.len() as u64, pb.as_ptr(), n as u64, out.as_mut_ptr(), status.as_mut_ptr()) };
        self.ctx.check(rc)?;
        Ok(out.iter().zip(status.iter()).map(|(c, s)| match *s {
            sys::CB2_NONE => None,
            sys::CB2_SOME => Some(contact(c)),
            other => reraise(other),
        }).collect())
    }

    /// The original single-pair signature, as a batch of one (no CPU fallback).
    pub fn intersection(&mut self, strategy: &CollisionStrategy, left: ShapeId, left_transform: &Transform3,
                        right: ShapeId, right_transform: &Transform3) -> Option<Contact<Point3<f32>>> {
        self.intersection_batch(strategy, &[left], std::slice::from_ref(left_transform), &[right],
                                std::slice::from_ref(right_transform)).expect("device error").pop().unwrap()
    }

    /// Batched `GJK3::distance` (gjk/mod.rs:271-321).
    pub fn distance_batch(&mut self, left: &[ShapeId], left_transform: &[Transform3],
                          right: &[ShapeId], right_transform: &[Transform3]) -> Result<Vec<Option<f32>>, Error> {
        let n = left.len();
        let l: Vec<u32> = left.iter().map(|s| s.0).collect();
        let r: Vec<u32> = right.iter().map(|s| s.0).collect();
        let lt: Vec<sys::cb2_xform> = left_transform.iter().map(xf).collect();
        let rt: Vec<sys::cb2_xform> = right_transform.iter().map(xf).collect();
        let mut out = vec![0f32; n];
        let mut status = vec![0u8; n];
        let rc = unsafe { sys::cb2_gjk3_distance(self.ctx.raw, &self.settings, l.as_ptr(), r.as_ptr(), lt.as_ptr(), rt.as_ptr(),
                                                 n as u64, out.as_mut_ptr(), status.as_mut_ptr()) };
        self.ctx.check(rc)?;
        Ok(out.iter().zip(status.iter()).map(|(d, s)| match *s {
            sys::CB2_NONE => None,
            sys::CB2_SOME => Some(*d),
            other => reraise(other),
        }).collect())
    }

    /// The original single-pair `GJK3::distance` signature, as a batch of one.
    pub fn distance(&mut self, left: ShapeId, left_transform: &Transform3, right: ShapeId,
                    right_transform: &Transform3) -> Option<f32> {
        self.distance_batch(&[left], std::slice::from_ref(left_transform), &[right],
                            std::slice::from_ref(right_transform)).expect("device error").pop().unwrap()
    }

    /// The original single-pair `GJK3::intersection_time_of_impact` signature, as a batch of one.
    /// The reference's loop has no iteration cap (gjk/mod.rs:204); a pair on which it would never
    /// return panics here after 1024 rounds instead of hanging.
    pub fn intersection_time_of_impact(&mut self, left: ShapeId, left_transform: std::ops::Range<&Transform3>,
                                       right: ShapeId, right_transform: std::ops::Range<&Transform3>)
                                       -> Option<Contact<Point3<f32>>> {
        let l = left_transform.start.clone()..left_transform.end.clone();
        let r = right_transform.start.clone()..right_transform.end.clone();
        match self.intersection_time_of_impact_batch(&[left], std::slice::from_ref(&l), &[right],
                                                     std::slice::from_ref(&r), 1024).expect("device error").pop().unwrap() {
            Ok(c) => c,
            Err(_) => panic!("intersection_time_of_impact did not converge (the reference loops forever)"),
        }
    }

    /// Batched `GJK3::intersect` (gjk/mod.rs:101-146): `Some(simplex)` is the tetrahedron that
    /// encloses the origin, in the reference's Vec order.
    pub fn intersect_batch(&mut self, left: &[ShapeId], left_transform: &[Transform3], right: &[ShapeId],
                           right_transform: &[Transform3]) -> Result<Vec<Option<Vec<SupportPoint>>>, Error> {
        let n = left.len();
        assert!(right.len() == n && left_transform.len() == n && right_transform.len() == n);
        let l: Vec<u32> = left.iter().map(|s| s.0).collect();
        let r: Vec<u32> = right.iter().map(|s| s.0).collect();
        let lt: Vec<sys::cb2_xform> = left_transform.iter().map(xf).collect();
        let rt: Vec<sys::cb2_xform> = right_transform.iter().map(xf).collect();
        let mut sx = vec![sys::cb2_support_point::default(); 4 * n];
        let mut status = vec![0u8; n];
        let rc = unsafe { sys::cb2_gjk3_intersect(self.ctx.raw, &self.settings, l.as_ptr(), r.as_ptr(), lt.as_ptr(),
                                                  rt.as_ptr(), n as u64, sx.as_mut_ptr(), status.as_mut_ptr()) };
        self.ctx.check(rc)?;
        Ok(status.iter().enumerate().map(|(i, s)| match *s {
            sys::CB2_NONE => None,
            sys::CB2_SOME => Some(sx[4 * i..4 * i + 4].iter().map(|p| SupportPoint {
                v: Vector3::new(p.v[0], p.v[1], p.v[2]),
                sup_a: Point3::new(p.sup_a[0], p.sup_a[1], p.sup_a[2]),
                sup_b: Point3::new(p.sup_b[0], p.sup_b[1], p.sup_b[2]) }).collect()),
            other => reraise(other),
        }).collect())
    }

    /// `GJK3::intersect` with the original single-pair signature.
    pub fn intersect(&mut self, left: ShapeId, left_transform: &Transform3, right: ShapeId,
                     right_transform: &Transform3) -> Option<Vec<SupportPoint>> {
        self.intersect_batch(&[left], std::slice::from_ref(left_transform), &[right],
                             std::slice::from_ref(right_transform)).expect("device error").pop().unwrap()
    }

    /// `GJK3::get_contact_manifold` (gjk/mod.rs:326-344): EPA3::process on a simplex from
    /// [`GJK3::intersect`]; fewer than four points give `None` (epa3d.rs:38-40).
    pub fn get_contact_manifold(&mut self, simplex: &mut Vec<SupportPoint>, left: ShapeId, left_transform: &Transform3,
                                right: ShapeId, right_transform: &Transform3) -> Option<Contact<Point3<f32>>> {
        let mut sx = [sys::cb2_support_point::default(); 4];
        for (o, p) in sx.iter_mut().zip(simplex.iter()) {
            *o = sys::cb2_support_point { v: [p.v.x, p.v.y, p.v.z], sup_a: [p.sup_a.x, p.sup_a.y, p.sup_a.z],
                                          sup_b: [p.sup_b.x, p.sup_b.y, p.sup_b.z] };
        }
        let len = simplex.len() as u32;
        let (lt, rt) = (xf(left_transform), xf(right_transform));
        let mut out = sys::cb2_contact::default();
        let mut status = 0u8;
        let rc = unsafe { sys::cb2_gjk3_get_contact_manifold(self.ctx.raw, &self.settings, sx.as_ptr(), &len, &left.0,
                                                             &right.0, &lt, &rt, 1, &mut out, &mut status) };
        self.ctx.check(rc).expect("device error");
        match status { sys::CB2_NONE => None, sys::CB2_SOME => Some(contact(&out)), other => reraise(other) }
    }

    /// Batched `GJK3::intersection_complex` (gjk/mod.rs:412-470) over composites uploaded with
    /// [`Context::upload_composites`].
    pub fn intersection_complex_batch(&mut self, strategy: &CollisionStrategy, left: &[ShapeId], left_transform: &[Transform3],
                                      right: &[ShapeId], right_transform: &[Transform3])
                                      -> Result<Vec<Option<Contact<Point3<f32>>>>, Error> {
        let n = left.len();
        let l: Vec<u32> = left.iter().map(|s| s.0).collect();
        let r: Vec<u32> = right.iter().map(|s| s.0).collect();
        let lt: Vec<sys::cb2_xform> = left_transform.iter().map(xf).collect();
        let rt: Vec<sys::cb2_xform> = right_transform.iter().map(xf).collect();
        let mut out = vec![sys::cb2_contact::default(); n];
        let mut status = vec![0u8; n];
        let strat = match strategy { CollisionStrategy::FullResolution => sys::CB2_FULL_RESOLUTION, CollisionStrategy::CollisionOnly => sys::CB2_COLLISION_ONLY };
        let rc = unsafe { sys::cb2_gjk3_intersection_complex(self.ctx.raw, strat, &self.settings, l.as_ptr(), r.as_ptr(),
                                                             lt.as_ptr(), rt.as_ptr(), n as u64, out.as_mut_ptr(), status.as_mut_ptr()) };
        self.ctx.check(rc)?;
        Ok(out.iter().zip(status.iter()).map(|(c, s)| match *s {
            sys::CB2_NONE => None,
            sys::CB2_SOME => Some(contact(c)),
            sys::CB2_PANIC_CMP_NAN => panic!("called `Option::unwrap()` on a `None` value (gjk/mod.rs:455-459)"),
            other => reraise(other),
        }).collect())
    }

    /// Batched `GJK3::distance_complex` (gjk/mod.rs:518-553).
    pub fn distance_complex_batch(&mut self, left: &[ShapeId], left_transform: &[Transform3], right: &[ShapeId],
                                  right_transform: &[Transform3]) -> Result<Vec<Option<f32>>, Error> {
        let n = left.len();
        let l: Vec<u32> = left.iter().map(|s| s.0).collect();
        let r: Vec<u32> = right.iter().map(|s| s.0).collect();
        let lt: Vec<sys::cb2_xform> = left_transform.iter().map(xf).collect();
        let rt: Vec<sys::cb2_xform> = right_transform.iter().map(xf).collect();
        let mut out = vec![0f32; n];
        let mut status = vec![0u8; n];
        let rc = unsafe { sys::cb2_gjk3_distance_complex(self.ctx.raw, &self.settings, l.as_ptr(), r.as_ptr(), lt.as_ptr(),
                                                         rt.as_ptr(), n as u64, out.as_mut_ptr(), status.as_mut_ptr()) };
        self.ctx.check(rc)?;
        Ok(out.iter().zip(status.iter()).map(|(d, s)| match *s {
            sys::CB2_NONE => None, sys::CB2_SOME => Some(*d), other => reraise(other) }).collect())
    }

    /// Batched `GJK3::intersection_complex_time_of_impact` (gjk/mod.rs:472-516).
    pub fn intersection_complex_time_of_impact_batch(&mut self, strategy: &CollisionStrategy, left: &[ShapeId],
                                                     left_transform: &[std::ops::Range<Transform3>], right: &[ShapeId],
                                                     right_transform: &[std::ops::Range<Transform3>], iter_cap: u32)
                                                     -> Result<Vec<Result<Option<Contact<Point3<f32>>>, usize>>, Error> {
        let n = left.len();
        let l: Vec<u32> = left.iter().map(|s| s.0).collect();
        let r: Vec<u32> = right.iter().map(|s| s.0).collect();
        let l0: Vec<sys::cb2_xform> = left_transform.iter().map(|t| xf(&t.start)).collect();
        let l1: Vec<sys::cb2_xform> = left_transform.iter().map(|t| xf(&t.end)).collect();
        let r0: Vec<sys::cb2_xform> = right_transform.iter().map(|t| xf(&t.start)).collect();
        let r1: Vec<sys::cb2_xform> = right_transform.iter().map(|t| xf(&t.end)).collect();
        let mut out = vec![sys::cb2_contact::default(); n];
        let mut status = vec![0u8; n];
        let strat = match strategy { CollisionStrategy::FullResolution => sys::CB2_FULL_RESOLUTION, CollisionStrategy::CollisionOnly => sys::CB2_COLLISION_ONLY };
        let rc = unsafe { sys::cb2_gjk3_time_of_impact_complex(self.ctx.raw, strat, &self.settings, l.as_ptr(), r.as_ptr(),
                                                               l0.as_ptr(), l1.as_ptr(), r0.as_ptr(), r1.as_ptr(), n as u64,
                                                               iter_cap, out.as_mut_ptr(), status.as_mut_ptr()) };
        self.ctx.check(rc)?;
        Ok(out.iter().zip(status.iter()).enumerate().map(|(i, (c, s))| match *s {
            sys::CB2_NONE => Ok(None),
            sys::CB2_SOME => Ok(Some(contact(c))),
            sys::CB2_NONCONVERGED => Err(i),
            other => reraise(other),
        }).collect())
    }

    /// Batched `GJK3::intersection_time_of_impact` (gjk/mod.rs:163-256); `Err(index)` inside the
    /// vector marks a pair whose loop did not converge within `iter_cap` (the reference hangs).
    pub fn intersection_time_of_impact_batch(&mut self, left: &[ShapeId], left_transform: &[std::ops::Range<Transform3>],
                                             right: &[ShapeId], right_transform: &[std::ops::Range<Transform3>], iter_cap: u32)
                                             -> Result<Vec<Result<Option<Contact<Point3<f32>>>, usize>>, Error> {
        let n = left.len();
        let l: Vec<u32> = left.iter().map(|s| s.0).collect();
        let r: Vec<u32> = right.iter().map(|s| s.0).collect();
        let l0: Vec<sys::cb2_xform> = left_transform.iter().map(|t| xf(&t.start)).collect();
        let l1: Vec<sys::cb2_xform> = left_transform.iter().map(|t| xf(&t.end)).collect();
        let r0: Vec<sys::cb2_xform> = right_transform.iter().map(|t| xf(&t.start)).collect();
        let r1: Vec<sys::cb2_xform> = right_transform.iter().map(|t| xf(&t.end)).collect();
        let mut out = vec![sys::cb2_contact::default(); n];
        let mut status = vec![0u8; n];
        let rc = unsafe { sys::cb2_gjk3_time_of_impact(self.ctx.raw, &self.settings, l.as_ptr(), r.as_ptr(), l0.as_ptr(), l1.as_ptr(),
                                                       r0.as_ptr(), r1.as_ptr(), n as u64, iter_cap, out.as_mut_ptr(), status.as_mut_ptr()) };
        self.ctx.check(rc)?;
        Ok(out.iter().zip(status.iter()).enumerate().map(|(i, (c, s))| match *s {
            sys::CB2_NONE => Ok(None),
            sys::CB2_SOME => Ok(Some(contact(c))),
            sys::CB2_NONCONVERGED => Err(i),
            other => reraise(other),
        }).collect())
    }
}

/// `SweepAndPrune3<f32, Aabb3<f32>>` (sweep_prune.rs:20-59) over a device context.
pub struct SweepAndPrune3 {
    sweep_axis: usize,
}

impl SweepAndPrune3 {
    pub fn new() -> Self { SweepAndPrune3 { sweep_axis: 0 } }
    pub fn with_sweep_axis(sweep_axis: usize) -> Self { SweepAndPrune3 { sweep_axis } }
    pub fn get_sweep_axis(&self) -> usize { self.sweep_axis }

    /// `find_collider_pairs` with the reference's own bound (sweep_prune.rs:76-81:
    /// `A: HasBound, A::Bound = Aabb3<f32>`); the extra parameter is the device context.  A NaN
    /// extent on the sweep axis is an `Err` (the reference's sort order is unspecified for it).
    pub fn find_collider_pairs<A>(&mut self, ctx: &mut Context, shapes: &mut [A]) -> Result<Vec<(usize, usize)>, Error>
    where A: collision::HasBound<Bound = collision::Aabb3<f32>> {
        self.find_collider_pairs_by(ctx, shapes, |s| s.bound().clone())
    }

    /// The same with a closure in place of `HasBound::bound()`.
    pub fn find_collider_pairs_by<A, F>(&mut self, ctx: &mut Context, shapes: &mut [A], bound: F)
                                        -> Result<Vec<(usize, usize)>, Error>
    where F: Fn(&A) -> collision::Aabb3<f32> {
        let n = shapes.len();
        if n <= 1 { return Ok(Vec::new()); }   // sweep_prune.rs:83-85: nothing sorted, axis unchanged
        let boxes: Vec<sys::cb2_aabb3> = shapes.iter().map(|s| { let b = bound(s);
            sys::cb2_aabb3 { min: [b.min.x, b.min.y, b.min.z], max: [b.max.x, b.max.y, b.max.z] } }).collect();
        let mut perm = vec![0u32; n];
        let mut axis = self.sweep_axis as u32;
        let mut cap = (8 * n) as u64;
        let mut pairs: Vec<u32>;
        let mut count = 0u64;
        loop {
            pairs = vec![0u32; 2 * cap as usize];
            let rc = unsafe { sys::cb2_sap3_find_collider_pairs(ctx.raw, boxes.as_ptr(), n as u64, &mut axis,
                                                                perm.as_mut_ptr(), pairs.as_mut_ptr(), cap, &mut count) };
            if rc == sys::CB2_ERR_NOSPC { cap = count; axis = self.sweep_axis as u32; continue; }
            ctx.check(rc)?;
            break;
        }
        apply_permutation(shapes, &perm);      // the reference sorts the caller's slice in place
        self.sweep_axis = axis as usize;       // always 0 for n >= 2 (sweep_prune.rs:254-277)
        Ok((0..count as usize).map(|k| (pairs[2 * k] as usize, pairs[2 * k + 1] as usize)).collect())
    }
}

/// Several GPUs of one box behind one handle (`cb2_create_multi`): host-buffer batches are split
/// into contiguous slices, one per device, and come back in place.
pub struct Multi {
    raw: *mut sys::cb2_multi,
}
unsafe impl Send for Multi {}

impl Multi {
    pub fn new(devices: &[i32]) -> Result<Self, Error> {
        let mut raw = std::ptr::null_mut();
        let rc = unsafe { sys::cb2_create_multi(devices.as_ptr(), devices.len() as i32, &mut raw) };
        if rc != 0 {
            let msg = unsafe { CStr::from_ptr(sys::cb2_last_error(std::ptr::null())) };
            return Err(Error(rc, msg.to_string_lossy().into_owned()));
        }
        Ok(Multi { raw })
    }
    fn check(&self, rc: i32) -> Result<(), Error> {
        if rc == 0 { return Ok(()); }
        let msg = unsafe { CStr::from_ptr(sys::cb2_multi_last_error(self.raw)) };
        Err(Error(rc, msg.to_string_lossy().into_owned()))
    }
    pub fn device_count(&self) -> usize { unsafe { sys::cb2_multi_device_count(self.raw) as usize } }

    /// Raw shape records (as [`Context::upload_shapes`] builds them) to every device.
    pub fn upload_shape_records(&mut self, shapes: &[sys::cb2_shape], verts_xyz: &[f32]) -> Result<(), Error> {
        let rc = unsafe { sys::cb2_multi_shapes_upload(self.raw, shapes.as_ptr(), shapes.len() as u32,
                                                       verts_xyz.as_ptr(), (verts_xyz.len() / 3) as u64) };
        self.check(rc)
    }

    /// Batched `GJK3::intersection` over all devices.
    pub fn intersection_batch(&mut self, settings: &sys::cb2_settings, strategy: &CollisionStrategy, left: &[ShapeId],
                              left_transform: &[Transform3], right: &[ShapeId], right_transform: &[Transform3])
                              -> Result<Vec<Option<Contact<Point3<f32>>>>, Error> {
        let n = left.len();
        let l: Vec<u32> = left.iter().map(|s| s.0).collect();
        let r: Vec<u32> = right.iter().map(|s| s.0).collect();
        let lt: Vec<sys::cb2_xform> = left_transform.iter().map(xf).collect();
        let rt: Vec<sys::cb2_xform> = right_transform.iter().map(xf).collect();
        let mut out = vec![sys::cb2_contact::default(); n];
        let mut status = vec![0u8; n];
        let strat = match strategy { CollisionStrategy::FullResolution => sys::CB2_FULL_RESOLUTION, CollisionStrategy::CollisionOnly => sys::CB2_COLLISION_ONLY };
        let rc = unsafe { sys::cb2_multi_gjk3_intersection(self.raw, strat, settings, l.as_ptr(), r.as_ptr(), lt.as_ptr(),
                                                           rt.as_ptr(), n as u64, out.as_mut_ptr(), status.as_mut_ptr()) };
        self.check(rc)?;
        Ok(out.iter().zip(status.iter()).map(|(c, s)| match *s {
            sys::CB2_NONE => None, sys::CB2_SOME => Some(contact(c)), other => reraise(other) }).collect())
    }
}

impl Drop for Multi {
    fn drop(&mut self) { unsafe { sys::cb2_destroy_multi(self.raw) } }
}

/// shapes[k] <- old shapes[perm[k]] with swaps only (`A` need not be `Clone`).
fn apply_permutation<A>(shapes: &mut [A], perm: &[u32]) {
    let mut done = vec![false; perm.len()];
    for start in 0..perm.len() {
        if done[start] { continue; }
        let mut k = start;
        loop {
            done[k] = true;
            let src = perm[k] as usize;
            if src == start { break; }
            shapes.swap(k, src);
            k = src;
        }
    }
}


// ---- 2-D: GJK2 / EPA2 over Primitive2 and Decomposed<Vector2<f32>, Basis2<f32>> -----------------------
pub type Transform2 = cgmath::Decomposed<cgmath::Vector2<f32>, cgmath::Basis2<f32>>;

/// `Primitive2<f32>` (primitive/primitive2.rs:20-33).
pub enum Shape2<'a> {
    Particle,
    Line { origin: cgmath::Point2<f32>, dest: cgmath::Point2<f32> },
    Circle { radius: f32 },
    Rectangle { dim: cgmath::Vector2<f32> },
    Square { dim: f32 },
    ConvexPolygon { vertices: &'a [cgmath::Point2<f32>] },
}

fn xf2(t: &Transform2) -> sys::cb2_xform2 {
    let m: &cgmath::Matrix2<f32> = t.rot.as_ref();   // column major: x = c0, y = c1
    sys::cb2_xform2 { scale: t.scale, c0x: m.x.x, c0y: m.x.y, c1x: m.y.x, c1y: m.y.y,
                      dx: t.disp.x, dy: t.disp.y, reserved: 0.0 }
}

fn contact2(c: &sys::cb2_contact2) -> Contact<cgmath::Point2<f32>> {
    let strategy = if c.strategy == sys::CB2_COLLISION_ONLY { CollisionStrategy::CollisionOnly }
                   else { CollisionStrategy::FullResolution };
    let mut out = Contact::new_with_point(strategy, cgmath::Vector2::new(c.normal[0], c.normal[1]),
                                          c.penetration_depth,
                                          cgmath::Point2::new(c.contact_point[0], c.contact_point[1]));
    out.time_of_impact = c.time_of_impact;
    out
}

impl Context {
    /// Replaces building `Circle::new` / `Rectangle::new` / ... values; ids index the 2-D table.
    pub fn upload_shapes2(&mut self, shapes: &[Shape2]) -> Result<Vec<ShapeId>, Error> {
        let mut recs = Vec::with_capacity(shapes.len());
        let mut verts: Vec<f32> = Vec::new();
        for s in shapes {
            let mut r = sys::cb2_shape2::default();
            match *s {
                Shape2::Particle => r.kind = sys::CB2_PARTICLE2,
                Shape2::Line { origin, dest } => { r.kind = sys::CB2_LINE2; r.p = [origin.x, origin.y, dest.x, dest.y]; }
                Shape2::Circle { radius } => { r.kind = sys::CB2_CIRCLE; r.p[0] = radius; }
                Shape2::Rectangle { dim } => { r.kind = sys::CB2_RECTANGLE; r.p[0] = dim.x; r.p[1] = dim.y; }
                Shape2::Square { dim } => { r.kind = sys::CB2_SQUARE; r.p[0] = dim; }
                Shape2::ConvexPolygon { vertices } => {
                    r.kind = sys::CB2_POLYGON;
                    r.vtx_off = (verts.len() / 2) as u32;
                    r.vtx_cnt = vertices.len() as u32;
                    for v in vertices { verts.push(v.x); verts.push(v.y); }
                }
            }
            recs.push(r);
        }
        let rc = unsafe { sys::cb2_shapes2_upload(self.raw, recs.as_ptr(), recs.len() as u32, verts.as_ptr(),
                                                  (verts.len() / 2) as u64) };
        self.check(rc)?;
        Ok((0..shapes.len() as u32).map(ShapeId).collect())
    }
}

/// `GJK2<f32>` (gjk/mod.rs:27) over a device context.
pub struct GJK2<'c> {
    ctx: &'c mut Context,
    settings: sys::cb2_settings,
}

impl<'c> GJK2<'c> {
    pub fn new(ctx: &'c mut Context) -> Self {
        let mut s = sys::cb2_settings { distance_tolerance: 0.0, continuous_tolerance: 0.0, epa_tolerance: 0.0, max_iterations: 0 };
        unsafe { sys::cb2_default_settings(&mut s) };
        GJK2 { ctx, settings: s }
    }

    /// Batched `GJK2::intersection` (gjk/mod.rs:362-391 -> EPA2::process, epa2d.rs:27-67).
    pub fn intersection_batch(&mut self, strategy: &CollisionStrategy, left: &[ShapeId], left_transform: &[Transform2],
                              right: &[ShapeId], right_transform: &[Transform2])
                              -> Result<Vec<Option<Contact<cgmath::Point2<f32>>>>, Error> {
        let n = left.len();
        assert!(right.len() == n && left_transform.len() == n && right_transform.len() == n);
        let l: Vec<u32> = left.iter().map(|s| s.0).collect();
        let r: Vec<u32> = right.iter().map(|s| s.0).collect();
        let lt: Vec<sys::cb2_xform2> = left_transform.iter().map(xf2).collect();
        let rt: Vec<sys::cb2_xform2> = right_transform.iter().map(xf2).collect();
        let mut out = vec![sys::cb2_contact2::default(); n];
        let mut status = vec![0u8; n];
        let strat = match strategy { CollisionStrategy::FullResolution => sys::CB2_FULL_RESOLUTION, CollisionStrategy::CollisionOnly => sys::CB2_COLLISION_ONLY };
        let rc = unsafe { sys::cb2_gjk2_intersection(self.ctx.raw, strat, &self.settings, l.as_ptr(), r.as_ptr(),
                                                     lt.as_ptr(), rt.as_ptr(), n as u64, out.as_mut_ptr(), status.as_mut_ptr()) };
        self.ctx.check(rc)?;
        Ok(out.iter().zip(status.iter()).map(|(c, s)| match *s {
            sys::CB2_NONE => None,
            sys::CB2_SOME => Some(contact2(c)),
            other => reraise(other),
        }).collect())
    }

    /// Batched `GJK2::distance` (gjk/mod.rs:271-321).
    pub fn distance_batch(&mut self, left: &[ShapeId], left_transform: &[Transform2],
                          right: &[ShapeId], right_transform: &[Transform2]) -> Result<Vec<Option<f32>>, Error> {
        let n = left.len();
        let l: Vec<u32> = left.iter().map(|s| s.0).collect();
        let r: Vec<u32> = right.iter().map(|s| s.0).collect();
        let lt: Vec<sys::cb2_xform2> = left_transform.iter().map(xf2).collect();
        let rt: Vec<sys::cb2_xform2> = right_transform.iter().map(xf2).collect();
        let mut out = vec![0f32; n];
        let mut status = vec![0u8; n];
        let rc = unsafe { sys::cb2_gjk2_distance(self.ctx.raw, &self.settings, l.as_ptr(), r.as_ptr(), lt.as_ptr(),
                                                 rt.as_ptr(), n as u64, out.as_mut_ptr(), status.as_mut_ptr()) };
        self.ctx.check(rc)?;
        Ok(out.iter().zip(status.iter()).map(|(d, s)| match *s {
            sys::CB2_NONE => None,
            sys::CB2_SOME => Some(*d),
            other => reraise(other),
        }).collect())
    }

    /// Batched `GJK2::intersection_time_of_impact` (gjk/mod.rs:163-256); `Err(i)` in the inner
    /// result marks an element on which the reference's uncapped loop would never return.
    pub fn intersection_time_of_impact_batch(&mut self, left: &[ShapeId], left_transform: &[std::ops::Range<Transform2>],
                                             right: &[ShapeId], right_transform: &[std::ops::Range<Transform2>], iter_cap: u32)
                                             -> Result<Vec<Result<Option<Contact<cgmath::Point2<f32>>>, usize>>, Error> {
        let n = left.len();
        let l: Vec<u32> = left.iter().map(|s| s.0).collect();
        let r: Vec<u32> = right.iter().map(|s| s.0).collect();
        let l0: Vec<sys::cb2_xform2> = left_transform.iter().map(|t| xf2(&t.start)).collect();
        let l1: Vec<sys::cb2_xform2> = left_transform.iter().map(|t| xf2(&t.end)).collect();
        let r0: Vec<sys::cb2_xform2> = right_transform.iter().map(|t| xf2(&t.start)).collect();
        let r1: Vec<sys::cb2_xform2> = right_transform.iter().map(|t| xf2(&t.end)).collect();
        let mut out = vec![sys::cb2_contact2::default(); n];
        let mut status = vec![0u8; n];
        let rc = unsafe { sys::cb2_gjk2_time_of_impact(self.ctx.raw, &self.settings, l.as_ptr(), r.as_ptr(), l0.as_ptr(),
                                                       l1.as_ptr(), r0.as_ptr(), r1.as_ptr(), n as u64, iter_cap,
                                                       out.as_mut_ptr(), status.as_mut_ptr()) };
        self.ctx.check(rc)?;
        Ok(out.iter().zip(status.iter()).enumerate().map(|(i, (c, s))| match *s {
            sys::CB2_NONE => Ok(None),
            sys::CB2_SOME => Ok(Some(contact2(c))),
            sys::CB2_NONCONVERGED => Err(i),
            other => reraise(other),
        }).collect())
    }
}
