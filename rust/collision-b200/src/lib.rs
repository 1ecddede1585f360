//! Safe wrapper that keeps the crate's own API shapes on top of `collision-b200-sys`.
//!
//! NOT COMPILED IN THIS REPOSITORY'S CI (no Rust toolchain in the build image); it is the
//! binding a maintainer of collision-rs would add, see INTEGRATION.md.  Mirrors:
//!   * `GJK3::{intersection, distance, intersection_time_of_impact}` (gjk/mod.rs:163-391) as
//!     batch methods plus batch-of-1 methods with the original signatures,
//!   * `CollisionStrategy`, `Contact<Point3<f32>>` (contact.rs:16-82),
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
        for s in shapes {
            recs.push(match s {
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
        let rc = unsafe { sys::cb2_shapes_upload(self.raw, recs.as_ptr(), recs.len() as u32,
                                                 verts.as_ptr(), (verts.len() / 3) as u64) };
        self.check(rc)?;
        Ok((0..recs.len() as u32).map(ShapeId).collect())
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

    /// `find_collider_pairs` (sweep_prune.rs:76-136).  `bound` extracts `HasBound::bound()`.
    pub fn find_collider_pairs<A, F>(&mut self, ctx: &mut Context, shapes: &mut [A], bound: F)
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
