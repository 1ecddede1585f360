"""Host-side mirror of the reference interface for the hot path, above the C ABI.

Names follow the crate (GJK3, SweepAndPrune3, CollisionStrategy, Contact, Sphere, Cuboid,
ConvexPolyhedron, Decomposed) so the parity tests read like the reference's own tests
(gjk/mod.rs:622-826, sweep_prune.rs:279-371).  Everything computes on the GPU through
libcollision_b200.so; there is no CPU implementation in this package.
"""
import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import abi
from .abi import (AABB_DT, CAPSULE, COLLISION_ONLY, CONTACT_DT, CUBE, CUBOID, CYLINDER,
                  FULL_RESOLUTION, PARTICLE, POLYHEDRON, POLYHEDRON_HE, QUAD, SHAPE_DT, SPHERE, XFORM_DT, Settings,
                  ptr)


class Cb2Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"cb2 error {code}: {msg}")
        self.code = code


class ReferencePanic(RuntimeError):
    """The reference would have panicked (or hung) on this input; .status says where."""

    def __init__(self, status):
        names = {2: "Plane::from_points(..).unwrap() (simplex3d.rs:146)",
                 3: "assert!(in_range(..)) (simplex3d.rs:150)",
                 4: "inverse_transform_vector(..).unwrap() (util.rs:18)",
                 5: "time-of-impact loop did not converge (gjk/mod.rs:204 has no cap)",
                 6: "faces[0] on an empty polytope (epa3d.rs:138)",
                 7: "device EPA scratch overflow",
                 8: "max_by(partial_cmp(..).unwrap()) on a NaN depth (gjk/mod.rs:455-459)",
                 9: "assert_ulps_ne!(inf, edge.distance) (epa2d.rs:154)",
                 10: "Basis2::invert(): Matrix2::invert().unwrap() on a singular matrix"}
        super().__init__(names.get(status, f"status {status}"))
        self.status = status


class Context:
    """One cb2_ctx (one GPU).  Batch methods take numpy arrays (host path) or raw device
    pointers (`*_dev`, for torch tensors: pass tensor.data_ptr())."""

    def __init__(self, device=0):
        self.lib = abi.load()
        h = C.c_void_p()
        rc = self.lib.cb2_create(int(device), C.byref(h))
        if rc != 0:
            raise Cb2Error(rc, self.lib.cb2_last_error(None).decode())
        self.h = h
        self.device = device
        self.n_shapes = 0

    def close(self):
        if getattr(self, "h", None):
            self.lib.cb2_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            raise Cb2Error(rc, self.lib.cb2_last_error(self.h).decode())

    @property
    def sm_count(self):
        return int(self.lib.cb2_device_sm_count(self.h))

    @property
    def kernel_launches(self):
        return int(self.lib.cb2_kernel_launches(self.h))

    def set_stage_timing(self, on=True):
        self._ck(self.lib.cb2_set_stage_timing(self.h, 1 if on else 0))

    def last_stage_ms(self):
        """ms of {GJK, EPA tier 0, EPA tiers 1+2, second chance} of the last timed device call."""
        out = (C.c_float * 4)()
        self._ck(self.lib.cb2_last_stage_ms(self.h, C.byref(out)))
        return [float(x) for x in out]

    # ---- shape table --------------------------------------------------------------------
    def upload_shapes(self, shapes, verts=None, faces=None):
        """faces: (F, 3) uint32 for HalfEdge polyhedra (kind POLYHEDRON_HE; their face range is
        stored in p[0], p[1] as uint32 bit patterns), vertex indices local to the owning shape."""
        shapes = np.ascontiguousarray(shapes, SHAPE_DT)
        nv = 0
        if verts is not None:
            verts = np.ascontiguousarray(verts, np.float32).reshape(-1, 3)
            nv = len(verts)
        if faces is None:
            self._ck(self.lib.cb2_shapes_upload(self.h, ptr(shapes), len(shapes), ptr(verts), nv))
        else:
            faces = np.ascontiguousarray(faces, np.uint32).reshape(-1, 3)
            self._ck(self.lib.cb2_shapes_upload_faces(self.h, ptr(shapes), len(shapes), ptr(verts),
                                                      nv, ptr(faces), len(faces)))
        self.n_shapes = len(shapes)

    # ---- 2-D narrow phase (GJK2 / EPA2) ---------------------------------------------------
    def upload_shapes2(self, shapes, verts=None):
        shapes = np.ascontiguousarray(shapes, abi.SHAPE2_DT)
        nv = 0
        if verts is not None:
            verts = np.ascontiguousarray(verts, np.float32).reshape(-1, 2)
            nv = len(verts)
        self._ck(self.lib.cb2_shapes2_upload(self.h, ptr(shapes), len(shapes), ptr(verts), nv))
        self.n_shapes2 = len(shapes)

    def intersection2(self, strategy, l_shape, r_shape, l_t, r_t, settings=None):
        settings = settings or Settings.default()
        n = len(l_shape)
        out = np.zeros(n, abi.CONTACT2_DT)
        status = np.zeros(n, np.uint8)
        self._ck(self.lib.cb2_gjk2_intersection(
            self.h, int(strategy), C.byref(settings), ptr(l_shape), ptr(r_shape), ptr(l_t),
            ptr(r_t), n, ptr(out), ptr(status)))
        return out, status

    def distance2(self, l_shape, r_shape, l_t, r_t, settings=None):
        settings = settings or Settings.default()
        n = len(l_shape)
        out = np.zeros(n, np.float32)
        status = np.zeros(n, np.uint8)
        self._ck(self.lib.cb2_gjk2_distance(
            self.h, C.byref(settings), ptr(l_shape), ptr(r_shape), ptr(l_t), ptr(r_t), n,
            ptr(out), ptr(status)))
        return out, status

    def time_of_impact2(self, l_shape, r_shape, l_t0, l_t1, r_t0, r_t1, settings=None,
                        iter_cap=1024):
        settings = settings or Settings.default()
        n = len(l_shape)
        out = np.zeros(n, abi.CONTACT2_DT)
        status = np.zeros(n, np.uint8)
        self._ck(self.lib.cb2_gjk2_time_of_impact(
            self.h, C.byref(settings), ptr(l_shape), ptr(r_shape), ptr(l_t0), ptr(l_t1),
            ptr(r_t0), ptr(r_t1), n, int(iter_cap), ptr(out), ptr(status)))
        return out, status

    def intersection2_dev(self, strategy, l_shape, r_shape, l_t, r_t, n, out, status, stream=0,
                          settings=None):
        settings = settings or Settings.default()
        self._ck(self.lib.cb2_gjk2_intersection_dev(
            self.h, int(strategy), C.byref(settings), l_shape, r_shape, l_t, r_t, n, out, status,
            stream))

    def distance2_dev(self, l_shape, r_shape, l_t, r_t, n, out, status, stream=0, settings=None):
        settings = settings or Settings.default()
        self._ck(self.lib.cb2_gjk2_distance_dev(
            self.h, C.byref(settings), l_shape, r_shape, l_t, r_t, n, out, status, stream))

    def time_of_impact2_dev(self, l_shape, r_shape, l_t0, l_t1, r_t0, r_t1, n, out, status,
                            stream=0, settings=None, iter_cap=1024):
        settings = settings or Settings.default()
        self._ck(self.lib.cb2_gjk2_time_of_impact_dev(
            self.h, C.byref(settings), l_shape, r_shape, l_t0, l_t1, r_t0, r_t1, n, int(iter_cap),
            out, status, stream))

    def world_aabbs2(self, shape, t):
        """(n, 4) f32: min.xy, max.xy of shape.compute_bound().transform_volume(&t) per body."""
        shape = np.ascontiguousarray(shape, np.uint32)
        t = np.ascontiguousarray(t, abi.XFORM2_DT)
        out = np.zeros((len(shape), 4), np.float32)
        self._ck(self.lib.cb2_world_aabbs2(self.h, ptr(shape), ptr(t), len(shape), ptr(out)))
        return out

    def world_aabbs2_dev(self, shape, t, n, out, stream=0):
        self._ck(self.lib.cb2_world_aabbs2_dev(self.h, shape, t, int(n), out, stream))

    def sap2_dev(self, aabbs2, n, axis, perm_out, pairs_out, cap, stream=0):
        npairs = C.c_uint64(0)
        rc = self.lib.cb2_sap2_find_collider_pairs_dev(self.h, aabbs2, int(n), int(axis), perm_out, pairs_out,
                                                       int(cap), C.byref(npairs), stream)
        if rc == abi.ERR_NOSPC:
            return rc, int(npairs.value)
        self._ck(rc)
        return 0, int(npairs.value)

    def intersection2_bodies_dev(self, strategy, body_shape, body_t, pair_body, n, out, status, stream=0,
                                 settings=None):
        settings = settings or Settings.default()
        self._ck(self.lib.cb2_gjk2_intersection_bodies_dev(
            self.h, int(strategy), C.byref(settings), body_shape, body_t, pair_body, int(n), out, status, stream))

    def upload_composites2(self, part_off, part_shape, part_local):
        """2-D composites: parts [part_off[c], part_off[c+1]) of (part_shape -> 2-D shape table,
        part_local: XFORM2_DT)."""
        part_off = np.ascontiguousarray(part_off, np.uint32)
        part_shape = np.ascontiguousarray(part_shape, np.uint32)
        part_local = np.ascontiguousarray(part_local, abi.XFORM2_DT)
        self._ck(self.lib.cb2_composites2_upload(self.h, ptr(part_off), ptr(part_shape), ptr(part_local),
                                                 len(part_off) - 1))

    def intersection_complex2(self, strategy, l_comp, r_comp, l_t, r_t, settings=None):
        settings = settings or Settings.default()
        n = len(l_comp)
        out = np.zeros(n, abi.CONTACT2_DT)
        status = np.zeros(n, np.uint8)
        self._ck(self.lib.cb2_gjk2_intersection_complex(
            self.h, int(strategy), C.byref(settings), ptr(l_comp), ptr(r_comp), ptr(l_t), ptr(r_t), n,
            ptr(out), ptr(status)))
        return out, status

    def distance_complex2(self, l_comp, r_comp, l_t, r_t, settings=None):
        settings = settings or Settings.default()
        n = len(l_comp)
        out = np.zeros(n, np.float32)
        status = np.zeros(n, np.uint8)
        self._ck(self.lib.cb2_gjk2_distance_complex(
            self.h, C.byref(settings), ptr(l_comp), ptr(r_comp), ptr(l_t), ptr(r_t), n, ptr(out), ptr(status)))
        return out, status

    def time_of_impact_complex2(self, strategy, l_comp, r_comp, l_t0, l_t1, r_t0, r_t1, settings=None,
                                iter_cap=1024):
        settings = settings or Settings.default()
        n = len(l_comp)
        out = np.zeros(n, abi.CONTACT2_DT)
        status = np.zeros(n, np.uint8)
        self._ck(self.lib.cb2_gjk2_time_of_impact_complex(
            self.h, int(strategy), C.byref(settings), ptr(l_comp), ptr(r_comp), ptr(l_t0), ptr(l_t1),
            ptr(r_t0), ptr(r_t1), n, int(iter_cap), ptr(out), ptr(status)))
        return out, status

    # ---- narrow phase, host buffers --------------------------------------------------------
    def intersection(self, strategy, l_shape, r_shape, l_t, r_t, settings=None, out=None,
                     status=None):
        settings = settings or Settings.default()
        n = len(l_shape)
        out = np.zeros(n, CONTACT_DT) if out is None else out
        status = np.zeros(n, np.uint8) if status is None else status
        self._ck(self.lib.cb2_gjk3_intersection(
            self.h, int(strategy), C.byref(settings), ptr(l_shape), ptr(r_shape), ptr(l_t),
            ptr(r_t), n, ptr(out), ptr(status)))
        return out, status

    def intersection_bodies(self, strategy, body_shape, body_t, pair_body, settings=None, out=None,
                            status=None):
        """GJK3::intersection over a pair list that indexes one body table (gjk/mod.rs:362-391 called
        as `intersection(&shape[a], &pose[a], &shape[b], &pose[b])` for the (a, b) of a broad phase):
        pair_body is an (n, 2) uint32 array."""
        settings = settings or Settings.default()
        pair_body = np.ascontiguousarray(pair_body, np.uint32).reshape(-1, 2)
        n = len(pair_body)
        out = np.zeros(n, CONTACT_DT) if out is None else out
        status = np.zeros(n, np.uint8) if status is None else status
        self._ck(self.lib.cb2_gjk3_intersection_bodies(
            self.h, int(strategy), C.byref(settings), ptr(body_shape), ptr(body_t), len(body_shape),
            ptr(pair_body), n, ptr(out), ptr(status)))
        return out, status

    def distance(self, l_shape, r_shape, l_t, r_t, settings=None):
        settings = settings or Settings.default()
        n = len(l_shape)
        out = np.zeros(n, np.float32)
        status = np.zeros(n, np.uint8)
        self._ck(self.lib.cb2_gjk3_distance(
            self.h, C.byref(settings), ptr(l_shape), ptr(r_shape), ptr(l_t), ptr(r_t), n,
            ptr(out), ptr(status)))
        return out, status

    def time_of_impact(self, l_shape, r_shape, l_t0, l_t1, r_t0, r_t1, settings=None,
                       iter_cap=1024):
        settings = settings or Settings.default()
        n = len(l_shape)
        out = np.zeros(n, CONTACT_DT)
        status = np.zeros(n, np.uint8)
        self._ck(self.lib.cb2_gjk3_time_of_impact(
            self.h, C.byref(settings), ptr(l_shape), ptr(r_shape), ptr(l_t0), ptr(l_t1),
            ptr(r_t0), ptr(r_t1), n, int(iter_cap), ptr(out), ptr(status)))
        return out, status

    def intersect(self, l_shape, r_shape, l_t, r_t, settings=None):
        """GJK::intersect (gjk/mod.rs:101-146): (simplex (n, 4) SupportPoints, status)."""
        from .abi import SUPPORT_POINT_DT
        settings = settings or Settings.default()
        n = len(l_shape)
        simplex = np.zeros((n, 4), SUPPORT_POINT_DT)
        status = np.zeros(n, np.uint8)
        self._ck(self.lib.cb2_gjk3_intersect(
            self.h, C.byref(settings), ptr(l_shape), ptr(r_shape), ptr(l_t), ptr(r_t), n,
            ptr(simplex), ptr(status)))
        return simplex, status

    def get_contact_manifold(self, simplex, l_shape, r_shape, l_t, r_t, simplex_len=None,
                             settings=None):
        """GJK::get_contact_manifold (gjk/mod.rs:326-344): EPA3::process on caller-held simplices."""
        settings = settings or Settings.default()
        n = len(l_shape)
        out = np.zeros(n, CONTACT_DT)
        status = np.zeros(n, np.uint8)
        self._ck(self.lib.cb2_gjk3_get_contact_manifold(
            self.h, C.byref(settings), ptr(simplex), ptr(simplex_len), ptr(l_shape), ptr(r_shape),
            ptr(l_t), ptr(r_t), n, ptr(out), ptr(status)))
        return out, status

    def world_aabbs(self, shape, t):
        n = len(shape)
        out = np.zeros(n, AABB_DT)
        self._ck(self.lib.cb2_world_aabbs(self.h, ptr(shape), ptr(t), n, ptr(out)))
        return out

    # ---- narrow phase, device pointers (ints) ------------------------------------------------
    def intersection_dev(self, strategy, l_shape, r_shape, l_t, r_t, n, out, status, stream=0,
                         settings=None):
        settings = settings or Settings.default()
        self._ck(self.lib.cb2_gjk3_intersection_dev(
            self.h, int(strategy), C.byref(settings), l_shape, r_shape, l_t, r_t, int(n), out,
            status, stream))

    def distance_dev(self, l_shape, r_shape, l_t, r_t, n, out, status, stream=0, settings=None):
        settings = settings or Settings.default()
        self._ck(self.lib.cb2_gjk3_distance_dev(
            self.h, C.byref(settings), l_shape, r_shape, l_t, r_t, int(n), out, status, stream))

    def time_of_impact_dev(self, l_shape, r_shape, l_t0, l_t1, r_t0, r_t1, n, out, status,
                           stream=0, settings=None, iter_cap=1024):
        settings = settings or Settings.default()
        self._ck(self.lib.cb2_gjk3_time_of_impact_dev(
            self.h, C.byref(settings), l_shape, r_shape, l_t0, l_t1, r_t0, r_t1, int(n),
            int(iter_cap), out, status, stream))

    def intersection_bodies_dev(self, strategy, body_shape, body_t, pair_body, n, out, status,
                                stream=0, settings=None):
        settings = settings or Settings.default()
        self._ck(self.lib.cb2_gjk3_intersection_bodies_dev(
            self.h, int(strategy), C.byref(settings), body_shape, body_t, pair_body, int(n), out,
            status, stream))

    def world_aabbs_dev(self, shape, t, n, out, stream=0):
        self._ck(self.lib.cb2_world_aabbs_dev(self.h, shape, t, int(n), out, stream))

    # ---- broad phase ------------------------------------------------------------------------------
    # ---- composite shapes: &[(Primitive, local transform)] (gjk/mod.rs:412-588) -----------------
    def upload_composites(self, part_off, part_shape, part_local):
        part_off = np.ascontiguousarray(part_off, np.uint32)
        part_shape = np.ascontiguousarray(part_shape, np.uint32)
        part_local = np.ascontiguousarray(part_local, XFORM_DT)
        assert len(part_shape) == len(part_local) == int(part_off[-1])
        self._ck(self.lib.cb2_composites_upload(self.h, ptr(part_off), ptr(part_shape),
                                                ptr(part_local), len(part_off) - 1))

    def intersection_complex(self, strategy, l_comp, r_comp, l_t, r_t, settings=None):
        n = len(l_comp)
        s = settings or Settings.default()
        out = np.zeros(n, CONTACT_DT)
        st = np.zeros(n, np.uint8)
        self._ck(self.lib.cb2_gjk3_intersection_complex(
            self.h, int(strategy), C.byref(s), ptr(np.ascontiguousarray(l_comp, np.uint32)),
            ptr(np.ascontiguousarray(r_comp, np.uint32)), ptr(np.ascontiguousarray(l_t, XFORM_DT)),
            ptr(np.ascontiguousarray(r_t, XFORM_DT)), n, ptr(out), ptr(st)))
        return out, st

    def distance_complex(self, l_comp, r_comp, l_t, r_t, settings=None):
        n = len(l_comp)
        s = settings or Settings.default()
        out = np.zeros(n, np.float32)
        st = np.zeros(n, np.uint8)
        self._ck(self.lib.cb2_gjk3_distance_complex(
            self.h, C.byref(s), ptr(np.ascontiguousarray(l_comp, np.uint32)),
            ptr(np.ascontiguousarray(r_comp, np.uint32)), ptr(np.ascontiguousarray(l_t, XFORM_DT)),
            ptr(np.ascontiguousarray(r_t, XFORM_DT)), n, ptr(out), ptr(st)))
        return out, st

    def time_of_impact_complex(self, strategy, l_comp, r_comp, l_t0, l_t1, r_t0, r_t1,
                               settings=None, iter_cap=1024):
        n = len(l_comp)
        s = settings or Settings.default()
        out = np.zeros(n, CONTACT_DT)
        st = np.zeros(n, np.uint8)
        a = [np.ascontiguousarray(x, XFORM_DT) for x in (l_t0, l_t1, r_t0, r_t1)]
        self._ck(self.lib.cb2_gjk3_time_of_impact_complex(
            self.h, int(strategy), C.byref(s), ptr(np.ascontiguousarray(l_comp, np.uint32)),
            ptr(np.ascontiguousarray(r_comp, np.uint32)), ptr(a[0]), ptr(a[1]), ptr(a[2]),
            ptr(a[3]), n, int(iter_cap), ptr(out), ptr(st)))
        return out, st

    def sap3(self, aabbs, axis=0, cap=None):
        """Host-buffer SweepAndPrune3::find_collider_pairs. Returns (perm, pairs, axis_next)."""
        aabbs = np.ascontiguousarray(aabbs, AABB_DT)
        n = len(aabbs)
        perm = np.zeros(max(n, 1), np.uint32)
        cap = int(cap) if cap is not None else max(16, 8 * n)
        while True:
            pairs = np.zeros((max(cap, 1), 2), np.uint32)
            ax = C.c_uint32(axis)
            npairs = C.c_uint64(0)
            rc = self.lib.cb2_sap3_find_collider_pairs(self.h, ptr(aabbs), n, C.byref(ax),
                                                       ptr(perm), ptr(pairs), cap,
                                                       C.byref(npairs))
            if rc == abi.ERR_NOSPC:
                cap = int(npairs.value)
                continue
            self._ck(rc)
            return perm[:n], pairs[:npairs.value], ax.value

    def sap2(self, aabbs2, axis=0, cap=None):
        """SweepAndPrune2::find_collider_pairs; aabbs2: (n, 4) f32 rows (min.x, min.y, max.x, max.y)."""
        b = np.ascontiguousarray(aabbs2, np.float32).reshape(-1, 4)
        n = len(b)
        perm = np.zeros(max(n, 1), np.uint32)
        cap = int(cap) if cap is not None else max(16, 8 * n)
        while True:
            pairs = np.zeros((max(cap, 1), 2), np.uint32)
            ax = C.c_uint32(axis)
            npairs = C.c_uint64(0)
            rc = self.lib.cb2_sap2_find_collider_pairs(self.h, ptr(b), n, C.byref(ax), ptr(perm),
                                                       ptr(pairs), cap, C.byref(npairs))
            if rc == abi.ERR_NOSPC:
                cap = int(npairs.value)
                continue
            self._ck(rc)
            return perm[:n], pairs[:npairs.value], ax.value

    def brute_force(self, aabbs, cap=None):
        """BruteForce::find_collider_pairs (brute_force.rs:20-39): (left, right) in the caller's
        indices, left ascending then right ascending."""
        aabbs = np.ascontiguousarray(aabbs, AABB_DT)
        n = len(aabbs)
        cap = int(cap) if cap is not None else max(16, 8 * n)
        while True:
            pairs = np.zeros((max(cap, 1), 2), np.uint32)
            npairs = C.c_uint64(0)
            rc = self.lib.cb2_brute_force_find_collider_pairs(self.h, ptr(aabbs), n, ptr(pairs), cap,
                                                              C.byref(npairs))
            if rc == abi.ERR_NOSPC:
                cap = int(npairs.value)
                continue
            self._ck(rc)
            return pairs[:npairs.value]

    def sap3_dev(self, aabbs, n, axis, perm_out, pairs_out, cap, stream=0):
        npairs = C.c_uint64(0)
        rc = self.lib.cb2_sap3_find_collider_pairs_dev(self.h, aabbs, int(n), int(axis), perm_out,
                                                       pairs_out, int(cap), C.byref(npairs), stream)
        if rc == abi.ERR_NOSPC:
            return rc, int(npairs.value)
        self._ck(rc)
        return 0, int(npairs.value)


    def sap3_unordered_dev(self, aabbs, n, axis, perm_out, pairs_out, cap, stream=0):
        """The pair SET (unspecified order) of SweepAndPrune3::find_collider_pairs (device pointers)."""
        npairs = C.c_uint64(0)
        rc = self.lib.cb2_sap3_find_collider_pairs_unordered_dev(self.h, aabbs, int(n), int(axis), perm_out,
                                                                 pairs_out, int(cap), C.byref(npairs), stream)
        if rc == abi.ERR_NOSPC:
            return rc, int(npairs.value)
        self._ck(rc)
        return 0, int(npairs.value)

    def sap3_shard_dev(self, aabbs, n, axis, shard, n_shards, perm_out, pairs_out, cap, stream=0):
        """This GPU's shard of SweepAndPrune3::find_collider_pairs (device pointers)."""
        npairs = C.c_uint64(0)
        rc = self.lib.cb2_sap3_find_collider_pairs_shard_dev(
            self.h, aabbs, int(n), int(axis), int(shard), int(n_shards), perm_out, pairs_out, int(cap),
            C.byref(npairs), stream)
        if rc == abi.ERR_NOSPC:
            return rc, int(npairs.value)
        self._ck(rc)
        return 0, int(npairs.value)


# ---- reference-shaped single-pair API (batch of 1; still the GPU) -----------------------------
class CollisionStrategy:
    """contact.rs:16-22"""
    FullResolution = FULL_RESOLUTION
    CollisionOnly = COLLISION_ONLY


@dataclass
class Contact:
    """contact.rs:30-46"""
    strategy: int
    normal: tuple
    penetration_depth: float
    contact_point: tuple
    time_of_impact: float

    @staticmethod
    def from_record(r):
        return Contact(int(r["strategy"]), tuple(np.float32(x) for x in r["normal"]),
                       np.float32(r["depth"]), tuple(np.float32(x) for x in r["point"]),
                       np.float32(r["toi"]))


@dataclass
class Sphere:
    """primitive/sphere.rs:14-24"""
    radius: float


@dataclass
class Cuboid:
    """primitive/cuboid.rs:18-42 (dim = full extents)"""
    x: float
    y: float
    z: float


@dataclass
class Cube:
    """primitive/cuboid.rs:144-158: Cuboid::new(dim, dim, dim)"""
    dim: float


@dataclass
class Quad:
    """primitive/quad.rs:24-57: rectangle in the z = 0 plane (dim = full extents)"""
    x: float
    y: float


@dataclass
class Cylinder:
    """primitive/cylinder.rs:21-36 (axis y)"""
    half_height: float
    radius: float


@dataclass
class Capsule:
    """primitive/capsule.rs:22-37 (axis y)"""
    half_height: float
    radius: float


@dataclass
class Particle3:
    """primitive/particle.rs: a point; Primitive3::Particle (primitive3.rs:24-25)"""


class ConvexPolyhedron:
    """primitive/polyhedron.rs:100-139: ConvexPolyhedron::new(vertices) (VertexOnly mode) or
    ConvexPolyhedron::new_with_faces(vertices, faces) (HalfEdge mode: hill-climb support)."""

    def __init__(self, vertices, faces=None):
        self.vertices = np.ascontiguousarray(vertices, np.float32).reshape(-1, 3)
        self.faces = None if faces is None else np.ascontiguousarray(faces, np.uint32).reshape(-1, 3)

    @staticmethod
    def new_with_faces(vertices, faces):
        return ConvexPolyhedron(vertices, faces)


@dataclass
class Decomposed:
    """cgmath Decomposed<Vector3<f32>, Quaternion<f32>>; rot = (s, x, y, z)."""
    disp: tuple = (0.0, 0.0, 0.0)
    rot: tuple = (1.0, 0.0, 0.0, 0.0)
    scale: float = 1.0

    def record(self):
        t = np.zeros(1, XFORM_DT)
        t["scale"] = self.scale
        t["q"] = self.rot
        t["d"] = self.disp
        return t


def _shape_table(prims):
    shapes = np.zeros(len(prims), SHAPE_DT)
    verts, faces = [], []
    off = foff = 0
    for i, p in enumerate(prims):
        if isinstance(p, Sphere):
            shapes[i]["kind"] = SPHERE
            shapes[i]["p"] = (p.radius, 0, 0)
        elif isinstance(p, Cuboid):
            shapes[i]["kind"] = CUBOID
            shapes[i]["p"] = (p.x, p.y, p.z)
        elif isinstance(p, Cube):
            shapes[i]["kind"] = CUBE
            shapes[i]["p"] = (p.dim, 0, 0)
        elif isinstance(p, Quad):
            shapes[i]["kind"] = QUAD
            shapes[i]["p"] = (p.x, p.y, 0)
        elif isinstance(p, Cylinder):
            shapes[i]["kind"] = CYLINDER
            shapes[i]["p"] = (p.half_height, p.radius, 0)
        elif isinstance(p, Capsule):
            shapes[i]["kind"] = CAPSULE
            shapes[i]["p"] = (p.half_height, p.radius, 0)
        elif isinstance(p, Particle3):
            shapes[i]["kind"] = PARTICLE
        elif isinstance(p, ConvexPolyhedron):
            shapes[i]["kind"] = POLYHEDRON
            shapes[i]["vtx_off"] = off
            shapes[i]["vtx_cnt"] = len(p.vertices)
            if p.faces is not None:
                shapes[i]["kind"] = POLYHEDRON_HE
                shapes[i:i + 1]["p"].view(np.uint32)[0, 0] = foff
                shapes[i:i + 1]["p"].view(np.uint32)[0, 1] = len(p.faces)
                faces.append(p.faces)
                foff += len(p.faces)
            verts.append(p.vertices)
            off += len(p.vertices)
        else:
            raise TypeError(f"unsupported primitive {type(p).__name__}")
    v = np.concatenate(verts) if verts else None
    f = np.concatenate(faces) if faces else None
    return shapes, v, f


class MultiContext:
    """One cb2_multi: several GPUs of one box behind one handle (cb2_create_multi).  Host-buffer
    batch calls are split into contiguous slices, one per device; results come back in place."""

    def __init__(self, devices):
        self.lib = abi.load()
        devs = (C.c_int * len(devices))(*[int(d) for d in devices])
        h = C.c_void_p()
        rc = self.lib.cb2_create_multi(devs, len(devices), C.byref(h))
        if rc != 0:
            raise Cb2Error(rc, self.lib.cb2_last_error(None).decode())
        self.h = h
        self.devices = list(devices)

    def close(self):
        if getattr(self, "h", None):
            self.lib.cb2_destroy_multi(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            raise Cb2Error(rc, self.lib.cb2_multi_last_error(self.h).decode())

    @property
    def device_count(self):
        return int(self.lib.cb2_multi_device_count(self.h))

    def upload_shapes(self, shapes, verts=None):
        shapes = np.ascontiguousarray(shapes, SHAPE_DT)
        v = None if verts is None else np.ascontiguousarray(verts, np.float32).reshape(-1, 3)
        self._ck(self.lib.cb2_multi_shapes_upload(self.h, ptr(shapes), len(shapes), ptr(v),
                                                  0 if v is None else len(v)))

    def intersection(self, strategy, l_shape, r_shape, l_t, r_t, settings=None, out=None, status=None):
        settings = settings or Settings.default()
        n = len(l_shape)
        out = np.zeros(n, CONTACT_DT) if out is None else out
        status = np.zeros(n, np.uint8) if status is None else status
        self._ck(self.lib.cb2_multi_gjk3_intersection(
            self.h, int(strategy), C.byref(settings), ptr(l_shape), ptr(r_shape), ptr(l_t), ptr(r_t), n,
            ptr(out), ptr(status)))
        return out, status

    def distance(self, l_shape, r_shape, l_t, r_t, settings=None):
        settings = settings or Settings.default()
        n = len(l_shape)
        out = np.zeros(n, np.float32)
        status = np.zeros(n, np.uint8)
        self._ck(self.lib.cb2_multi_gjk3_distance(
            self.h, C.byref(settings), ptr(l_shape), ptr(r_shape), ptr(l_t), ptr(r_t), n, ptr(out),
            ptr(status)))
        return out, status

    def time_of_impact(self, l_shape, r_shape, l_t0, l_t1, r_t0, r_t1, settings=None, iter_cap=1024):
        settings = settings or Settings.default()
        n = len(l_shape)
        out = np.zeros(n, CONTACT_DT)
        status = np.zeros(n, np.uint8)
        self._ck(self.lib.cb2_multi_gjk3_time_of_impact(
            self.h, C.byref(settings), ptr(l_shape), ptr(r_shape), ptr(l_t0), ptr(l_t1), ptr(r_t0),
            ptr(r_t1), n, int(iter_cap), ptr(out), ptr(status)))
        return out, status

    def sap3(self, aabbs, axis=0, cap=None):
        aabbs = np.ascontiguousarray(aabbs, AABB_DT)
        n = len(aabbs)
        perm = np.zeros(max(n, 1), np.uint32)
        cap = int(cap) if cap is not None else max(16, 8 * n)
        while True:
            pairs = np.zeros((max(cap, 1), 2), np.uint32)
            ax = C.c_uint32(axis)
            npairs = C.c_uint64(0)
            rc = self.lib.cb2_multi_sap3_find_collider_pairs(self.h, ptr(aabbs), n, C.byref(ax), ptr(perm),
                                                             ptr(pairs), cap, C.byref(npairs))
            if rc == abi.ERR_NOSPC:
                cap = int(npairs.value)
                continue
            self._ck(rc)
            return perm[:n], pairs[:npairs.value], ax.value


_I0 = np.array([0], np.uint32)
_I1 = np.array([1], np.uint32)


class GJK3:
    """GJK<SimplexProcessor3<f32>, EPA3<f32>, f32> (gjk/mod.rs:30,44-84), batch-of-1 wrappers."""

    def __init__(self, ctx, distance_tolerance=1e-6, continuous_tolerance=1e-6,
                 epa_tolerance=1e-5, max_iterations=100):
        self.ctx = ctx
        self.settings = Settings(np.float32(distance_tolerance), np.float32(continuous_tolerance),
                                 np.float32(epa_tolerance), int(max_iterations))

    @staticmethod
    def new(ctx):
        return GJK3(ctx)

    @staticmethod
    def new_with_settings(ctx, distance_tolerance, continuous_tolerance, epa_tolerance,
                          max_iterations):
        return GJK3(ctx, distance_tolerance, continuous_tolerance, epa_tolerance, max_iterations)

    def _upload(self, left, right):
        shapes, verts, faces = _shape_table([left, right])
        self.ctx.upload_shapes(shapes, verts, faces)

    def intersection(self, strategy, left, left_transform, right, right_transform):
        """gjk/mod.rs:362-391 -> Option<Contact>"""
        self._upload(left, right)
        out, st = self.ctx.intersection(strategy, _I0, _I1, left_transform.record(),
                                        right_transform.record(), self.settings)
        if st[0] > 1:
            raise ReferencePanic(int(st[0]))
        return Contact.from_record(out[0]) if st[0] == 1 else None

    def intersect(self, left, left_transform, right, right_transform):
        """gjk/mod.rs:101-146 -> Option<Simplex>: the tetrahedron's four SupportPoints as a
        (4,) record array (fields v, sup_a, sup_b), or None"""
        self._upload(left, right)
        sx, st = self.ctx.intersect(_I0, _I1, left_transform.record(), right_transform.record(),
                                    self.settings)
        if st[0] > 1:
            raise ReferencePanic(int(st[0]))
        return sx[0].copy() if st[0] == 1 else None

    def get_contact_manifold(self, simplex, left, left_transform, right, right_transform):
        """gjk/mod.rs:326-344 -> Option<Contact>: EPA3::process on a simplex from `intersect`
        (a simplex of fewer than four points gives None, epa3d.rs:38-40)"""
        from .abi import SUPPORT_POINT_DT
        self._upload(left, right)
        held = np.zeros((1, 4), SUPPORT_POINT_DT)
        k = min(len(simplex), 4)
        held[0, :k] = simplex[:k]
        out, st = self.ctx.get_contact_manifold(held, _I0, _I1, left_transform.record(),
                                                right_transform.record(),
                                                np.array([len(simplex)], np.uint32), self.settings)
        if st[0] > 1:
            raise ReferencePanic(int(st[0]))
        return Contact.from_record(out[0]) if st[0] == 1 else None

    def distance(self, left, left_transform, right, right_transform):
        """gjk/mod.rs:271-321 -> Option<f32>"""
        self._upload(left, right)
        out, st = self.ctx.distance(_I0, _I1, left_transform.record(), right_transform.record(),
                                    self.settings)
        if st[0] > 1:
            raise ReferencePanic(int(st[0]))
        return np.float32(out[0]) if st[0] == 1 else None

    def intersection_time_of_impact(self, left, left_range, right, right_range, iter_cap=1024):
        """gjk/mod.rs:163-256; ranges are (start, end) pairs -> Option<Contact>"""
        self._upload(left, right)
        out, st = self.ctx.time_of_impact(_I0, _I1, left_range[0].record(), left_range[1].record(),
                                          right_range[0].record(), right_range[1].record(),
                                          self.settings, iter_cap)
        if st[0] > 1:
            raise ReferencePanic(int(st[0]))
        return Contact.from_record(out[0]) if st[0] == 1 else None


# ---- 2-D mirror: Primitive2 (primitive/primitive2.rs:20-33), Decomposed<Vector2, Basis2>, GJK2 --------
@dataclass
class Particle2:
    """primitive/particle.rs:39"""


@dataclass
class Line2:
    """line.rs:21-28 as a Primitive (primitive/line.rs:7-26)"""
    origin: tuple
    dest: tuple


@dataclass
class Circle:
    """primitive/circle.rs:14-24"""
    radius: float


@dataclass
class Rectangle:
    """primitive/rectangle.rs:18-42 (full extents)"""
    x: float
    y: float


@dataclass
class Square:
    """primitive/rectangle.rs:122-134"""
    dim: float


class ConvexPolygon:
    """primitive/polygon.rs:17-24: vertices in order around the polygon"""

    def __init__(self, vertices):
        self.vertices = np.ascontiguousarray(vertices, np.float32).reshape(-1, 2)


@dataclass
class Basis2:
    """cgmath Basis2<f32>: a column-major Matrix2 (c0x, c0y, c1x, c1y)."""
    mat: tuple = (1.0, 0.0, 0.0, 1.0)

    @staticmethod
    def from_angle(rad):
        """Rotation2::from_angle(Rad(t)) = Matrix2::new(cos t, sin t, -sin t, cos t), in f32."""
        s, c = np.sin(np.float32(rad), dtype=np.float32), np.cos(np.float32(rad), dtype=np.float32)
        return Basis2((c, s, -s, c))


@dataclass
class Decomposed2:
    """cgmath Decomposed<Vector2<f32>, Basis2<f32>>"""
    disp: tuple = (0.0, 0.0)
    rot: Basis2 = None
    scale: float = 1.0

    def record(self):
        t = np.zeros(1, abi.XFORM2_DT)
        t["scale"] = self.scale
        t["m"] = (self.rot or Basis2()).mat
        t["d"] = self.disp
        return t


def _shape_table2(prims):
    shapes = np.zeros(len(prims), abi.SHAPE2_DT)
    verts = []
    off = 0
    for i, p in enumerate(prims):
        if isinstance(p, Particle2):
            shapes[i]["kind"] = abi.PARTICLE2
        elif isinstance(p, Line2):
            shapes[i]["kind"] = abi.LINE2
            shapes[i]["p"] = tuple(p.origin) + tuple(p.dest)
        elif isinstance(p, Circle):
            shapes[i]["kind"] = abi.CIRCLE
            shapes[i]["p"] = (p.radius, 0, 0, 0)
        elif isinstance(p, Rectangle):
            shapes[i]["kind"] = abi.RECTANGLE
            shapes[i]["p"] = (p.x, p.y, 0, 0)
        elif isinstance(p, Square):
            shapes[i]["kind"] = abi.SQUARE
            shapes[i]["p"] = (p.dim, 0, 0, 0)
        elif isinstance(p, ConvexPolygon):
            shapes[i]["kind"] = abi.POLYGON
            shapes[i]["vtx_off"] = off
            shapes[i]["vtx_cnt"] = len(p.vertices)
            verts.append(p.vertices)
            off += len(p.vertices)
        else:
            raise TypeError(f"unsupported 2-D primitive {type(p).__name__}")
    return shapes, (np.concatenate(verts) if verts else None)


@dataclass
class Contact2:
    """Contact<Point2<f32>> (contact.rs:30-46)"""
    strategy: int
    normal: tuple
    penetration_depth: float
    contact_point: tuple
    time_of_impact: float

    @staticmethod
    def from_record(r):
        return Contact2(int(r["strategy"]), tuple(np.float32(x) for x in r["normal"]),
                        np.float32(r["depth"]), tuple(np.float32(x) for x in r["point"]),
                        np.float32(r["toi"]))


class GJK2:
    """GJK<SimplexProcessor2<f32>, EPA2<f32>, f32> (gjk/mod.rs:27,44-84), batch-of-1 wrappers."""

    def __init__(self, ctx, distance_tolerance=1e-6, continuous_tolerance=1e-6,
                 epa_tolerance=1e-5, max_iterations=100):
        self.ctx = ctx
        self.settings = Settings(np.float32(distance_tolerance), np.float32(continuous_tolerance),
                                 np.float32(epa_tolerance), int(max_iterations))

    @staticmethod
    def new(ctx):
        return GJK2(ctx)

    @staticmethod
    def new_with_settings(ctx, distance_tolerance, continuous_tolerance, epa_tolerance,
                          max_iterations):
        return GJK2(ctx, distance_tolerance, continuous_tolerance, epa_tolerance, max_iterations)

    def _upload(self, left, right):
        self.ctx.upload_shapes2(*_shape_table2([left, right]))

    def intersect(self, left, left_transform, right, right_transform):
        """gjk/mod.rs:101-146 -> whether Some(simplex) (the simplex itself stays on the device)"""
        return self.intersection(CollisionStrategy.CollisionOnly, left, left_transform, right,
                                 right_transform) is not None

    def intersection(self, strategy, left, left_transform, right, right_transform):
        """gjk/mod.rs:362-391 -> Option<Contact>"""
        self._upload(left, right)
        out, st = self.ctx.intersection2(strategy, _I0, _I1, left_transform.record(),
                                         right_transform.record(), self.settings)
        if st[0] > 1:
            raise ReferencePanic(int(st[0]))
        return Contact2.from_record(out[0]) if st[0] == 1 else None

    def distance(self, left, left_transform, right, right_transform):
        """gjk/mod.rs:271-321 -> Option<f32>"""
        self._upload(left, right)
        out, st = self.ctx.distance2(_I0, _I1, left_transform.record(), right_transform.record(),
                                     self.settings)
        if st[0] > 1:
            raise ReferencePanic(int(st[0]))
        return np.float32(out[0]) if st[0] == 1 else None

    def intersection_time_of_impact(self, left, left_range, right, right_range, iter_cap=1024):
        """gjk/mod.rs:163-256; ranges are (start, end) pairs -> Option<Contact>"""
        self._upload(left, right)
        out, st = self.ctx.time_of_impact2(_I0, _I1, left_range[0].record(), left_range[1].record(),
                                           right_range[0].record(), right_range[1].record(),
                                           self.settings, iter_cap)
        if st[0] > 1:
            raise ReferencePanic(int(st[0]))
        return Contact2.from_record(out[0]) if st[0] == 1 else None


class BruteForce:
    """algorithm/broad_phase/brute_force.rs:8-39"""

    def __init__(self, ctx):
        self.ctx = ctx

    def find_collider_pairs(self, aabbs):
        return [tuple(int(x) for x in p) for p in self.ctx.brute_force(aabbs)]


class SweepAndPrune3:
    """SweepAndPrune<Variance3<f32, Aabb3<f32>>> (sweep_prune.rs:16, 34-136)."""

    def __init__(self, ctx, sweep_axis=0):
        self.ctx = ctx
        self.sweep_axis = sweep_axis

    @staticmethod
    def new(ctx):
        return SweepAndPrune3(ctx, 0)

    @staticmethod
    def with_sweep_axis(ctx, sweep_axis):
        return SweepAndPrune3(ctx, sweep_axis)

    def get_sweep_axis(self):
        return self.sweep_axis

    def find_collider_pairs(self, shapes, bound=lambda s: s):
        """Sorts `shapes` (a list) in place like the reference and returns [(a, b), ...] over
        the sorted list.  bound(shape) -> (min_xyz, max_xyz) plays HasBound::bound()."""
        n = len(shapes)
        if n <= 1:
            return []
        a = np.zeros(n, AABB_DT)
        for i, s in enumerate(shapes):
            mn, mx = bound(s)
            a[i]["min"] = mn
            a[i]["max"] = mx
        perm, pairs, axis = self.ctx.sap3(a, self.sweep_axis)
        shapes[:] = [shapes[int(i)] for i in perm]
        self.sweep_axis = axis
        return [(int(x), int(y)) for x, y in pairs]
