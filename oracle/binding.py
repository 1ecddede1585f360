"""ctypes binding of the CPU oracle (oracle/collision_oracle.cpp).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs — never by the product package.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_BUILD = os.path.join(_HERE, "_build")

SHAPE_DT = np.dtype([("kind", "<u4"), ("p", "<f4", (3,)), ("vtx_off", "<u4"), ("vtx_cnt", "<u4")])
XFORM_DT = np.dtype([("scale", "<f4"), ("q", "<f4", (4,)), ("d", "<f4", (3,))])
CONTACT_DT = np.dtype([("strategy", "<u4"), ("normal", "<f4", (3,)), ("depth", "<f4"),
                       ("point", "<f4", (3,)), ("toi", "<f4")])
AABB_DT = np.dtype([("min", "<f4", (3,)), ("max", "<f4", (3,))])
SUPPORT_POINT_DT = np.dtype([("v", "<f4", (3,)), ("sup_a", "<f4", (3,)), ("sup_b", "<f4", (3,))])


class Settings(C.Structure):
    _fields_ = [("distance_tolerance", C.c_float), ("continuous_tolerance", C.c_float),
                ("epa_tolerance", C.c_float), ("max_iterations", C.c_uint32)]

    @staticmethod
    def default():
        return Settings(np.float32(1e-6), np.float32(1e-6), np.float32(1e-5), 100)


def build(force=False):
    """make -C oracle (gcc only). Building the checker is not using it."""
    out = os.path.join(_BUILD, "liboracle.so")
    stale = force
    for lib, src in (("liboracle.so", "collision_oracle.cpp"), ("liboracle2d.so", "collision_oracle2d.cpp")):
        o, c = os.path.join(_BUILD, lib), os.path.join(_HERE, src)
        stale = stale or not os.path.exists(o) or os.path.getmtime(o) < os.path.getmtime(c)
    if stale:
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return out


def _p(a):
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.c_void_p)


class Oracle:
    def __init__(self, count_ops=False):
        build()
        name = "liboracle_count.so" if count_ops else "liboracle.so"
        self.lib = C.CDLL(os.path.join(_BUILD, name))
        self.lib.orc_sap3_find_collider_pairs.restype = C.c_uint64
        self.lib.orc_brute_force_count.restype = C.c_uint64
        self.lib.orc_brute_force_pairs.restype = C.c_uint64
        self.lib.orc_ops_reset.restype = C.c_uint64
        self.lib.orc_closest_point_to_origin.restype = C.c_uint8
        self.lib.orc_epa3_process.restype = C.c_uint8
        self.lib.orc_gjk3_intersect.restype = C.c_uint8
        self.count_ops = count_ops

    def hw_threads(self):
        return int(self.lib.orc_hw_threads())

    # ---- unit hooks ---------------------------------------------------------------
    def support_point(self, shape, verts, xform, direction):
        out = np.zeros(3, np.float32)
        st = C.c_uint8(0)
        d = np.asarray(direction, np.float32)
        self.lib.orc_support_point(_p(shape), _p(verts), _p(xform), _p(d), _p(out), C.byref(st))
        return out, st.value

    def reduce_to_closest_feature(self, pts, v=(0, 0, 0)):
        buf = np.zeros((8, 3), np.float32)
        pts = np.asarray(pts, np.float32).reshape(-1, 3)
        buf[:len(pts)] = pts
        n = C.c_uint32(len(pts))
        vv = np.asarray(v, np.float32).copy()
        hit = self.lib.orc_reduce_to_closest_feature(_p(buf), C.byref(n), _p(vv))
        return bool(hit), buf[:n.value].copy(), vv

    def check_side(self, pts, above, ignore_ab):
        buf = np.zeros((8, 3), np.float32)
        pts = np.asarray(pts, np.float32).reshape(-1, 3)
        buf[:len(pts)] = pts
        n = C.c_uint32(len(pts))
        vv = np.zeros(3, np.float32)
        self.lib.orc_check_side(_p(buf), C.byref(n), int(above), int(ignore_ab), _p(vv))
        return buf[:n.value].copy(), vv

    def closest_point_to_origin(self, pts):
        buf = np.zeros((8, 3), np.float32)
        pts = np.asarray(pts, np.float32).reshape(-1, 3)
        buf[:len(pts)] = pts
        n = C.c_uint32(len(pts))
        out = np.zeros(3, np.float32)
        st = self.lib.orc_closest_point_to_origin(_p(buf), C.byref(n), _p(out))
        return st, out, buf[:n.value].copy()

    def closest_point_on_edge(self, start, end, point):
        out = np.zeros(3, np.float32)
        a, b, c = (np.asarray(x, np.float32) for x in (start, end, point))
        self.lib.orc_closest_point_on_edge(_p(a), _p(b), _p(c), _p(out))
        return out

    def barycentric(self, p, a, b, c):
        out = np.zeros(3, np.float32)
        p, a, b, c = (np.asarray(x, np.float32) for x in (p, a, b, c))
        self.lib.orc_barycentric(_p(p), _p(a), _p(b), _p(c), _p(out))
        return out

    def plane_from_points(self, a, b, c):
        out = np.zeros(4, np.float32)
        a, b, c = (np.asarray(x, np.float32) for x in (a, b, c))
        ok = self.lib.orc_plane_from_points(_p(a), _p(b), _p(c), _p(out))
        return (out[:3].copy(), out[3]) if ok else None

    def polytope(self, verts, add=()):
        verts = np.asarray(verts, np.float32).reshape(-1, 3)
        add = np.asarray(add, np.float32).reshape(-1, 3)
        idx = np.zeros((512, 3), np.uint32)
        nd = np.zeros((512, 4), np.float32)
        nf, closest, nv = C.c_uint32(0), C.c_uint32(0), C.c_uint32(0)
        self.lib.orc_polytope(_p(verts), C.c_uint32(len(verts)), _p(add), C.c_uint32(len(add)),
                              _p(idx), _p(nd), C.byref(nf), C.byref(closest), C.byref(nv))
        return idx[:nf.value].copy(), nd[:nf.value].copy(), closest.value, nv.value

    def remove_or_add_edge(self, edges, e):
        buf = np.zeros((64, 2), np.uint32)
        edges = np.asarray(edges, np.uint32).reshape(-1, 2)
        buf[:len(edges)] = edges
        n = C.c_uint32(len(edges))
        self.lib.orc_remove_or_add_edge(_p(buf), C.byref(n), C.c_uint32(e[0]), C.c_uint32(e[1]))
        return [tuple(int(x) for x in r) for r in buf[:n.value]]

    def epa3_process(self, simplex, l, lt, r, rt, verts=None, settings=None):
        settings = settings or Settings.default()
        s = np.asarray(simplex, np.float32).reshape(-1, 3)
        out = np.zeros(1, CONTACT_DT)
        st = self.lib.orc_epa3_process(_p(s), C.c_uint32(len(s)), _p(l), _p(lt), _p(r), _p(rt),
                                       _p(verts), C.byref(settings), _p(out))
        return st, out[0]

    def gjk3_intersect(self, l, lt, r, rt, verts=None, settings=None):
        settings = settings or Settings.default()
        s = np.zeros((8, 3), np.float32)
        n = C.c_uint32(0)
        st = self.lib.orc_gjk3_intersect(_p(l), _p(lt), _p(r), _p(rt), _p(verts),
                                         C.byref(settings), _p(s), C.byref(n))
        return st, s[:n.value].copy()

    def aabb3_intersects(self, a, b):
        a = np.asarray(a, np.float32).reshape(6)
        b = np.asarray(b, np.float32).reshape(6)
        return bool(self.lib.orc_aabb3_intersects(_p(a), _p(b)))

    def build_half_edges(self, shapes, verts, faces):
        """ConvexPolyhedron::new_with_faces for every kind-8 shape (call before any batch that
        uses them). faces: (F, 3) uint32, vertex indices local to the owning shape."""
        faces = np.ascontiguousarray(faces, np.uint32)
        self._faces = faces  # keep alive
        rc = self.lib.orc_build_half_edges(_p(shapes), C.c_uint32(len(shapes)), _p(verts),
                                           C.c_uint64(0 if verts is None else len(verts)), _p(faces))
        if rc != 0:
            raise ValueError("half-edge construction failed (degenerate face / open ring)")

    # ---- batch ---------------------------------------------------------------------
    def intersection(self, strategy, shapes, verts, l_shape, r_shape, l_t, r_t, settings=None,
                     nthreads=1, want_stats=False):
        settings = settings or Settings.default()
        n = len(l_shape)
        out = np.zeros(n, CONTACT_DT)
        status = np.zeros(n, np.uint8)
        stats = np.zeros((n, 4), np.uint32) if want_stats else None
        self.lib.orc_gjk3_intersection_batch(
            C.c_uint32(strategy), C.byref(settings), _p(shapes), _p(verts), _p(l_shape),
            _p(r_shape), _p(l_t), _p(r_t), C.c_uint64(n), _p(out), _p(status), _p(stats),
            int(nthreads))
        return (out, status, stats) if want_stats else (out, status)

    def intersect_batch(self, shapes, verts, l_shape, r_shape, l_t, r_t, settings=None, nthreads=1):
        """GJK::intersect per pair: ((n, 4) SupportPoints, status)"""
        settings = settings or Settings.default()
        n = len(l_shape)
        sx = np.zeros((n, 4), SUPPORT_POINT_DT)
        status = np.zeros(n, np.uint8)
        self.lib.orc_gjk3_intersect_batch(C.byref(settings), _p(shapes), _p(verts), _p(l_shape),
                                          _p(r_shape), _p(l_t), _p(r_t), C.c_uint64(n), _p(sx),
                                          _p(status), int(nthreads))
        return sx, status

    def manifold_batch(self, simplex, shapes, verts, l_shape, r_shape, l_t, r_t, simplex_len=None,
                       settings=None, nthreads=1):
        """GJK::get_contact_manifold per pair: (contacts, status)"""
        settings = settings or Settings.default()
        n = len(l_shape)
        out = np.zeros(n, CONTACT_DT)
        status = np.zeros(n, np.uint8)
        self.lib.orc_gjk3_manifold_batch(C.byref(settings), _p(simplex), _p(simplex_len), _p(shapes),
                                         _p(verts), _p(l_shape), _p(r_shape), _p(l_t), _p(r_t),
                                         C.c_uint64(n), _p(out), _p(status), int(nthreads))
        return out, status

    def distance(self, shapes, verts, l_shape, r_shape, l_t, r_t, settings=None, nthreads=1,
                 want_iters=False):
        settings = settings or Settings.default()
        n = len(l_shape)
        out = np.zeros(n, np.float32)
        status = np.zeros(n, np.uint8)
        iters = np.zeros(n, np.uint32) if want_iters else None
        self.lib.orc_gjk3_distance_batch(
            C.byref(settings), _p(shapes), _p(verts), _p(l_shape), _p(r_shape), _p(l_t), _p(r_t),
            C.c_uint64(n), _p(out), _p(status), _p(iters), int(nthreads))
        return (out, status, iters) if want_iters else (out, status)

    def time_of_impact(self, shapes, verts, l_shape, r_shape, l_t0, l_t1, r_t0, r_t1,
                       settings=None, iter_cap=1024, nthreads=1, want_iters=False):
        settings = settings or Settings.default()
        n = len(l_shape)
        out = np.zeros(n, CONTACT_DT)
        status = np.zeros(n, np.uint8)
        iters = np.zeros(n, np.uint32) if want_iters else None
        self.lib.orc_gjk3_toi_batch(
            C.byref(settings), _p(shapes), _p(verts), _p(l_shape), _p(r_shape), _p(l_t0),
            _p(l_t1), _p(r_t0), _p(r_t1), C.c_uint64(n), C.c_uint32(iter_cap), _p(out),
            _p(status), _p(iters), int(nthreads))
        return (out, status, iters) if want_iters else (out, status)

    def complex(self, op, strategy, shapes, verts, comp_off, comp_shape, comp_local, l_comp, r_comp,
                l_t0, r_t0, l_t1=None, r_t1=None, settings=None, iter_cap=1024, nthreads=1):
        """op 0: intersection_complex, 1: distance_complex, 2: intersection_complex_time_of_impact
        (gjk/mod.rs:412-588). Returns (contacts | distances, status)."""
        settings = settings or Settings.default()
        n = len(l_comp)
        out_c = np.zeros(n, CONTACT_DT)
        out_d = np.zeros(n, np.float32)
        status = np.zeros(n, np.uint8)
        self.lib.orc_gjk3_complex_batch(
            C.c_uint32(op), C.c_uint32(strategy), C.byref(settings), _p(shapes), _p(verts),
            _p(comp_off), _p(comp_shape), _p(comp_local), _p(l_comp), _p(r_comp), _p(l_t0),
            _p(l_t1 if l_t1 is not None else l_t0), _p(r_t0), _p(r_t1 if r_t1 is not None else r_t0),
            C.c_uint64(n), C.c_uint32(iter_cap), _p(out_c), _p(out_d), _p(status), int(nthreads))
        return (out_d if op == 1 else out_c), status

    def world_aabbs(self, shapes, verts, shape, t, nthreads=1):
        n = len(shape)
        out = np.zeros(n, AABB_DT)
        self.lib.orc_world_aabbs(_p(shapes), _p(verts), _p(shape), _p(t), C.c_uint64(n), _p(out),
                                 int(nthreads))
        return out

    def sap3_mt(self, aabbs, axis=0, cap=None, nthreads=None):
        """sap3 with the sweep split over host threads (same pairs, same order)."""
        n = len(aabbs)
        nthreads = nthreads or self.hw_threads()
        perm = np.zeros(max(n, 1), np.uint32)
        cap = int(cap) if cap is not None else 16 * n + 1024
        pairs = np.zeros((max(cap, 1), 2), np.uint32)
        ax = C.c_uint32(axis)
        self.lib.orc_sap3_find_collider_pairs_mt.restype = C.c_uint64
        cnt = self.lib.orc_sap3_find_collider_pairs_mt(_p(aabbs), C.c_uint64(n), C.byref(ax), _p(perm),
                                                       _p(pairs), C.c_uint64(cap), int(nthreads))
        if cnt > cap:  # denser than the first guess: once more with the exact room
            return self.sap3_mt(aabbs, axis, int(cnt), nthreads)
        return perm[:n].copy(), pairs[:int(cnt)].copy(), int(cnt), ax.value

    def sap3(self, aabbs, axis=0, cap=None):
        n = len(aabbs)
        perm = np.zeros(max(n, 1), np.uint32)
        ax = C.c_uint32(axis)
        if cap is None:
            cnt = self.lib.orc_sap3_find_collider_pairs(_p(aabbs), C.c_uint64(n), C.byref(ax),
                                                        _p(perm), None, C.c_uint64(0))
            cap = int(cnt)
            ax = C.c_uint32(axis)
        pairs = np.zeros((max(cap, 1), 2), np.uint32)
        cnt = self.lib.orc_sap3_find_collider_pairs(_p(aabbs), C.c_uint64(n), C.byref(ax),
                                                    _p(perm), _p(pairs), C.c_uint64(cap))
        return perm[:n].copy(), pairs[:min(int(cnt), cap)].copy(), int(cnt), ax.value

    def brute_force_count(self, aabbs):
        return int(self.lib.orc_brute_force_count(_p(aabbs), C.c_uint64(len(aabbs))))

    def brute_force_pairs(self, aabbs):
        """BruteForce::find_collider_pairs (brute_force.rs:20-39)."""
        n = int(self.lib.orc_brute_force_pairs(_p(aabbs), C.c_uint64(len(aabbs)), None, C.c_uint64(0)))
        pairs = np.zeros((max(n, 1), 2), np.uint32)
        self.lib.orc_brute_force_pairs(_p(aabbs), C.c_uint64(len(aabbs)), _p(pairs), C.c_uint64(n))
        return pairs[:n]

    def ops_reset(self):
        return int(self.lib.orc_ops_reset())


# ---- 2-D narrow phase (oracle/collision_oracle2d.cpp) ------------------------------------------------
SHAPE2_DT = np.dtype([("kind", "<u4"), ("p", "<f4", (4,)), ("vtx_off", "<u4"), ("vtx_cnt", "<u4"),
                      ("reserved", "<u4")])
XFORM2_DT = np.dtype([("scale", "<f4"), ("m", "<f4", (4,)), ("d", "<f4", (2,)), ("reserved", "<f4")])
CONTACT2_DT = np.dtype([("strategy", "<u4"), ("normal", "<f4", (2,)), ("depth", "<f4"),
                        ("point", "<f4", (2,)), ("toi", "<f4")])


class Oracle2D:
    """GJK2 / EPA2 restatement; same array contract as the cb2_gjk2_* entry points."""

    def __init__(self):
        build()
        self.lib = C.CDLL(os.path.join(_BUILD, "liboracle2d.so"))
        for f in ("orc2_closest_edge", "orc2_epa2_process", "orc2_gjk2_intersect"):
            getattr(self.lib, f).restype = C.c_uint8

    # unit hooks
    def support_point(self, shape, verts, xform, direction):
        out = np.zeros(2, np.float32)
        st = C.c_uint8(0)
        d = np.asarray(direction, np.float32)
        self.lib.orc2_support_point(_p(shape), _p(verts), _p(xform), _p(d), _p(out), C.byref(st))
        return out, st.value

    def reduce_to_closest_feature(self, pts, d=(0, 0)):
        buf = np.zeros((8, 2), np.float32)
        pts = np.asarray(pts, np.float32).reshape(-1, 2)
        buf[:len(pts)] = pts
        n = C.c_uint32(len(pts))
        dd = np.asarray(d, np.float32).copy()
        hit = self.lib.orc2_reduce_to_closest_feature(_p(buf), C.byref(n), _p(dd))
        return bool(hit), buf[:n.value].copy(), dd

    def closest_point_to_origin(self, pts):
        buf = np.zeros((8, 2), np.float32)
        pts = np.asarray(pts, np.float32).reshape(-1, 2)
        buf[:len(pts)] = pts
        n = C.c_uint32(len(pts))
        out = np.zeros(2, np.float32)
        self.lib.orc2_closest_point_to_origin(_p(buf), C.byref(n), _p(out))
        return out, buf[:n.value].copy()

    def closest_point_on_edge(self, start, end, point):
        out = np.zeros(2, np.float32)
        self.lib.orc2_closest_point_on_edge(_p(np.asarray(start, np.float32)), _p(np.asarray(end, np.float32)),
                                            _p(np.asarray(point, np.float32)), _p(out))
        return out

    def closest_edge(self, pts):
        pts = np.ascontiguousarray(np.asarray(pts, np.float32).reshape(-1, 2))
        out = np.zeros(3, np.float32)
        idx = C.c_uint32(0)
        st = self.lib.orc2_closest_edge(_p(pts), C.c_uint32(len(pts)), _p(out), C.byref(idx))
        return st, out[:2].copy(), float(out[2]), idx.value

    def epa2_process(self, simplex, l, lt, r, rt, verts=None, settings=None):
        settings = settings or Settings.default()
        pts = np.ascontiguousarray(np.asarray(simplex, np.float32).reshape(-1, 2))
        out = np.zeros(1, CONTACT2_DT)
        st = self.lib.orc2_epa2_process(_p(pts), C.c_uint32(len(pts)), _p(l), _p(lt), _p(r), _p(rt),
                                        _p(verts), C.byref(settings), _p(out))
        return st, out[0]

    def gjk2_intersect(self, l, lt, r, rt, verts=None, settings=None):
        settings = settings or Settings.default()
        buf = np.zeros((8, 2), np.float32)
        n = C.c_uint32(0)
        st = self.lib.orc2_gjk2_intersect(_p(l), _p(lt), _p(r), _p(rt), _p(verts), C.byref(settings),
                                          _p(buf), C.byref(n))
        return st, buf[:n.value].copy()

    # batches
    def intersection(self, strategy, shapes, verts, l_shape, r_shape, l_t, r_t, settings=None,
                     nthreads=1, iters=False):
        settings = settings or Settings.default()
        n = len(l_shape)
        out = np.zeros(n, CONTACT2_DT)
        status = np.zeros(n, np.uint8)
        it = np.zeros((n, 2), np.uint32) if iters else None
        self.lib.orc2_gjk2_intersection_batch(C.c_uint32(strategy), C.byref(settings), _p(shapes), _p(verts),
                                              _p(l_shape), _p(r_shape), _p(l_t), _p(r_t), C.c_uint64(n),
                                              _p(out), _p(status), _p(it), C.c_int(nthreads))
        return (out, status, it) if iters else (out, status)

    def distance(self, shapes, verts, l_shape, r_shape, l_t, r_t, settings=None, nthreads=1, iters=False):
        settings = settings or Settings.default()
        n = len(l_shape)
        out = np.zeros(n, np.float32)
        status = np.zeros(n, np.uint8)
        it = np.zeros(n, np.uint32) if iters else None
        self.lib.orc2_gjk2_distance_batch(C.byref(settings), _p(shapes), _p(verts), _p(l_shape), _p(r_shape),
                                          _p(l_t), _p(r_t), C.c_uint64(n), _p(out), _p(status), _p(it),
                                          C.c_int(nthreads))
        return (out, status, it) if iters else (out, status)

    def time_of_impact(self, shapes, verts, l_shape, r_shape, l_t0, l_t1, r_t0, r_t1, settings=None,
                       iter_cap=0, nthreads=1, iters=False):
        settings = settings or Settings.default()
        n = len(l_shape)
        out = np.zeros(n, CONTACT2_DT)
        status = np.zeros(n, np.uint8)
        it = np.zeros(n, np.uint32) if iters else None
        self.lib.orc2_gjk2_toi_batch(C.byref(settings), _p(shapes), _p(verts), _p(l_shape), _p(r_shape),
                                     _p(l_t0), _p(l_t1), _p(r_t0), _p(r_t1), C.c_uint64(n),
                                     C.c_uint32(iter_cap), _p(out), _p(status), _p(it), C.c_int(nthreads))
        return (out, status, it) if iters else (out, status)

    def complex(self, op, strategy, shapes, verts, comp_off, comp_shape, comp_local, l_comp, r_comp, l_t0, r_t0,
                l_t1=None, r_t1=None, settings=None, iter_cap=0, nthreads=1):
        """op 0 intersection_complex, 1 distance_complex, 2 intersection_complex_time_of_impact."""
        settings = settings or Settings.default()
        n = len(l_comp)
        out_c = np.zeros(n, CONTACT2_DT)
        out_d = np.zeros(n, np.float32)
        status = np.zeros(n, np.uint8)
        self.lib.orc2_gjk2_complex_batch(C.c_uint32(op), C.c_uint32(strategy), C.byref(settings), _p(shapes),
                                         _p(verts), _p(comp_off), _p(comp_shape), _p(comp_local), _p(l_comp),
                                         _p(r_comp), _p(l_t0), _p(l_t1), _p(r_t0), _p(r_t1), C.c_uint64(n),
                                         C.c_uint32(iter_cap), _p(out_c), _p(out_d), _p(status), C.c_int(nthreads))
        return (out_d if op == 1 else out_c), status

    def world_aabbs(self, shapes, verts, shape, t):
        out = np.zeros((len(shape), 4), np.float32)  # min.xy, max.xy
        self.lib.orc2_world_aabbs(_p(shapes), _p(verts), _p(shape), _p(t), C.c_uint64(len(shape)), _p(out))
        return out
