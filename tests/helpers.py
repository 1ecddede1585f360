"""Fixture builders mirroring the reference's test helpers
(transform_3d: gjk/mod.rs:609-620, epa3d.rs:366-377, sphere.rs:226-232)."""
import ctypes
import ctypes.util

import numpy as np

from oracle.binding import SHAPE_DT, XFORM_DT, AABB_DT

_libm = ctypes.CDLL(ctypes.util.find_library("m"))
_libm.sinf.restype = ctypes.c_float
_libm.sinf.argtypes = [ctypes.c_float]
_libm.cosf.restype = ctypes.c_float
_libm.cosf.argtypes = [ctypes.c_float]


def shape(kind, p=(0, 0, 0), off=0, cnt=0):
    s = np.zeros(1, SHAPE_DT)
    s["kind"] = kind
    s["p"] = np.asarray(p, np.float32)
    s["vtx_off"] = off
    s["vtx_cnt"] = cnt
    return s


def sphere(r):
    return shape(0, (r, 0, 0))


def cuboid(x, y, z):
    return shape(1, (x, y, z))


def polyhedron(off, cnt):
    return shape(2, (0, 0, 0), off, cnt)


def transform_3d(x, y, z, angle_z=0.0, scale=1.0):
    """Decomposed{disp:(x,y,z), rot: Quaternion::from_angle_z(Rad(angle_z)), scale}.
    cgmath: from_angle_z(t) = (cos(t*0.5), 0, 0, sin(t*0.5)) in f32 (libm sinf/cosf like Rust)."""
    h = np.float32(np.float32(angle_z) * np.float32(0.5))
    t = np.zeros(1, XFORM_DT)
    t["scale"] = scale
    t["q"] = (_libm.cosf(h), 0.0, 0.0, _libm.sinf(h))
    t["d"] = (x, y, z)
    return t


def aabb(mn, mx):
    a = np.zeros(1, AABB_DT)
    a["min"] = mn
    a["max"] = mx
    return a


def polyhedron_he(off, cnt, f_off, f_cnt):
    """ConvexPolyhedron::new_with_faces (kind 8): the face range rides in p[0], p[1] as uint32
    bit patterns (include/collision_b200.h)."""
    s = shape(8, (0, 0, 0), off, cnt)
    s["p"].view(np.uint32)[0, 0] = f_off
    s["p"].view(np.uint32)[0, 1] = f_cnt
    return s


# ---- 2-D fixtures (transform: gjk/mod.rs:600-606, epa2d.rs:271-277, circle.rs:187-193) ----------
def shape2(kind, p=(0, 0, 0, 0), off=0, cnt=0):
    from oracle.binding import SHAPE2_DT
    s = np.zeros(1, SHAPE2_DT)
    s["kind"] = kind
    s["p"] = np.asarray(tuple(p) + (0,) * (4 - len(p)), np.float32)
    s["vtx_off"] = off
    s["vtx_cnt"] = cnt
    return s


def particle2():
    return shape2(0)


def line2(ox, oy, dx, dy):
    return shape2(1, (ox, oy, dx, dy))


def circle(r):
    return shape2(2, (r,))


def rectangle(x, y):
    return shape2(3, (x, y))


def square(d):
    return shape2(4, (d,))


def polygon(off, cnt):
    return shape2(5, (), off, cnt)


def transform_2d(x, y, angle=0.0, scale=1.0):
    """Decomposed{disp:(x,y), rot: Rotation2::from_angle(Rad(angle)), scale}.
    cgmath: Matrix2::from_angle(t) = Matrix2::new(cos, sin, -sin, cos) (column major), f32 libm."""
    from oracle.binding import XFORM2_DT
    a = np.float32(angle)
    s, c = _libm.sinf(a), _libm.cosf(a)
    t = np.zeros(1, XFORM2_DT)
    t["scale"] = scale
    t["m"] = (c, s, -s, c)
    t["d"] = (x, y)
    return t
