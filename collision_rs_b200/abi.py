"""ctypes binding of the C ABI in include/collision_b200.h.

The shared library (collision_rs_b200/csrc -> libcollision_b200.so) is the
product; this module only marshals numpy / raw pointers into it.  There is no
CPU fallback: if the library is missing, `load()` raises.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# CB2_LIB_PATH: an alternative build of the same library (A/B experiments, tools/ab_variants.sh)
LIB_PATH = os.environ.get("CB2_LIB_PATH") or os.path.join(_HERE, "libcollision_b200.so")

# ---- POD mirrors (numpy structured dtypes; layouts asserted against sizeof in tests) ----
SHAPE_DT = np.dtype([("kind", "<u4"), ("p", "<f4", (3,)), ("vtx_off", "<u4"), ("vtx_cnt", "<u4")])
XFORM_DT = np.dtype([("scale", "<f4"), ("q", "<f4", (4,)), ("d", "<f4", (3,))])  # q = (s,x,y,z)
CONTACT_DT = np.dtype([("strategy", "<u4"), ("normal", "<f4", (3,)), ("depth", "<f4"),
                       ("point", "<f4", (3,)), ("toi", "<f4")])
AABB_DT = np.dtype([("min", "<f4", (3,)), ("max", "<f4", (3,))])
# SupportPoint<Point3<f32>> (minkowski/mod.rs:16-23)
SUPPORT_POINT_DT = np.dtype([("v", "<f4", (3,)), ("sup_a", "<f4", (3,)), ("sup_b", "<f4", (3,))])
assert SUPPORT_POINT_DT.itemsize == 36
assert SHAPE_DT.itemsize == 24 and XFORM_DT.itemsize == 32
assert CONTACT_DT.itemsize == 36 and AABB_DT.itemsize == 24
# 2-D narrow phase (cb2_shape2 / cb2_xform2 / cb2_contact2); m = Basis2 columns (c0x, c0y, c1x, c1y)
SHAPE2_DT = np.dtype([("kind", "<u4"), ("p", "<f4", (4,)), ("vtx_off", "<u4"), ("vtx_cnt", "<u4"),
                      ("reserved", "<u4")])
XFORM2_DT = np.dtype([("scale", "<f4"), ("m", "<f4", (4,)), ("d", "<f4", (2,)), ("reserved", "<f4")])
CONTACT2_DT = np.dtype([("strategy", "<u4"), ("normal", "<f4", (2,)), ("depth", "<f4"),
                        ("point", "<f4", (2,)), ("toi", "<f4")])
assert SHAPE2_DT.itemsize == 32 and XFORM2_DT.itemsize == 32 and CONTACT2_DT.itemsize == 28
PARTICLE2, LINE2, CIRCLE, RECTANGLE, SQUARE, POLYGON = range(6)
PANIC_EPA2_EDGE, PANIC_INV_ROT = 9, 10

SPHERE, CUBOID, POLYHEDRON, PARTICLE, QUAD, CYLINDER, CAPSULE, CUBE, POLYHEDRON_HE = range(9)
FULL_RESOLUTION, COLLISION_ONLY = 0, 1
(NONE, SOME, PANIC_UNWRAP_PLANE, PANIC_ASSERT_RANGE, PANIC_INV_SCALE, NONCONVERGED,
 PANIC_EPA_EMPTY) = range(7)
BAD_INDEX = 11
OK, ERR_NO_DEVICE, ERR_CUDA, ERR_ARG, ERR_NOSPC, ERR_UNSUPPORTED, ERR_NAN = 0, -1, -2, -3, -4, -5, -6


class Settings(C.Structure):
    """cb2_settings == GJK::new_with_settings arguments (gjk/mod.rs:71-84)."""
    _fields_ = [("distance_tolerance", C.c_float), ("continuous_tolerance", C.c_float),
                ("epa_tolerance", C.c_float), ("max_iterations", C.c_uint32)]

    @staticmethod
    def default():
        # gjk/mod.rs:22-24, epa/mod.rs:15-16
        return Settings(np.float32(1e-6), np.float32(1e-6), np.float32(1e-5), 100)


# every symbol include/collision_b200.h declares
SYMBOLS = [
    "cb2_abi_version", "cb2_default_settings", "cb2_create", "cb2_destroy", "cb2_last_error",
    "cb2_device_sm_count", "cb2_kernel_launches", "cb2_set_stage_timing", "cb2_last_stage_ms",
    "cb2_shapes_upload", "cb2_shapes_upload_faces",
    "cb2_gjk3_intersection", "cb2_gjk3_distance", "cb2_gjk3_time_of_impact",
    "cb2_gjk3_intersection_dev", "cb2_gjk3_distance_dev", "cb2_gjk3_time_of_impact_dev",
    "cb2_gjk3_intersection_bodies", "cb2_gjk3_intersection_bodies_dev",
    "cb2_gjk3_intersect", "cb2_gjk3_get_contact_manifold",
    "cb2_gjk3_intersect_dev", "cb2_gjk3_get_contact_manifold_dev",
    "cb2_create_multi", "cb2_destroy_multi", "cb2_multi_device_count", "cb2_multi_ctx",
    "cb2_multi_last_error", "cb2_multi_shapes_upload", "cb2_multi_gjk3_intersection",
    "cb2_multi_gjk3_distance", "cb2_multi_gjk3_time_of_impact", "cb2_multi_sap3_find_collider_pairs",
    "cb2_composites_upload", "cb2_gjk3_intersection_complex", "cb2_gjk3_distance_complex",
    "cb2_gjk3_time_of_impact_complex",
    "cb2_sap3_find_collider_pairs", "cb2_sap3_find_collider_pairs_dev",
    "cb2_sap3_find_collider_pairs_shard_dev", "cb2_sap3_find_collider_pairs_unordered_dev",
    "cb2_sap2_find_collider_pairs",
    "cb2_brute_force_find_collider_pairs", "cb2_brute_force_find_collider_pairs_dev",
    "cb2_world_aabbs", "cb2_world_aabbs_dev",
    "cb2_shapes2_upload", "cb2_gjk2_intersection", "cb2_gjk2_distance", "cb2_gjk2_time_of_impact",
    "cb2_gjk2_intersection_dev", "cb2_gjk2_distance_dev", "cb2_gjk2_time_of_impact_dev",
    "cb2_composites2_upload", "cb2_gjk2_intersection_complex", "cb2_gjk2_distance_complex",
    "cb2_gjk2_time_of_impact_complex",
    "cb2_world_aabbs2", "cb2_world_aabbs2_dev",
    "cb2_sap2_find_collider_pairs_dev", "cb2_gjk2_intersection_bodies_dev",
]

_lib = None


def load():
    """dlopen the CUDA library. Raises (loudly) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; "
            "g.build()'` (nvcc, sm_100a). There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    vp, u32, u64, i32 = C.c_void_p, C.c_uint32, C.c_uint64, C.c_int
    lib.cb2_abi_version.restype = u32
    lib.cb2_default_settings.argtypes = [C.POINTER(Settings)]
    lib.cb2_default_settings.restype = None
    lib.cb2_create.argtypes = [i32, C.POINTER(vp)]
    lib.cb2_destroy.argtypes = [vp]
    lib.cb2_destroy.restype = None
    lib.cb2_last_error.argtypes = [vp]
    lib.cb2_last_error.restype = C.c_char_p
    lib.cb2_device_sm_count.argtypes = [vp]
    lib.cb2_kernel_launches.argtypes = [vp]
    lib.cb2_kernel_launches.restype = u64
    lib.cb2_set_stage_timing.argtypes = [vp, i32]
    lib.cb2_last_stage_ms.argtypes = [vp, C.POINTER(C.c_float * 4)]
    lib.cb2_shapes_upload.argtypes = [vp, vp, u32, vp, u64]
    lib.cb2_shapes_upload_faces.argtypes = [vp, vp, u32, vp, u64, vp, u64]
    sp = C.POINTER(Settings)
    lib.cb2_gjk3_intersection.argtypes = [vp, u32, sp, vp, vp, vp, vp, u64, vp, vp]
    lib.cb2_gjk3_distance.argtypes = [vp, sp, vp, vp, vp, vp, u64, vp, vp]
    lib.cb2_gjk3_time_of_impact.argtypes = [vp, sp, vp, vp, vp, vp, vp, vp, u64, u32, vp, vp]
    lib.cb2_gjk3_intersection_dev.argtypes = [vp, u32, sp, vp, vp, vp, vp, u64, vp, vp, vp]
    lib.cb2_gjk3_distance_dev.argtypes = [vp, sp, vp, vp, vp, vp, u64, vp, vp, vp]
    lib.cb2_gjk3_time_of_impact_dev.argtypes = [vp, sp, vp, vp, vp, vp, vp, vp, u64, u32, vp, vp, vp]
    lib.cb2_gjk3_intersection_bodies.argtypes = [vp, u32, sp, vp, vp, u64, vp, u64, vp, vp]
    lib.cb2_gjk3_intersection_bodies_dev.argtypes = [vp, u32, sp, vp, vp, vp, u64, vp, vp, vp]
    lib.cb2_create_multi.argtypes = [C.POINTER(i32), i32, C.POINTER(vp)]
    lib.cb2_destroy_multi.argtypes = [vp]
    lib.cb2_destroy_multi.restype = None
    lib.cb2_multi_device_count.argtypes = [vp]
    lib.cb2_multi_ctx.argtypes = [vp, i32]
    lib.cb2_multi_ctx.restype = vp
    lib.cb2_multi_last_error.argtypes = [vp]
    lib.cb2_multi_last_error.restype = C.c_char_p
    lib.cb2_multi_shapes_upload.argtypes = [vp, vp, u32, vp, u64]
    lib.cb2_multi_gjk3_intersection.argtypes = [vp, u32, sp, vp, vp, vp, vp, u64, vp, vp]
    lib.cb2_multi_gjk3_distance.argtypes = [vp, sp, vp, vp, vp, vp, u64, vp, vp]
    lib.cb2_multi_gjk3_time_of_impact.argtypes = [vp, sp, vp, vp, vp, vp, vp, vp, u64, u32, vp, vp]
    lib.cb2_multi_sap3_find_collider_pairs.argtypes = [vp, vp, u64, C.POINTER(u32), vp, vp, u64,
                                                       C.POINTER(u64)]
    lib.cb2_gjk3_intersect.argtypes = [vp, sp, vp, vp, vp, vp, u64, vp, vp]
    lib.cb2_gjk3_get_contact_manifold.argtypes = [vp, sp, vp, vp, vp, vp, vp, vp, u64, vp, vp]
    lib.cb2_gjk3_intersect_dev.argtypes = [vp, sp, vp, vp, vp, vp, u64, vp, vp, vp]
    lib.cb2_gjk3_get_contact_manifold_dev.argtypes = [vp, sp, vp, vp, vp, vp, vp, vp, u64, vp, vp, vp]
    lib.cb2_composites_upload.argtypes = [vp, vp, vp, vp, u32]
    lib.cb2_gjk3_intersection_complex.argtypes = [vp, u32, sp, vp, vp, vp, vp, u64, vp, vp]
    lib.cb2_gjk3_distance_complex.argtypes = [vp, sp, vp, vp, vp, vp, u64, vp, vp]
    lib.cb2_gjk3_time_of_impact_complex.argtypes = [vp, u32, sp, vp, vp, vp, vp, vp, vp, u64, u32,
                                                    vp, vp]
    lib.cb2_sap3_find_collider_pairs.argtypes = [vp, vp, u64, C.POINTER(u32), vp, vp, u64,
                                                 C.POINTER(u64)]
    lib.cb2_sap3_find_collider_pairs_dev.argtypes = [vp, vp, u64, u32, vp, vp, u64,
                                                     C.POINTER(u64), vp]
    lib.cb2_sap3_find_collider_pairs_shard_dev.argtypes = [vp, vp, u64, u32, u32, u32, vp, vp, u64,
                                                           C.POINTER(u64), vp]
    lib.cb2_sap3_find_collider_pairs_unordered_dev.argtypes = [vp, vp, u64, u32, vp, vp, u64,
                                                               C.POINTER(u64), vp]
    lib.cb2_sap2_find_collider_pairs.argtypes = [vp, vp, u64, C.POINTER(u32), vp, vp, u64,
                                                 C.POINTER(u64)]
    lib.cb2_brute_force_find_collider_pairs.argtypes = [vp, vp, u64, vp, u64, C.POINTER(u64)]
    lib.cb2_brute_force_find_collider_pairs_dev.argtypes = [vp, vp, u64, vp, u64, C.POINTER(u64), vp]
    lib.cb2_world_aabbs.argtypes = [vp, vp, vp, u64, vp]
    lib.cb2_world_aabbs_dev.argtypes = [vp, vp, vp, u64, vp, vp]
    lib.cb2_shapes2_upload.argtypes = [vp, vp, u32, vp, u64]
    lib.cb2_gjk2_intersection.argtypes = [vp, u32, sp, vp, vp, vp, vp, u64, vp, vp]
    lib.cb2_gjk2_distance.argtypes = [vp, sp, vp, vp, vp, vp, u64, vp, vp]
    lib.cb2_gjk2_time_of_impact.argtypes = [vp, sp, vp, vp, vp, vp, vp, vp, u64, u32, vp, vp]
    lib.cb2_gjk2_intersection_dev.argtypes = [vp, u32, sp, vp, vp, vp, vp, u64, vp, vp, vp]
    lib.cb2_gjk2_distance_dev.argtypes = [vp, sp, vp, vp, vp, vp, u64, vp, vp, vp]
    lib.cb2_gjk2_time_of_impact_dev.argtypes = [vp, sp, vp, vp, vp, vp, vp, vp, u64, u32, vp, vp, vp]
    lib.cb2_composites2_upload.argtypes = [vp, vp, vp, vp, u32]
    lib.cb2_gjk2_intersection_complex.argtypes = [vp, u32, sp, vp, vp, vp, vp, u64, vp, vp]
    lib.cb2_gjk2_distance_complex.argtypes = [vp, sp, vp, vp, vp, vp, u64, vp, vp]
    lib.cb2_gjk2_time_of_impact_complex.argtypes = [vp, u32, sp, vp, vp, vp, vp, vp, vp, u64, u32, vp, vp]
    lib.cb2_world_aabbs2.argtypes = [vp, vp, vp, u64, vp]
    lib.cb2_world_aabbs2_dev.argtypes = [vp, vp, vp, u64, vp, vp]
    lib.cb2_sap2_find_collider_pairs_dev.argtypes = [vp, vp, u64, u32, vp, vp, u64, C.POINTER(u64), vp]
    lib.cb2_gjk2_intersection_bodies_dev.argtypes = [vp, u32, sp, vp, vp, vp, u64, vp, vp, vp]
    _lib = lib
    return lib


def ptr(a):
    """Host pointer of a C-contiguous numpy array (or None)."""
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"], "array must be C-contiguous"
    return a.ctypes.data_as(C.c_void_p)
