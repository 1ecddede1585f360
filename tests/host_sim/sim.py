"""ctypes binding of tests/host_sim/epa_sim.cpp (TEST INFRASTRUCTURE: the CUDA kernels' per-lane
source compiled for the CPU through csrc/cb2_hostsim.h)."""
import ctypes as C
import os
import subprocess

import numpy as np

from oracle.binding import CONTACT_DT, Settings, _p

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "epa_sim.cpp")
_OUT = os.path.join(_HERE, "_build_epa_sim.so")
_CSRC = os.path.join(_HERE, "..", "..", "collision_rs_b200", "csrc")


def build():
    deps = [_SRC] + [os.path.join(_CSRC, f) for f in os.listdir(_CSRC) if f.endswith((".cuh", ".h"))]
    if os.path.exists(_OUT) and all(os.path.getmtime(_OUT) >= os.path.getmtime(d) for d in deps):
        return _OUT
    cuda_inc = os.environ.get("CUDA_HOME", "/usr/local/cuda") + "/include"
    subprocess.check_call([
        "g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math",
        "-I" + cuda_inc, "-Wno-unknown-pragmas", "-o", _OUT, _SRC])
    return _OUT


class Sim:
    def __init__(self):
        self.lib = C.CDLL(build())

    def force_literal_horizon(self, on):
        """team kernel: take the literal remove_or_add_edge replay (the non-manifold path) always"""
        self.lib.sim_force_literal_horizon(C.c_int(int(on)))

    def intersection_full(self, shapes, verts, l_shape, r_shape, l_t, r_t, settings=None, caps=0,
                          warps=3, dump_cap0=None, prefetch=True, team_w=0):
        settings = settings or Settings.default()
        n = len(l_shape)
        out = np.zeros(n, CONTACT_DT)
        status = np.zeros(n, np.uint8)
        tiers = np.zeros(3, np.uint32)
        nv = 0 if verts is None else len(verts)
        rc = self.lib.sim_gjk3_intersection_full(
            _p(shapes), C.c_uint32(len(shapes)), _p(verts), C.c_uint64(nv), _p(l_shape), _p(r_shape),
            _p(l_t), _p(r_t), C.c_uint64(n), C.byref(settings), C.c_int(caps), C.c_int(warps),
            C.c_uint32(n if dump_cap0 is None else dump_cap0), C.c_int(int(prefetch)), C.c_int(team_w), _p(out), _p(status), _p(tiers))
        assert rc == 0, rc
        return out, status, tiers
