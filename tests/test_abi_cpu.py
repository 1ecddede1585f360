"""CPU-side checks of the boundary: the C-ABI library loads, exports every symbol the header
declares, its POD layouts match the numpy mirrors, and it fails LOUDLY without a GPU
(no CPU fallback).  No compute calls here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from collision_rs_b200 import abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(abi.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    return abi.load()


def test_header_symbols_are_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "collision_b200.h")).read()
    declared = set(re.findall(r"\b(cb2_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(abi.SYMBOLS), declared ^ set(abi.SYMBOLS)
    for s in declared:
        assert hasattr(lib, s), s


def test_abi_version_and_defaults(lib):
    assert lib.cb2_abi_version() == 1
    s = abi.Settings()
    lib.cb2_default_settings(C.byref(s))
    d = abi.Settings.default()
    # gjk/mod.rs:22-24, epa/mod.rs:15-16
    assert (s.distance_tolerance, s.continuous_tolerance, s.epa_tolerance, s.max_iterations) == \
        (d.distance_tolerance, d.continuous_tolerance, d.epa_tolerance, 100)
    assert np.float32(s.epa_tolerance) == np.float32(1e-5)


def test_pod_layouts():
    assert abi.SHAPE_DT.itemsize == 24 and abi.XFORM_DT.itemsize == 32
    assert abi.CONTACT_DT.itemsize == 36 and abi.AABB_DT.itemsize == 24
    assert C.sizeof(abi.Settings) == 16
    assert abi.CONTACT_DT.fields["depth"][1] == 16 and abi.CONTACT_DT.fields["toi"][1] == 32


def test_2d_pod_layouts_and_oracle_mirrors():
    from oracle import binding as ob
    assert abi.SHAPE2_DT.itemsize == 32 and abi.XFORM2_DT.itemsize == 32 and abi.CONTACT2_DT.itemsize == 28
    assert abi.CONTACT2_DT.fields["depth"][1] == 12 and abi.CONTACT2_DT.fields["toi"][1] == 24
    # the oracle's ctypes mirrors are the same layouts (the two sides never share code)
    for a, b in ((abi.SHAPE_DT, ob.SHAPE_DT), (abi.XFORM_DT, ob.XFORM_DT), (abi.CONTACT_DT, ob.CONTACT_DT),
                 (abi.SHAPE2_DT, ob.SHAPE2_DT), (abi.XFORM2_DT, ob.XFORM2_DT), (abi.CONTACT2_DT, ob.CONTACT2_DT)):
        assert a == b


def test_rust_sys_crate_declares_every_entry_point_and_pod():
    """rust/collision-b200-sys cannot be compiled here (no toolchain); at least keep it in step with
    the header: every function and every struct the header declares is declared there too."""
    hdr = open(os.path.join(ROOT, "include", "collision_b200.h")).read()
    rs = open(os.path.join(ROOT, "rust", "collision-b200-sys", "src", "lib.rs")).read()
    for fn in set(re.findall(r"\b(cb2_[a-z0-9_]+)\s*\(", hdr)):
        assert re.search(r"pub fn %s\s*\(" % fn, rs), fn
    for st in set(re.findall(r"typedef struct (cb2_[a-z0-9_]+) \{", hdr)):
        assert re.search(r"pub struct %s\b" % st, rs), st
    # struct sizes: count the 4-byte fields of each #[repr(C)] mirror
    sizes = {"cb2_shape": 24, "cb2_xform": 32, "cb2_contact": 36, "cb2_settings": 16, "cb2_aabb3": 24,
             "cb2_aabb2": 16, "cb2_shape2": 32, "cb2_xform2": 32, "cb2_contact2": 28,
             "cb2_support_point": 36}
    for name, size in sizes.items():
        body = re.search(r"pub struct %s \{(.*?)\n\}" % name, rs, re.S).group(1)
        n = 0
        for ty in re.findall(r"pub \w+:\s*([^,\n]+),", body):
            m = re.match(r"\[(\w+);\s*(\d+)\]", ty.strip())
            n += int(m.group(2)) if m else 1
        assert 4 * n == size, (name, n)


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    h = C.c_void_p()
    rc = lib.cb2_create(0, C.byref(h))
    assert rc == abi.ERR_NO_DEVICE and not h.value
    assert b"no CPU fallback" in lib.cb2_last_error(None)
    from collision_rs_b200.host import Cb2Error, Context
    with pytest.raises(Cb2Error):
        Context(0)


def test_product_never_touches_the_oracle():
    """The oracle is test infrastructure: nothing under the product package may import, include,
    link or dlopen anything from oracle/."""
    pkg = os.path.join(ROOT, "collision_rs_b200")
    bad = re.compile(r"(import\s+oracle|from\s+oracle|oracle/|oracle\.binding|liboracle|orc_[a-z])")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp")) or f == "Makefile":
                src = open(os.path.join(dp, f), errors="ignore").read()
                assert not bad.search(src), os.path.join(dp, f)
