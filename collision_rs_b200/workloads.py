"""Synthetic inputs for the five BASELINE.json configs.

Pure numpy, deterministic per seed; consumed identically by the oracle and by the
CUDA path.  Rotation = normalised 4-vector of N(0,1); scale = 1; everything f32.
"""
import numpy as np

from .abi import AABB_DT, CUBOID, POLYHEDRON, SHAPE_DT, SPHERE, XFORM_DT


def random_xforms(rng, n, spread, scale=1.0):
    t = np.zeros(n, XFORM_DT)
    q = rng.standard_normal((n, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    t["scale"] = scale
    t["q"] = q.astype(np.float32)
    t["d"] = rng.uniform(-spread, spread, (n, 3)).astype(np.float32)
    return t


def cuboid_shapes(rng, n, lo=0.5, hi=2.0):
    s = np.zeros(n, SHAPE_DT)
    s["kind"] = CUBOID
    s["p"] = rng.uniform(lo, hi, (n, 3)).astype(np.float32)
    return s


def mixed_shapes(rng, n):
    """Sphere(r U[0.25,1]) w.p. 1/2 else Cuboid(U[0.5,2]^3)."""
    s = cuboid_shapes(rng, n)
    is_sphere = rng.random(n) < 0.5
    r = rng.uniform(0.25, 1.0, n).astype(np.float32)
    s["kind"][is_sphere] = SPHERE
    s["p"][is_sphere, 0] = r[is_sphere]
    s["p"][is_sphere, 1:] = 0
    return s


def cfg1_cuboid_pairs(n=10_000, seed=12345, spread=0.75):
    """cfg 1: GJK3+EPA3 FullResolution on random Cuboid-Cuboid pairs (~91 % hits)."""
    rng = np.random.default_rng(seed)
    shapes = cuboid_shapes(rng, 2 * n)
    l_shape = np.arange(0, n, dtype=np.uint32)
    r_shape = np.arange(n, 2 * n, dtype=np.uint32)
    return dict(shapes=shapes, verts=None, l_shape=l_shape, r_shape=r_shape,
                l_t=random_xforms(rng, n, spread), r_t=random_xforms(rng, n, spread))


def cfg2_mixed_pairs(n=1_000_000, seed=4242, spread=1.5):
    """cfg 2: CollisionOnly + distance on mixed Sphere/Cuboid pairs (~25 % hits)."""
    rng = np.random.default_rng(seed)
    shapes = mixed_shapes(rng, 2 * n)
    l_shape = np.arange(0, n, dtype=np.uint32)
    r_shape = np.arange(n, 2 * n, dtype=np.uint32)
    return dict(shapes=shapes, verts=None, l_shape=l_shape, r_shape=r_shape,
                l_t=random_xforms(rng, n, spread), r_t=random_xforms(rng, n, spread))


def polyhedron_table(n_shapes=4096, seed=777, sizes=(32, 64, 128, 256)):
    """Shape table of convex polyhedra: V in sizes, vertices = unit-sphere directions x
    per-axis radii U[0.5,1] (all points extreme => convex), VertexOnly mode."""
    rng = np.random.default_rng(seed)
    cnt = rng.choice(np.asarray(sizes, np.uint32), n_shapes).astype(np.uint32)
    off = np.zeros(n_shapes, np.uint32)
    off[1:] = np.cumsum(cnt)[:-1]
    total = int(cnt.sum())
    v = rng.standard_normal((total, 3))
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    radii = rng.uniform(0.5, 1.0, (n_shapes, 3))
    v *= np.repeat(radii, cnt, axis=0)
    shapes = np.zeros(n_shapes, SHAPE_DT)
    shapes["kind"] = POLYHEDRON
    shapes["vtx_off"] = off
    shapes["vtx_cnt"] = cnt
    return shapes, np.ascontiguousarray(v.astype(np.float32))


def cfg3_polyhedron_pairs(n=4_000_000, seed=777, spread=0.75, n_shapes=4096,
                          sizes=(32, 64, 128, 256), shard=0):
    """cfg 3: GJK3+EPA3 on ConvexPolyhedron pairs (32-256 verts), ~80-90 % hits.
    `shard` perturbs the pair stream (one shard per GPU) but not the shape table."""
    shapes, verts = polyhedron_table(n_shapes, seed, sizes)
    rng = np.random.default_rng([seed, 1, shard])
    l_shape = rng.integers(0, n_shapes, n).astype(np.uint32)
    r_shape = rng.integers(0, n_shapes, n).astype(np.uint32)
    return dict(shapes=shapes, verts=verts, l_shape=l_shape, r_shape=r_shape,
                l_t=random_xforms(rng, n, spread), r_t=random_xforms(rng, n, spread))


def cfg4_aabbs(n=1_000_000, seed=2024, extent=200.0):
    """cfg 4: centres U[0,extent)^3, half-extents U[0.5,1.5]^3; bodies are Cuboids with those
    half-extents at identity rotation (so AABB == shape bound). extent scales as cbrt(n) to
    keep density (and so ~4 pairs per box) when n is reduced for tests."""
    rng = np.random.default_rng(seed)
    c = rng.uniform(0.0, extent, (n, 3)).astype(np.float32)
    h = rng.uniform(0.5, 1.5, (n, 3)).astype(np.float32)
    a = np.zeros(n, AABB_DT)
    a["min"] = c - h
    a["max"] = c + h
    shapes = np.zeros(n, SHAPE_DT)
    shapes["kind"] = CUBOID
    shapes["p"] = (a["max"] - a["min"])
    t = np.zeros(n, XFORM_DT)
    t["scale"] = 1.0
    t["q"][:, 0] = 1.0
    t["d"] = c
    return dict(aabbs=a, shapes=shapes, body_t=t)


def cfg5_toi_pairs(n=1_000_000, seed=99, shard=0):
    """cfg 5: continuous sweep on Cuboid pairs; start disp U[-3,3]^3, aimed end poses
    (~75 % hits); rotation unchanged over the step."""
    rng = np.random.default_rng([seed, shard])
    shapes = cuboid_shapes(rng, 2 * n)
    l_shape = np.arange(0, n, dtype=np.uint32)
    r_shape = np.arange(n, 2 * n, dtype=np.uint32)
    l0 = random_xforms(rng, n, 3.0)
    r0 = random_xforms(rng, n, 3.0)
    l1 = l0.copy()
    r1 = r0.copy()
    aim = (r0["d"] - l0["d"]) * rng.uniform(0.5, 1.5, (n, 1)).astype(np.float32)
    l1["d"] = l0["d"] + aim + rng.uniform(-1, 1, (n, 3)).astype(np.float32)
    r1["d"] = r0["d"] + rng.uniform(-1, 1, (n, 3)).astype(np.float32)
    return dict(shapes=shapes, verts=None, l_shape=l_shape, r_shape=r_shape,
                l_t0=l0, l_t1=l1, r_t0=r0, r_t1=r1)


def all_primitive3_shapes(rng, n):
    """Every Primitive3 variant but the polyhedron: Sphere, Cuboid, Particle, Quad, Cylinder,
    Capsule, Cube in equal shares (kinds 0,1,3..7 of include/collision_b200.h)."""
    s = np.zeros(n, SHAPE_DT)
    kinds = np.array([0, 1, 3, 4, 5, 6, 7], np.uint32)
    s["kind"] = kinds[rng.integers(0, len(kinds), n)]
    a = rng.uniform(0.4, 1.6, (n, 3)).astype(np.float32)
    s["p"] = a
    return s


# ---- 2-D narrow phase (GJK2 / EPA2; SURVEY §8 row N4, not a BASELINE config) ------------------------
def random_xforms2(rng, n, spread, scale=1.0):
    from .abi import XFORM2_DT
    t = np.zeros(n, XFORM2_DT)
    a = rng.uniform(-np.pi, np.pi, n).astype(np.float32)
    s, c = np.sin(a, dtype=np.float32), np.cos(a, dtype=np.float32)
    t["scale"] = scale
    t["m"] = np.stack([c, s, -s, c], axis=1)  # Rotation2::from_angle: columns (cos, sin), (-sin, cos)
    t["d"] = rng.uniform(-spread, spread, (n, 2)).astype(np.float32)
    return t


def shapes2_table(n_shapes=4096, seed=31, kinds=(0, 1, 2, 3, 4, 5), poly_sizes=(3, 5, 9, 10, 17, 64)):
    """A table of Primitive2 values of every kind; polygons are convex, counter-clockwise, with
    vertex counts on both sides of the reference's 10-vertex switch (polygon.rs:37)."""
    from .abi import SHAPE2_DT
    rng = np.random.default_rng(seed)
    shapes = np.zeros(n_shapes, SHAPE2_DT)
    kind = rng.choice(np.asarray(kinds, np.uint32), n_shapes)
    shapes["kind"] = kind
    shapes["p"] = 0
    verts = []
    off = 0
    for i in range(n_shapes):
        k = kind[i]
        if k == 1:  # Line2(origin, dest)
            shapes[i]["p"] = rng.uniform(-1, 1, 4).astype(np.float32)
        elif k == 2:  # Circle(radius)
            shapes[i]["p"][0] = np.float32(rng.uniform(0.25, 1.0))
        elif k == 3:  # Rectangle(dim)
            shapes[i]["p"][:2] = rng.uniform(0.5, 2.0, 2).astype(np.float32)
        elif k == 4:  # Square(dim)
            shapes[i]["p"][0] = np.float32(rng.uniform(0.5, 2.0))
        elif k == 5:
            v = int(rng.choice(poly_sizes))
            ang = np.sort(rng.uniform(0, 2 * np.pi, v))
            rad = rng.uniform(0.5, 1.0, 2)
            pts = np.stack([np.cos(ang) * rad[0], np.sin(ang) * rad[1]], axis=1).astype(np.float32)
            shapes[i]["vtx_off"] = off
            shapes[i]["vtx_cnt"] = v
            verts.append(pts)
            off += v
    return shapes, (np.concatenate(verts) if verts else None)


def cfg2d_mixed_pairs(n=1_000_000, seed=32, spread=1.0, n_shapes=4096, kinds=(0, 1, 2, 3, 4, 5),
                      scale=1.0):
    """Random pairs over a mixed Primitive2 table, Decomposed<Vector2, Basis2> transforms."""
    rng = np.random.default_rng(seed)
    shapes, verts = shapes2_table(n_shapes, seed + 1, kinds)
    return dict(shapes=shapes, verts=verts,
                l_shape=rng.integers(0, n_shapes, n, dtype=np.uint32),
                r_shape=rng.integers(0, n_shapes, n, dtype=np.uint32),
                l_t=random_xforms2(rng, n, spread, scale), r_t=random_xforms2(rng, n, spread, scale))


def cfg2d_toi_pairs(n=1_000_000, seed=33, n_shapes=4096, kinds=(0, 1, 2, 3, 4, 5)):
    """Start poses 3 apart on average, the left body moves towards the right one (cfg 5's recipe)."""
    rng = np.random.default_rng(seed)
    w = cfg2d_mixed_pairs(n, seed, 3.0, n_shapes, kinds)
    l0, r0 = w.pop("l_t"), w.pop("r_t")
    l1, r1 = l0.copy(), r0.copy()
    gap = r0["d"] - l0["d"]
    l1["d"] = (l0["d"] + gap * rng.uniform(0.5, 1.5, (n, 1)) + rng.uniform(-1, 1, (n, 2))).astype(np.float32)
    r1["d"] = (r0["d"] + rng.uniform(-1, 1, (n, 2))).astype(np.float32)
    w.update(l_t0=l0, l_t1=l1, r_t0=r0, r_t1=r1)
    return w
