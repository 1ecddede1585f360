import sys, os, time
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np
from collision_rs_b200 import workloads as W
from collision_rs_b200.host import Context
from test_gpu_parity import _he_table
op, n = sys.argv[1], int(sys.argv[2])
shapes, verts, faces = _he_table(24, seed=21)
prim = W.mixed_shapes(np.random.default_rng(2), 8)
shapes = np.concatenate([shapes, prim])
ctx = Context(0)
ctx.upload_shapes(shapes, verts, faces)
rng = np.random.default_rng(22)
l = rng.integers(0, len(shapes), n).astype(np.uint32)
r = rng.integers(0, len(shapes), n).astype(np.uint32)
lt, rt = W.random_xforms(rng, n, 0.7), W.random_xforms(rng, n, 0.7)
t = time.time()
if op == "gjk":
    out, st = ctx.intersection(1, l, r, lt, rt)
elif op == "full":
    out, st = ctx.intersection(0, l, r, lt, rt)
elif op == "dist":
    out, st = ctx.distance(l, r, lt, rt)
else:
    out = ctx.world_aabbs(l, lt); st = np.zeros(1, np.uint8)
print(op, n, "ok", round(time.time() - t, 3), np.bincount(st), flush=True)
