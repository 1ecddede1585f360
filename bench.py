#!/usr/bin/env python
"""bench.py — headline benchmark of the B200 collision pipeline.

Metric (BASELINE.json): "GJK3+EPA3 contact pairs/s at 1/2/4/8 B200 vs host CPU; SAP AABBs/s".
Workload at every N: BASELINE.json configs[2], "GJK3+EPA3 on 4M ConvexPolyhedron pairs
(32-256 verts each)" — the config the pairs/s metric is quoted on at 1/2/4/8 GPUs.  A step is
one GJK3::intersection(FullResolution) pass over one batch of 4M pairs per GPU (weak scaling:
every rank owns its own 4M-pair shard; the shape table is broadcast once over NCCL; no
data-path collective).  Other configs (cfg1 cuboids, cfg2 mixed, cfg4 SAP, cfg5 TOI) are
reported in the `extra` object of the same JSON line at N=1; at N>1 `extra.multi` holds what only
exists on several GPUs: the device-resident NCCL all-gather of the contact records (`gather_ms`),
the strong-scaling pass (the 4M pairs of ONE batch split over the ranks), cfg 5 (8M time-of-impact
pairs per GPU: 64M over 8) and cfg 4 with the pair list sharded over the GPUs.

  python bench.py --gpus N --steps K --warmup W            (torchrun for N > 1)
  python bench.py --impl reference ...                      (CPU oracle port, all host threads)
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "GJK3+EPA3 FullResolution contact pairs/s"
UNIT = "pairs/s"
WORKLOAD = "cfg3: GJK3+EPA3 FullResolution on ConvexPolyhedron pairs (32-256 verts), 4096-shape table"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--pairs", type=int, default=4_000_000, help="pairs per GPU per step")
    ap.add_argument("--no-extra", action="store_true", help="skip the secondary configs")
    ap.add_argument("--shard", type=int, default=None,
                    help="pair-stream shard to generate (default: the rank); diagnosis only")
    ap.add_argument("--cpu-sample", type=int, default=4_000_000,
                    help="pairs of the workload the CPU oracle port is timed on")
    return ap.parse_args()


def config_of(a):
    """The same `config` object in both arms (the driver compares them): the workload, not the run."""
    return {"workload": WORKLOAD, "pairs_per_gpu_per_step": a.pairs, "seed": 777,
            "l2": "inputs+outputs 109 B/pair: 436 MB per 4M-pair step > 126 MB L2 (no flush needed)"}


# ---------------------------------------------------------------------------------------------
def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "20"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1]))
                mx.append(float(c[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                "sw_power_cap"), c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.f.name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)),
                "reasons": sorted(reasons), "samples": len(sm)}


def algorithmic_flops_per_pair(w, sample=20_000):
    """Algorithmic flops = f32 + - * / sqrt the REFERENCE algorithm executes (oracle op counter),
    measured on a prefix sample of this very workload.  Returns (whole pass, GJK stage only):
    the CollisionOnly strategy stops after GJK::intersect, so the difference is EPA3::process."""
    from oracle.binding import Oracle
    oc = Oracle(count_ops=True)
    m = min(sample, len(w["l_shape"]))
    out = []
    for strategy in (0, 1):
        oc.ops_reset()
        oc.intersection(strategy, w["shapes"], w["verts"], w["l_shape"][:m], w["r_shape"][:m],
                        w["l_t"][:m], w["r_t"][:m], nthreads=1)
        out.append(oc.ops_reset() / m)
    return out[0], out[1]


def counted_flops(op, w, sample=20_000):
    """f32 + - * / sqrt per pair the reference executes for `op` on a prefix of workload `w`
    (oracle op-counter build)."""
    from oracle.binding import Oracle
    oc = Oracle(count_ops=True)
    m = min(sample, len(w["l_shape"]))
    oc.ops_reset()
    if op == "distance":
        oc.distance(w["shapes"], w.get("verts"), w["l_shape"][:m], w["r_shape"][:m], w["l_t"][:m], w["r_t"][:m])
    elif op == "toi":
        oc.time_of_impact(w["shapes"], w.get("verts"), w["l_shape"][:m], w["r_shape"][:m], w["l_t0"][:m],
                          w["l_t1"][:m], w["r_t0"][:m], w["r_t1"][:m])
    else:
        oc.intersection(1 if op == "intersection1" else 0, w["shapes"], w.get("verts"), w["l_shape"][:m],
                        w["r_shape"][:m], w["l_t"][:m], w["r_t"][:m], nthreads=1)
    return oc.ops_reset() / m


def kernel_rooflines(n, stage_ms, flops_pair, flops_gjk_pair, extra, fp32_peak, hbm_peak):
    """One roofline entry per kernel the north star names.  `ms` is measured live in this run (CUDA
    events on the launching stream); `achieved` = ALGORITHMIC flops (what the reference executes,
    oracle op counter) / ms; `executed` = the FP32 instructions the kernel really issues, per unit
    from the committed ncu counters (profiles/r2_kernels.json: smsp__sass_thread_inst_executed_op_
    {fadd,fmul,ffma}_pred_on) scaled to this run's units and time.  Parity forbids FMA, so the
    reachable FP32 ceiling is peak / 2 (frac_of_unfused_peak)."""
    prof = {}
    pj = os.path.join(ROOT, "profiles", "r2_kernels.json")
    if os.path.exists(pj):
        prof = json.load(open(pj))
    pk, units = prof.get("kernels", {}), prof.get("units", {})

    def executed(keys, sec, live_units, ms):
        recs = [r for k in keys for name, r in pk.items() if name == k or (k.endswith(",") and name.startswith(k))]
        if not recs or not units.get(sec) or not ms:
            return None
        per_unit = sum(r["fadd"] + r["fmul"] + 2 * r["ffma"] for r in recs) / units[sec]
        tf = per_unit * live_units / (ms * 1e-3) / 1e12
        inst = sum(r["thread_inst_executed"] for r in recs)
        return {"fp32_flops_per_unit": per_unit, "tflops": tf, "frac": tf / fp32_peak if fp32_peak else None,
                "frac_of_unfused_peak": 2 * tf / fp32_peak if fp32_peak else None,
                "fp32_share_of_thread_instructions": sum(r["fadd"] + r["fmul"] + r["ffma"] for r in recs) / inst if inst else None,
                "threads_per_instruction": [r["threads_per_inst"] for r in recs],
                "ipc": [r["ipc"] for r in recs], "warps_active_pct": [r["warps_active_pct"] for r in recs],
                "dram_bytes_per_unit": sum(r["dram_bytes"] for r in recs) / units[sec],
                "source": "profiles/r2_kernels.json (ncu counters of tools/profile_pass.py)"}

    def entry(kernel, ref, workload, units_live, ms, alg_flops_unit, keys, sec):
        if not ms:
            return None
        ach = alg_flops_unit * units_live / (ms * 1e-3) / 1e12 if alg_flops_unit else None
        return {"kernel": kernel, "replaces": ref, "workload": workload, "units": units_live, "ms": ms,
                "bound": "fp32", "unit": "TFLOP/s", "algorithmic_flops_per_unit": alg_flops_unit,
                "achieved": ach, "peak": fp32_peak, "frac": ach / fp32_peak if ach and fp32_peak else None,
                "frac_of_unfused_peak": 2 * ach / fp32_peak if ach and fp32_peak else None,
                "executed": executed(keys, sec, units_live, ms)}

    out = [
        entry("k_gjk_hits_p", "GJK::intersect, gjk/mod.rs:101-146", "cfg3", n, stage_ms[0], flops_gjk_pair,
              ["cfg3:k_gjk_hits_p<1>"], "cfg3"),
        entry("k_epa tier 0", "EPA3::process, epa3d.rs:27-175", "cfg3", n, stage_ms[1], None,
              ["cfg3:k_epa<4,"], "cfg3"),
        entry("k_epa (all tiers)", "EPA3::process, epa3d.rs:27-175", "cfg3", n, stage_ms[1] + stage_ms[2],
              flops_pair - flops_gjk_pair,
              ["cfg3:k_epa<4,", "cfg3:k_epa<8,72,", "cfg3:k_epa<8,208,"], "cfg3"),
    ]
    x = extra.get("cfg2_mixed_distance")
    if x:
        out.append(entry("k_distance_p", "GJK::distance, gjk/mod.rs:271-321", "cfg2", x["pairs"], x["ms"],
                         x.get("algorithmic_flops_per_pair"), ["cfg2:k_distance_p"], "cfg2"))
    x = extra.get("cfg2_mixed_collision_only")
    if x:
        out.append(entry("k_gjk_hits_p (CollisionOnly)", "GJK::intersect, gjk/mod.rs:101-146", "cfg2", x["pairs"],
                         x["ms"], x.get("algorithmic_flops_per_pair"), ["cfg2:k_gjk_hits_p<0>"], "cfg2"))
    x = extra.get("cfg5_time_of_impact")
    if x:
        out.append(entry("k_time_of_impact_p", "GJK::intersection_time_of_impact, gjk/mod.rs:163-256", "cfg5",
                         x["pairs"], x["ms"], x.get("algorithmic_flops_per_pair"), ["cfg5:k_time_of_impact_p"], "cfg5"))
    x = extra.get("cfg4_sap")
    if x and pk:
        # HBM-bound: algorithmic bytes of the whole call against the live call time, and k_grid_sweep's
        # share of the call from the ncu launch list
        sap = {k: r for k, r in pk.items() if k.startswith("cfg4:") and ("sap" in k or "grid" in k or "Radix" in k or "Scan" in k)}
        tot = sum(r["ms"] * r["launches_in_pass"] / 2 for r in sap.values()) or None
        sw = pk.get("cfg4:k_grid_sweep")
        out.append({"kernel": "SweepAndPrune3 call (sorts + k_grid_sweep)", "replaces": "sweep_prune.rs:76-136",
                    "workload": "cfg4", "units": x["aabbs"], "ms": x["ms"], "bound": "hbm", "unit": "GB/s",
                    "algorithmic_bytes": x["algorithmic_bytes"],
                    "achieved": x["algorithmic_bytes"] / (x["ms"] * 1e-3) / 1e9, "peak": hbm_peak,
                    "frac": x["hbm_frac"],
                    "executed": {"dram_bytes_per_call": sum(r["dram_bytes"] * r["launches_in_pass"] / 2 for r in sap.values()),
                                 "k_grid_sweep_share_of_call": (sw["ms"] / tot) if sw and tot else None,
                                 "k_grid_sweep_ipc": sw["ipc"] if sw else None,
                                 "radix_sort_share_of_call": (sum(r["ms"] * r["launches_in_pass"] / 2 for k, r in sap.items() if "Radix" in k) / tot) if tot else None,
                                 "source": "profiles/r2_kernels.json"}})
    x = extra.get("gjk2_mixed_primitive2_full_resolution")
    if x:
        out.append(entry("k_gjk2_intersection + k_epa2", "GJK2::intersection, gjk/mod.rs:362-391 / epa2d.rs:21-157",
                         "gjk2", x["pairs"], x["ms"], None,
                         ["gjk2:k_gjk2_intersection", "gjk2:k_epa2<12,1>", "gjk2:k_epa2<104,0>"], "gjk2"))
    return [e for e in out if e]


def profiled_traffic(n):
    """dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel (the three k_epa tier
    launches of one cfg-3 pass) from the committed ncu counters, per pair, scaled to this launch
    (profiles/r2_kernels.json, taken at 4 M pairs; round 1's r1_traffic.json as a fallback)."""
    p2 = os.path.join(ROOT, "profiles", "r2_kernels.json")
    if os.path.exists(p2):
        d = json.load(open(p2))
        k = [r for name, r in d.get("kernels", {}).items() if name.startswith("cfg3:k_epa")]
        units = d.get("units", {}).get("cfg3")
        if k and units:
            return sum(r["dram_bytes"] for r in k) / units * n
    p = os.path.join(ROOT, "profiles", "r1_traffic.json")
    if not os.path.exists(p):
        return None
    d = json.load(open(p))
    return d["k_epa_dram_bytes_per_pair"] * n


def cpu_baseline(w, sample, threads=None):
    from oracle.binding import Oracle
    o = Oracle()
    threads = threads or o.hw_threads()
    m = min(sample, len(w["l_shape"]))
    args = (0, w["shapes"], w["verts"], w["l_shape"][:m], w["r_shape"][:m], w["l_t"][:m],
            w["r_t"][:m])
    o.intersection(0, w["shapes"], w["verts"], w["l_shape"][:2048], w["r_shape"][:2048],
                   w["l_t"][:2048], w["r_t"][:2048], nthreads=threads)  # warm-up
    t = time.perf_counter()
    o.intersection(*args, nthreads=threads)
    dt = time.perf_counter() - t
    out = {"value": m / dt, "unit": UNIT, "cores": threads, "kind": "port",
           "sample": f"first {m} pairs of the same workload, oracle port "
                     f"(g++ -O2 -ffp-contract=off), {threads} threads, {dt:.2f} s"}
    # BASELINE.json configs[0]: the reference's own CPU-runnable case, as specified — 10 k random
    # Cuboid-Cuboid pairs, FullResolution, ONE thread (best of 3 passes)
    from collision_rs_b200 import workloads as W
    w1 = W.cfg1_cuboid_pairs(10_000)
    best = None
    for _ in range(3):
        t = time.perf_counter()
        o.intersection(0, w1["shapes"], None, w1["l_shape"], w1["r_shape"], w1["l_t"], w1["r_t"], nthreads=1)
        d1 = time.perf_counter() - t
        best = d1 if best is None or d1 < best else best
    out["cfg1_single_thread"] = {"pairs": 10_000, "ms": best * 1e3, "pairs_per_s": 10_000 / best,
                                 "what": "GJK3+EPA3 FullResolution on 10k Cuboid-Cuboid pairs, oracle port, 1 thread"}
    return out


# ---------------------------------------------------------------------------------------------
def run_reference(a):
    """--impl reference: the reference algorithm's CPU path (the oracle port: the Rust crate
    cannot be built here or on the box — no rustc, cgmath not vendored) on all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from collision_rs_b200 import workloads as W
    from oracle.binding import Oracle
    o = Oracle()
    threads = o.hw_threads()
    m = min(a.cpu_sample, a.pairs)
    w = W.cfg3_polyhedron_pairs(m)
    args = (0, w["shapes"], w["verts"], w["l_shape"], w["r_shape"], w["l_t"], w["r_t"])
    for _ in range(a.warmup):
        o.intersection(0, w["shapes"], w["verts"], w["l_shape"][:4096], w["r_shape"][:4096],
                       w["l_t"][:4096], w["r_t"][:4096], nthreads=threads)
    t = time.perf_counter()
    for _ in range(a.steps):
        o.intersection(*args, nthreads=threads)
    dt = time.perf_counter() - t
    v = m * a.steps / dt
    sample = (f"{m} pairs/step of the same workload (bounded sample of the 4M-pair batch), "
              f"oracle port on {threads} host threads")
    emit({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": a.gpus,
        "steps": a.steps, "warmup": a.warmup, "ms_per_step": dt / a.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": config_of(a),
        "run": {"pairs_per_step": m},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0})


# ---------------------------------------------------------------------------------------------
def emit(obj):
    """The ONE JSON line goes to the real stdout; everything else (NCCL banners, library chatter)
    was rerouted to stderr by main()."""
    os.write(_REAL_STDOUT, (json.dumps(obj) + "\n").encode())


_REAL_STDOUT = 1


def main():
    global _REAL_STDOUT
    a = parse()
    # keep stdout clean for the single JSON line: libraries (NCCL prints its version banner on
    # stdout) write to fd 1, so fd 1 is pointed at stderr and the JSON goes to a saved copy
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    if a.impl == "reference":
        return run_reference(a)

    import torch
    import torch.distributed as dist
    from collision_rs_b200 import workloads as W
    from collision_rs_b200.abi import CONTACT_DT, SHAPE_DT, XFORM_DT
    from collision_rs_b200.host import Context

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU fallback "
                         "(use --impl reference for the CPU oracle port)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    # staging buffers and the threads that feed them live next to the GPU (eight ranks on one socket's
    # memory controller cost 17 % end to end in round 1)
    from collision_rs_b200.sharding import bind_to_gpu_numa
    all_cpus = os.sched_getaffinity(0)
    numa = bind_to_gpu_numa(local)
    numa_cpus = f"{numa[0]}-{numa[-1]} ({len(numa)})" if numa else None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    ctx = Context(local)

    # ---- inputs: shape table from rank 0 over NCCL, pair shard per rank --------------------
    n = a.pairs
    w = W.cfg3_polyhedron_pairs(n, shard=rank if a.shard is None else a.shard)
    shapes, verts = w["shapes"], w["verts"]
    if world > 1:
        # the shape / vertex tables travel once over NCCL (NVLink); every rank rebuilds its
        # support cells from them.  No collective on the data path (collision_rs_b200/sharding.py).
        from collision_rs_b200.sharding import broadcast_shape_table
        shapes, verts = broadcast_shape_table(shapes if rank == 0 else None,
                                              verts if rank == 0 else None, src=0, device=dev)
        assert np.array_equal(shapes, w["shapes"]) and np.array_equal(verts, w["verts"])
    ctx.upload_shapes(shapes, verts)

    def pinned(arr):
        t = torch.from_numpy(arr.view(np.uint8).reshape(-1)).pin_memory()
        return t

    h_in = {k: pinned(w[k]) for k in ("l_shape", "r_shape", "l_t", "r_t")}
    d_in = {k: v.to(dev, non_blocking=True) for k, v in h_in.items()}
    d_out = torch.empty(n * CONTACT_DT.itemsize, dtype=torch.uint8, device=dev)
    d_st = torch.empty(n, dtype=torch.uint8, device=dev)
    h_out = torch.empty(n * CONTACT_DT.itemsize, dtype=torch.uint8).pin_memory()
    h_st = torch.empty(n, dtype=torch.uint8).pin_memory()
    torch.cuda.synchronize()
    stream = torch.cuda.current_stream()

    def step_dev():
        ctx.intersection_dev(0, d_in["l_shape"].data_ptr(), d_in["r_shape"].data_ptr(),
                             d_in["l_t"].data_ptr(), d_in["r_t"].data_ptr(), n,
                             d_out.data_ptr(), d_st.data_ptr(), stream.cuda_stream)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing (value) ------------------------------------------------------------
    for _ in range(max(a.warmup, 3)):
        step_dev()
    barrier()
    # per-stage CUDA events (on the launching stream) of one more untimed pass: the dominant
    # kernel's live duration for the roofline
    ctx.set_stage_timing(True)
    stage_runs = []
    for _ in range(3):
        step_dev()
        stage_runs.append(ctx.last_stage_ms())
    ctx.set_stage_timing(False)
    stage_ms = [float(x) for x in np.mean(np.asarray(stage_runs), axis=0)]
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    l0 = ctx.kernel_launches
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(a.steps + 1)]
    ev[0].record(stream)
    for k in range(a.steps):
        step_dev()
        ev[k + 1].record(stream)
    barrier()
    launches = ctx.kernel_launches - l0
    ms_total = ev[0].elapsed_time(ev[-1])
    per_step = [ev[k].elapsed_time(ev[k + 1]) for k in range(a.steps)]
    clocks = sampler.stop() if sampler else None
    t_ms = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    per_rank_ms = [ms_total / a.steps]
    if world > 1:
        allt = [torch.zeros_like(t_ms) for _ in range(world)]
        dist.all_gather(allt, t_ms)
        per_rank_ms = [float(x.item()) / a.steps for x in allt]
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms_total = float(t_ms.item())
    value = n * world * a.steps / (ms_total * 1e-3)

    # ---- end to end through the host-buffer C ABI (pinned host in, pinned host out) -----------------
    from collision_rs_b200.abi import Settings
    W_SETTINGS = Settings.default()

    def step_e2e():
        ctx._ck(ctx.lib.cb2_gjk3_intersection(
            ctx.h, 0, C.byref(W_SETTINGS), h_in["l_shape"].data_ptr(), h_in["r_shape"].data_ptr(),
            h_in["l_t"].data_ptr(), h_in["r_t"].data_ptr(), n, h_out.data_ptr(), h_st.data_ptr()))

    step_e2e()
    barrier()
    e2e_steps = max(2, min(a.steps, 5))
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        step_e2e()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    t_e = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
    e2e_value = n * world * e2e_steps / float(t_e.item())
    h2d = sum(int(v.numel()) for v in h_in.values())
    d2h = int(h_out.numel() + h_st.numel())
    # the host path and the device path must agree bit for bit
    assert torch.equal(h_st, d_st.cpu()) and torch.equal(h_out, d_out.cpu())
    hits = int((h_st == 1).sum())
    bad = int((h_st > 1).sum())

    # ---- the same through the body-table call (cb2_gjk3_intersection_bodies): a quarter as many bodies as
    #      pairs, poses drawn like cfg 3's, so the pair statistics match; only 8 bytes per pair go in ----
    bodies_e2e = None
    if not a.no_extra:
        rngb = np.random.default_rng([777, 9, rank])
        nb = max(n // 4, 2)
        b_shape = pinned(rngb.integers(0, len(shapes), nb).astype(np.uint32))
        b_t = pinned(W.random_xforms(rngb, nb, 0.75))
        b_pairs = pinned(rngb.integers(0, nb, (n, 2)).astype(np.uint32))

        def step_bodies():
            ctx._ck(ctx.lib.cb2_gjk3_intersection_bodies(
                ctx.h, 0, C.byref(W_SETTINGS), b_shape.data_ptr(), b_t.data_ptr(), nb, b_pairs.data_ptr(), n,
                h_out.data_ptr(), h_st.data_ptr()))

        step_bodies()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            step_bodies()
        torch.cuda.synchronize()
        t_b = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t_b, op=dist.ReduceOp.MAX)
        bodies_e2e = {"value": n * world * e2e_steps / float(t_b.item()), "unit": UNIT, "bodies_per_gpu": nb,
                      "pairs_per_gpu": n, "hit_fraction": float((h_st == 1).sum()) / n,
                      "h2d_bytes_per_step": int(b_shape.numel() + b_t.numel() + b_pairs.numel()),
                      "d2h_bytes_per_step": d2h,
                      "what": "cb2_gjk3_intersection_bodies on pinned host buffers: body table uploaded per "
                              "call, 8 bytes per pair streamed"}

    # ---- N > 1: what only exists on several GPUs (every rank takes part; rank 0 reports) -------------
    multi = {}
    if world > 1:
        multi = multi_gpu_extras(a, ctx, torch, dist, dev, stream, rank, world, d_out, d_st, n,
                                 ms_total / a.steps, shapes, verts)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel: k_epa (EPA3::process, both shared-memory tiers) -------------
    flops_pair, flops_gjk_pair = algorithmic_flops_per_pair(w)
    step_ms = float(np.mean(per_step))  # one intersection pass per step; events bracket it
    epa_ms = stage_ms[1] + stage_ms[2]
    epa_flops = (flops_pair - flops_gjk_pair) * n
    achieved = epa_flops / (epa_ms * 1e-3) / 1e12
    peak_lib = C.CDLL(os.path.join(ROOT, "collision_rs_b200", "libcb2_peak.so"))
    peak_lib.cb2_peak_fp32_tflops.restype = C.c_double
    fp32_peak = float(peak_lib.cb2_peak_fp32_tflops(local))
    hbm_peak, hbm_src = peaks()
    bytes_pair = 2 * 4 + 2 * XFORM_DT.itemsize + CONTACT_DT.itemsize + 1
    hbm_achieved = bytes_pair * n / (step_ms * 1e-3) / 1e9
    roofline = {
        "bound": "fp32", "kernel": "k_epa (EPA3::process; its three shared-memory tier launches)",
        "achieved": achieved, "peak": fp32_peak, "unit": "TFLOP/s",
        "frac": achieved / fp32_peak if fp32_peak > 0 else None,
        "traffic": profiled_traffic(n),
        "peak_source": "FFMA micro-benchmark measured in this run (libcb2_peak.so; "
                       "MEASURED_PEAKS.json has no FP32 figure); parity forbids FMA so the "
                       "reachable ceiling is peak/2",
        "frac_of_unfused_peak": 2 * achieved / fp32_peak if fp32_peak > 0 else None,
        "algorithmic_flops_per_pair": {"pass": flops_pair, "gjk": flops_gjk_pair,
                                       "epa": flops_pair - flops_gjk_pair},
        "how": "algorithmic flops = f32 + - * / sqrt the reference executes for these pairs "
               "(oracle op counter, incl. its O(V) vertex scans that the support cells avoid)",
        "kernel_ms": epa_ms,
        "stage_ms": {"k_gjk_hits": stage_ms[0], "k_epa_tier0": stage_ms[1],
                     "k_epa_tiers1_2": stage_ms[2], "second_chance": stage_ms[3]},
        "step": {"ms": step_ms, "achieved": flops_pair * n / (step_ms * 1e-3) / 1e12,
                 "frac": flops_pair * n / (step_ms * 1e-3) / 1e12 / fp32_peak if fp32_peak > 0 else None},
        "hbm": {"achieved": hbm_achieved, "peak": hbm_peak, "unit": "GB/s",
                "frac": hbm_achieved / hbm_peak, "bytes_per_pair": bytes_pair,
                "peak_source": hbm_src},
    }
    os.sched_setaffinity(0, all_cpus)  # the CPU baseline gets every host core again
    cpu = cpu_baseline(w, a.cpu_sample)

    extra = {}
    if not a.no_extra and world == 1:
        extra = extras(ctx, torch, dev, stream)
    if multi:
        extra["multi"] = multi
    roofline["kernels"] = kernel_rooflines(n, stage_ms, flops_pair, flops_gjk_pair, extra, fp32_peak, hbm_peak)

    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps,
        "warmup": max(a.warmup, 3), "ms_per_step": ms_total / a.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_of(a),
        "run": {"hit_fraction": hits / n, "non_ok_status": bad, "ms_per_step_per_rank": per_rank_ms,
                "numa_cpus": numa_cpus},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "steps": e2e_steps,
                "how": "cb2_gjk3_intersection on pinned host buffers: 768 K-pair chunks through one "
                       "copy-in stream, one high-priority stream for the GJK stages (running ahead), "
                       "three streams for the EPA tiers and one copy-out stream over eight staging "
                       "buffers, all inside the timed call"},
        "e2e_bodies": bodies_e2e,
        "gpu_launches": launches,
        "roofline": roofline, "cpu_baseline": cpu, "extra": extra,
    }
    emit(out)
    if world > 1:
        dist.destroy_process_group()


def multi_gpu_extras(a, ctx, torch, dist, dev, stream, rank, world, d_out, d_st, n, step_ms, shapes, verts):
    """Collective-bearing and sharded passes (all ranks call this; times are the max over ranks).
    * gather: the per-rank contact + status records all-gathered ON THE DEVICE over NCCL — what a
      caller pays to have the whole contact list on every GPU (north_star: "gather contact lists");
    * strong: ONE 4M-pair batch split over the ranks (plus its gather);
    * cfg5: 8M time-of-impact pairs per GPU (64M over 8);
    * cfg4: the AABB table broadcast over NCCL, every GPU computes its shard of the pair list
      (cb2_sap3_find_collider_pairs_shard_dev) and runs the narrow phase on it."""
    from collision_rs_b200 import workloads as W
    from collision_rs_b200.abi import AABB_DT, CONTACT_DT
    from collision_rs_b200.sharding import gather_contacts_dev, shard_range
    s = stream.cuda_stream
    res = {}

    def dv(arr):
        return torch.from_numpy(arr.view(np.uint8).reshape(-1).copy()).to(dev)

    def max_ms(ms):
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def timed(fn, reps=3, warm=1):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(reps):
            fn()
        e1.record(stream)
        torch.cuda.synchronize()
        return max_ms(e0.elapsed_time(e1) / reps)

    # (a) device all-gather of this step's records: 37 B per pair from every rank to every rank
    total = n * world
    g_ms = timed(lambda: gather_contacts_dev(d_out, d_st, total))
    gathered = (CONTACT_DT.itemsize + 1) * total
    res["gather"] = {"what": "NCCL all_gather_into_tensor of contact (36 B) + status (1 B) records, device to device",
                     "pairs": total, "bytes_received_per_rank": gathered, "gather_ms": g_ms,
                     "algbw_gbs": gathered / (g_ms * 1e-3) / 1e9,
                     "step_ms": step_ms, "pairs_per_s_with_gather": total / ((step_ms + g_ms) * 1e-3)}
    # (b) strong scaling: one batch of a.pairs split over the ranks
    tot = a.pairs
    b, e = shard_range(tot, rank, world)
    w = W.cfg3_polyhedron_pairs(tot, shard=0)
    m = e - b
    d = {k: dv(np.ascontiguousarray(w[k][b:e])) for k in ("l_shape", "r_shape", "l_t", "r_t")}
    out = torch.empty(m * CONTACT_DT.itemsize, dtype=torch.uint8, device=dev)
    st = torch.empty(m, dtype=torch.uint8, device=dev)
    ms = timed(lambda: ctx.intersection_dev(0, d["l_shape"].data_ptr(), d["r_shape"].data_ptr(),
                                            d["l_t"].data_ptr(), d["r_t"].data_ptr(), m,
                                            out.data_ptr(), st.data_ptr(), s), reps=5, warm=2)
    g2 = timed(lambda: gather_contacts_dev(out, st, tot))
    res["strong"] = {"pairs_total": tot, "pairs_per_rank": m, "ms": ms, "pairs_per_s": tot / ms * 1e3,
                     "gather_ms": g2, "pairs_per_s_with_gather": tot / (ms + g2) * 1e3}
    del d, out, st, w
    # (c) cfg 5: 8M time-of-impact pairs per GPU
    n5 = 8_000_000
    w = W.cfg5_toi_pairs(n5, shard=rank)
    ctx.upload_shapes(w["shapes"])
    d = {k: dv(w[k]) for k in ("l_shape", "r_shape", "l_t0", "l_t1", "r_t0", "r_t1")}
    out = torch.empty(n5 * CONTACT_DT.itemsize, dtype=torch.uint8, device=dev)
    st = torch.empty(n5, dtype=torch.uint8, device=dev)
    ms = timed(lambda: ctx.time_of_impact_dev(d["l_shape"].data_ptr(), d["r_shape"].data_ptr(),
                                              d["l_t0"].data_ptr(), d["l_t1"].data_ptr(),
                                              d["r_t0"].data_ptr(), d["r_t1"].data_ptr(), n5,
                                              out.data_ptr(), st.data_ptr(), s), reps=3, warm=1)
    res["cfg5_time_of_impact"] = {"pairs_total": n5 * world, "pairs_per_rank": n5, "ms": ms,
                                  "pairs_per_s": n5 * world / ms * 1e3}
    del d, out, st, w
    # (d) cfg 4: broadcast the AABB table, shard the pair list, narrow phase on each shard
    nb = 1_000_000
    w = W.cfg4_aabbs(nb)
    boxes = dv(w["aabbs"]) if rank == 0 else torch.zeros(nb * AABB_DT.itemsize, dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()
    b_ms = timed(lambda: dist.broadcast(boxes, 0), reps=3, warm=1)
    ctx.upload_shapes(w["shapes"])
    perm = torch.empty(nb, dtype=torch.int32, device=dev)
    cap = 8 * nb
    pairs = torch.empty(2 * cap, dtype=torch.int32, device=dev)
    cnt = {}

    def sap():
        rc, c = ctx.sap3_shard_dev(boxes.data_ptr(), nb, 0, rank, world, perm.data_ptr(), pairs.data_ptr(), cap, s)
        assert rc == 0
        cnt["n"] = c

    sap_ms = timed(sap, reps=5, warm=2)
    mine = int(cnt["n"])
    t = torch.tensor([mine], dtype=torch.int64, device=dev)
    dist.all_reduce(t)
    p64 = perm.long()
    body_shape = dv(np.arange(nb, dtype=np.uint32)).view(torch.int32)[p64].contiguous()
    body_t = dv(w["body_t"]).view(-1, 32)[p64].contiguous()
    outp = torch.empty(max(mine, 1) * CONTACT_DT.itemsize, dtype=torch.uint8, device=dev)
    stp = torch.empty(max(mine, 1), dtype=torch.uint8, device=dev)
    np_ms = timed(lambda: ctx.intersection_bodies_dev(0, body_shape.data_ptr(), body_t.data_ptr(),
                                                      pairs.data_ptr(), mine, outp.data_ptr(), stp.data_ptr(), s))
    res["cfg4_sap_sharded"] = {"aabbs": nb, "pairs_total": int(t.item()), "pairs_this_rank": mine,
                               "broadcast_ms": b_ms, "sap_shard_ms": sap_ms, "aabbs_per_s": nb / sap_ms * 1e3,
                               "narrow_on_shard_ms": np_ms, "narrow_pairs_per_s": int(t.item()) / np_ms * 1e3}
    return res


def extras(ctx, torch, dev, stream):
    """Secondary configs (N=1): cfg1 cuboids, cfg2 mixed, cfg4 SAP (+ chained narrow phase), cfg5 TOI.
    Device-resident, CUDA-event timed, 3 warm-ups + 5 timed passes each."""
    from collision_rs_b200 import workloads as W
    from collision_rs_b200.abi import AABB_DT, CONTACT_DT

    def dv(arr):
        return torch.from_numpy(arr.view(np.uint8).reshape(-1).copy()).to(dev)

    def timed(fn, reps=5, warm=3):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(reps):
            fn()
        e1.record(stream)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    res = {}
    s = stream.cuda_stream
    # cfg1 at GPU scale: 4M cuboid pairs, FullResolution
    n = 4_000_000
    w = W.cfg1_cuboid_pairs(n)
    ctx.upload_shapes(w["shapes"])
    d = {k: dv(w[k]) for k in ("l_shape", "r_shape", "l_t", "r_t")}
    out = torch.empty(n * CONTACT_DT.itemsize, dtype=torch.uint8, device=dev)
    st = torch.empty(n, dtype=torch.uint8, device=dev)
    ms = timed(lambda: ctx.intersection_dev(0, d["l_shape"].data_ptr(), d["r_shape"].data_ptr(),
                                            d["l_t"].data_ptr(), d["r_t"].data_ptr(), n,
                                            out.data_ptr(), st.data_ptr(), s))
    res["cfg1_cuboid_full_resolution"] = {"pairs": n, "ms": ms, "pairs_per_s": n / ms * 1e3}
    # cfg2: 1M mixed, CollisionOnly + distance
    n = 1_000_000
    w = W.cfg2_mixed_pairs(n)
    ctx.upload_shapes(w["shapes"])
    d = {k: dv(w[k]) for k in ("l_shape", "r_shape", "l_t", "r_t")}
    dist_out = torch.empty(n, dtype=torch.float32, device=dev)
    ms_c = timed(lambda: ctx.intersection_dev(1, d["l_shape"].data_ptr(), d["r_shape"].data_ptr(),
                                              d["l_t"].data_ptr(), d["r_t"].data_ptr(), n,
                                              out.data_ptr(), st.data_ptr(), s))
    ms_d = timed(lambda: ctx.distance_dev(d["l_shape"].data_ptr(), d["r_shape"].data_ptr(),
                                          d["l_t"].data_ptr(), d["r_t"].data_ptr(), n,
                                          dist_out.data_ptr(), st.data_ptr(), s))
    res["cfg2_mixed_collision_only"] = {"pairs": n, "ms": ms_c, "pairs_per_s": n / ms_c * 1e3,
                                        "algorithmic_flops_per_pair": counted_flops("intersection1", w)}
    res["cfg2_mixed_distance"] = {"pairs": n, "ms": ms_d, "pairs_per_s": n / ms_d * 1e3,
                                  "algorithmic_flops_per_pair": counted_flops("distance", w)}
    # ... and per type pair (Sphere-Cuboid distance is dominated by the 100-iteration cap)
    kl, kr = w["shapes"]["kind"][w["l_shape"]], w["shapes"]["kind"][w["r_shape"]]
    names = {0: "sphere", 1: "cuboid"}
    for a_ in (0, 1):
        for b_ in (0, 1):
            sel = np.nonzero((kl == a_) & (kr == b_))[0]
            m = len(sel)
            db = {k: dv(np.ascontiguousarray(w[k][sel])) for k in ("l_shape", "r_shape", "l_t", "r_t")}
            ms_c = timed(lambda: ctx.intersection_dev(1, db["l_shape"].data_ptr(), db["r_shape"].data_ptr(),
                                                      db["l_t"].data_ptr(), db["r_t"].data_ptr(), m,
                                                      out.data_ptr(), st.data_ptr(), s))
            hit = float((st[:m] == 1).float().mean())
            ms_d = timed(lambda: ctx.distance_dev(db["l_shape"].data_ptr(), db["r_shape"].data_ptr(),
                                                  db["l_t"].data_ptr(), db["r_t"].data_ptr(), m,
                                                  dist_out.data_ptr(), st.data_ptr(), s))
            res[f"cfg2_{names[a_]}_{names[b_]}"] = {
                "pairs": m, "collision_only_pairs_per_s": m / ms_c * 1e3, "hit_fraction": hit,
                "distance_pairs_per_s": m / ms_d * 1e3, "distance_some_fraction": float((st[:m] == 1).float().mean())}
    # cfg4: SAP on 1M AABBs feeding the narrow phase
    n = 1_000_000
    w = W.cfg4_aabbs(n)
    ctx.upload_shapes(w["shapes"])
    boxes = dv(w["aabbs"])
    perm = torch.empty(n, dtype=torch.int32, device=dev)
    cap = 8 * n
    pairs = torch.empty(2 * cap, dtype=torch.int32, device=dev)
    box_sorted = torch.empty_like(boxes)
    np_holder = {}

    def sap():
        rc, cnt = ctx.sap3_dev(boxes.data_ptr(), n, 0, perm.data_ptr(), pairs.data_ptr(), cap, s)
        np_holder["n"] = cnt
        assert rc == 0

    ms_sap = timed(sap)
    npairs = np_holder["n"]
    hbm_peak, _ = peaks()
    alg_bytes = 24 * n + 4 * n + 8 * npairs
    res["cfg4_sap"] = {"aabbs": n, "pairs": npairs, "ms": ms_sap, "aabbs_per_s": n / ms_sap * 1e3,
                       "algorithmic_bytes": alg_bytes,
                       "hbm_frac": alg_bytes / (ms_sap * 1e-3) / 1e9 / hbm_peak}
    def sap_set():
        rc, cnt = ctx.sap3_unordered_dev(boxes.data_ptr(), n, 0, perm.data_ptr(), pairs.data_ptr(), cap, s)
        assert rc == 0 and cnt == npairs

    ms_set = timed(sap_set)
    res["cfg4_sap_pair_set"] = {"aabbs": n, "pairs": npairs, "ms": ms_set, "aabbs_per_s": n / ms_set * 1e3,
                                "what": "same pairs, unspecified order (feeds the narrow phase): no final pair sort"}
    sap()  # leave the ordered list in `pairs` for the chained narrow phase below
    body_shape = dv(np.arange(n, dtype=np.uint32))
    body_t = dv(w["body_t"])
    # bodies indexed in SORTED order: permute shapes / transforms once, like the reference's slice
    p64 = perm.long()
    body_shape_s = body_shape.view(torch.int32)[p64].contiguous()
    body_t_s = body_t.view(-1, 32)[p64].contiguous()
    outp = torch.empty(npairs * CONTACT_DT.itemsize, dtype=torch.uint8, device=dev)
    stp = torch.empty(npairs, dtype=torch.uint8, device=dev)
    ms_np = timed(lambda: ctx.intersection_bodies_dev(0, body_shape_s.data_ptr(),
                                                      body_t_s.data_ptr(), pairs.data_ptr(),
                                                      npairs, outp.data_ptr(), stp.data_ptr(), s))
    res["cfg4_narrow_on_sap_pairs"] = {"pairs": npairs, "ms": ms_np,
                                       "pairs_per_s": npairs / ms_np * 1e3,
                                       "contacts": int((stp == 1).sum())}
    # cfg5: TOI, 8M cuboid pairs on one GPU (64M over 8)
    n = 8_000_000
    w = W.cfg5_toi_pairs(n)
    ctx.upload_shapes(w["shapes"])
    d = {k: dv(w[k]) for k in ("l_shape", "r_shape", "l_t0", "l_t1", "r_t0", "r_t1")}
    out = torch.empty(n * CONTACT_DT.itemsize, dtype=torch.uint8, device=dev)
    st = torch.empty(n, dtype=torch.uint8, device=dev)
    ms = timed(lambda: ctx.time_of_impact_dev(d["l_shape"].data_ptr(), d["r_shape"].data_ptr(),
                                              d["l_t0"].data_ptr(), d["l_t1"].data_ptr(),
                                              d["r_t0"].data_ptr(), d["r_t1"].data_ptr(), n,
                                              out.data_ptr(), st.data_ptr(), s), reps=3, warm=1)
    res["cfg5_time_of_impact"] = {"pairs": n, "ms": ms, "pairs_per_s": n / ms * 1e3,
                                  "nonconverged": int((st == 5).sum()), "iter_cap": 1024,
                                  "algorithmic_flops_per_pair": counted_flops("toi", w)}
    # 2-D narrow phase (not a BASELINE config): GJK2 + EPA2 on 4M mixed Primitive2
    # pairs and on rectangles (cfg 1's 2-D twin)
    from collision_rs_b200.abi import CONTACT2_DT
    n = 4_000_000
    for name, kinds, spread in (("gjk2_mixed_primitive2", (0, 1, 2, 3, 4, 5), 1.0),
                                ("gjk2_rectangles", (3, 4), 0.75)):
        w = W.cfg2d_mixed_pairs(n, seed=11, spread=spread, kinds=kinds)
        ctx.upload_shapes2(w["shapes"], w["verts"])
        d = {k: dv(w[k]) for k in ("l_shape", "r_shape", "l_t", "r_t")}
        out2 = torch.empty(n * CONTACT2_DT.itemsize, dtype=torch.uint8, device=dev)
        st2 = torch.empty(n, dtype=torch.uint8, device=dev)
        ms = timed(lambda: ctx.intersection2_dev(0, d["l_shape"].data_ptr(), d["r_shape"].data_ptr(),
                                                 d["l_t"].data_ptr(), d["r_t"].data_ptr(), n,
                                                 out2.data_ptr(), st2.data_ptr(), s))
        res[name + "_full_resolution"] = {"pairs": n, "ms": ms, "pairs_per_s": n / ms * 1e3,
                                          "hit_fraction": float((st2 == 1).float().mean())}
    return res


if __name__ == "__main__":
    main()
