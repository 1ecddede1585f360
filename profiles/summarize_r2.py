"""Per-kernel counters of the round-2 profile pass -> profiles/r2_kernels.json + a readable table.

    python profiles/summarize_r2.py gpurun_out/r2_metrics.csv

Input: `ncu --metrics <list> --csv` of `tools/profile_pass.py` (tools/r2_gpu_metrics.sh): every
BASELINE config once at its full size.  Each config runs twice; the LAST launch of every kernel
within a config section is the one reported (warm tables, workspaces allocated).  Executed FP32 work
= fadd + fmul + 2 x ffma thread instructions (the parity path never fuses, so ffma comes from
address arithmetic of the compiler and the support-cell build only); XU = MUFU (rcp / rsqrt seeds
of the IEEE divide and square root sequences)."""
import csv
import json
import os
import re
import sys
from collections import OrderedDict, defaultdict

path = sys.argv[1]
rows = [r for r in csv.reader(open(path)) if len(r) >= 15 and r[0].isdigit()]
launch = OrderedDict()
for r in rows:
    lid = int(r[0])
    d = launch.setdefault(lid, {"name": r[4], "block": r[7], "grid": r[8]})
    try:
        d[r[12]] = float(r[14].replace(",", ""))
    except ValueError:
        d[r[12]] = r[14]
    d["unit:" + r[12]] = r[13]


def short(name):
    n = re.sub(r"\(.*", "", name)
    n = n.replace("void ", "").replace("cb2::", "")
    n = re.sub(r"(unsigned )?(int|long|char|bool)\)", "", n)
    return n.replace("(int)", "").replace("(bool)", "").replace(" ", "")


# config sections: tools/profile_pass.py launches a one-element torch fill before each config
order = (sys.argv[2].split(",") if len(sys.argv) > 2 else ["cfg3", "cfg2", "cfg5", "cfg4", "gjk2"])
sec = -1
by_kernel = defaultdict(list)
have_marks = any("FillFunctor" in d["name"] and d["grid"].startswith("(1,") for d in launch.values())
for lid, d in launch.items():
    if "FillFunctor" in d["name"] and d["grid"].startswith("(1,"):
        sec += 1
        continue
    if d["name"].startswith("void at::") or d["name"].startswith("at::"):
        continue  # torch's own index / copy kernels of the harness
    tag = order[sec] if have_marks and 0 <= sec < len(order) else "all"
    by_kernel[tag + ":" + short(d["name"])].append(d)

PEAK_FP32 = float(os.environ.get("CB2_FP32_PEAK_TFLOPS", "72.5"))   # FFMA peak measured by libcb2_peak.so
peaks = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "MEASURED_PEAKS.json"))) \
    if os.path.exists(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "MEASURED_PEAKS.json")) else {}
PEAK_HBM = float(peaks.get("hbm_gbs", 6541.8))

out = OrderedDict()
print(f"{'kernel':58s} {'n':>3s} {'ms':>9s} {'regs':>4s} {'occ%':>5s} {'IPC':>5s} {'thr/inst':>8s} {'issue%':>6s} "
      f"{'FP32 TF/s':>9s} {'%peak':>6s} {'%unfused':>8s} {'XU G/s':>7s} {'DRAM GB/s':>9s} {'%HBM':>5s}")
for k, ls in by_kernel.items():
    d = ls[-1]  # last launch: warm
    t_ns = d.get("gpu__time_duration.sum", 0.0)
    unit = d.get("unit:gpu__time_duration.sum", "ns")
    t_s = t_ns * {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}.get(unit, 1e-9)
    if t_s <= 0:
        continue
    fadd = d.get("smsp__sass_thread_inst_executed_op_fadd_pred_on.sum", 0.0)
    fmul = d.get("smsp__sass_thread_inst_executed_op_fmul_pred_on.sum", 0.0)
    ffma = d.get("smsp__sass_thread_inst_executed_op_ffma_pred_on.sum", 0.0)
    xu = d.get("smsp__inst_executed_pipe_xu.sum", 0.0)
    flops = fadd + fmul + 2 * ffma
    dr = d.get("dram__bytes_read.sum", 0.0) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(d.get("unit:dram__bytes_read.sum", "byte"), 1)
    dw = d.get("dram__bytes_write.sum", 0.0) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(d.get("unit:dram__bytes_write.sum", "byte"), 1)
    inst = d.get("smsp__inst_executed.sum", 0.0)
    tinst = d.get("smsp__thread_inst_executed.sum", 0.0)
    rec = {
        "launches_in_pass": len(ls), "ms": t_s * 1e3, "grid": d["grid"], "block": d["block"],
        "registers": d.get("launch__registers_per_thread"),
        "warps_active_pct": d.get("sm__warps_active.avg.pct_of_peak_sustained_active"),
        "ipc": d.get("sm__inst_executed.avg.per_cycle_elapsed"),
        "threads_per_inst": tinst / inst if inst else None,
        "issue_active_pct": d.get("smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "inst_executed": inst, "thread_inst_executed": tinst,
        "fadd": fadd, "fmul": fmul, "ffma": ffma, "xu_warp_inst": xu,
        "executed_fp32_tflops": flops / t_s / 1e12,
        "executed_fp32_frac_of_ffma_peak": flops / t_s / 1e12 / PEAK_FP32,
        "executed_fp32_frac_of_unfused_peak": flops / t_s / 1e12 / (PEAK_FP32 / 2),
        "fp32_share_of_thread_inst": (fadd + fmul + ffma) / tinst if tinst else None,
        "dram_bytes": dr + dw, "dram_gbs": (dr + dw) / t_s / 1e9, "dram_frac_of_hbm_peak": (dr + dw) / t_s / 1e9 / PEAK_HBM,
        "l1_hit_pct": d.get("l1tex__t_sector_hit_rate.pct"), "l2_hit_pct": d.get("lts__t_sector_hit_rate.pct"),
    }
    out[k] = rec
    print(f"{k[:58]:58s} {len(ls):3d} {rec['ms']:9.3f} {int(rec['registers'] or 0):4d} {rec['warps_active_pct'] or 0:5.1f} "
          f"{rec['ipc'] or 0:5.2f} {rec['threads_per_inst'] or 0:8.1f} {rec['issue_active_pct'] or 0:6.1f} "
          f"{rec['executed_fp32_tflops']:9.2f} {100 * rec['executed_fp32_frac_of_ffma_peak']:6.1f} "
          f"{100 * rec['executed_fp32_frac_of_unfused_peak']:8.1f} {xu / t_s / 1e9:7.1f} {rec['dram_gbs']:9.1f} "
          f"{100 * rec['dram_frac_of_hbm_peak']:5.1f}")
dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "r2_kernels.json")
UNITS = {"cfg3": int(os.environ.get("CB2_PROFILE_PAIRS", 4_000_000)), "cfg2": 1_000_000, "cfg5": 8_000_000,
         "cfg4": 1_000_000, "gjk2": 4_000_000}
json.dump({"source": os.path.basename(path), "fp32_ffma_peak_tflops": PEAK_FP32, "hbm_peak_gbs": PEAK_HBM,
           "units": UNITS,
           "note": "last launch of each kernel in tools/profile_pass.py (every BASELINE config at full size, "
                   "ncu --clock-control none, cold-cache serialised launches: use shares and per-unit counts, "
                   "not absolute times)", "kernels": out}, open(dst, "w"), indent=1)
print("wrote", dst)
