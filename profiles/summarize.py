#!/usr/bin/env python
"""Turn the raw captures in gpurun_out/ into the committed summaries of this round:
  profiles/r1_launches_cfg3.txt     per-launch times of one cfg-3 step (shares of the step)
  profiles/r1_launches_sap.txt      per-launch times of one SweepAndPrune3 call (1 M boxes)
  profiles/r1_ncu_kernels.csv       the --set full metrics the roofline quotes, per kernel
  profiles/r1_k_epa_by_line.txt     stall samples / instructions by source line of k_epa (tier 0)
  profiles/r1_traffic.json          dram bytes per pair of k_epa (bench.py's roofline.traffic)
usage: python profiles/summarize.py [tag]   (tag: r1)"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
GO = os.path.join(ROOT, "gpurun_out")
tag = sys.argv[1] if len(sys.argv) > 1 else "r1"


def run(cmd, **kw):
    return subprocess.run(cmd, capture_output=True, text=True, **kw).stdout


def launches(csv_in, first, occ, out_name, title):
    txt = run([sys.executable, os.path.join(OUT, "launch_summary.py"), csv_in, first, str(occ)])
    with open(os.path.join(OUT, out_name), "w") as f:
        f.write(title + "\n" + txt)
    return txt


WANT = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
    "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
    "smsp__sass_thread_inst_executed_op_fadd_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_fmul_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
]


def raw_rows(rep):
    txt = run(["ncu", "-i", rep, "--page", "raw", "--csv"])
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    return hdr, units, rows[2:]


def kernels_csv(reps):
    out = [["report", "kernel"] + WANT]
    units_row = None
    traffic = {}
    for rep in reps:
        hdr, units, rows = raw_rows(rep)
        ix = {h: i for i, h in enumerate(hdr)}
        if units_row is None:
            units_row = ["", "(unit)"] + [units[ix[w]] if w in ix else "" for w in WANT]
            out.append(units_row)
        for r in rows:
            name = r[ix["Kernel Name"]]
            out.append([os.path.basename(rep), name] + [r[ix[w]] if w in ix else "" for w in WANT])
            traffic.setdefault(name, []).append(r)
    with open(os.path.join(OUT, f"{tag}_ncu_kernels.csv"), "w", newline="") as f:
        csv.writer(f).writerows(out)
    return out


def main():
    launches(os.path.join(GO, f"launches_{tag}.csv"), "k_gjk_hits_p", 4, f"{tag}_launches_cfg3.txt",
             "# one cfg-3 step (1 M polyhedron pairs, FullResolution): ncu --metrics "
             "gpu__time_duration.sum --clock-control none, bench.py --steps 2 --warmup 3 "
             "--pairs 1000000 --no-extra")
    launches(os.path.join(GO, f"launches_sap_{tag}.csv"), "k_sap_keys", 0, f"{tag}_launches_sap.txt",
             "# one SweepAndPrune3::find_collider_pairs call, cfg 4 (1 M boxes, 3.95 M pairs): "
             "tools/sap_probe.py under the same ncu pass")
    reps = [os.path.join(GO, f"prof_{tag}.ncu-rep"), os.path.join(GO, f"prof_sap_{tag}.ncu-rep"),
            os.path.join(GO, f"prof_gjk2_{tag}.ncu-rep")]
    rows = kernels_csv([r for r in reps if os.path.exists(r)])
    # dram traffic of k_epa per pair (all tier launches of one pass, 1 M pairs)
    hdr = rows[0]
    ki, ri, wi, ti = hdr.index("kernel"), hdr.index("dram__bytes_read.sum"), hdr.index(
        "dram__bytes_write.sum"), hdr.index("gpu__time_duration.sum")
    units = rows[1]
    scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
    epa = [r for r in rows[2:] if "k_epa<" in r[ki]]
    tot = sum(float(r[ri]) * scale[units[ri]] + float(r[wi]) * scale[units[wi]] for r in epa)
    json.dump({"k_epa_dram_bytes_per_pair": tot / 1e6, "pairs_profiled": 1_000_000,
               "launches_summed": len(epa),
               "source": f"profiles/{tag}_ncu_kernels.csv (ncu --set full, dram__bytes_read.sum + "
                         "dram__bytes_write.sum of the k_epa launches of one cfg-3 pass)"},
              open(os.path.join(OUT, f"{tag}_traffic.json"), "w"), indent=1)
    # per-line view of the dominant kernel
    rep = reps[0]
    sass = run(["ncu", "-i", rep, "--page", "source", "--csv"])
    lines = sass.splitlines()
    starts = [i for i, l in enumerate(lines) if l.startswith('"Kernel Name"')]
    names = [lines[i] for i in starts]
    k = [i for i, nme in enumerate(names) if "k_epa<(int)4, (int)40" in nme][0]
    seg = lines[starts[k]:starts[k + 1] if k + 1 < len(starts) else len(lines)]
    tmp = "/tmp/_sass_seg.csv"
    open(tmp, "w").write("\n".join(seg) + "\n")
    run(["cuobjdump", "-xelf", "cb2_epa.sm_100a.cubin",
         os.path.join(ROOT, "collision_rs_b200", "libcollision_b200.so")], cwd="/tmp")
    by_line = run([sys.executable, os.path.join(OUT, "ncu_by_line.py"), tmp,
                   "/tmp/cb2_epa.sm_100a.cubin", "k_epaILi4ELi40", "40", "outer"])
    with open(os.path.join(OUT, f"{tag}_k_epa_by_line.txt"), "w") as f:
        f.write("# k_epa tier 0: warp-stall samples and executed instructions charged to the line "
                "of the kernel body\n# (ncu --set full --import-source on; profiles/ncu_by_line.py ... outer)\n")
        f.write(by_line)
    print(open(os.path.join(OUT, f"{tag}_launches_cfg3.txt")).read())
    print(open(os.path.join(OUT, f"{tag}_launches_sap.txt")).read())
    print(open(os.path.join(OUT, f"{tag}_traffic.json")).read())
    print(by_line[:1500])


if __name__ == "__main__":
    main()
