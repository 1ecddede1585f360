#!/bin/bash
# Round-1 captures (run on the GPU box through gpurun); summaries: python profiles/summarize.py r1
B="python bench.py --steps 2 --warmup 3 --pairs 1000000 --no-extra"
$B > gpurun_out/plain.log 2>&1 || { echo "bench failed"; tail -5 gpurun_out/plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1.csv $B > gpurun_out/ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_epa|k_gjk_hits' --launch-skip 4 --launch-count 4 \
    -f -o gpurun_out/prof_r1 $B > gpurun_out/ncu2.log 2>&1
python tools/sap_probe.py 1000000 2 > gpurun_out/sap_plain.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_sap_r1.csv \
    python tools/sap_probe.py 1000000 2 > gpurun_out/ncu_sap.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_grid_sweep' --launch-skip 3 --launch-count 1 \
    -f -o gpurun_out/prof_sap_r1 python tools/sap_probe.py 1000000 2 > gpurun_out/ncu_sap2.log 2>&1
python tools/gjk2_probe.py 1000000 > gpurun_out/gjk2_plain.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_epa2|k_gjk2_intersection' --launch-skip 2 --launch-count 2 \
    -f -o gpurun_out/prof_gjk2_r1 python tools/gjk2_probe.py 1000000 > gpurun_out/ncu_gjk2.log 2>&1
ls -la gpurun_out/*.ncu-rep; tail -2 gpurun_out/plain.log | cut -c1-300; cat gpurun_out/sap_plain.log; tail -5 gpurun_out/gjk2_plain.log
