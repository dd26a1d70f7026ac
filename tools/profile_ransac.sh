#!/bin/bash
# Two-view evidence: config-3 bench line, RANSAC probe, ncu --set full of the four two-view kernels
set -u
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out
python bench.py --workload tum > $O/bench_final_tum_1gpu.json 2> $O/bench_final_tum_1gpu.err
python tools/ransac_probe.py 512 > $O/ransac_probe_final.log 2>&1
ncu --clock-control none --set full --import-source on -k regex:essential --launch-skip 4 --launch-count 4 -o /tmp/full_ransac -f python tools/ransac_probe.py 512 > $O/ncu_full_ransac.log 2>&1
ncu -i /tmp/full_ransac.ncu-rep --page raw --csv > $O/raw_full_ransac.csv 2>> $O/ncu_full_ransac.log
head -3 $O/ransac_probe_final.log | tail -1
