#!/bin/bash
# One `ncu --set full` capture of the level-0 launch of one kernel inside a short bench run, digested to the counters that
# say what bounds it (issue slots, pipes, stall reasons, occupancy, DRAM bytes).  usage: bash tools/ncu_kernel.sh <kernel-regex> [out-tag]
# Results: gpurun_out/ncu_<tag>.txt (+ the .ncu-rep).  Never a timing source: ncu replays each kernel ~40 times.
set -u
cd "${GRAFT_REPO_ROOT:-.}"
K=$1; TAG=${2:-$1}
CMD="python bench.py --frames 256 --steps 1 --warmup 3 --no-cpu-baseline --no-extra-legs"
# skip the warm-up launches of this kernel: the (3 x launches-per-step + 1)-th launch is the timed step's level 0
PER=$(python - <<PY
print({"blur7": 8, "fast9": 8, "pyr_down": 7}.get("$K".split("_kernel")[0].replace("_mask", ""), 1))
PY
)
SKIP=$((3 * PER))
ncu --clock-control none --set full --import-source on -k regex:$K --launch-skip $SKIP --launch-count 1 -o gpurun_out/ncu_$TAG -f $CMD > gpurun_out/ncu_$TAG.log 2>&1
ncu -i gpurun_out/ncu_$TAG.ncu-rep --page raw --csv > /tmp/raw_$TAG.csv 2>/dev/null
ncu -i gpurun_out/ncu_$TAG.ncu-rep --page source --csv > gpurun_out/src_$TAG.csv 2>/dev/null
python - /tmp/raw_$TAG.csv > gpurun_out/ncu_$TAG.txt <<'PY'
import csv, sys
rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
head, units, body = rows[0], rows[1], rows[2:]
r = body[0]
want = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__occupancy_limit", "gpu__time_duration.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput", "sm__warps_active.avg.pct", "smsp__issue_active.avg.pct",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed", "sm__inst_executed_pipe_", "sm__pipe_", "smsp__average_warp", "smsp__warp_issue_stalled",
        "smsp__average_warps_issue_stalled", "l1tex__t_sector_hit_rate", "lts__t_sector_hit_rate", "l1tex__data_bank_conflicts", "smsp__warps_eligible", "sm__maximum_warps",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed_op_shared"]
for i, h in enumerate(head):
    if any(w in h for w in want):
        v = r[i]
        try:
            if float(v.replace(",", "")) == 0.0: continue
        except ValueError:
            pass
        print(f"{h} [{units[i]}] = {v}")
PY
tail -3 gpurun_out/ncu_$TAG.log
