#!/bin/bash
# Round-2 evidence on one B200 (everything lands in gpurun_out/; the committed copies under profiles/ are made from these):
#   r02_bench_final.json          default `python bench.py` line (headline + mode_reference / config3 / hamming_sweep legs)
#   r02_bench_final_refarm.json   `python bench.py --impl reference`
#   r02_launches_orb.csv          ncu launch list (gpu__time_duration.sum, --clock-control none) of a 256-frame run
#   r02_raw_full_orb.csv          ncu --set full of the launches of one resident step (summarised by tools/ncu_summary.py)
set -u
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out
NCU="ncu --clock-control none"
python bench.py > $O/r02_bench_final.out 2> $O/r02_bench_final.err; grep '^{' $O/r02_bench_final.out > $O/r02_bench_final.json
python bench.py --impl reference --steps 3 --warmup 1 > $O/r02_bench_final_refarm.json 2>> $O/r02_bench_final.err
CMD="python bench.py --frames 256 --steps 1 --warmup 3 --no-cpu-baseline --no-extra-legs"
$NCU --metrics gpu__time_duration.sum -c 3000 --csv --log-file $O/r02_launches_orb.csv $CMD > $O/r02_ncu_launches.log 2>&1
SKIP=$(python - "$O/r02_launches_orb.csv" <<'PY'
import csv, sys
rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))][1:]
names = [r[4] for r in rows]
starts = [i for i, n in enumerate(names) if "pyr_down" in n and (i == 0 or "pyr_down" not in names[i - 1])]
print(starts[3], starts[4] - starts[3] if len(starts) > 4 else len(names) - starts[3])  # three warm-up steps, then the first timed one; launches per step
PY
)
PER=${SKIP#* }
SKIP=${SKIP% *}
echo "launch-skip $SKIP, $PER launches" > $O/r02_profile.log
$NCU --set full --import-source on --launch-skip $SKIP --launch-count $PER -o /tmp/r02_full_orb -f $CMD > $O/r02_ncu_full.log 2>&1
ncu -i /tmp/r02_full_orb.ncu-rep --page raw --csv > $O/r02_raw_full_orb.csv 2>> $O/r02_profile.log
python tools/ncu_summary.py $O/r02_raw_full_orb.csv $O/r02_orb_ncu_full_summary.csv --json $O/r02_ncu_dram_per_frame.json --frames 256 --note "ncu --set full --clock-control none, one resident step of bench.py --frames 256 (ORB mode), round 2" >> $O/r02_profile.log 2>&1
# tensor / TMEM counters of the matcher, whatever this ncu calls them
head -1 $O/r02_raw_full_orb.csv | tr ',' '\n' | grep -inE "tensor|tmem|pipe_tc|utc" | head -40 > $O/r02_ncu_tensor_metric_names.txt
tail -3 $O/r02_profile.log; tail -c 400 $O/r02_bench_final.err
