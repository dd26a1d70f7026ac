#!/bin/bash
# Round evidence on one B200: bench lines (both modes + their CPU reference arms), ncu launch lists, one ncu --set full
# capture of a complete step per mode and of the RANSAC kernels.  Everything lands in gpurun_out/; the summaries that
# are committed under profiles/ are made from these files with tools/ncu_summary.py.
set -u
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out
NCU="ncu --clock-control none"
MODES="${MODES:-orb reference}"   # MODES=reference bash tools/profile_round.sh refreshes one mode only
for MODE in $MODES; do
  SFX=$([ $MODE = orb ] && echo orb || echo refmode)
  python bench.py --mode $MODE > $O/bench_final_$SFX.json 2> $O/bench_final_$SFX.err
  python bench.py --impl reference --mode $MODE --steps 3 --warmup 1 > $O/bench_final_${SFX}_refarm.json 2>> $O/bench_final_$SFX.err
done
for MODE in $MODES; do
  CMD="python bench.py --mode $MODE --frames 256 --steps 1 --warmup 3 --no-cpu-baseline"
  $NCU --metrics gpu__time_duration.sum -c 2000 --csv --log-file $O/launches_$MODE.csv $CMD > $O/ncu_launches_$MODE.log 2>&1
  FIRST=$([ $MODE = orb ] && echo pyr_down_kernel || echo fast_mask_kernel)
  PER=$([ $MODE = orb ] && echo 31 || echo 11)
  # first launch of the timed resident step = the (3 warm-up steps + 1)-th step that starts with $FIRST
  SKIP=$(python - "$O/launches_$MODE.csv" $FIRST $MODE <<'PY'
import csv, sys
rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))][1:]
names = [r[4] for r in rows]
starts = [i for i, n in enumerate(names) if sys.argv[2] in n and (i == 0 or sys.argv[2] not in names[i - 1])]
if sys.argv[3] == "orb":  # pyr_down launches come in runs of 7: keep the first of each run
    starts = [i for k, i in enumerate(starts)]
print(starts[3])
PY
)
  echo "mode $MODE: launch-skip $SKIP, $PER launches" >> $O/profile_round.log
  $NCU --set full --import-source on --launch-skip $SKIP --launch-count $PER -o /tmp/full_$MODE -f $CMD > $O/ncu_full_$MODE.log 2>&1
  ncu -i /tmp/full_$MODE.ncu-rep --page raw --csv > $O/raw_full_$MODE.csv 2>> $O/profile_round.log
done
# (the two-view kernels: tools/profile_ransac.sh)
tail -3 $O/profile_round.log
