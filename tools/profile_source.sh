#!/bin/bash
# Source-level ncu pages (per-line executed instructions and stall samples) of the level-0 launches of the three
# heaviest extractor kernels and of the matcher; results land in gpurun_out/src_<kernel>.csv
set -u
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out
CMD="python bench.py --frames 256 --steps 1 --warmup 3 --no-cpu-baseline"
for K in fast9_mask:24 blur7:24 pyr_down:21 match_kernel:3 orb_describe:3; do
  NAME=${K%%:*}; SKIP=${K##*:}
  ncu --clock-control none --set full --import-source on -k regex:$NAME --launch-skip $SKIP --launch-count 1 -o /tmp/src_$NAME -f $CMD > $O/ncu_src_$NAME.log 2>&1
  ncu -i /tmp/src_$NAME.ncu-rep --page source --csv > $O/src_$NAME.csv 2>> $O/ncu_src_$NAME.log
  wc -l $O/src_$NAME.csv
done
