#!/bin/bash
# One point of the scaling curve on an N-GPU box, launched exactly as the driver does:  bash tools/scale_run.sh N
# Writes gpurun_out/r02_scale_${N}gpu.json (our arm), r02_scale_${N}gpu_refarm.json (--impl reference) and, for N > 1,
# r02_xfer_probe_${N}gpu.json (copies only: the host-link ceiling of the end-to-end leg).
set -u
cd "${GRAFT_REPO_ROOT:-.}"
N=${1:-1}
O=gpurun_out
mkdir -p $O
if [ "$N" -eq 1 ]; then
  python bench.py --gpus 1 --steps 20 --warmup 5 --no-extra-legs 2> $O/scale_${N}.err | grep '^{' > $O/r02_scale_${N}gpu.json
  python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 2>> $O/scale_${N}.err | grep '^{' > $O/r02_scale_${N}gpu_refarm.json
else
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 2> $O/scale_${N}.err | grep '^{' > $O/r02_scale_${N}gpu.json
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 3 --warmup 1 2>> $O/scale_${N}.err | grep '^{' > $O/r02_scale_${N}gpu_refarm.json
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 tools/xfer_probe.py --ranks 2>> $O/scale_${N}.err | grep "^{" > $O/r02_xfer_probe_${N}gpu.json
fi
python - "$O/r02_scale_${N}gpu.json" "$O/r02_scale_${N}gpu_refarm.json" <<'PY'
import json, sys
b = json.loads(open(sys.argv[1]).read()); r = json.loads(open(sys.argv[2]).read())
e = b["e2e"]
print("N", b["n_gpus"], "resident", round(b["value"]), "ms", round(b["ms_per_step"], 3), "e2e", round(e["value"]), "link GB/s", round(e.get("host_link_gbs_all_ranks", 0), 1),
      "ref", round(r.get("value", 0), 1), "clocks", b.get("clocks"))
PY
tail -c 300 $O/scale_${N}.err
