P=$((20000 + RANDOM % 20000))
for N in 8 4; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((P+N)) bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r02_scale_${N}gpu.out 2> gpurun_out/r02_scale_${N}gpu.err; echo rc=$?
grep '^{' gpurun_out/r02_scale_${N}gpu.out > gpurun_out/r02_scale_${N}gpu.json
python - <<PY
import json
d=json.load(open('gpurun_out/r02_scale_${N}gpu.json'))
print($N, 'value', round(d['value']), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), round(d['e2e']['ms_per_step'],2), 'link GB/s', round(d['e2e']['host_link_gbs_all_ranks'],1), d['stats']['numa'])
PY
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((P+N+20)) tools/xfer_probe.py --ranks 2>/dev/null | grep '^{' > gpurun_out/r02_xfer_probe_${N}gpu.json; cat gpurun_out/r02_xfer_probe_${N}gpu.json
done
nvidia-smi topo -m > gpurun_out/r02_topo_8gpu.txt 2>&1; lscpu | grep -E "Model name|^CPU\(s\)|NUMA|Socket" >> gpurun_out/r02_topo_8gpu.txt; head -12 gpurun_out/r02_topo_8gpu.txt
