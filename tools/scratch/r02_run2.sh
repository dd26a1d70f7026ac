P=$((20000 + RANDOM % 20000))
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $P bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02_b2_2gpu.json 2> gpurun_out/r02_b2_2gpu.err; echo rc=$?; tail -c 1500 gpurun_out/r02_b2_2gpu.err; wc -c gpurun_out/r02_b2_2gpu.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_b2_2gpu.json'))
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['host_link_gbs_all_ranks'])
print(d['stats'])
PY
