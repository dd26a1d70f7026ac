P=$((20000 + RANDOM % 20000))
for N in 2 4 8; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((P+N)) bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r02_scale_${N}gpu.out 2> gpurun_out/r02_scale_${N}gpu.err; echo rc=$?
grep '^{' gpurun_out/r02_scale_${N}gpu.out > gpurun_out/r02_scale_${N}gpu.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((P+N+20)) tools/xfer_probe.py --ranks 2>/dev/null | grep '^{' > gpurun_out/r02_xfer_probe_${N}gpu.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((P+N+40)) bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/r02_scale_${N}gpu_refarm.json 2>/dev/null
done
python bench.py --gpus 1 --steps 5 --warmup 3 --no-extra-legs > gpurun_out/r02_scale_1gpu.out 2>/dev/null; grep '^{' gpurun_out/r02_scale_1gpu.out > gpurun_out/r02_scale_1gpu.json
python tools/xfer_probe.py --ranks 2>/dev/null | grep '^{' > gpurun_out/r02_xfer_probe_1gpu.json
python - <<'PY'
import json
for N in (1,2,4,8):
    d=json.load(open(f'gpurun_out/r02_scale_{N}gpu.json')); x=json.load(open(f'gpurun_out/r02_xfer_probe_{N}gpu.json'))
    print(N, 'value', round(d['value']), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), round(d['e2e']['ms_per_step'],2), 'link', round(d['e2e']['host_link_gbs_all_ranks'],1), '| probe ceiling fps', round(x['frames_per_s_ceiling']), 'GB/s', round(x['aggregate_gbs'],1))
PY
