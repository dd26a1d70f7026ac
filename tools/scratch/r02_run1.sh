set -x
python -m pytest tests/test_gpu_sharding.py tests/test_gpu_scale_parity.py -q -k "sharded or dense or counts_device or config3_full" > gpurun_out/r02_t4_tests.log 2>&1; tail -15 gpurun_out/r02_t4_tests.log
python bench.py --steps 5 --warmup 3 > gpurun_out/r02_b1.json 2> gpurun_out/r02_b1.err; tail -c 1500 gpurun_out/r02_b1.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_b1.json'))
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e'])
print('cpu', d.get('cpu_baseline'))
for k in ('mode_reference','config3'):
    x=d.get(k,{}); print(k, x.get('value'), x.get('e2e',{}).get('value'), x.get('cpu_baseline',{}).get('value'), x.get('error'))
print('sweep', [ (r['n'], round(r['kernel_gcmp_s'])) for r in d.get('hamming_sweep',{}).get('rows',[])], d.get('hamming_sweep',{}).get('error'))
PY
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_b1_ref.json 2>> gpurun_out/r02_b1.err; head -c 700 gpurun_out/r02_b1_ref.json
