python -m pytest tests -m gpu -x -q -k "match or knn or config5 or config4 or sequence or pipeline or orb_parity or sharded" > gpurun_out/r02_t5_tests.log 2>&1; tail -5 gpurun_out/r02_t5_tests.log
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extra-legs > gpurun_out/r02_b3.json 2> gpurun_out/r02_b3.err; tail -c 500 gpurun_out/r02_b3.err
python - <<'PY'
import json
d=[json.loads(l) for l in open('gpurun_out/r02_b3.json') if l.startswith('{')][-1]
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'])
for k in d['kernels']: print(k['kernel'], round(k['ms_per_step'],3), round(k['frac'],3), round(k['achieved'],1))
PY
python tools/hamming_sweep.py --quick | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print([(r['n'], round(r['kernel_gcmp_s']), r['spot_check']) for r in d['rows']])"
