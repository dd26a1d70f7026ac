python -m pytest tests -m gpu -x -q -k "orb and not config" > gpurun_out/r02_t6_tests.log 2>&1; tail -5 gpurun_out/r02_t6_tests.log
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extra-legs 2>/dev/null | python -c "
import json,sys
d=[json.loads(l) for l in sys.stdin if l.startswith('{')][-1]
print('value', round(d['value']), 'ms', round(d['ms_per_step'],3), 'prof', round(d['stats']['ms_per_step_profiled_pass'],3), 'e2e', round(d['e2e']['value']))
for k in d['kernels']: print(k['kernel'], round(k['ms_per_step'],3), round(k['frac'],3), round(k['achieved'],1))"
