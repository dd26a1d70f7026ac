python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_sharding.py -x -q 2>&1 | tail -4
for C in 0 125 100 50; do
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extra-legs --lane-chunk $C 2>/dev/null | python -c "
import json,sys
d=[json.loads(l) for l in sys.stdin if l.startswith('{')][-1]
print('chunk $C value', round(d['value']), 'ms', round(d['ms_per_step'],3), 'prof', round(d['stats']['ms_per_step_profiled_pass'],3), 'e2e', round(d['e2e']['value']), round(d['e2e']['ms_per_step'],2))"
done
SLAMCU_ONE_LANE=1 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extra-legs 2>/dev/null | python -c "
import json,sys
d=[json.loads(l) for l in sys.stdin if l.startswith('{')][-1]
print('one lane value', round(d['value']), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), round(d['e2e']['ms_per_step'],2))"
