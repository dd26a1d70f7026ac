for T in 2048 1024 512 256; do
SLAMCU_MATCH_TILE=$T python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extra-legs 2>/dev/null | python -c "
import json,sys
d=[json.loads(l) for l in sys.stdin if l.startswith('{')][-1]
k=[k for k in d['kernels'] if k['kernel']=='match'][0]
print('tile $T value', round(d['value']), 'ms', round(d['ms_per_step'],3), 'match', round(k['ms_per_step'],3), round(k['achieved'],1), round(k['frac'],3))"
done
