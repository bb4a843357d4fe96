run() { python tools/run_case.py --nz 64 $* 2>&1 | python -c "
import sys,json
t=sys.stdin.read()
try:
    d=json.loads(t.strip().splitlines()[-1]); print(d['kernel'], round(d['ms'],4), round(d['gbs']), d['rejected'], d['slow_voxels'], d.get('parity_first_planes'))
except Exception as e: print('ERR', t[-300:])"; }
for n in 4 12 24 48 64 96 128; do run --n $n --scaled --check; done
run --n 48 --check; run --n 48 --scaled --var stat_mean --check; run --n 48 --scaled --mad --check; run --n 12 --median --check
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -8
