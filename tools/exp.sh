run() { python tools/run_case.py $* 2>&1 | python -c "
import sys,json
t=sys.stdin.read()
try:
    d=json.loads(t.strip().splitlines()[-1]); print(d['kernel'], d['nz'], round(d['ms'],4), round(d['gbs']), d['rejected'], d['slow_voxels'], d.get('parity_first_planes'))
except Exception as e: print('ERR', t[-300:])"; }
for f in 1 33 32 0; do echo flags=$f; MPDAF_B200_WARP_FLAGS=$f run --n 48 --nz 64 --scaled --check; done
for f in 0 32 33; do echo N=12 flags=$f; MPDAF_B200_WARP_FLAGS=$f run --n 12 --nz 64 --scaled --check; done
for f in 1 33; do echo N=24 flags=$f; MPDAF_B200_WARP_FLAGS=$f run --n 24 --nz 64 --scaled --check; done
for f in 1 33; do echo N=96 flags=$f; MPDAF_B200_WARP_FLAGS=$f run --n 96 --nz 64 --scaled --check; done
