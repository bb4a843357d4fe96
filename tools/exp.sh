python tools/e2e_probe.py --nz 32 2>&1 | tail -4
MPDAF_B200_LIBNAME=_ctools python bench.py --steps 3 --warmup 3 --nz 128 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('e2e', d['e2e']['value'], 'dropin', d['dropin_files'], 'parity', d['parity']['ok'])"
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
