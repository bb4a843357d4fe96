python tools/e2e_probe.py --nz 32 2>&1 | tail -7
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
