timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -4
python bench.py > gpurun_out/bench_r01f.json 2> gpurun_out/bench_r01f.err; python -c "
import json; d=json.load(open('gpurun_out/bench_r01f.json')); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['parity']['ok'], d['cpu_baseline']['value'])"
