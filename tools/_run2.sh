python tools/pair_check.py > gpurun_out/s3_pair_check.txt 2>&1; tail -1 gpurun_out/s3_pair_check.txt
rm -f gpurun_out/s3_cases.jsonl
for fl in 2 1 0; do
MPDAF_B200_PAIR_FLAGS=$fl timeout 300 python tools/run_case.py --n 48 --nz 64 --scaled --var propagate --check --steps 5 --warmup 2 2>&1 | tail -1 | tee -a gpurun_out/s3_cases.jsonl
done
for args in "--n 48 --nz 64 --scaled --var stat_mean --check" "--n 24 --nz 128 --scaled --var propagate --check" "--n 48 --nz 64 --scaled --var propagate --mad --check" "--n 48 --nz 32 --scaled --var propagate --f64 --check" "--n 64 --nz 48 --scaled --var stat_mean --check" "--n 32 --nz 96 --scaled --var propagate --check" "--n 12 --nz 256 --scaled --var propagate --check"; do
  timeout 300 python tools/run_case.py $args --steps 5 --warmup 2 2>&1 | tail -1 | tee -a gpurun_out/s3_cases.jsonl
done
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/s3_pytest.txt 2>&1; tail -3 gpurun_out/s3_pytest.txt
