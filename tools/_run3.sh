rm -f gpurun_out/s4_cases.jsonl
for fl in 2 16 32 48 64 96 192; do
echo "flags $fl" | tee -a gpurun_out/s4_cases.jsonl
MPDAF_B200_PAIR_FLAGS=$fl timeout 300 python tools/run_case.py --n 48 --nz 64 --scaled --var propagate --check --steps 5 --warmup 2 2>&1 | tail -1 | cut -c1-120 | tee -a gpurun_out/s4_cases.jsonl
done
for fl in 2 48 64; do
echo "flags $fl" | tee -a gpurun_out/s4_cases.jsonl
MPDAF_B200_PAIR_FLAGS=$fl timeout 300 python tools/run_case.py --n 24 --nz 128 --scaled --var propagate --check --steps 5 --warmup 2 2>&1 | tail -1 | cut -c1-120 | tee -a gpurun_out/s4_cases.jsonl
MPDAF_B200_PAIR_FLAGS=$fl timeout 300 python tools/run_case.py --n 48 --nz 32 --scaled --var propagate --f64 --steps 5 --warmup 2 2>&1 | tail -1 | cut -c1-120 | tee -a gpurun_out/s4_cases.jsonl
done
timeout 300 python tools/run_case.py --n 64 --nz 48 --scaled --var stat_mean --check --steps 5 --warmup 2 2>&1 | tail -1 | cut -c1-120 | tee -a gpurun_out/s4_cases.jsonl
