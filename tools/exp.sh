MPDAF_B200_LIBNAME=_ctools_tfinal python tools/run_case.py --n 48 --nz 64 --scaled --steps 1 --warmup 1 | grep -E "block (0|400) warp (0|13) |kernel" | tail -5 | cut -c1-300
