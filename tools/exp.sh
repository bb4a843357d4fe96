for f in 229 5 1; do
MPDAF_B200_LIBNAME=_ctools_abl MPDAF_B200_WARP_FLAGS=$f ncu --set full --clock-control none --import-source on -k regex:sigclip_warp -c 1 -o gpurun_out/abl_f$f -f python tools/run_case.py --n 48 --nz 64 --scaled --steps 1 --warmup 0 > gpurun_out/ncu_abl_$f.log 2>&1; tail -2 gpurun_out/ncu_abl_$f.log
done
