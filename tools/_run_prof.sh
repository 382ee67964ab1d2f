timeout 120 python tools/store_bench.py 200000 2504 2 > gpurun_out/sb_small.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"donor_frames_kernel" -c 1 -o gpurun_out/prof_df2 -f python tools/store_bench.py 200000 2504 1 > gpurun_out/ncu_df2.log 2>&1
tail -1 gpurun_out/sb_small.log
