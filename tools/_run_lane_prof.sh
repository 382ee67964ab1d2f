set -e
timeout 200 python tools/store_bench.py 1100000 2504 3 > gpurun_out/sb_lane.log 2>&1
timeout 120 python tools/store_bench.py 200000 2504 2 > gpurun_out/sb_lane_small.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"donor_frames_lane|pack_alleles" -c 2 -o gpurun_out/prof_lane -f python tools/store_bench.py 200000 2504 1 > gpurun_out/ncu_lane.log 2>&1
tail -3 gpurun_out/sb_lane.log
