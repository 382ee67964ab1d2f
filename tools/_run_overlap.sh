HB_DONOR_SPLIT=warp timeout 200 python tools/overlap_probe.py > gpurun_out/overlap_warp.log 2>&1
HB_DONOR_SPLIT=warp timeout 200 python tools/overlap_probe.py 1100000 2504 hi > gpurun_out/overlap_warp_hi.log 2>&1
timeout 200 python tools/overlap_probe.py > gpurun_out/overlap_lane.log 2>&1
tail -2 gpurun_out/overlap_warp.log gpurun_out/overlap_warp_hi.log gpurun_out/overlap_lane.log
