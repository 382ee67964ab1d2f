#!/bin/bash
# ncu --set full capture of kernels 3 and 4b at the bench shape (1.1 M x 2,504, chunk 1,075); run under gpurun.
# usage: tools/ncu_k4.sh <tag>     -> gpurun_out/<tag>_plain.log, <tag>_k4b.ncu-rep, <tag>_k3.ncu-rep
tag=${1:-r02}
python tools/store_bench.py 1100000 2504 3 > gpurun_out/${tag}_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:donor_frames -s 1 -c 1 -o gpurun_out/${tag}_k4b -f \
    python tools/store_bench.py 1100000 2504 2 > gpurun_out/${tag}_ncu_k4b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:decode_gt -s 0 -c 1 -o gpurun_out/${tag}_k3 -f \
    python tools/store_bench.py 1100000 2504 1 > gpurun_out/${tag}_ncu_k3.log 2>&1
cat gpurun_out/${tag}_plain.log
