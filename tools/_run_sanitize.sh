timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_store_gpu.py -x -q -m gpu -k "random_vcf or degenerate or tiny or lane_per_frame" > gpurun_out/sanitize_store.log 2>&1
echo "rc=$?"; grep -c "Invalid\|out of bounds" gpurun_out/sanitize_store.log; tail -5 gpurun_out/sanitize_store.log
