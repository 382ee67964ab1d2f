timeout 600 python -m pytest tests/test_store_gpu.py tests/test_convert_gpu.py -x -q -m gpu 2>&1 | tail -3
for pad in 0; do
HB_DF_PAD=$pad timeout 200 python tools/store_bench.py 1100000 2504 3 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('pad $pad ms_frames', d['ms_frames'], 'ratio', d['ratio'])"
done
