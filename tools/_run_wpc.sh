timeout 600 python -m pytest tests/test_store_gpu.py tests/test_convert_gpu.py -x -q -m gpu 2>&1 | tail -2
for w in auto 4 8 12; do
  if [ $w = auto ]; then unset HB_DF_WPC; else export HB_DF_WPC=$w; fi
  timeout 200 python tools/store_bench.py 1100000 2504 3 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('wpc $w ms_frames', d['ms_frames'])"
done
