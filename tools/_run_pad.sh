for pad in 0 1808 4400; do
  echo "pad $pad"; HB_DF_PAD=$pad timeout 200 python tools/store_bench.py 1100000 2504 2 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_frames'])"
done
