timeout 500 python -m pytest tests/test_inflate_gpu.py tests/test_parse_gpu.py -x -q -m gpu 2>&1 | tail -3
timeout 500 python tools/e2e_breakdown.py 2>&1 | tail -1
