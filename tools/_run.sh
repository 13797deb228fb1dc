python tools/stage_bench.py 40 640 480 dprefetch 2>&1 | head -1
python tools/stage_bench.py 40 640 480 dprefetch 2>&1 | head -1
python tools/stage_bench.py 20 1280 720 c3 0.002 0x80000 2>&1 | head -1
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
