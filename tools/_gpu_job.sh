python -m pytest tests -m gpu -x -q -k "batch" 2>&1 | tail -3
python tools/bench_cfg5_e2e.py --scale 0.0625 2>&1 | tail -1 | cut -c1-700
