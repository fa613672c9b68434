set -x
python -m pytest tests -m gpu -x -q -k "batch or split" 2>&1 | tail -5
python tools/bench_configs.py --only cfg5,short > gpurun_out/configs_s12.txt 2>&1
tail -12 gpurun_out/configs_s12.txt
