set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/hbm_probe.py
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_s13.json 2> gpurun_out/bench_s13.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref_s13.json 2>> gpurun_out/bench_s13.err
python tools/bench_configs.py --only cfg2,cfg3,cfg4,cfg5,short > gpurun_out/configs_s13.txt 2>&1
python -c "import __graft_entry__ as g; g.smoke()"
