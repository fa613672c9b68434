set -x
N=${1:-2}
nproc; python -c "import os; print(len(os.sched_getaffinity(0)))"
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
grep bench gpurun_out/bench_n$N.err
python bench.py --steps 5 --warmup 3 --skip-cpu > gpurun_out/bench_n1b.json 2> gpurun_out/bench_n1b.err
