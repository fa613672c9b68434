for v in 0 1 2 3 4; do
echo "variant $v"
BN_BATCH_VARIANT=$v python tools/bench_configs.py --only cfg5,short 2>&1 | grep "encode_batch" | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('  ',d['kernel'][:24], d['ms'], d['frac_of_measured_peak'])"
done
