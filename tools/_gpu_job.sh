for v in 0 1 4 5 6; do
echo "variant $v"
BN_BATCH_VARIANT=$v timeout 300 python tools/bench_configs.py --only cfg5,short 2>&1 | grep "encode_batch\|rror" | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l); print('  ',d['kernel'][:24], d['ms'], d['frac_of_measured_peak'])
    except Exception: print(l[:200])"
done
BN_BATCH_VARIANT=1 timeout 300 python -m pytest tests -m gpu -x -q -k "batch" 2>&1 | tail -2
