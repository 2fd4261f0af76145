mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
echo "=== diag"; timeout 200 python scripts/diag_e2e.py 2>&1 | grep sync_each
echo "=== bench train"
timeout 900 python bench.py --steps 100 --warmup 10 --no-cpu-baseline 2>gpurun_out/bench_train.err > gpurun_out/bench_train_r1e.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_train_r1e.json'))
print({k:d[k] for k in ('value','ms_per_step')}, 'e2e', d['e2e']['value'], d['breakdown_ms'], 'render', d.get('render',{}).get('value'))
PY
tail -3 gpurun_out/bench_train.err
