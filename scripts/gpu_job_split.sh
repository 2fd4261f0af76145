mkdir -p gpurun_out
for sp in 0 3 4; do
echo "=== TNF_FORWARD_SPLIT=$sp"
TNF_FORWARD_SPLIT=$sp timeout 600 python -m pytest tests/test_train_gpu.py -m gpu -x -q 2>&1 | tail -2
TNF_FORWARD_SPLIT=$sp timeout 600 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-render 2>/dev/null > gpurun_out/bench_split_$sp.json
python - <<PY
import json
d=json.load(open('gpurun_out/bench_split_$sp.json'))
print({k:d[k] for k in ('value','ms_per_step')}, d['breakdown_ms']['forward_ms'], 'e2e', d['e2e']['value'])
PY
done
