# Round 2, job B: first run of the tcgen05 weight-gradient backward
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_train_gpu.py -x -q 2>&1 | tail -15
echo "=== bench train"
timeout 600 python bench.py --no-render 2>gpurun_out/bench_train_b.err > gpurun_out/bench_train_b.json; tail -3 gpurun_out/bench_train_b.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_train_b.json'))
print('train', {k:d[k] for k in ('value','ms_per_step','gpu_launches')}, 'e2e', d['e2e']['value'], d['breakdown_ms'])
PY
