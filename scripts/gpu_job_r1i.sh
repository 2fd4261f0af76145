mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_train_gpu.py -m gpu -x -q -s 2>&1 | grep -E "d loss|pose_adj|passed|failed|Error|assert" | head -20
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
echo "=== bench train"
timeout 900 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-render 2>gpurun_out/bench_train.err > gpurun_out/bench_train_r1i.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_train_r1i.json'))
print({k:d[k] for k in ('value','ms_per_step')}, 'e2e', d['e2e']['value'], d['breakdown_ms'])
PY
