mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
echo "=== bench train"
timeout 900 python bench.py --steps 100 --warmup 10 --no-cpu-baseline 2>gpurun_out/bench_train.err | tee gpurun_out/bench_train_r1d.json | cut -c1-400
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_train_r1d.json'))
print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['breakdown_ms'], d.get('render',{}).get('value'))
PY
tail -3 gpurun_out/bench_train.err
bash scripts/gpu_job_launches.sh
