mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
echo "=== bench train"
timeout 900 python bench.py --steps 100 --warmup 10 --no-cpu-baseline 2>gpurun_out/bench_train.err > gpurun_out/bench_train_r1h.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_train_r1h.json'))
print({k:d[k] for k in ('value','ms_per_step')}, 'e2e', d['e2e']['value'], d['breakdown_ms'], 'render', d.get('render',{}).get('value'))
PY
echo "=== bench render"
timeout 900 python bench.py --mode render --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench_render.err > gpurun_out/bench_render_r1h.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_render_r1h.json'))
print({k:d[k] for k in ('value','ms_per_step')}, 'e2e', d['e2e']['value'], d['e2e_ray_bundle']['value'], d['roofline']['frac'], d['roofline']['kernel_ms'])
PY
tail -3 gpurun_out/bench_render.err
