mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_surface_gpu.py tests/test_train_gpu.py -q 2>&1 | tail -15
echo "=== bench train"
timeout 600 python bench.py --no-render --no-cpu-baseline 2>gpurun_out/bench_train_c.err > gpurun_out/bench_train_c.json; tail -3 gpurun_out/bench_train_c.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_train_c.json'))
print('train', {k:d[k] for k in ('value','ms_per_step','gpu_launches')}, 'e2e', d['e2e']['value'], d['breakdown_ms'])
PY
echo "=== ncu full (field backward)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'tnf_backward_field' -s 6 -c 1 -o gpurun_out/bwd_field_r2c -f python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-render > gpurun_out/ncu_bwd_c.log 2>&1; tail -2 gpurun_out/ncu_bwd_c.log
