mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
echo "=== bench train"
timeout 900 python bench.py --steps 200 --warmup 20 2>gpurun_out/bench_train.err > gpurun_out/bench_train_r1f.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_train_r1f.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, 'e2e', d['e2e']['value'], d['breakdown_ms'], 'render', d.get('render',{}).get('value'), d['roofline']['kernel'], d['roofline']['frac'], d.get('cpu_baseline'))
PY
tail -3 gpurun_out/bench_train.err
echo "=== ncu full (all train kernels, one launch each from a late step)"
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:'tnf_(forward|backward_prop|backward_field|wgrad|adam|losses)' -s 60 -c 7 -o gpurun_out/train_kernels_r1f -f python bench.py --steps 14 --warmup 3 --no-cpu-baseline --no-render > gpurun_out/ncu_full_train.log 2>&1
tail -2 gpurun_out/ncu_full_train.log
ls -la gpurun_out | tail -5
