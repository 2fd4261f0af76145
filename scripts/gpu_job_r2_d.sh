mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -25
echo "=== bench default"
timeout 900 python bench.py 2>gpurun_out/bench_d.err > gpurun_out/bench_d.json; grep -E "Elapsed|Error|error" gpurun_out/bench_d.err | tail -5
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_d.json'))
print('train', {k:d[k] for k in ('value','ms_per_step','gpu_launches')}, 'e2e', d['e2e']['value'])
print(d['breakdown_ms'])
print('roofline', {k:d['roofline'][k] for k in ('kernel','frac','achieved','kernel_ms')})
print('render', {k:v for k,v in d['render'].items() if k in ('value','ms_per_frame')}, d['render']['roofline']['frac'], d['render']['e2e']['value'], d['render'].get('torch_cuda_baseline'))
print('torch', json.dumps(d.get('torch_cuda_baseline'))[:900])
print('cpu', d.get('cpu_baseline'))
PY
