# Round-end measurement on one B200: GPU tests, smoke, both bench modes, 8192-ray config, reference arm, ncu launch list + full captures.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "=== bench train (default)"
timeout 900 python bench.py 2>gpurun_out/bench_train.err > gpurun_out/r2_bench_train_final.json; tail -2 gpurun_out/bench_train.err
echo "=== bench render"
timeout 900 python bench.py --mode render --steps 20 --warmup 3 2>gpurun_out/bench_render.err > gpurun_out/r2_bench_render_final.json; tail -2 gpurun_out/bench_render.err
echo "=== bench train 8192 rays"
timeout 600 python bench.py --rays 8192 --no-cpu-baseline --no-torch-cuda-baseline --no-render 2>/dev/null > gpurun_out/r2_bench_train_8192.json
echo "=== reference arm"
timeout 600 python bench.py --impl reference 2>/dev/null > gpurun_out/r2_bench_reference_final.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_train_final.json'))
print('train', {k:d[k] for k in ('value','ms_per_step','gpu_launches')}, 'e2e', d['e2e']['value'], 'plugin', d['e2e_plugin_autograd']['value'])
print('  ', d['breakdown_ms'])
print('  roofline', d['roofline']['kernel'], round(d['roofline']['frac'],3), 'traffic', d['roofline'].get('traffic'), 'cpu', d['cpu_baseline']['value'])
print('  torch', {k:(v['value'] if isinstance(v,dict) else v) for k,v in d['torch_cuda_baseline'].items() if k in ('fp32','fp16_autocast_gradscaler','ratio')})
r=d['render']; print('  render', r['value'], r['ms_per_frame'], r['roofline']['frac'], r['e2e']['value'], r.get('torch_cuda_baseline',{}).get('ratio'))
d=json.load(open('gpurun_out/r2_bench_render_final.json'))
print('render', {k:d[k] for k in ('value','ms_per_step')}, 'e2e', d['e2e']['value'], d['e2e_ray_bundle']['value'], round(d['roofline']['frac'],3), d['roofline']['kernel_ms'])
d=json.load(open('gpurun_out/r2_bench_train_8192.json'))
print('train 8192', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'])
d=json.load(open('gpurun_out/r2_bench_reference_final.json'))
print('reference', d['value'], d['cpu_baseline']['cores'])
PY
echo "=== ncu launch list (train)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:tnf_ -c 200 --csv --log-file gpurun_out/r2_launches_train_final.csv python bench.py --steps 16 --warmup 3 --no-cpu-baseline --no-render --no-torch-cuda-baseline > gpurun_out/ncu_train.log 2>&1
echo "=== ncu full (train kernels)"
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:'tnf_(forward|backward_prop|backward_field|adam|losses)' -s 72 -c 8 -o gpurun_out/r2_train_kernels_final -f python bench.py --steps 14 --warmup 3 --no-cpu-baseline --no-render --no-torch-cuda-baseline > gpurun_out/ncu_full_train.log 2>&1
echo "=== ncu full (render forward, 640000 rays: proposal launch + field launch)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tnf_forward -s 8 -c 2 -o gpurun_out/r2_render_forward_final -f python bench.py --mode render --steps 2 --warmup 3 --no-cpu-baseline --no-torch-cuda-baseline > gpurun_out/ncu_full_render.log 2>&1
ls -la gpurun_out | grep r2_
