# forward experiments on one GPU: paired 16-byte loads (variant library) x eval split
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_render_gpu.py tests/test_golden_gpu.py tests/test_fullsize_gpu.py -m gpu -q --tb=short 2>&1 | grep -v "^$" | tail -8
echo "--- same tests, paired-load library"
TNF_B200_LIB=$PWD/build/exp/libtnf_pair.so timeout 600 python -m pytest tests/test_render_gpu.py tests/test_golden_gpu.py tests/test_fullsize_gpu.py tests/test_train_gpu.py -m gpu -q --tb=short 2>&1 | grep -v "^$" | tail -8
for lib in default pair; do for es in 1 0; do
  echo "=== lib=$lib TNF_EVAL_SPLIT=$es"
  if [ $lib = pair ]; then export TNF_B200_LIB=$PWD/build/exp/libtnf_pair.so; else unset TNF_B200_LIB; fi
  TNF_EVAL_SPLIT=$es timeout 300 python bench.py --mode render --steps 12 --warmup 3 --no-cpu-baseline --no-torch-cuda-baseline 2>gpurun_out/fwd_${lib}_$es.err >gpurun_out/fwd_${lib}_$es.json
  python -c "
import json
d=[json.loads(x) for x in open('gpurun_out/fwd_${lib}_$es.json') if x.startswith('{')][-1]
print('render', d['value'], d['ms_per_step'], 'kernel_ms', d['roofline']['kernel_ms'], 'e2e', d['e2e']['value'])"
done
  TNF_EVAL_SPLIT=1 timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline --no-torch-cuda-baseline --no-render 2>gpurun_out/fwdt_${lib}.err >gpurun_out/fwdt_${lib}.json
  python -c "
import json
d=[json.loads(x) for x in open('gpurun_out/fwdt_${lib}.json') if x.startswith('{')][-1]
print('train', d['value'], d['ms_per_step'], d['breakdown_ms'])" | cut -c1-330
done
