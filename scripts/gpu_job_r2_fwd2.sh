mkdir -p gpurun_out
for sp in 4 5; do
  echo "=== TNF_FORWARD_SPLIT=$sp"
  TNF_FORWARD_SPLIT=$sp timeout 300 python bench.py --mode render --steps 12 --warmup 3 --no-cpu-baseline --no-torch-cuda-baseline 2>gpurun_out/fwd2_$sp.err >gpurun_out/fwd2_$sp.json
  python -c "
import json
d=[json.loads(x) for x in open('gpurun_out/fwd2_$sp.json') if x.startswith('{')][-1]
print('render', d['value'], d['ms_per_step'], 'kernel_ms', d['roofline']['kernel_ms'], 'e2e', d['e2e']['value'])"
  TNF_FORWARD_SPLIT=$sp timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline --no-torch-cuda-baseline 2>gpurun_out/fwd2t_$sp.err >gpurun_out/fwd2t_$sp.json
  python -c "
import json
d=[json.loads(x) for x in open('gpurun_out/fwd2t_$sp.json') if x.startswith('{')][-1]
print('train', d['value'], d['ms_per_step'], d['breakdown_ms']['forward_ms'], 'quick render', d['render']['value'], d['render']['ms_per_frame'])"
done
