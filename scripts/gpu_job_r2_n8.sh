# 8 GPUs: exchange flavours side by side (phases timed), 8192-ray config, render, single-GPU reference on the same box
mkdir -p gpurun_out
N=${N:-8}
nvidia-smi --query-gpu=name --format=csv,noheader | sort | uniq -c
TNF_TEST_ALL_GPUS=1 timeout 300 python -m pytest tests/test_peer_gpu.py -q -s --tb=short 2>&1 | grep -v "^$" | tail -12
run() {  # name, env..., -- bench args
  name=$1; shift
  echo "=== $name"
  env "$@" timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus $N --steps 200 --warmup 20 --no-cpu-baseline --no-torch-cuda-baseline $EXTRA 2>gpurun_out/n${N}_$name.err >gpurun_out/n${N}_$name.json
  tail -2 gpurun_out/n${N}_$name.err | cut -c1-300
  python - <<PY
import json
try:
    d=[json.loads(x) for x in open('gpurun_out/n${N}_$name.json') if x.startswith('{')][-1]
    print({k:d.get(k) for k in ('value','unit','ms_per_step','exchange_phases_ms','param_checksum_all_ranks_equal','exchange_barrier_timeouts')}, 'e2e', (d.get('e2e') or {}).get('value'))
    print('  exchange:', d.get('exchange'), ' clocks', d.get('clocks'))
except Exception as e:
    print('no json', e)
PY
}
EXTRA="--no-render"
run multimem TNF_PEER_GATHER=multimem TNF_PEER_TIMING=1
run push TNF_PEER_GATHER=push TNF_PEER_TIMING=1
run pull TNF_PEER_GATHER=pull TNF_PEER_TIMING=1
run auto_notiming TNF_PEER_GATHER=auto
EXTRA="--no-render --rays 8192"
run auto_8192 TNF_PEER_GATHER=auto
EXTRA="--mode render"
run render TNF_PEER_GATHER=auto
echo "=== single GPU on this box"
timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline --no-torch-cuda-baseline --no-render 2>gpurun_out/n${N}_single.err >gpurun_out/n${N}_single.json
python -c "
import json
d=[json.loads(x) for x in open('gpurun_out/n${N}_single.json') if x.startswith('{')][-1]
print(d['value'], d['ms_per_step'])"
