# N GPUs: pipelined exchange (field slice under the next proposal pass) against the single-block exchange
mkdir -p gpurun_out
N=${N:-2}
if [ -z "$SKIP_TESTS" ]; then
timeout 900 python -m pytest tests/test_peer_gpu.py -q -s --tb=short 2>&1 | grep -v "^$" | tail -30
fi
run() {
  name=$1; shift
  echo "=== $name"
  env "$@" timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus $N --steps 200 --warmup 20 --no-cpu-baseline --no-torch-cuda-baseline --no-render $EXTRA 2>gpurun_out/p${N}_$name.err >gpurun_out/p${N}_$name.json
  tail -2 gpurun_out/p${N}_$name.err | cut -c1-300
  python - <<PY
import json
try:
    d=[json.loads(x) for x in open('gpurun_out/p${N}_$name.json') if x.startswith('{')][-1]
    print({k:d.get(k) for k in ('value','ms_per_step','param_checksum_all_ranks_equal','exchange_barrier_timeouts')}, 'e2e', (d.get('e2e') or {}).get('value'))
    print('  exchange:', str(d.get('exchange'))[:160])
except Exception as e:
    print('no json', e)
PY
}
for g in $GATHERS; do
run pipe_$g TNF_PEER_GATHER=$g TNF_PEER_PIPELINE=1
run block_$g TNF_PEER_GATHER=$g TNF_PEER_PIPELINE=0
done
if [ -n "$SINGLE" ]; then
for sp in 0 4 3; do
echo "=== single GPU TNF_FORWARD_SPLIT=$sp"
TNF_FORWARD_SPLIT=$sp timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline --no-torch-cuda-baseline --no-render 2>gpurun_out/p_single_$sp.err >gpurun_out/p_single_$sp.json
python -c "
import json
d=[json.loads(x) for x in open('gpurun_out/p_single_$sp.json') if x.startswith('{')][-1]
print(d['value'], d['ms_per_step'], d.get('breakdown_ms',{}).get('forward_ms'))"
done
fi
# knob sweep of the pipelined exchange: "name ENV=... ENV=..." per line in $SWEEP_FILE
if [ -n "$SWEEP_FILE" ]; then
while read -r name envs; do
  [ -z "$name" ] && continue
  run $name $envs
done < "$SWEEP_FILE"
fi
