# 2 GPUs: fused exchange tests (IPC + symmetric memory / multicast), concat tests, 2-GPU bench lines
mkdir -p gpurun_out
nvidia-smi topo -m 2>/dev/null | head -6
timeout 600 python -m pytest tests/test_peer_gpu.py tests/test_concat_gpu.py -q -s --tb=short 2>&1 | grep -v "^$" | tail -40
for mode in auto ipc; do
echo "=== bench 2 GPUs backend=$mode"
TNF_PEER_BACKEND=$mode TNF_PEER_TIMING=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 10 --no-render 2>gpurun_out/bench_n2_$mode.err > gpurun_out/bench_n2_$mode.json; tail -2 gpurun_out/bench_n2_$mode.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_n2_$mode.json'))
print({k:d.get(k) for k in ('value','ms_per_step','exchange','exchange_phases_ms','param_checksum_all_ranks_equal','exchange_barrier_timeouts')}, 'e2e', d['e2e']['value'])
PY
done
