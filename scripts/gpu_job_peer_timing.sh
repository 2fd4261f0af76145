N=${1:-2}
for mode in push pull; do
TNF_PEER_GATHER=$mode timeout 300 python -m pytest tests/test_peer_gpu.py -m gpu -q 2>&1 | tail -1
TNF_PEER_GATHER=$mode TNF_PEER_TIMING=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 60 --warmup 10 --no-cpu-baseline 2>/dev/null > gpurun_out/bench_peer_timing_n${N}_$mode.json
TNF_PEER_GATHER=$mode timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 60 --warmup 10 --no-cpu-baseline 2>/dev/null > gpurun_out/bench_peer_n${N}_$mode.json
python - <<PY
import json
d=json.load(open('gpurun_out/bench_peer_timing_n${N}_$mode.json'))
e=json.load(open('gpurun_out/bench_peer_n${N}_$mode.json'))
print("$mode", {k:e[k] for k in ('value','ms_per_step','n_gpus')}, d['exchange_phases_ms'])
PY
done
