N=${1:-8}
for mode in pull; do
TNF_PEER_GATHER=$mode TNF_PEER_TIMING=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 60 --warmup 10 --no-cpu-baseline 2>gpurun_out/pull.err > gpurun_out/bench_peer_timing_n${N}_${mode}2.json
python - <<PY
import json
d=json.load(open('gpurun_out/bench_peer_timing_n${N}_${mode}2.json'))
print("$mode", {k:d[k] for k in ('value','ms_per_step','n_gpus')}, d['exchange_phases_ms'], d['config']['final_losses'])
PY
done
tail -2 gpurun_out/pull.err | cut -c1-200
