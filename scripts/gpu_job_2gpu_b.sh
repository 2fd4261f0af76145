mkdir -p gpurun_out
echo "=== peer tests"
timeout 600 python -m pytest tests/test_peer_gpu.py -m gpu -x -q 2>&1 | tail -15
for ex in peer nccl; do
echo "=== train N=2 exchange=$ex"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 10 --no-cpu-baseline --exchange $ex 2>gpurun_out/bench_train_n2_$ex.err > gpurun_out/bench_train_n2_$ex.json; tail -2 gpurun_out/bench_train_n2_$ex.err | cut -c1-300
python - <<PY
import json
d=json.load(open('gpurun_out/bench_train_n2_$ex.json'))
print({k:d[k] for k in ('value','ms_per_step','n_gpus','exchange','gpu_launches')}, 'e2e', d['e2e']['value'], d['config']['final_losses'])
PY
done
