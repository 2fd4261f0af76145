mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_peer_gpu.py -m gpu -q 2>&1 | tail -2
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 10 2>gpurun_out/n2.err > gpurun_out/bench_train_n2_final.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_train_n2_final.json'))
print({k:d[k] for k in ('value','ms_per_step','n_gpus','exchange','gpu_launches')}, 'e2e', d['e2e']['value'])
PY
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 2>/dev/null | cut -c1-160
