mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 4 --steps 60 --warmup 10 2>gpurun_out/n4.err > gpurun_out/bench_train_n4_final.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_train_n4_final.json'))
print({k:d[k] for k in ('value','ms_per_step','n_gpus','exchange','gpu_launches')}, 'e2e', d['e2e']['value'])
PY
tail -2 gpurun_out/n4.err | cut -c1-200
