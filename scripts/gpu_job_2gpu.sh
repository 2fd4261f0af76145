mkdir -p gpurun_out
nvidia-smi -L
echo "=== train N=2"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 60 --warmup 10 --no-cpu-baseline 2>gpurun_out/bench_train_n2.err > gpurun_out/bench_train_n2.json; tail -2 gpurun_out/bench_train_n2.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_train_n2.json'))
print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, 'e2e', d['e2e']['value'])
PY
echo "=== render N=2"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --mode render --steps 6 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench_render_n2.err > gpurun_out/bench_render_n2.json; tail -2 gpurun_out/bench_render_n2.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_render_n2.json'))
print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, 'e2e', d['e2e']['value'])
PY
echo "=== reference arm N=2 (rank 0 only)"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 2>/dev/null | cut -c1-300
