mkdir -p gpurun_out
nvidia-smi -L | wc -l
N=${1:-8}
echo "=== train N=$N exchange=peer"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 60 --warmup 10 --no-cpu-baseline 2>gpurun_out/bench_train_n$N.err > gpurun_out/bench_train_n$N.json; tail -2 gpurun_out/bench_train_n$N.err | cut -c1-300
python - <<PY
import json
d=json.load(open('gpurun_out/bench_train_n$N.json'))
print({k:d[k] for k in ('value','ms_per_step','n_gpus','exchange','gpu_launches')}, 'e2e', d['e2e']['value'])
PY
echo "=== render N=$N"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --mode render --steps 6 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench_render_n$N.err > gpurun_out/bench_render_n$N.json; tail -2 gpurun_out/bench_render_n$N.err | cut -c1-300
python - <<PY
import json
d=json.load(open('gpurun_out/bench_render_n$N.json'))
print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, 'e2e', d['e2e']['value'])
PY
