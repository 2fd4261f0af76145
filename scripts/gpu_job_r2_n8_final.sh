# 8 GPUs, round end: the default line (train + render) and the 8192-rays-per-GPU shape of configs[3]
mkdir -p gpurun_out
N=8
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus $N --steps 300 --warmup 30 > gpurun_out/r2_bench_train_n8_final.json 2> gpurun_out/n8f.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29556 bench.py --gpus $N --steps 200 --warmup 20 --rays 8192 --no-render > gpurun_out/r2_bench_train_n8_8192.json 2> gpurun_out/n8f2.err
python - <<'PY'
import json
for f in ("r2_bench_train_n8_final","r2_bench_train_n8_8192"):
    try:
        d=[json.loads(x) for x in open(f"gpurun_out/{f}.json") if x.startswith("{")][-1]
        print(f, d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "render", (d.get("render") or {}).get("value"), (d.get("render") or {}).get("ms_per_frame"), d["param_checksum_all_ranks_equal"], d["exchange_barrier_timeouts"])
        print("   ranks", [(round(r["ms_per_step"],4), r["sm_mhz"], r["reasons"]) for r in d["ranks"]])
    except Exception as e:
        print(f, "failed", e)
PY
tail -3 gpurun_out/n8f.err | cut -c1-200
