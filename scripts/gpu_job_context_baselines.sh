# One B200: the PyTorch-CUDA context numbers for north_star's ">= 10x the reference PyTorch-CUDA path" target
# (the oracle port run eagerly in fp32 on the GPU), next to the product's own numbers.  Not part of the default bench.
mkdir -p gpurun_out
echo "=== train + torch-cuda training baseline"
timeout 900 python bench.py --torch-cuda-baseline --no-render 2>gpurun_out/ctx_train.err > gpurun_out/ctx_train.json; tail -2 gpurun_out/ctx_train.err
echo "=== render + torch-cuda render baseline"
timeout 900 python bench.py --mode render --steps 20 --warmup 3 --torch-cuda-baseline 2>gpurun_out/ctx_render.err > gpurun_out/ctx_render.json; tail -2 gpurun_out/ctx_render.err
python - <<'PY'
import json
for name in ("train", "render"):
    d = json.load(open(f"gpurun_out/ctx_{name}.json"))
    t = d.get("torch_cuda_baseline", {})
    print(name, "ours", d["value"], d["unit"], "| torch-cuda", t.get("value"), t.get("unit"), t.get("error"),
          "| ratio", (d["value"] / t["value"]) if t.get("value") else None)
PY
