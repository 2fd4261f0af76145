mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
echo "=== diag"
timeout 300 python scripts/diag_e2e.py 2>&1 | tail -14
echo "=== bench train"
timeout 900 python bench.py --steps 100 --warmup 10 2>gpurun_out/bench_train.err | tee gpurun_out/bench_train_r1c.json | cut -c1-2500
tail -3 gpurun_out/bench_train.err
