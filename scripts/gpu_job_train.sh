mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_train_gpu.py -m gpu -q --tb=short -s 2>&1 > gpurun_out/pytest_train.log; grep -E "rel-L2|^(FAILED|E  )|passed|failed" gpurun_out/pytest_train.log | cut -c1-220 | head -90
echo "=== bench train tc"; timeout 900 python bench.py --steps 100 --warmup 10 2>gpurun_out/bench_train.err | tee gpurun_out/bench_train.json | cut -c1-3000
tail -5 gpurun_out/bench_train.err
