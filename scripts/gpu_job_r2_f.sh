mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_concat_gpu.py -q --tb=line 2>&1 | grep -v "^$" | tail -30
