mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_camera_post_gpu.py tests/test_fullsize_gpu.py tests/test_render_gpu.py tests/test_train_gpu.py -m gpu -q --tb=short 2>&1 | grep -v "^$" | tail -150
