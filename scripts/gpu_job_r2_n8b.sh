mkdir -p gpurun_out
TNF_TEST_ALL_GPUS=1 timeout 240 python -m pytest tests/test_peer_gpu.py -q -s --tb=short -k "auto" 2>&1 | grep -v "^$\|NCCL version" | tail -12
N=8 SKIP_TESTS=1 GATHERS="" SWEEP_FILE=scripts/sweep_n8.txt bash scripts/gpu_job_r2_pipe.sh
