mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
echo "=== bench tc_fp16"; timeout 600 python bench.py --steps 20 --warmup 3 --torch-cuda-baseline 2>gpurun_out/bench_tc.err | tee gpurun_out/bench_tc.json
echo "=== bench fp32"; timeout 600 python bench.py --steps 5 --warmup 3 --precision fp32 --no-cpu-baseline 2>gpurun_out/bench_fp32.err | tee gpurun_out/bench_fp32.json
echo "=== ncu launches"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
tail -3 gpurun_out/launches_r1.csv
echo "=== ncu full"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tnf_forward -s 3 -c 1 -o gpurun_out/fwd_tc_r1 -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
