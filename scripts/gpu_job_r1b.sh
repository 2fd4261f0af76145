mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
echo "=== e2e host profile"
timeout 600 python scripts/profile_e2e.py > gpurun_out/profile_e2e.txt 2>&1; head -70 gpurun_out/profile_e2e.txt
echo "=== ncu full (train kernels)"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'tnf_(forward|backward_prop|backward_field)' -s 9 -c 3 -o gpurun_out/train_kernels_r1b -f python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-render > gpurun_out/ncu_full_train.log 2>&1
tail -3 gpurun_out/ncu_full_train.log
ls -la gpurun_out
