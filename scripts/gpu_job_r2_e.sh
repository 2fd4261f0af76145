mkdir -p gpurun_out
timeout 600 python scripts/profile_e2e.py > gpurun_out/profile_e2e_r2.txt 2>&1; head -70 gpurun_out/profile_e2e_r2.txt
