# Round 2, job A: tcgen05 probe, nerfstudio install probe, PyTorch-CUDA context baselines (never run in round 1)
mkdir -p gpurun_out
echo "=== probe_umma"
timeout 120 ./build/probe_umma 2>&1 | tee gpurun_out/probe_umma.txt
echo "=== nerfstudio install probe"
(timeout 40 python -m pip --retries 0 --timeout 5 download --no-deps -d /tmp/ns nerfstudio==1.1.5 2>&1 | tail -3) > gpurun_out/nerfstudio_probe.txt; cat gpurun_out/nerfstudio_probe.txt
python -c "import nerfstudio" 2>&1 | tail -1 | tee -a gpurun_out/nerfstudio_probe.txt
nproc | tee gpurun_out/nproc.txt
echo "=== context baselines"
bash scripts/gpu_job_context_baselines.sh 2>&1 | tail -12
