mkdir -p gpurun_out
for p in tc_fp16 fp32; do timeout 600 python scripts/quality_synthetic.py --steps 3000 --precision $p 2>&1 | tail -1 | tee gpurun_out/quality_$p.json; done
