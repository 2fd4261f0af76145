mkdir -p gpurun_out
echo "=== ncu launch list (train)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:tnf_ -c 160 --csv --log-file gpurun_out/launches_train.csv python bench.py --steps 16 --warmup 3 --no-cpu-baseline --no-render > gpurun_out/ncu_train.log 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open('gpurun_out/launches_train.csv')) if len(r) > 10 and r[0].isdigit()]
agg = collections.defaultdict(list)
for r in rows:
    agg[r[4].split('(')[0][:60]].append(float(r[-1]))
for k, v in agg.items():
    print(f"{k:62s} n={len(v):3d} avg={sum(v)/len(v)/1e3:9.1f} us  min={min(v)/1e3:9.1f} max={max(v)/1e3:9.1f}")
PY
echo "=== ncu full (train kernels)"
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:'tnf_(forward|backward_prop|backward_field|wgrad|adam|losses)' -s 60 -c 7 -o gpurun_out/train_kernels_r1j -f python bench.py --steps 14 --warmup 3 --no-cpu-baseline --no-render > gpurun_out/ncu_full_train.log 2>&1
tail -1 gpurun_out/ncu_full_train.log
