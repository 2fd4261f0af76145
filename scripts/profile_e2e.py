"""Host-side profile of one training iteration through the plugin API (the `e2e` leg of bench.py):
cProfile over N steps + CPU-only launch time of TrainEngine.step.  Run on a GPU box."""
import cProfile
import io
import pstats
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch  # noqa: E402

import bench  # noqa: E402
from thermo_nerf_b200 import FusedAdam, RayBundle  # noqa: E402
from thermo_nerf_b200.engine import TrainEngine  # noqa: E402

device = torch.device("cuda", 0)
R = 4096
model = bench.build_b200_model(device, "tc_fp16")
model.train()
groups = model.get_param_groups()
opts = [FusedAdam(groups["proposal_networks"], lr=1e-2, eps=1e-15), FusedAdam(groups["fields"], lr=1e-2, eps=1e-15)]
cbs = model.get_training_callbacks()
host = bench.train_batches(8, R, device, 0, pin=True)
loss_host = torch.empty(1, dtype=torch.float32).pin_memory()


def e2e_step(i):
    for c in cbs:
        if c.where_to_run == ["BEFORE_TRAIN_ITERATION"]:
            c.run_callback(i)
    ho, hd, hc, hrgb, hth = host[i % 8]
    rb = RayBundle(origins=ho.to(device, non_blocking=True), directions=hd.to(device, non_blocking=True),
                   camera_indices=hc.to(device, non_blocking=True).view(-1, 1))
    batch = {"image": hrgb.to(device, non_blocking=True), "thermal": hth.view(-1, 1)}
    for o in opts:
        o.zero_grad()
    out = model(rb)
    metrics = model.get_metrics_dict(out, batch)
    ld = model.get_loss_dict(out, batch, metrics)
    loss = sum(ld.values())
    loss.backward()
    for o in opts:
        o.step()
    for c in cbs:
        if c.where_to_run == ["AFTER_TRAIN_ITERATION"]:
            c.run_callback(i)
    loss_host.copy_(loss.detach().reshape(1), non_blocking=True)
    torch.cuda.current_stream().synchronize()


for i in range(10):
    e2e_step(i)
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(40):
    e2e_step(10 + i)
print(f"e2e plugin step: {(time.perf_counter() - t0) / 40 * 1e3:.3f} ms/step")
pr = cProfile.Profile()
pr.enable()
for i in range(40):
    e2e_step(50 + i)
pr.disable()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(40)
print(s.getvalue()[:7000])
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(30)
print(s.getvalue()[:6000])

engine = TrainEngine(model, world_size=1)
dev_batches = bench.train_batches(8, R, device, 0)
for i in range(10):
    engine.step(*dev_batches[i % 8])
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(50):
    engine.step(*dev_batches[i % 8])
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"engine.step: host launch {(t1 - t0) / 50 * 1e3:.3f} ms/step, with device drain {(t2 - t0) / 50 * 1e3:.3f} ms/step")
