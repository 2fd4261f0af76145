"""Diagnostic: where does the host time of the plugin-API training step go?"""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
import bench
from thermo_nerf_b200 import FusedAdam, RayBundle
import thermo_nerf_b200.functional as F

device = torch.device("cuda", 0)
R = 4096
model = bench.build_b200_model(device, "tc_fp16")
model.train()
groups = model.get_param_groups()
opts = [FusedAdam(groups["proposal_networks"], lr=1e-2, eps=1e-15), FusedAdam(groups["fields"], lr=1e-2, eps=1e-15)]
host = bench.train_batches(8, R, device, 0, pin=True)
T = {}
def tick(name, t0):
    t1 = time.perf_counter(); T[name] = T.get(name, 0.0) + (t1 - t0); return t1

def step(i, sync_each=False):
    t = time.perf_counter()
    ho, hd, hc, hrgb, hth = host[i % 8]
    rb = RayBundle(origins=ho.to(device, non_blocking=True), directions=hd.to(device, non_blocking=True),
                   camera_indices=hc.to(device, non_blocking=True).view(-1, 1))
    batch = {"image": hrgb.to(device, non_blocking=True), "thermal": hth.view(-1, 1)}
    for o in opts: o.zero_grad()
    t = tick("h2d+zero_grad", t)
    out = model(rb)
    if sync_each: torch.cuda.synchronize()
    t = tick("forward", t)
    metrics = model.get_metrics_dict(out, batch); ld = model.get_loss_dict(out, batch, metrics); loss = sum(ld.values())
    if sync_each: torch.cuda.synchronize()
    t = tick("losses", t)
    loss.backward()
    if sync_each: torch.cuda.synchronize()
    t = tick("backward", t)
    for o in opts: o.step()
    if sync_each: torch.cuda.synchronize()
    t = tick("adam", t)
    torch.cuda.synchronize()
    t = tick("final sync", t)

for i in range(10): step(i)
for sync_each in (False, True):
    T.clear()
    s0 = torch.cuda.memory_stats()
    for i in range(40): step(10 + i, sync_each)
    s1 = torch.cuda.memory_stats()
    print("sync_each", sync_each, {k: round(v / 40 * 1e3, 3) for k, v in T.items()}, "ms/step")
    for k in ("num_device_alloc", "num_device_free", "num_alloc_retries", "num_sync_all_streams"):
        print("   ", k, s1.get(k, 0) - s0.get(k, 0))
print("reserved MiB", torch.cuda.memory_reserved() >> 20, "allocated MiB", torch.cuda.memory_allocated() >> 20)
# isolate the multiply
g = torch.empty((R, 256), device=device); u = torch.ones((), device=device)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(100): y = g * u
t1 = time.perf_counter(); torch.cuda.synchronize()
print("g*u host time per call (us):", (t1 - t0) / 100 * 1e6)
