#!/usr/bin/env python
"""Benchmark of the ThermoNeRF volumetric-render hot path on B200.

    python bench.py --gpus N --steps K --warmup W            # our arm (libtnf_b200.so)
    python bench.py --impl reference --gpus N --steps K ...  # reference arm: the path's own
                                                             # PyTorch implementation (oracle port)
                                                             # on the box's host cores

One "step" = one pass of the hot path over one batch of synthetic ThermoScenes-shaped rays:
  --mode train  : (default) one full training iteration at 4096 rays/batch per GPU
                  (BASELINE.json configs[1]; metric training rays/s = world * rays / iteration time).
                  The line also carries `render` (configs[4] shape: 800x800 frames, Mpix/s, roofline, e2e - on every
                  rank when N > 1), `torch_cuda_baseline` (the reference's PyTorch path on the same GPU, fp32 and fp16
                  autocast + GradScaler), `rays_8192` (configs[2] / the per-GPU shape of configs[3], with its own
                  PyTorch-CUDA baseline), `cpu_baseline`, and for N > 1 `ranks`, `param_checksum_all_ranks_equal`,
                  `exchange_barrier_timeouts`.
  --mode render : one 800x800 frame (640 000 rays) through get_outputs_for_camera_ray_bundle
                  (BASELINE.json configs[4]; metric render Mpix/s, 1 ray = 1 pixel)
  --rays 8192   : the full contract run at 8192 rays per batch per GPU
Weights are random "trained-like" (no datasets/checkpoints offline); data is synthetic.
Prints ONE JSON line on rank 0 (contract in the task statement).
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402

HW = 800
FOCAL = 1111.1
NUM_IMAGES = 100
ALGO_BYTES_PER_RAY = 161_876  # fp32 hash-gather bytes (161 792) + ray I/O (84)
ALGO_FLOP_PER_RAY = 1.709e6


def env_int(name: str, default: int) -> int:
    return int(os.environ.get(name, default))


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi sampled DURING the timed region (B200_PROFILING.md clocks line)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int) -> None:
        self.gpu_index = gpu_index
        self.proc = None
        self.path = None

    def start(self) -> None:
        try:
            f = tempfile.NamedTemporaryFile(prefix="tnf_clocks_", suffix=".csv", delete=False)
            self.path = f.name
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.gpu_index)], stdout=f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self) -> dict:
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        try:
            for line in open(self.path):
                p = [x.strip() for x in line.split(",")]
                if len(p) < 9:
                    continue
                try:
                    sm.append(float(p[1]))
                    mx.append(float(p[2]))
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            sm.sort()
            out.update(sm_mhz=sm[len(sm) // 2], sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


# ----------------------------------------------------------------------------- model / data
def randomise_trained_like(model, seed: int = 0) -> None:
    """Synthetic 'trained-like' weights: N(0, 0.5^2) hash entries and sharpened density heads so
    that the proposal PDFs are peaked (samples cluster as they do around real surfaces)."""
    g = torch.Generator().manual_seed(seed + 1234)
    with torch.no_grad():
        encs = [model.field.mlp_base.encoder] + [p.encoding for p in model.proposal_networks]
        for enc in encs:
            enc.hash_table.copy_(torch.randn(enc.hash_table.shape, generator=g) * 0.5)
        for p in model.proposal_networks:
            p.mlp_base[1].layers[1].weight.mul_(6.0)
        model.field.mlp_base.mlp.layers[1].weight[0].mul_(6.0)


def build_b200_model(device, precision: str, camera_optimizer_mode: str = "off"):
    """Camera optimiser off: pose refinement stays in PyTorch upstream of the path and is not part of
    the measured workload (the oracle arm is configured the same way, oracle_train_setup)."""
    from thermo_nerf_b200 import ThermalNerfModel, ThermalNerfModelConfig

    torch.manual_seed(0)
    cfg = ThermalNerfModelConfig(precision=precision, camera_optimizer_mode=camera_optimizer_mode)
    model = ThermalNerfModel(cfg, {"thermal": []}, torch.tensor([[-1.0, -1, -1], [1, 1, 1]]), NUM_IMAGES)
    randomise_trained_like(model, 0)
    return model.to(device).eval()


def frame_bundles(n_frames: int, device, rank: int, world: int):
    """``n_frames`` distinct orbit views per rank.  Frames shard with no collective (rank r renders frames r, r+world,
    ... of a trajectory, thermo_nerf_b200.dist.shard_frames); for the scaling measurement every rank gets the SAME
    views as the single-GPU run: the cost of a frame depends on the view (measured 21.6 - 28 ms over an orbit of the
    synthetic scene), so striping different views over the ranks would report the scene's load imbalance through the
    max over ranks rather than anything about the system."""
    from thermo_nerf_b200 import orbit_cameras

    cams = orbit_cameras(max(n_frames, 1), hw=HW, focal=FOCAL, device=device)
    return [cams.generate_rays(i) for i in range(n_frames)]


class L2Flusher:
    def __init__(self, device, nbytes: int = 256 << 20) -> None:
        self.buf = torch.empty(nbytes // 4, dtype=torch.float32, device=device)

    def __call__(self) -> None:
        self.buf.fill_(1.0)


# ----------------------------------------------------------------------------- synthetic training data
def train_batches(n_batches: int, rays: int, device, rank: int, pin: bool = False):
    """ThermoScenes-shaped random pixel batches: 100 pinhole cameras 800x800 on a sphere,
    uniform (image, y, x) draws seeded per rank (nerfstudio DDP: every rank draws its own rays), analytic GT."""
    from thermo_nerf_b200 import sphere_cameras
    from thermo_nerf_b200.dist import rank_seed

    cams = sphere_cameras(NUM_IMAGES, hw=HW, focal=FOCAL)
    out = []
    for b in range(n_batches):
        g = torch.Generator().manual_seed(rank_seed(77, rank, b))
        cam = torch.randint(0, NUM_IMAGES, (rays,), generator=g)
        ys = torch.randint(0, HW, (rays,), generator=g)
        xs = torch.randint(0, HW, (rays,), generator=g)
        r = cams.generate_pixel_rays(cam, ys, xs)
        gt_rgb = (0.5 + 0.4 * torch.sin(r.directions * 7.0)).float()
        gt_th = (0.5 + 0.4 * torch.cos(r.directions[:, :1] * 5.0 + r.directions[:, 1:2] * 3.0)).float()
        gt_th = gt_th + 0.01 * torch.rand(gt_th.shape, generator=g)
        t = (r.origins, r.directions, r.camera_indices.reshape(-1), gt_rgb.contiguous(), gt_th.reshape(-1).contiguous())
        out.append(tuple(x.pin_memory() for x in t) if pin else tuple(x.to(device) for x in t))
    return out


KERNEL_NAMES = {"forward": "tnf_forward_kernel", "backward_prop": "tnf_backward_prop_kernel",
                "backward_field": "tnf_backward_field_kernel_tc", "wgrad": "tnf_wgrad_kernel_fp32",
                "adam": "tnf_adam_kernel"}

TRAIN_WORKLOAD = ("thermal-nerf training iteration, ThermoScenes double_robot-shaped synthetic rays: {rays} rays/batch per "
                  "GPU (BASELINE configs[1]), samples 256/96/48, forward + losses (rgb, thermal, interlevel, distortion) + "
                  "backward + Adam(lr 1e-2, eps 1e-15, exp. decay) over all 19.4M parameters; each rank draws its own "
                  "rays; gradient mean over ranks fused with Adam over NVLink peer memory (--exchange nccl: NCCL "
                  "all-reduce, then Adam)")


# ----------------------------------------------------------------------------- reference arm
def oracle_train_setup(rays: int):
    from oracle import OracleConfig, OracleThermalNerf

    model = OracleThermalNerf(OracleConfig(camera_optimizer_mode="off"), NUM_IMAGES, seed=0)
    randomise_trained_like(model, 0)
    field = [p for n, p in model.named_parameters() if n.startswith("field.")]
    props = [p for n, p in model.named_parameters() if n.startswith("proposal_networks.")]
    opts = [torch.optim.Adam(field, lr=1e-2, eps=1e-15), torch.optim.Adam(props, lr=1e-2, eps=1e-15)]
    batch = train_batches(1, rays, "cpu", 0)[0]
    return model, opts, batch


class OracleSchedule:
    """ProposalNetworkSampler's update schedule (thermal_nerf_model.py:152-161): the proposal networks carry
    gradients - and their optimiser steps - only on "updated" iterations, exactly as in the product's TrainEngine."""

    def __init__(self) -> None:
        self.steps_since_update = 0

    def updated(self, step: int) -> bool:
        import numpy as np

        prev = max(step - 1, 0)  # the sampler's _step is set by step_cb after an iteration: it lags by one
        sched = float(np.clip(np.interp(prev, [0, 5000], [0, 5]), 1, 5))
        return self.steps_since_update > sched or prev < 10


def oracle_train_step(model, opts, batch, step: int, sched: "OracleSchedule" = None, scaler=None) -> float:
    """One full iteration of the PyTorch restatement: forward, 4 losses, autograd backward, torch Adam.  `scaler`
    (a torch GradScaler) switches on the reference's mixed precision: fp16 autocast + GradScaler
    (config_thermal_nerf.py:22 -> nerfstudio Trainer)."""
    from oracle import OracleRays

    o, d, cam, gt_rgb, gt_th = batch
    updated = True if sched is None else sched.updated(step)
    model.set_anneal_for_step(step)
    for op in opts:
        op.zero_grad()
    with torch.autocast(device_type=o.device.type, dtype=torch.float16, enabled=scaler is not None):
        out = model.get_outputs(OracleRays(o, d, cam.reshape(-1, 1)), training=True, prop_grad=updated)
        ld = model.get_loss_dict(out, gt_rgb, gt_th.reshape(-1, 1), training=True)
        loss = sum(ld.values())
    field_opt, prop_opt = opts
    if scaler is not None:
        scaler.scale(loss).backward()
        scaler.step(field_opt)
        if updated:
            scaler.step(prop_opt)
        scaler.update()
    else:
        loss.backward()
        field_opt.step()
        if updated:
            prop_opt.step()
    if sched is not None:
        if updated:
            sched.steps_since_update = 0
        sched.steps_since_update += 1
    return float(loss.detach())


def run_reference(args, rank: int, world: int, emit) -> None:
    """The path's own PyTorch implementation (nerfstudio torch semantics; oracle port since
    nerfstudio is not installable offline) on the host cores, bounded sample per step."""
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    if args.mode == "train":
        sample = args.ref_rays
        model, opts, batch = oracle_train_setup(sample)
        sched = OracleSchedule()
        for i in range(args.warmup):
            oracle_train_step(model, opts, batch, i, sched)
        t0 = time.perf_counter()
        for i in range(args.steps):
            oracle_train_step(model, opts, batch, args.warmup + i, sched)
        dt = time.perf_counter() - t0
        val = sample * args.steps / dt
        desc = (f"{sample}-ray batch per step (full iteration: forward, 4 losses, autograd backward, torch Adam; proposal "
                f"networks updated on the sampler's schedule), fp32, torch {torch.__version__} on {cores} host threads")
        line = {"impl": "reference", "metric": "train_rays_per_s", "value": val, "unit": "rays/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
                "config": {"workload": TRAIN_WORKLOAD.format(rays=args.rays), "sample": desc},
                "cpu_baseline": {"value": val, "unit": "rays/s", "cores": cores, "kind": "port", "sample": desc},
                "e2e": {"value": val, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        emit(line)
        return
    from oracle import OracleConfig, OracleThermalNerf, make_synthetic_rays

    sample = args.ref_rays
    model = OracleThermalNerf(OracleConfig(), NUM_IMAGES, seed=0)
    randomise_trained_like(model, 0)
    rays = make_synthetic_rays(sample, num_images=NUM_IMAGES, seed=1, contiguous_pixels=True)
    with torch.no_grad():
        for _ in range(args.warmup):
            model.get_outputs(rays, training=False)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            model.get_outputs(rays, training=False)
        dt = time.perf_counter() - t0
    mpix = sample * args.steps / dt / 1e6
    desc = f"{sample} contiguous rays of an 800x800 frame per step, eval forward, fp32, torch {torch.__version__}"
    line = {
        "impl": "reference", "metric": "render_mpix_per_s", "value": mpix, "unit": "Mpix/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
        "config": {"workload": "render 800x800 ThermoScenes-shaped frame, eval chunk 65536, samples 256/96/48",
                   "sample": desc},
        "cpu_baseline": {"value": mpix, "unit": "Mpix/s", "cores": cores, "kind": "port", "sample": desc},
        "e2e": {"value": mpix, "unit": "Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


def hbm_peak():
    peaks_path = ROOT / "MEASURED_PEAKS.json"
    if peaks_path.exists():
        try:
            return float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


# ----------------------------------------------------------------------------- our arm
def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--mode", default="train", choices=["train", "render"])
    ap.add_argument("--precision", default="tc_fp16", choices=["tc_fp16", "fp32"])
    ap.add_argument("--rays", type=int, default=4096, help="training rays per batch per GPU")
    ap.add_argument("--ref-rays", type=int, default=None, help="rays per step of the CPU reference sample")
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"],
                    help="multi-GPU gradient exchange: 'peer' = reduce-scatter + Adam + all-gather fused in one "
                         "kernel over NVLink peer memory (default), 'nccl' = NCCL all-reduce, then Adam")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-render", action="store_true", help="train mode: skip the short render measurement")
    ap.add_argument("--no-torch-cuda-baseline", action="store_true",
                    help="skip the PyTorch-CUDA baseline (the restatement run eagerly on the same GPU, fp32 and fp16 "
                         "autocast + GradScaler: north_star's >= 10x target)")
    ap.add_argument("--torch-cuda-baseline", action="store_true", help=argparse.SUPPRESS)  # the default now
    args = ap.parse_args()
    ref = args.impl == "reference"
    if args.steps is None:
        args.steps = (3 if ref else 200) if args.mode == "train" else (3 if ref else 20)
    if args.warmup is None:
        args.warmup = 1 if ref else (20 if args.mode == "train" else 3)
    if args.ref_rays is None:
        args.ref_rays = 4096
    if not ref:
        args.warmup = max(args.warmup, 3)

    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    # stdout carries exactly one JSON line: anything a library prints there (NCCL's version banner does)
    # is sent to stderr by pointing fd 1 at fd 2 for the duration of the run
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    def emit(line: dict) -> None:
        os.write(json_fd, (json.dumps(line) + "\n").encode())

    if ref:
        run_reference(args, rank, world, emit)
        return

    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py (impl b200) needs a CUDA device: there is no CPU fallback for the product path")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def gather_ranks(obj):
        """[obj of rank 0, obj of rank 1, ...] on every rank (small python objects: per-rank time and clocks)."""
        if world == 1:
            return [obj]
        out = [None] * world
        dist.all_gather_object(out, obj)
        return out

    ctx = dict(args=args, rank=rank, world=world, local=local, device=device, barrier=barrier,
               max_over_ranks=max_over_ranks, gather_ranks=gather_ranks)
    line = bench_train(ctx) if args.mode == "train" else bench_render(ctx)
    if rank == 0:
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def bench_train(ctx) -> dict:
    """BASELINE configs[1]: full training iterations at `--rays` rays per batch per GPU."""
    args, rank, world, device = ctx["args"], ctx["rank"], ctx["world"], ctx["device"]
    barrier, max_over_ranks = ctx["barrier"], ctx["max_over_ranks"]
    from thermo_nerf_b200 import FusedAdam, ModelTensors, RayBundle
    from thermo_nerf_b200 import functional as F
    from thermo_nerf_b200.dist import allreduce_mean_grads_
    from thermo_nerf_b200.engine import TrainEngine

    R = args.rays
    model = build_b200_model(device, args.precision)
    model.train()
    try:
        engine = TrainEngine(model, world_size=world, peer_fused=(world > 1 and args.exchange == "peer"))
    except RuntimeError as e:
        # CUDA IPC / peer access not available between these processes (PeerArena reports it on every rank at once):
        # run the NCCL exchange instead and say so in the output line
        if not (world > 1 and args.exchange == "peer" and "PeerArena" in str(e)):
            raise
        print(f"[bench] {e}; falling back to --exchange nccl", file=sys.stderr, flush=True)
        model = build_b200_model(device, args.precision)
        model.train()
        engine = TrainEngine(model, world_size=world, peer_fused=False)
    n_distinct = 8
    batches = train_batches(n_distinct, R, device, rank)

    # ---- device-resident throughput ("value"): inputs already in HBM
    for i in range(args.warmup):
        engine.step(*batches[i % n_distinct])
    barrier()
    if engine.arena is not None and os.environ.get("TNF_PEER_TIMING"):
        engine.arena.timing = []  # per-phase events of the exchange (slightly perturbs the step: diagnostic runs only)
    sampler = ClockSampler(ctx["local"])
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    prop_steps_before = engine.prop_steps
    e0.record()
    for i in range(args.steps):
        losses = engine.step(*batches[i % n_distinct])
    e1.record()
    barrier()
    prop_steps_timed = engine.prop_steps - prop_steps_before
    clocks = sampler.stop()
    my_ms = e0.elapsed_time(e1) / args.steps
    ms_per_step = max_over_ranks(e0.elapsed_time(e1)) / args.steps
    ranks = ctx["gather_ranks"]({"ms_per_step": my_ms, "sm_mhz": clocks["sm_mhz"], "reasons": clocks["reasons"]})
    value = world * R / (ms_per_step * 1e-3)
    final_losses = [float(x) for x in losses.tolist()]

    # ---- per-kernel breakdown of one iteration (CUDA events between the launches, same stream)
    breakdown = kernel_breakdown(engine, batches, device) if rank == 0 else {}
    roofline = None
    if breakdown:
        peak, peak_src = hbm_peak()
        n_params = sum(p.numel() for p in engine.params)
        # algorithmic bytes per launch: 8-byte table cells, 8 corners per level lookup
        algo = {
            "forward": R * ALGO_BYTES_PER_RAY,
            # re-gather of the two proposal levels + scatter (read-modify-write) of their table gradients
            "backward_prop": R * (256 + 96) * 5 * 8 * 8 * 3,
            # saved fp16 features (64 B) + field outputs (20 B) in, table-gradient scatter as read-modify-write
            # (16 levels x 8 corners x 8 B x 2); the weight gradients never leave the SM (tensor memory)
            "backward_field": R * 48 * (64 + 20 + 16 * 8 * 8 * 2),
            "adam": n_params * 28,
        }
        if args.precision == "fp32":  # exact mode: (X, dY) rows staged in HBM for a separate weight-gradient pass
            algo["backward_field"] = R * 48 * (128 + 20 + 16 * 8 * 8 * 2 + 2752)
            algo["wgrad"] = R * 48 * (2752 + 128) + R * (48 + 64) * 4
        algo = {k: v for k, v in algo.items() if breakdown.get(k + "_ms")}
        dom = max(algo, key=lambda k: breakdown.get(k + "_ms", 0.0))
        ach = algo[dom] / (breakdown[dom + "_ms"] * 1e-3) / 1e9
        traffic = None
        tp = ROOT / "profiles" / "train_traffic.json"
        if tp.exists():
            try:
                traffic = json.load(open(tp)).get(dom, {}).get("dram_bytes_per_launch")
            except Exception:
                traffic = None
        roofline = {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": traffic,
                    "kernel": KERNEL_NAMES[dom], "kernel_ms": breakdown[dom + "_ms"], "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": algo[dom],
                    "note": "algorithmic bytes per launch as defined in DESIGN.md section 4; the hash tables and "
                            "their gradients are L2-resident within a launch, so DRAM traffic sits below the "
                            "algorithmic gather/scatter bytes",
                    "all_kernels": {KERNEL_NAMES[k]: {"ms": breakdown[k + "_ms"],
                                                      "algorithmic_GBps": algo[k] / (breakdown[k + "_ms"] * 1e-3) / 1e9}
                                    for k in algo if breakdown.get(k + "_ms")}}

    # ---- end to end with HOST buffers, every step: pinned host batch -> H2D -> full iteration -> D2H of the loss,
    #      host synchronised (the trainer reads the loss).  Two routes over the same kernels:
    #      (1) TrainEngine.step_host - the package's train-iteration call (what Trainer.train_iteration does around
    #          the reference model, pipeline_tracking.py:47-59), C-ABI launches without autograd; for world > 1 it
    #          uses the same peer-memory exchange as the device-resident loop.  This is `e2e`.
    #      (2) the nerfstudio plugin surface under autograd: model(ray_bundle) -> get_metrics_dict -> get_loss_dict
    #          -> loss.backward() -> FusedAdam.step (NCCL all-reduce of the gradients for world > 1).  Reported as
    #          `e2e_plugin_autograd`; its step is bound by PyTorch's per-call host time, not by the kernels.
    host = train_batches(n_distinct, R, device, rank, pin=True)
    h2d = sum(t.numel() * t.element_size() for t in host[0])
    loss_slots = [torch.empty(4, dtype=torch.float32).pin_memory() for _ in range(2)]
    loss_events = [torch.cuda.Event(), torch.cuda.Event()]
    seen = []

    def engine_e2e_step(i: int) -> None:
        # step i: H2D of its batch, the iteration, D2H of its 4 losses - all enqueued; then the host reads the
        # losses of step i-1, which has finished by now or finishes while step i runs (one-step lag, as an
        # asynchronous logger reads them): every step's losses reach the host and are read inside the timed region
        ls = engine.step_host(*host[i % n_distinct])
        loss_slots[i & 1].copy_(ls, non_blocking=True)
        loss_events[i & 1].record()
        if i > 0:
            loss_events[(i - 1) & 1].synchronize()
            seen.append(float(loss_slots[(i - 1) & 1][0]))

    def drain(i_last: int) -> None:
        loss_events[i_last & 1].synchronize()
        seen.append(float(loss_slots[i_last & 1][0]))

    n_e2e = max(args.steps // 2, 10)
    n_warm = max(args.warmup // 2, 3)
    for i in range(n_warm):
        engine_e2e_step(i)
    drain(n_warm - 1)
    barrier()
    seen.clear()
    t0 = time.perf_counter()
    for i in range(n_e2e):
        engine_e2e_step(i)
    drain(n_e2e - 1)
    torch.cuda.synchronize()
    t_e2e = max_over_ranks(time.perf_counter() - t0)
    barrier()
    assert len(seen) == n_e2e  # the first timed call also reads a (stale) slot: i == 0 reads nothing, drain reads the last
    e2e_val = world * R * n_e2e / t_e2e
    checksum = param_checksum_all_ranks_equal(engine, world, device)

    model2 = build_b200_model(device, args.precision)
    model2.train()
    groups = model2.get_param_groups()
    opts = [FusedAdam(groups["proposal_networks"], lr=1e-2, eps=1e-15), FusedAdam(groups["fields"], lr=1e-2, eps=1e-15)]
    cbs = model2.get_training_callbacks()
    all_params = ModelTensors.from_module(model2).param_list()  # fixed order = the backward's gradient arena order
    loss1_host = torch.empty(1, dtype=torch.float32).pin_memory()

    def plugin_e2e_step(i: int) -> None:
        for c in cbs:
            if c.where_to_run == ["BEFORE_TRAIN_ITERATION"]:
                c.run_callback(i)
        ho, hd, hc, hrgb, hth = host[i % n_distinct]
        rb = RayBundle(origins=ho.to(device, non_blocking=True), directions=hd.to(device, non_blocking=True),
                       camera_indices=hc.to(device, non_blocking=True).view(-1, 1))
        batch = {"image": hrgb.to(device, non_blocking=True), "thermal": hth.view(-1, 1)}  # thermal GT stays on the
        for o in opts:                                                                       # host until the loss
            o.zero_grad()
        out = model2(rb)
        metrics = model2.get_metrics_dict(out, batch)
        ld = model2.get_loss_dict(out, batch, metrics)
        loss = sum(ld.values())
        loss.backward()
        if world > 1:
            allreduce_mean_grads_(all_params, world_size=world)
        for o in opts:
            o.step()
        for c in cbs:
            if c.where_to_run == ["AFTER_TRAIN_ITERATION"]:
                c.run_callback(i)
        loss1_host.copy_(loss.detach().reshape(1), non_blocking=True)
        torch.cuda.current_stream().synchronize()

    n_plugin = max(args.steps // 8, 10)
    for i in range(3):
        plugin_e2e_step(i)
    barrier()
    t0 = time.perf_counter()
    for i in range(n_plugin):
        plugin_e2e_step(10 + i)
    torch.cuda.synchronize()
    t_plugin = max_over_ranks(time.perf_counter() - t0)
    barrier()
    plugin_val = world * R * n_plugin / t_plugin

    line = {
        "metric": "train_rays_per_s", "value": value, "unit": "rays/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None,
        "dtype": ("fp32 hashing/sampling/compositing/Adam + fp16-operand forward and bf16-operand backward mma "
                  "(fp32 accumulate) in the field MLPs" if args.precision == "tc_fp16" else "fp32"),
        "data": "synthetic",
        "config": {"workload": TRAIN_WORKLOAD.format(rays=R), "rays_per_batch_per_gpu": R,
                   "l2": "working set per step (tables + gradients + Adam state = 296 MiB) exceeds the 126 MB L2; "
                         f"{n_distinct} distinct ray batches cycle",
                   "weights": "random trained-like init, full-size tables (field 2^19x16, proposals 2^17x5)",
                   "final_losses": dict(zip(F.LOSS_NAMES, final_losses))},
        "clocks": clocks, "ranks": ranks,
        "exchange": ("none (1 GPU)" if world == 1 else
                     ({"push": "peer-memory fused reduce-scatter + Adam + all-gather (tnf_peer_adam_step)",
                       "pull": "peer-memory fused reduce-scatter + Adam (tnf_peer_adam_reduce), pull all-gather "
                               "(tnf_peer_gather_params)",
                       "multimem": "NVSwitch multicast: in-switch reduction (multimem.ld_reduce) + Adam + broadcast "
                                   "(multimem.st) in one kernel (tnf_peer_adam_multimem)"}[engine.arena.gather]
                      + (", arenas mapped by torch symmetric memory" if engine.arena._symm is not None
                         else ", arenas mapped by CUDA IPC")
                      if engine.arena is not None else "NCCL all-reduce (mean) + tnf_adam_step")),
        "gpu_launches": None,
        "exchange_phases_ms": engine.arena.timing_summary() if engine.arena is not None else None,
        # barrier waits that gave up because a peer never arrived (must be 0; a non-zero count invalidates the run)
        "exchange_barrier_timeouts": engine.arena.timeouts() if engine.arena is not None else None,
        # every rank's parameter arena hashed after the timed steps: DDP semantics require bit-identical replicas
        "param_checksum_all_ranks_equal": checksum,
        "e2e": {"value": e2e_val, "unit": "rays/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 16, "steps": n_e2e,
                "note": "TrainEngine.step_host: pinned host rays + GT -> H2D -> forward, losses, backward, (peer exchange,) "
                        "Adam through the C ABI -> D2H of the 4 losses; the host reads every step's losses inside the "
                        "timed region, one step behind the launch (the read of step i-1 overlaps step i)"},
        "e2e_plugin_autograd": {
            "value": plugin_val, "unit": "rays/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4, "steps": n_plugin,
            "note": "nerfstudio plugin surface under autograd: pinned host rays+GT -> H2D -> model(ray_bundle) -> "
                    "get_metrics_dict -> get_loss_dict -> loss.backward() -> FusedAdam.step -> D2H loss; bound by "
                    "PyTorch's per-call host time (autograd engine, tensor bookkeeping), not by the kernels"},
    }
    # launches of OUR kernels per engine step: forward (proposal launch + field launch in tensor-core mode, one fused
    # launch in fp32 mode or with TNF_FORWARD_SPLIT=0) + clip + losses + backward_field (+ the fp32 mode's
    # weight-gradient pass) + adam, and backward_prop + a second adam launch on update steps
    # (memsets and, on the NCCL path, the all-reduce kernel are not ours)
    fwd = 2 if (args.precision == "tc_fp16" and os.environ.get("TNF_FORWARD_SPLIT", "4") != "0") else 1
    per_step = fwd + 3 + (1 if args.precision == "fp32" else 0)
    if engine.arena is not None and engine.pipelined:
        # field slice: 2 barrier kernels + exchange kernel every step; proposal slice on update steps: backward_prop +
        # exchange kernel + closing barrier, + its opening barrier when the field slice started early (otherwise that
        # barrier is the field slice's too)
        line["gpu_launches"] = int(args.steps * (per_step + 3) + prop_steps_timed * (4 if engine._early else 3))
        line["exchange"] += (" - pipelined (tnf_peer_adam_range): proposal slice on the critical path, field slice on a "
                             f"side stream ({engine._side_ctas} CTAs) started "
                             + ("under the proposal backward" if engine._early else "after the backward")
                             + ", next field level waits for it (tnf_render_forward_staged)")
    elif engine.arena is not None:  # unpipelined peer exchange: 2 barrier kernels + fused Adam (+ gather kernel when pulling)
        line["gpu_launches"] = int(args.steps * (per_step + (4 if engine.arena.gather == "pull" else 3)) + prop_steps_timed)
    else:
        line["gpu_launches"] = int(args.steps * (per_step + 1) + prop_steps_timed * 2)
    if roofline:
        line["roofline"] = roofline
        line["breakdown_ms"] = breakdown
    if not args.no_render:
        # every rank renders the same views (frames shard with no collective); aggregate = world * pixels / slowest rank
        r = quick_render(model, device, world == 1 and not args.no_torch_cuda_baseline, ctx)
        if rank == 0:
            line["render"] = r
    if rank == 0 and world == 1 and not args.no_torch_cuda_baseline:
        del engine, model2, opts
        torch.cuda.empty_cache()
        line["torch_cuda_baseline"] = torch_cuda_train_baseline(device, args.rays, value)
        if args.rays != 8192:
            # BASELINE configs[2] / the per-GPU shape of configs[3]: ours and the PyTorch-CUDA baseline at 8192 rays
            line["rays_8192"] = train_at_8192(device, args.precision)
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline_train(min(args.ref_rays, 4096))
    return line


def param_checksum_all_ranks_equal(engine, world: int, device):
    """True when every rank holds bit-identical parameters after the timed steps (None on one GPU): a 64-bit sum of
    the parameter arena's bit patterns per rank, gathered and compared."""
    if world == 1:
        return None
    import torch.distributed as dist

    with torch.no_grad():
        acc = torch.zeros((), dtype=torch.int64, device=device)
        for p in engine.params:
            acc = acc + p.detach().contiguous().view(torch.int32).to(torch.int64).sum()
        mine = acc.reshape(1)
    allv = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(allv, mine)
    return bool(all(int(v.item()) == int(allv[0].item()) for v in allv))


def torch_cuda_train_baseline(device, rays: int, ours_rays_per_s: float) -> dict:
    """north_star's target: ">= 10x the reference PyTorch-CUDA path".  A stock install of the reference runs the
    nerfstudio torch implementation on CUDA (tinycudann is not in uv.lock) under fp16 autocast + GradScaler
    (mixed_precision=True, config_thermal_nerf.py:22): the PyTorch restatement (oracle port) is timed here the same
    way - eager, full training iterations, proposal networks on the sampler's schedule, torch Adam - in fp32 and in
    fp16 autocast.  `ratio` = this run's device-resident rays/s over the baseline's."""
    out = {"kind": "oracle port (nerfstudio-1.1.5 torch semantics), eager PyTorch on the same GPU, torch.optim.Adam",
           "rays_per_batch": rays}
    for tag, amp in (("fp32", False), ("fp16_autocast_gradscaler", True)):
        try:
            model, _, batch = oracle_train_setup(rays)
            model = model.to(device)
            field = [p for n, p in model.named_parameters() if n.startswith("field.")]
            props = [p for n, p in model.named_parameters() if n.startswith("proposal_networks.")]
            opts = [torch.optim.Adam(field, lr=1e-2, eps=1e-15), torch.optim.Adam(props, lr=1e-2, eps=1e-15)]
            batch = tuple(t.to(device) for t in batch)
            scaler = torch.amp.GradScaler("cuda") if amp else None
            sched = OracleSchedule()
            for i in range(3):
                oracle_train_step(model, opts, batch, i, sched, scaler)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 10
            a.record()
            for i in range(reps):
                oracle_train_step(model, opts, batch, 3 + i, sched, scaler)
            b.record()
            torch.cuda.synchronize()
            ms = a.elapsed_time(b) / reps
            out[tag] = {"value": rays / (ms * 1e-3), "unit": "rays/s", "ms_per_step": ms,
                        "ratio": ours_rays_per_s / (rays / (ms * 1e-3)),
                        "sample": f"{reps} full iterations of {rays} rays (each reads the loss back, as the trainer's "
                                  "logging does)"}
            del model, opts, batch
            torch.cuda.empty_cache()
        except Exception as e:  # context only: never fail the bench on it
            out[tag] = {"error": repr(e)[:300]}
    best = max((v["value"] for v in out.values() if isinstance(v, dict) and "value" in v), default=None)
    if best:
        out["value"], out["unit"] = best, "rays/s"
        out["ms_per_step"] = rays / best * 1e3
        out["ratio"] = ours_rays_per_s / best  # against the FASTER of the two baselines
    return out


def train_at_8192(device, precision: str) -> dict:
    """Device-resident training rate at 8192 rays per batch (one GPU) next to the PyTorch-CUDA baseline of the same
    batch size: a short run (20 + 100 iterations); the full contract run is `bench.py --rays 8192`."""
    from thermo_nerf_b200.engine import TrainEngine

    R = 8192
    try:
        model = build_b200_model(device, precision)
        model.train()
        engine = TrainEngine(model, world_size=1)
        batches = train_batches(4, R, device, 0)
        for i in range(20):
            engine.step(*batches[i % 4])
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(100):
            engine.step(*batches[i % 4])
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 100
        value = R / (ms * 1e-3)
        del engine, model, batches
        torch.cuda.empty_cache()
        return {"value": value, "unit": "rays/s", "ms_per_step": ms, "rays_per_batch": R, "steps": 100,
                "torch_cuda_baseline": torch_cuda_train_baseline(device, R, value)}
    except Exception as e:  # context only
        return {"error": repr(e)[:300]}


def kernel_breakdown(engine, batches, device) -> dict:
    """Times every kernel of one iteration with CUDA events on the launching stream (median of 11).  The three
    backward kernels share one C entry point; tnf_backward_stage_mask runs them one at a time."""
    from thermo_nerf_b200 import _lib as L
    from thermo_nerf_b200 import functional as F

    tc = engine.model._precision() == L.PRECISION_TC_FP16
    stages = (("backward_prop", 1), ("backward_field", 2)) + (() if tc else (("wgrad", 4),))
    names = ["forward", "losses"] + [n for n, _ in stages] + ["adam"]
    acc = {n: [] for n in names}
    lib = L.load()
    reps = 11

    def timed(name, fn):
        # a ~0.3 ms device-side spin first: the host prepares and enqueues the launch while the GPU is still
        # busy, so the two events bracket the kernel alone and not the Python/ctypes time before the launch
        torch.cuda._sleep(600_000)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        r = fn()
        b.record()
        acc[name].append((a, b))
        return r

    for rep in range(reps):
        o, d, cam, gt_rgb, gt_th = batches[rep % len(batches)]
        R = o.shape[0]
        cfg = engine.cfg
        jitter = torch.rand((3, R), device=device)
        kw = dict(num_samples=(*cfg.num_proposal_samples_per_ray, cfg.num_nerf_samples_per_ray),
                  near_plane=cfg.near_plane, far_plane=cfg.far_plane, anneal=engine.anneal(engine.step_count),
                  appearance_mode=L.APPEARANCE_LOOKUP, precision=engine.model._precision())
        res = timed("forward", lambda: F.render_forward(engine.tensors, o, d, cam, None, None, jitter, training=True,
                                                        return_samples=True, save_for_backward=True, **kw))
        losses, g = timed("losses", lambda: F.losses_forward_backward(
            res["weights_list"], res["sdist_list"], res["rgb"], res["thermal"], gt_rgb, gt_th))
        res["_workspace"] = engine._ws
        gout = {"rgb": g["rgb"], "thermal": g["thermal"], "weights_list": g["weights_list"]}
        try:
            for nm, mask in stages:
                lib.tnf_backward_stage_mask(mask)
                timed(nm, lambda: F.render_backward(engine.tensors, res["_model_struct"], o, d, cam, None, None, jitter,
                                                    res, gout, list(engine.grads)))
        finally:
            lib.tnf_backward_stage_mask(7)
        n = len(engine.params)
        # lr = 0: timing only, parameters unchanged
        if engine.arena is None:
            timed("adam", lambda: F.adam_step(engine.params, engine.grads, engine.exp_avg, engine.exp_avg_sq,
                                              [0.0] * n, step=1000, eps=1e-15, zero_grads=True))
        else:
            engine.grad_arena.zero_()
        torch.cuda.synchronize()
    acc = {k: [a.elapsed_time(b) for a, b in v] for k, v in acc.items() if v}
    # median: the interval between two events also contains any host stall between the launches (GC pause,
    # allocator growth), which an average would book as kernel time
    out = {k + "_ms": sorted(v)[len(v) // 2] for k, v in acc.items()}
    out["note"] = ("one kernel per entry, except forward: proposal launch + field launch in tensor-core mode, plus the 3 us "
                   "depth-clip pass; in tensor-core mode the field "
                   "backward accumulates the weight gradients in tensor memory - there is no separate weight-gradient "
                   "kernel); the proposal backward runs only on the sampler's update steps (every 2nd step in the first "
                   "1000 iterations, every 6th after 5000)")
    return out


def quick_render(model, device, with_torch_baseline: bool = True, ctx=None) -> dict:
    """The second headline metric inside the default line (BASELINE configs[4] shape, so that the driver's record
    carries it): 800x800 frames in eval mode - device-resident value, roofline of the forward kernel on algorithmic
    bytes, end to end through Renderer.render([RGB, THERMAL], camera) with uint8 frames on the host, and the
    PyTorch-CUDA baseline of the same frame.  The full contract run is bench.py --mode render."""
    from thermo_nerf_b200 import PinholeCameras, RenderedImageModality, Renderer, orbit_cameras

    model.eval()
    bundles = frame_bundles(2, device, 0, 1)
    flush = L2Flusher(device)
    with torch.no_grad():
        for i in range(2):
            model.get_outputs_for_camera_ray_bundle(bundles[i % 2])
        evs = []
        for i in range(5):
            flush()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            model.get_outputs_for_camera_ray_bundle(bundles[i % 2])
            b.record()
            evs.append((a, b))
        torch.cuda.synchronize()
    world = ctx["world"] if ctx else 1
    slowest = ctx["max_over_ranks"] if ctx else (lambda x: x)
    ms = slowest(sum(a.elapsed_time(b) for a, b in evs) / len(evs))
    value = world * HW * HW / (ms * 1e-3) / 1e6
    peak, peak_src = hbm_peak()
    achieved = HW * HW * ALGO_BYTES_PER_RAY / (ms * 1e-3) / 1e9  # per GPU
    out = {"metric": "render_mpix_per_s", "value": value, "unit": "Mpix/s", "ms_per_frame": ms, "n_gpus": world,
           "workload": "800x800 frame, rgb+thermal+depth+accumulation in one pass, L2 flushed between frames; "
                       "full contract run: bench.py --mode render",
           "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                        "kernel": "tnf_forward_kernel", "peak_source": peak_src,
                        "note": "frame time (forward + 3 us clip pass) on algorithmic bytes, 161 876 B per ray"}}
    # end to end: the reference-facing call of render_video_script.py, one camera per step
    cams = orbit_cameras(2, hw=HW, focal=FOCAL)
    one = [PinholeCameras(cams.camera_to_worlds[i:i + 1], cams.fx, cams.fy, cams.cx, cams.cy, cams.width, cams.height)
           for i in range(2)]
    lut = torch.rand((256, 3), generator=torch.Generator().manual_seed(5)).numpy()
    renderer = Renderer(model)
    mods = [RenderedImageModality.RGB, RenderedImageModality.THERMAL]
    for i in range(2):
        renderer.render(mods, one[i % 2], thermal_color_map=lut)
    t = 0.0
    for i in range(4):
        flush()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        renderer.render(mods, one[i % 2], thermal_color_map=lut)
        t += time.perf_counter() - t0
    t = slowest(t)
    out["e2e"] = {"value": world * HW * HW * 4 / t / 1e6, "unit": "Mpix/s", "h2d_bytes_per_step": 72,
                  "d2h_bytes_per_step": 2 * HW * HW * 3,
                  "note": "Renderer.render([RGB, THERMAL], one camera): rays generated in the kernel, uint8 + colour map "
                          "on the device, two uint8 frames D2H"}
    if with_torch_baseline:
        out["torch_cuda_baseline"] = torch_cuda_baseline(device, value)
    model.train()
    return out


def cpu_baseline_train(sample: int) -> dict:
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    model, opts, batch = oracle_train_setup(sample)
    sched = OracleSchedule()
    oracle_train_step(model, opts, batch, 0, sched)
    reps = 2
    t0 = time.perf_counter()
    for i in range(reps):
        oracle_train_step(model, opts, batch, 1 + i, sched)
    dt = (time.perf_counter() - t0) / reps
    return {"value": sample / dt, "unit": "rays/s", "cores": cores, "kind": "port",
            "sample": f"{reps} full training iterations of {sample} rays (forward, losses, autograd backward, torch "
                      f"Adam over all parameters), fp32 oracle port on CPU"}


def bench_render(ctx) -> dict:
    args, rank, world, device = ctx["args"], ctx["rank"], ctx["world"], ctx["device"]
    barrier, max_over_ranks, local = ctx["barrier"], ctx["max_over_ranks"], ctx["local"]
    from thermo_nerf_b200 import RayBundle

    model = build_b200_model(device, args.precision)
    n_distinct = 4
    bundles = frame_bundles(n_distinct, device, rank, world)
    flush = L2Flusher(device)
    rays_per_step = HW * HW
    out_keys = ("rgb", "thermal", "depth", "accumulation")

    def step(i: int):
        return model.get_outputs_for_camera_ray_bundle(bundles[i % n_distinct])

    # ---- device-resident throughput ("value")
    for i in range(args.warmup):
        step(i)
    barrier()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    sampler = ClockSampler(local)
    sampler.start()
    for i in range(args.steps):
        flush()
        evs[i][0].record()
        step(i)
        evs[i][1].record()
    barrier()
    clocks = sampler.stop()
    my_ms = sum(a.elapsed_time(b) for a, b in evs) / args.steps
    ms_total = max_over_ranks(sum(a.elapsed_time(b) for a, b in evs))
    ms_per_step = ms_total / args.steps
    # every rank renders the same views: what differs between them is the GPU (clocks under an 8-GPU load, binning)
    ranks = ctx["gather_ranks"]({"ms_per_step": my_ms, "sm_mhz": clocks["sm_mhz"], "reasons": clocks["reasons"]})
    value = world * rays_per_step / (ms_per_step * 1e-3) / 1e6

    # ---- dominant kernel alone (roofline): tnf_render_forward on device-resident flat rays
    from thermo_nerf_b200 import functional as F
    from thermo_nerf_b200 import _lib as L

    flat = bundles[0].flatten()
    o, d = flat.origins.contiguous(), flat.directions.contiguous()
    kw = dict(near_plane=0.0, far_plane=1000.0, appearance_mode=L.APPEARANCE_MEAN,
              precision=L.PRECISION_FP32 if args.precision == "fp32" else L.PRECISION_TC_FP16,
              depth_clip_chunk=1 << 16)
    for _ in range(3):
        F.render_forward(model.tensors(), o, d, **kw)
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(10)]
    for a, b in kev:
        flush()
        a.record()
        F.render_forward(model.tensors(), o, d, **kw)
        b.record()
    torch.cuda.synchronize()
    k_ms = sum(a.elapsed_time(b) for a, b in kev) / len(kev)
    peak, peak_src = hbm_peak()
    achieved = rays_per_step * ALGO_BYTES_PER_RAY / (k_ms * 1e-3) / 1e9
    traffic = None
    tp = ROOT / "profiles" / "forward_traffic.json"
    if tp.exists():
        try:
            traffic = json.load(open(tp)).get(args.precision, {}).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "kernel": "tnf_forward_kernel", "kernel_ms": k_ms, "peak_source": peak_src,
                "algorithmic_bytes_per_ray": ALGO_BYTES_PER_RAY,
                "note": "tables (74 MiB fp32) are L2-resident, so DRAM traffic is far below algorithmic gather bytes"}

    # ---- end to end through the public API with HOST buffers (pinned) in the timed region
    host_in = []
    for b in bundles:
        f = b.flatten()
        host_in.append((f.origins.cpu().pin_memory(), f.directions.cpu().pin_memory(),
                        f.camera_indices.cpu().pin_memory()))
    host_out = {k: torch.empty((HW, HW, 3 if k == "rgb" else 1), dtype=torch.float32).pin_memory() for k in out_keys}
    h2d = sum(t.numel() * t.element_size() for t in host_in[0])

    def e2e_step(i: int):
        ho, hd, hc = host_in[i % n_distinct]
        rb = RayBundle(origins=ho.to(device, non_blocking=True).view(HW, HW, 3),
                       directions=hd.to(device, non_blocking=True).view(HW, HW, 3),
                       camera_indices=hc.to(device, non_blocking=True).view(HW, HW, 1))
        out = model.get_outputs_for_camera_ray_bundle(rb)
        for k in out_keys:
            host_out[k].copy_(out[k], non_blocking=True)
        torch.cuda.current_stream().synchronize()  # the caller consumes the frame on the host

    def time_e2e(fn) -> float:
        for i in range(args.warmup):
            fn(i)
        barrier()
        t = 0.0
        for i in range(args.steps):
            flush()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            fn(i)
            t += time.perf_counter() - t0
        barrier()
        return world * rays_per_step * args.steps / max_over_ranks(t) / 1e6

    e2e_bundle = time_e2e(e2e_step)

    # ---- end to end through the reference-facing call: Renderer.render([RGB, THERMAL], cameras) - the default
    #      modalities of render_video_script.py - one camera per step: the camera (72 B of kernel arguments) goes
    #      in, the two uint8 frames come back to pinned host memory
    from thermo_nerf_b200 import PinholeCameras, RenderedImageModality, Renderer, orbit_cameras
    all_cams = orbit_cameras(n_distinct, hw=HW, focal=FOCAL)
    mine = range(n_distinct)  # the same views on every rank, see frame_bundles
    one = [PinholeCameras(all_cams.camera_to_worlds[i:i + 1], all_cams.fx, all_cams.fy, all_cams.cx, all_cams.cy,
                          all_cams.width, all_cams.height) for i in mine]
    g = torch.Generator().manual_seed(5)
    lut = torch.rand((256, 3), generator=g).numpy()  # stand-in colour table (matplotlib's magma is not installed)
    renderer = Renderer(model)
    mods = [RenderedImageModality.RGB, RenderedImageModality.THERMAL]

    def renderer_step(i: int):
        renderer.render(mods, one[i % n_distinct], thermal_color_map=lut)

    e2e_val = time_e2e(renderer_step)
    d2h = 2 * HW * HW * 3
    h2d_cam = 72

    line = {
        "metric": "render_mpix_per_s", "value": value, "unit": "Mpix/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None,
        "dtype": "fp32 hash/sampling/compositing" + (" + fp16-operand/fp32-accumulate mma field MLPs"
                                                      if args.precision == "tc_fp16" else " + fp32 MLPs"),
        "data": "synthetic",
        "config": {"workload": "render 800x800 ThermoScenes-shaped frame per step (640000 rays, samples 256/96/48, "
                               "eval chunk 65536), rgb+thermal+depth+accumulation in one pass; frames shard "
                               "over ranks with no collective (every rank renders the same 4 views: equal work per "
                               "rank, frame cost depends on the view)",
                   "l2": "flushed between timed iterations (256 MiB write)", "rays_per_second": value * 1e6,
                   "weights": "random trained-like, full-size tables (field 2^19x16, proposals 2^17x5)"},
        "clocks": clocks, "ranks": ranks, "gpu_launches": 2 * args.steps,  # forward + depth-clip pass per frame (device-resident loop)
        "e2e": {"value": e2e_val, "unit": "Mpix/s", "h2d_bytes_per_step": h2d_cam, "d2h_bytes_per_step": d2h,
                "note": "Renderer.render([RGB, THERMAL], one camera): rays generated in the kernel from the camera, "
                        "uint8 conversion + colour map on the device, two uint8 frames D2H into pinned memory"},
        "e2e_ray_bundle": {"value": e2e_bundle, "unit": "Mpix/s", "h2d_bytes_per_step": h2d,
                           "d2h_bytes_per_step": sum(t.numel() * t.element_size() for t in host_out.values()),
                           "note": "pinned host rays -> H2D -> get_outputs_for_camera_ray_bundle -> D2H float "
                                   "rgb/thermal/depth/accumulation"},
        "roofline": roofline,
    }

    if rank == 0 and world == 1 and not args.no_torch_cuda_baseline:
        line["torch_cuda_baseline"] = torch_cuda_baseline(device, value)
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(args.ref_rays)
    return line


def cpu_baseline(sample: int) -> dict:
    from oracle import OracleConfig, OracleThermalNerf, make_synthetic_rays

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    model = OracleThermalNerf(OracleConfig(), NUM_IMAGES, seed=0)
    randomise_trained_like(model, 0)
    rays = make_synthetic_rays(sample, num_images=NUM_IMAGES, seed=1, contiguous_pixels=True)
    with torch.no_grad():
        model.get_outputs(rays, training=False)
        t0 = time.perf_counter()
        reps = 2
        for _ in range(reps):
            model.get_outputs(rays, training=False)
        dt = (time.perf_counter() - t0) / reps
    return {"value": sample / dt / 1e6, "unit": "Mpix/s", "cores": cores, "kind": "port",
            "sample": f"{sample} contiguous rays of an 800x800 frame, eval forward fp32, oracle port on CPU, "
                      f"{reps} reps"}


def torch_cuda_baseline(device, ours_mpix_per_s: float = None, frame: bool = True) -> dict:
    """The PyTorch restatement run eagerly on the GPU (what a stock install of the reference executes on CUDA;
    tinycudann is absent from uv.lock): one 65536-ray eval chunk and, with `frame`, one whole 800x800 frame in
    eval_num_rays_per_chunk = 65536 slices as nerfstudio's get_outputs_for_camera_ray_bundle does, fp32 and fp16
    autocast."""
    from oracle import OracleConfig, OracleRays, OracleThermalNerf, make_synthetic_rays

    out = {"kind": "oracle port (nerfstudio-1.1.5 torch semantics), eager PyTorch on the same GPU"}
    try:
        model = OracleThermalNerf(OracleConfig(), NUM_IMAGES, seed=0)
        randomise_trained_like(model, 0)
        model = model.to(device)
        chunk = 1 << 16
        r = make_synthetic_rays(chunk, num_images=NUM_IMAGES, seed=1, contiguous_pixels=True)
        rays = OracleRays(r.origins.to(device), r.directions.to(device), r.camera_indices.to(device))
        for tag, amp in (("fp32", False), ("fp16_autocast", True)):
            with torch.no_grad(), torch.autocast(device_type="cuda", dtype=torch.float16, enabled=amp):
                for _ in range(2):
                    model.get_outputs(rays, training=False)
                torch.cuda.synchronize()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                reps = 5
                a.record()
                for _ in range(reps):
                    model.get_outputs(rays, training=False)
                b.record()
                torch.cuda.synchronize()
                ms = a.elapsed_time(b) / reps
                out[tag] = {"value": chunk / (ms * 1e-3) / 1e6, "unit": "Mpix/s", "sample": "one 65536-ray eval chunk",
                            "ms_per_chunk": ms}
                if frame:
                    n = HW * HW
                    a.record()
                    for s0 in range(0, n, chunk):
                        k = min(chunk, n - s0)
                        model.get_outputs(OracleRays(rays.origins[:k], rays.directions[:k], rays.camera_indices[:k]),
                                          training=False)
                    b.record()
                    torch.cuda.synchronize()
                    fms = a.elapsed_time(b)
                    out[tag]["frame_800x800"] = {"value": n / (fms * 1e-3) / 1e6, "unit": "Mpix/s", "ms_per_frame": fms}
        best = max(v["value"] for v in out.values() if isinstance(v, dict) and "value" in v)
        out["value"], out["unit"] = best, "Mpix/s"
        if ours_mpix_per_s:
            out["ratio"] = ours_mpix_per_s / best  # against the FASTER of the two baselines
    except Exception as e:  # context only: never fail the bench on it
        out["error"] = repr(e)[:300]
    return out


if __name__ == "__main__":
    main()
