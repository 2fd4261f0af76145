#!/usr/bin/env python
"""RGB PSNR / thermal MAE of a model trained with the B200 path on a synthetic ThermoScenes-shaped scene
(the real dataset is not available offline): a random 'trained-like' teacher model renders train and held-out
views; a freshly initialised student is trained on the teacher's train views with the device pixel sampler and
the fused engine, and evaluated on the held-out views exactly as the reference's Evaluator does
(evaluator.py:47-106: PSNR of rgb, thermal_metrics.mae_thermal of the de-normalised thermal image).

    python scripts/quality_synthetic.py [--steps 3000] [--precision tc_fp16|fp32|oracle_fp32|oracle_fp16]  -> one JSON line

``oracle_*``: the student is the PyTorch restatement of the reference (oracle port) trained eagerly on the GPU with
torch.optim.Adam on the same batches, schedule and learning-rate decay - fp32, or fp16 autocast + GradScaler as the
reference's mixed_precision=True - and evaluated through the same fp32 render of its weights.  It answers whether
the B200 path's arithmetic (fp16 forward / bf16 backward operands in tensor-core mode) costs quality against the
reference's own arithmetic, not just against the path's fp32 mode.
"""
import argparse
import json
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch  # noqa: E402

from thermo_nerf_b200 import DevicePixelSampler, ThermalNerfModel, ThermalNerfModelConfig, sphere_cameras  # noqa: E402
from thermo_nerf_b200.engine import TrainEngine  # noqa: E402


def build(seed, precision, n_img, trained_like):
    torch.manual_seed(seed)
    cfg = ThermalNerfModelConfig(precision=precision, camera_optimizer_mode="off", max_temperature=40.0,
                                 min_temperature=10.0)
    m = ThermalNerfModel(cfg, {"thermal": []}, torch.tensor([[-1.0, -1, -1], [1, 1, 1]]), n_img)
    if trained_like:
        g = torch.Generator().manual_seed(seed + 1)
        with torch.no_grad():
            # smooth teacher: only the coarse half of the levels carries signal, densities moderately sharp
            t = m.field.mlp_base.encoder.hash_table
            t.copy_(torch.randn(t.shape, generator=g) * 0.5)
            t.view(16, -1, 2)[8:] *= 0.05
            m.field.mlp_base.mlp.layers[1].weight[0].mul_(4.0)
            # colourful, thermally varied surfaces: amplify the geo features and both heads, centre the thermal output
            m.field.mlp_base.mlp.layers[1].weight[1:].mul_(4.0)
            for lin in (m.field.mlp_head.layers[0], m.field.mlp_head.layers[2], m.field.mlp_thermal.layers[0],
                        m.field.mlp_thermal.layers[1]):
                lin.weight.mul_(5.0)
            m.field.field_head_thermal.net.weight.mul_(3.0)
            # thermal = w . sigmoid(.) + b with sigmoid(.) centred on 1/2: centre the output on mid-range
            m.field.field_head_thermal.net.bias.fill_(0.5 - 0.5 * float(m.field.field_head_thermal.net.weight.sum()))
    return m.cuda()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=3000)
    ap.add_argument("--rays", type=int, default=4096)
    ap.add_argument("--precision", default="tc_fp16")
    ap.add_argument("--hw", type=int, default=160)
    args = ap.parse_args()
    n_train, n_test, hw = 60, 6, args.hw
    cams = sphere_cameras(n_train + n_test, hw=hw, focal=1.4 * hw)
    teacher = build(1, "fp32", n_train + n_test, True).eval()
    with torch.no_grad():
        frames = [teacher.get_outputs_for_camera(cams, i) for i in range(n_train + n_test)]
    rgb = torch.stack([f["rgb"] for f in frames])          # [N,H,W,3]
    th = torch.stack([f["thermal"] for f in frames])       # [N,H,W,1]
    test_ids = list(range(0, n_train + n_test, (n_train + n_test) // n_test))[:n_test]
    train_ids = [i for i in range(n_train + n_test) if i not in test_ids]
    sampler = DevicePixelSampler(rgb[train_ids], th[train_ids], cams.camera_to_worlds[train_ids], cams.fx, cams.fy,
                                 cams.cx, cams.cy, device="cuda", seed=0)
    oracle_arm = args.precision.startswith("oracle")
    student = build(2, "fp32" if oracle_arm else args.precision, len(train_ids), False).train()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    if oracle_arm:
        sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
        from bench import OracleSchedule, oracle_train_step
        from oracle import OracleConfig, OracleThermalNerf
        from thermo_nerf_b200.engine import exponential_decay_lr

        om = OracleThermalNerf(OracleConfig(camera_optimizer_mode="off"), len(train_ids), seed=0)
        missing, unexpected = om.load_state_dict(student.state_dict(), strict=False)  # the student's initialisation
        assert not [k for k in missing if "camera_optimizer" not in k], missing
        om = om.cuda().train()
        field = [p for n, p in om.named_parameters() if n.startswith("field.")]
        props = [p for n, p in om.named_parameters() if n.startswith("proposal_networks.")]
        opts = [torch.optim.Adam(field, lr=1e-2, eps=1e-15), torch.optim.Adam(props, lr=1e-2, eps=1e-15)]
        scaler = torch.amp.GradScaler("cuda") if args.precision == "oracle_fp16" else None
        sched = OracleSchedule()
        for step in range(args.steps):
            rb, batch = sampler.sample(args.rays)
            for op in opts:
                for g in op.param_groups:
                    g["lr"] = exponential_decay_lr(step)
            oracle_train_step(om, opts, (rb.origins, rb.directions, rb.camera_indices.reshape(-1), batch["image"],
                                         batch["thermal"].reshape(-1)), step, sched, scaler)
        student.load_state_dict(om.state_dict(), strict=False)
        student._tensors = None
    else:
        eng = TrainEngine(student)
        for step in range(args.steps):
            rb, batch = sampler.sample(args.rays)
            eng.step(rb.origins, rb.directions, rb.camera_indices.reshape(-1), batch["image"],
                     batch["thermal"].reshape(-1))
    torch.cuda.synchronize()
    train_s = time.perf_counter() - t0
    student.eval()
    psnr, mae = [], []
    with torch.no_grad():
        for i in test_ids:
            out = student.get_outputs_for_camera(cams, i)
            mse = torch.mean((out["rgb"] - rgb[i]) ** 2)
            psnr.append(float(-10 * torch.log10(mse)))
            mae.append(float(student.mae_thermal(th[i], out["thermal"])))
    print(json.dumps({"metric": "held-out RGB PSNR (dB) / thermal MAE (deg C, range 10-40) on a synthetic teacher scene",
                      "precision": args.precision, "steps": args.steps, "rays_per_batch": args.rays,
                      "train_views": len(train_ids), "test_views": len(test_ids), "resolution": hw,
                      "psnr_db": sum(psnr) / len(psnr), "thermal_mae_degC": sum(mae) / len(mae),
                      "scene_stats": {"rgb_std": float(rgb.std()), "thermal_std_degC": float(th.std()) * 30.0,
                                      "thermal_min_max_norm": [float(th.min()), float(th.max())],
                                      "mean_accumulation": float(torch.stack([f["accumulation"] for f in frames]).mean())},
                      "train_seconds": train_s, "train_rays_per_s_incl_sampler": args.steps * args.rays / train_s}))


if __name__ == "__main__":
    main()
