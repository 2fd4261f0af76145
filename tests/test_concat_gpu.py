"""The CUDA path of the reference's `concat_nerf` model type against the oracle's restatement
(OracleConfig(head="concat"), itself checked against the reference's own ConcatNerfModel / ConcatNerfactoTField /
RGBTRenderer code in tests/test_oracle_concat_cpu.py and tests/golden/reference_concat_wiring.pt):
one 4-channel RGBT colour head (thermo_nerf/rgb_concat/concat_field.py:65-75), no background term in the composite
(rgb_concat/rgbt_renderer.py:63-71), the noise-blended loss (rgb_concat/concat_nerfacto_model.py:197-211) and its
gradients.  Tolerances: fp32 mode 2e-4 abs on outputs, rel-L2 5e-3 per gradient tensor (the test sharpens the colour
head, so a sample that lands in the neighbouring hash cell moves the table gradient more than in the thermal tests);
tensor-core mode 2e-2 abs, rel-L2 6e-2 and cosine >= 0.998 (bf16 backward operands)."""

from __future__ import annotations

import pytest
import torch

pytestmark = pytest.mark.gpu

from tests.helpers import make_trained_like, oracle_config  # noqa: E402


def _pair(precision="fp32", num_images=6, seed=7, log2_field=14, log2_prop=11):
    from oracle import OracleThermalNerf
    from thermo_nerf_b200 import ConcatNerfModel, ConcatNerfModelConfig

    ocfg = oracle_config(log2_field=log2_field, log2_prop=log2_prop, head="concat", camera_optimizer_mode="off")
    oracle = OracleThermalNerf(ocfg, num_images, seed=seed)
    make_trained_like(oracle, seed)
    with torch.no_grad():  # spread the temperature channel
        oracle.field.mlp_head.layers[1].weight.mul_(3.0)
        oracle.field.mlp_head.layers[2].weight.mul_(6.0)
    cfg = ConcatNerfModelConfig(log2_hashmap_size=log2_field, precision=precision, camera_optimizer_mode="off",
                                proposal_net_args_list=[dict(a, use_linear=False) for a in ocfg.proposal_net_args_list])
    model = ConcatNerfModel(cfg, torch.tensor([[-1.0, -1, -1], [1, 1, 1]]), num_images)
    missing, unexpected = model.load_state_dict(oracle.state_dict(), strict=False)
    assert set(missing) <= {"device_indicator_param"} and not unexpected, (missing, unexpected)
    assert not any("thermal" in k for k in model.state_dict())
    assert model.field.mlp_head.layers[2].weight.shape == (4, 64)
    return oracle, model.to("cuda:0")


@pytest.mark.parametrize("precision,tol", [("fp32", 2e-4), ("tc_fp16", 2e-2)])
def test_concat_eval_and_train_forward_match_oracle(precision, tol):
    from oracle import make_synthetic_rays
    from thermo_nerf_b200 import RayBundle

    oracle, model = _pair(precision)
    R = 300
    rays = make_synthetic_rays(R, num_images=6, seed=2)
    rb = lambda: RayBundle(origins=rays.origins.cuda(), directions=rays.directions.cuda(),  # noqa: E731
                           camera_indices=rays.camera_indices.cuda())
    model.eval()
    with torch.no_grad():
        ref = oracle.get_outputs(rays, training=False)
        out = model(rb())
    assert out["rgb"].shape == (R, 4) and "thermal" not in out
    for k in ("rgb", "accumulation"):
        err = (out[k].cpu() - ref[k]).abs().max().item()
        assert err <= tol, (k, err)
    assert float(ref["rgb"][:, 3].std()) > 0.01  # the temperature channel has contrast
    # no background term: every channel is bounded by the accumulation
    assert bool((out["rgb"] <= out["accumulation"] + 1e-5).all())
    rel = ((out["expected_depth"].cpu() - ref["expected_depth"]).abs() / ref["expected_depth"].abs().clamp_min(1e-3)).max()
    assert rel <= max(10 * tol, 1e-4)
    # whole-frame path (rays generated in the kernel) returns the 4-channel image too
    from thermo_nerf_b200 import sphere_cameras

    cams = sphere_cameras(2, hw=24, focal=30.0)
    with torch.no_grad():
        a = model.get_outputs_for_camera(cams, 1)
        b = model.get_outputs_for_camera_ray_bundle(cams.generate_rays(1).to("cuda:0"))
    # (PinholeCameras.generate_rays is torch arithmetic, the kernel generates the same rays to the last ulp or so)
    assert a["rgb"].shape == (24, 24, 4) and torch.allclose(a["rgb"], b["rgb"], atol=5e-3) and a["img"] is a["rgb"]
    assert float((a["rgb"] - b["rgb"]).abs().mean()) < 1e-4


@pytest.mark.parametrize("precision,rel_tol,cos_min", [("fp32", 5e-3, 0.9999), ("tc_fp16", 6e-2, 0.998)])
def test_concat_loss_and_gradients_match_oracle_autograd(precision, rel_tol, cos_min):
    from oracle import make_synthetic_rays
    from thermo_nerf_b200 import _lib as L
    from thermo_nerf_b200 import functional as F

    oracle, model = _pair(precision)
    R = 320
    rays = make_synthetic_rays(R, num_images=6, seed=5)
    g = torch.Generator().manual_seed(9)
    jitter = torch.rand((3, R, 1), generator=g)
    gt = torch.rand((R, 4), generator=g)
    noise = torch.rand((R, 4), generator=g)
    oracle.train()
    oracle.zero_grad()
    ref_out = oracle.get_outputs(rays, training=True, jitter=jitter)
    ref_ld = oracle.get_loss_dict(ref_out, gt, training=True, background_noise=noise)
    sum(ref_ld.values()).backward()
    model.train()
    model.zero_grad()
    out = F.render(model.tensors(), rays.origins.cuda(), rays.directions.cuda(), rays.camera_indices.cuda(), None, None,
                   jitter.cuda().reshape(3, -1), num_samples=(256, 96, 48), near_plane=0.05, far_plane=1000.0,
                   anneal=1.0, appearance_mode=L.APPEARANCE_LOOKUP,
                   precision=L.PRECISION_FP32 if precision == "fp32" else L.PRECISION_TC_FP16, head_mode=L.HEAD_CONCAT)
    ld = F.losses(out, gt[:, :3].cuda(), gt[:, 3].cuda(), concat_noise=noise.cuda())
    assert set(ld) == {"rgb_loss", "interlevel_loss", "distortion_loss"} == set(ref_ld)
    sum(ld.values()).backward()
    torch.cuda.synchronize()
    for k in ld:
        a, b = float(ld[k].detach()), float(ref_ld[k].detach())
        assert abs(a - b) <= (5e-4 if precision == "fp32" else 2e-2) * max(abs(b), 1e-3), (k, a, b)
    ref_grads = dict(oracle.named_parameters())
    checked = 0
    for name, p in model.named_parameters():
        if name not in ref_grads or ref_grads[name].grad is None or name.startswith("camera_optimizer"):
            continue
        assert p.grad is not None, name
        a, b = p.grad.detach().cpu().flatten(), ref_grads[name].grad.flatten()
        if float(b.norm()) < 1e-12:
            continue
        rel = float((a - b).norm() / b.norm())
        cos = float(torch.dot(a, b) / (a.norm() * b.norm()).clamp_min(1e-30))
        assert rel <= rel_tol and cos >= cos_min, (name, rel, cos)
        checked += 1
    assert checked >= 20
    # the RGBT head's fourth row received a gradient (the temperature channel trains through the colour head)
    assert float(model.field.mlp_head.layers[2].weight.grad[3].abs().sum()) > 0


def test_concat_model_training_step_through_the_plugin_surface_and_engine():
    from oracle import make_synthetic_rays
    from thermo_nerf_b200 import FusedAdam, RayBundle
    from thermo_nerf_b200.engine import TrainEngine

    _, model = _pair("tc_fp16", log2_field=12, log2_prop=10)
    model.train()
    R = 256
    rays = make_synthetic_rays(R, num_images=6, seed=11)
    gt = (0.5 + 0.4 * torch.sin(rays.directions[:, [0, 1, 2, 0]] * 5.0)).cuda()
    opt = FusedAdam(model.parameters(), lr=1e-2, eps=1e-15)
    hist = []
    for _ in range(12):
        opt.zero_grad()
        out = model(RayBundle(origins=rays.origins.cuda(), directions=rays.directions.cuda(),
                              camera_indices=rays.camera_indices.cuda()))
        m = model.get_metrics_dict(out, {"image": gt})
        ld = model.get_loss_dict(out, {"image": gt}, m)
        assert set(ld) == {"rgb_loss", "interlevel_loss", "distortion_loss"}
        sum(ld.values()).backward()
        opt.step()
        hist.append(float(ld["rgb_loss"]))
    assert hist[-1] < 0.8 * hist[0], hist
    eng = TrainEngine(model)
    l0 = eng.step(rays.origins.cuda(), rays.directions.cuda(), rays.camera_indices.cuda().reshape(-1), gt[:, :3].contiguous(),
                  gt[:, 3].contiguous())
    assert torch.isfinite(l0).all() and float(l0[3]) == 0.0
    model.eval()
    H = W = 16
    metrics, images = model.get_image_metrics_and_images(
        {"rgb": torch.rand(H, W, 4).cuda(), "accumulation": torch.rand(H, W, 1).cuda(), "depth": torch.rand(H, W, 1).cuda(),
         "prop_depth_0": torch.rand(H, W, 1).cuda(), "prop_depth_1": torch.rand(H, W, 1).cuda()},
        {"image": torch.rand(H, W, 4)}, threshold=0.3)
    assert set(metrics) == {"psnr", "ssim", "lpips", "mae_thermal_foreground", "mae_thermal"}
    assert images["img"].shape == (H, 2 * W, 4)
