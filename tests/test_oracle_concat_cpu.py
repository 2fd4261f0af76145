"""Oracle restatement of the reference's `concat_nerf` baseline: one 4-channel RGBT colour
head (rgb_concat/concat_field.py:65-75), RGBTRenderer with its default "random" background
(rgb_concat/rgbt_renderer.py:24,63-71), and the loss of rgb_concat/concat_nerfacto_model.py:197-233.
CPU only; the CUDA path for this head is not built yet."""

import pytest
import torch

from tests.helpers import make_trained_like, oracle_config
from oracle import OracleThermalNerf, make_synthetic_rays
from oracle import nerfstudio_math as M


def concat_oracle(seed=0, **kw):
    cfg = oracle_config(log2_field=12, log2_prop=10, num_samples=(32, 16, 12), head="concat", **kw)
    o = OracleThermalNerf(cfg, 6, seed=seed)
    make_trained_like(o, seed)
    return o


def test_concat_field_has_one_four_channel_head_and_no_temperature_head():
    o = concat_oracle()
    keys = set(o.state_dict())
    assert o.field.mlp_head.layers[-1].weight.shape == (4, 64)
    assert o.field.mlp_head.layers[0].weight.shape == (64, 16 + 15 + 32)
    assert not any("mlp_thermal" in k or "field_head_thermal" in k for k in keys)
    # everything else is the thermal model's skeleton
    t = OracleThermalNerf(oracle_config(log2_field=12, log2_prop=10, num_samples=(32, 16, 12)), 6, seed=0)
    shared = {k for k in t.state_dict() if "mlp_thermal" not in k and "field_head_thermal" not in k}
    assert shared == keys
    for k in shared - {"field.mlp_head.layers.2.weight", "field.mlp_head.layers.2.bias"}:
        assert t.state_dict()[k].shape == o.state_dict()[k].shape, k


def test_concat_outputs_shapes_and_keys():
    o = concat_oracle()
    rays = make_synthetic_rays(40, num_images=6, seed=3)
    with torch.no_grad():
        out = o.get_outputs(rays, training=False)
    assert out["rgb"].shape == (40, 4)
    assert "thermal" not in out
    for k in ("accumulation", "depth", "expected_depth", "prop_depth_0", "prop_depth_1"):
        assert out[k].shape == (40, 1), k
    assert out["field_rgb"].shape == (40, 12, 4)


def test_concat_render_is_the_plain_weighted_sum():
    """"random" background: no background term, so every channel is bounded by the accumulation
    (sigmoid outputs lie in (0,1)) - unlike last_sample compositing, which always sums to a convex combination."""
    o = concat_oracle()
    rays = make_synthetic_rays(64, num_images=6, seed=4)
    with torch.no_grad():
        out = o.get_outputs(rays, training=True, jitter=torch.rand(3, 64, 1, generator=torch.Generator().manual_seed(1)))
    w, c = out["weights_list"][-1], out["field_rgb"]
    want = (w * c).sum(-2)
    assert torch.equal(out["rgb"], want)
    assert (out["rgb"] <= out["accumulation"] + 1e-6).all()
    # scalar loop restatement of one ray
    r = 17
    acc = [0.0] * 4
    for s in range(c.shape[1]):
        for ch in range(4):
            acc[ch] += float(w[r, s, 0]) * float(c[r, s, ch])
    assert torch.allclose(out["rgb"][r], torch.tensor(acc), atol=1e-6)


def test_concat_eval_sanitises_and_clamps():
    vals = torch.tensor([[[0.5, float("nan"), 2.0, -1.0], [0.5, 0.5, 2.0, -1.0]]])
    w = torch.tensor([[[0.6], [0.4]]])
    out = M.render_rgbt_no_background(vals, w, training=False)
    assert torch.allclose(out, torch.tensor([[0.5, 0.2, 1.0, 0.0]]))
    out_t = M.render_rgbt_no_background(vals[..., [0, 2, 3]], w, training=True)
    assert torch.allclose(out_t, torch.tensor([[0.5, 2.0, -1.0]]))  # training: no clamp


def test_concat_loss_blends_noise_into_the_prediction_only():
    o = concat_oracle()
    R = 48
    rays = make_synthetic_rays(R, num_images=6, seed=5)
    g = torch.Generator().manual_seed(2)
    jitter = torch.rand(3, R, 1, generator=g)
    gt = torch.rand(R, 4, generator=g)
    noise = torch.rand(R, 4, generator=g)
    out = o.get_outputs(rays, training=True, jitter=jitter)
    loss = o.get_loss_dict(out, gt, training=True, background_noise=noise)
    assert set(loss) == {"rgb_loss", "interlevel_loss", "distortion_loss"}
    want = ((out["rgb"] + noise * (1 - out["accumulation"]) - gt) ** 2).mean()
    assert torch.allclose(loss["rgb_loss"], want, atol=1e-7)
    # eval: colour term only
    assert set(o.get_loss_dict(out, gt, training=False, background_noise=noise)) == {"rgb_loss"}
    # the noise term carries gradient into the accumulation (hence the densities), the GT does not move
    total = sum(loss.values())
    total.backward()
    head = o.field.mlp_head.layers[-1].weight.grad
    assert head is not None and head.shape == (4, 64) and (head.abs().sum(dim=1) > 0).all()
    assert o.field.mlp_base.encoder.hash_table.grad.abs().sum() > 0
    for p in o.proposal_networks:
        assert p.encoding.hash_table.grad.abs().sum() > 0  # interlevel loss


def test_concat_fourth_channel_is_independent_of_the_temperature_switches():
    """pass_thermal_gradients only exists for the separate temperature head."""
    a = concat_oracle(pass_thermal_gradients=True)
    b = concat_oracle(pass_thermal_gradients=False)
    assert a.field.pass_thermal_gradients is False and b.field.pass_thermal_gradients is False
    rays = make_synthetic_rays(16, num_images=6, seed=6)
    with torch.no_grad():
        assert torch.equal(a.get_outputs(rays)["rgb"], b.get_outputs(rays)["rgb"])


# ---- against the reference's own ConcatNerfModel code (tests/golden/make_reference_concat_golden.py) ----
def test_concat_oracle_matches_the_reference_concat_model():
    """rgb_concat/concat_nerfacto_model.py (populate_modules, get_loss_dict, get_metrics_dict), concat_field.py and
    rgbt_renderer.py executed from the reference over the nerfstudio stand-ins: module tree, outputs, loss (with the same
    torch.rand_like draw for the "random" background), gradients."""
    from pathlib import Path

    from oracle import OracleConfig, OracleRays

    gold = torch.load(Path(__file__).parent / "golden" / "reference_concat_wiring.pt", weights_only=True)
    mini = gold["mini"]
    cfg = OracleConfig(log2_hashmap_size=mini["log2_hashmap_size"],
                       num_proposal_samples_per_ray=tuple(mini["num_proposal_samples_per_ray"]),
                       num_nerf_samples_per_ray=mini["num_nerf_samples_per_ray"],
                       proposal_net_args_list=[{k: v for k, v in a.items() if k != "use_linear"}
                                               for a in mini["proposal_net_args_list"]], head="concat")
    o = OracleThermalNerf(cfg, gold["num_images"], seed=0)
    assert gold["load_unexpected"] == [] and set(gold["load_missing"]) <= {"device_indicator_param"}
    ref_keys = {k: tuple(v) for k, v in gold["reference_state_dict_keys"].items() if k != "device_indicator_param"}
    assert ref_keys == {k: tuple(v.shape) for k, v in o.state_dict().items()}
    assert ref_keys["field.mlp_head.layers.2.weight"] == (4, 64)
    assert gold["background_color"] == "random"  # RGBTRenderer() default (concat_nerfacto_model.py:181)
    o.load_state_dict(gold["state_dict"], strict=True)

    def rays():
        return OracleRays(gold["origins"].clone(), gold["directions"].clone(), gold["camera_indices"].clone())

    tr = gold["train"]
    o.anneal = tr["anneal"]
    out = o.get_outputs(rays(), training=True, jitter=gold["jitter"])
    assert tr["output_keys"] == ["accumulation", "depth", "expected_depth", "prop_depth_0", "prop_depth_1",
                                 "ray_samples_list", "rgb", "weights_list"]
    for k, v in tr["outputs"].items():
        assert out[k].shape == v.shape and torch.allclose(out[k], v, atol=1e-6, rtol=1e-6), k
    torch.manual_seed(tr["noise_seed"])
    noise = torch.rand_like(out["rgb"])
    loss = o.get_loss_dict(out, gold["batch"]["image"], training=True, background_noise=noise)
    assert set(loss) == set(tr["loss"]) == {"rgb_loss", "interlevel_loss", "distortion_loss"}
    for k, v in tr["loss"].items():
        assert torch.allclose(loss[k], v, atol=1e-7, rtol=1e-5), (k, loss[k], v)
    assert tr["metric_keys"] == ["distortion", "psnr"]
    mse = torch.mean((out["rgb"] - gold["batch"]["image"]) ** 2)  # psnr over all four channels (:240)
    assert torch.allclose(-10 * torch.log10(mse), tr["psnr"], rtol=1e-5)
    o.zero_grad()
    sum(loss.values()).backward()
    grads = {k: p.grad for k, p in o.named_parameters()}
    for k, g in tr["grads"].items():
        assert ((grads[k] - g).abs().max() / g.abs().max().clamp_min(1e-12)) < 1e-4, k
    o.anneal = 1.0  # the eval vectors were taken with a fresh sampler's anneal
    with torch.no_grad():
        ev = o.get_outputs(rays(), training=False)
    assert gold["eval"]["output_keys"] == ["accumulation", "depth", "expected_depth", "prop_depth_0", "prop_depth_1", "rgb"]
    for k, v in gold["eval"]["outputs"].items():
        assert torch.allclose(ev[k], v, atol=1e-6, rtol=1e-6), k


def test_concat_image_metrics_take_the_temperature_from_channel_three():
    """concat_nerfacto_model.py:251-324: psnr / ssim / lpips / MAE are all computed on channel 3 of the 4-channel image
    (lpips on its 3-fold repetition); psnr and the MAE are the real formulas, ssim / lpips the stand-in markers."""
    from pathlib import Path

    gold = torch.load(Path(__file__).parent / "golden" / "reference_concat_wiring.pt", weights_only=True)
    im = gold["image_metrics"]
    gt = torch.moveaxis(im["batch"]["image"], -1, 0)[None][:, 3][None]
    pr = torch.moveaxis(im["outputs"]["rgb"], -1, 0)[None][:, 3][None]
    assert gt.shape == (1, 1, 16, 14)
    m = im["metrics"]
    assert set(m) == {"psnr", "ssim", "lpips", "mae_thermal_foreground", "mae_thermal"}
    assert m["psnr"] == pytest.approx(float(-10 * torch.log10(torch.mean((gt - pr) ** 2))), rel=1e-5)
    tmax, tmin = 33.085, 13.896
    assert m["mae_thermal"] == pytest.approx(M.mae_thermal(gt, pr, False, tmax, tmin).item(), rel=1e-5)
    assert m["mae_thermal_foreground"] == pytest.approx(M.mae_thermal(gt, pr, False, tmax, tmin, threshold=0.4).item(), rel=1e-5)
    assert m["ssim"] == pytest.approx(float(1.0 - (gt - 0.5 * pr).abs().sum() / 1000.0), rel=1e-5)
    g3, p3 = torch.repeat_interleave(gt, 3, dim=1), torch.repeat_interleave(pr, 3, dim=1)
    assert m["lpips"] == pytest.approx(float(((g3 - 0.25 * p3) ** 2).sum() / 1000.0), rel=1e-5)
    assert im["image_keys"] == ["accumulation", "depth", "img", "prop_depth_0", "prop_depth_1"]
    assert im["image_shapes"]["img"] == [16, 28, 4]  # ground truth | prediction side by side, all four channels
