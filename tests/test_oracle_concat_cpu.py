"""Oracle restatement of the reference's `concat_nerf` baseline: one 4-channel RGBT colour
head (rgb_concat/concat_field.py:65-75), RGBTRenderer with its default "random" background
(rgb_concat/rgbt_renderer.py:24,63-71), and the loss of rgb_concat/concat_nerfacto_model.py:197-233.
CPU only; the CUDA path for this head is not built yet."""

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
