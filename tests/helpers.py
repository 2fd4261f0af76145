"""Shared test helpers: build an oracle model (CPU) and the B200 model with identical
weights, and compare output dicts."""

from __future__ import annotations

from typing import Dict, Sequence, Tuple

import torch

from oracle import OracleConfig, OracleThermalNerf


def oracle_config(log2_field=15, log2_prop=12, num_samples: Sequence[int] = (256, 96, 48), contraction=True,
                  **kw) -> OracleConfig:
    cfg = OracleConfig(
        log2_hashmap_size=log2_field,
        num_proposal_samples_per_ray=tuple(num_samples[:2]),
        num_nerf_samples_per_ray=num_samples[2],
        proposal_net_args_list=[
            {"hidden_dim": 16, "log2_hashmap_size": log2_prop, "num_levels": 5, "max_res": 128},
            {"hidden_dim": 16, "log2_hashmap_size": log2_prop, "num_levels": 5, "max_res": 256},
        ],
        disable_scene_contraction=not contraction,
        **kw,
    )
    return cfg


def make_trained_like(oracle: OracleThermalNerf, seed: int = 0) -> None:
    """Non-trivial densities: N(0, 0.5^2) hash entries, sharpened density outputs."""
    g = torch.Generator().manual_seed(seed + 1234)
    with torch.no_grad():
        for enc in [oracle.field.mlp_base.encoder] + [p.encoding for p in oracle.proposal_networks]:
            enc.hash_table.copy_(torch.randn(enc.hash_table.shape, generator=g) * 0.5)
        for p in oracle.proposal_networks:
            p.mlp_base[1].layers[1].weight.mul_(6.0)
        oracle.field.mlp_base.mlp.layers[1].weight[0].mul_(6.0)
        oracle.camera_optimizer.pose_adjustment.copy_(
            torch.randn(oracle.camera_optimizer.pose_adjustment.shape, generator=g) * 1e-2)


def add_thermal_contrast(oracle: OracleThermalNerf, target_std: float = 0.08, s0: float = 6.0, s1: float = 4.0) -> None:
    """A freshly initialised temperature head is almost constant (about -0.05, flattened to 0 by the eval clamp), which
    would make a forward parity bound on the temperature image vacuous.  Spread the temperature MLP's response, then
    calibrate the head (an affine map of its output is an affine map of the rendered image wherever accumulation is 1)
    on a probe batch so that rendered temperatures sit around 0.5 with a standard deviation of ``target_std``.

    The temperature then depends on the geometry features through a gain of a few hundred, and with random tables
    ("trained-like": every level, including the 2048-cell one, holds N(0, 0.5) features) those features move by ~1e-3
    relative when a sample position moves by one fp32 ulp.  Measured on B200, fp32 mode: max temperature error
    ~8e-4 x the image's standard deviation (rgb, which barely depends on the geometry features: 2e-7)."""
    if not hasattr(oracle.field, "mlp_thermal"):
        return
    from oracle import make_synthetic_rays

    head = oracle.field.field_head_thermal.net
    with torch.no_grad():
        oracle.field.mlp_thermal.layers[0].weight.mul_(s0)
        oracle.field.mlp_thermal.layers[1].weight.mul_(s1)
        probe = make_synthetic_rays(256, num_images=oracle.field.embedding_appearance.embedding.num_embeddings, seed=1234)
        mode, oracle.camera_optimizer.mode = oracle.camera_optimizer.mode, "off"  # same calibration in every pose mode
        image = oracle.get_outputs(probe, training=True)["thermal"].float()
        oracle.camera_optimizer.mode = mode
        gain = target_std / max(float(image.std()), 1e-6)
        head.weight.mul_(gain)
        head.bias.copy_(head.bias * gain + (0.5 - gain * float(image.mean())))


def make_pair(log2_field=15, log2_prop=12, num_images=8, seed=0, trained_like=True, device="cuda:0",
              num_samples: Sequence[int] = (256, 96, 48), contraction=True, precision="fp32", thermal_contrast=True,
              contrast_args=(), **kw):
    from thermo_nerf_b200 import ThermalNerfModel, ThermalNerfModelConfig

    cfg = oracle_config(log2_field, log2_prop, num_samples, contraction, **kw)
    oracle = OracleThermalNerf(cfg, num_images, seed=seed)
    if trained_like:
        make_trained_like(oracle, seed)
        if thermal_contrast:
            add_thermal_contrast(oracle, *contrast_args)
    mcfg = ThermalNerfModelConfig(
        log2_hashmap_size=log2_field,
        num_proposal_samples_per_ray=tuple(num_samples[:2]),
        num_nerf_samples_per_ray=num_samples[2],
        proposal_net_args_list=[dict(a, use_linear=False) for a in cfg.proposal_net_args_list],
        disable_scene_contraction=not contraction,
        use_average_appearance_embedding=cfg.use_average_appearance_embedding,
        camera_optimizer_mode=cfg.camera_optimizer_mode,
        pass_thermal_gradients=cfg.pass_thermal_gradients,
        precision=precision,
    )
    model = ThermalNerfModel(mcfg, {"thermal": []}, torch.tensor([[-1.0, -1, -1], [1, 1, 1]]), num_images)
    missing, unexpected = model.load_state_dict(oracle.state_dict(), strict=False)
    assert set(missing) <= {"device_indicator_param"}, missing
    assert not unexpected, unexpected
    model = model.to(device)
    model.eval()
    return oracle, model


def compare_outputs(out: Dict[str, torch.Tensor], ref: Dict[str, torch.Tensor], tol: float,
                    depth_rtol: float = None, median_bad_frac: float = 0.02, thermal_contrast: bool = False) -> float:
    """Asserts parity; returns the worst abs error over rgb/thermal/accumulation."""
    worst = 0.0
    if thermal_contrast:
        # the bound on the temperature image below is only meaningful if that image has contrast
        assert float(ref["thermal"].float().std()) > 0.02, "test weights give an almost constant temperature image"
    for k in ("rgb", "thermal", "accumulation"):
        a, b = out[k].detach().float().cpu(), ref[k].detach().float().cpu()
        assert a.shape == b.shape, (k, a.shape, b.shape)
        err = (a - b).abs().max().item()
        # temperature: plus 2e-3 of the image's standard deviation, the fp32 sample-position sensitivity described in
        # add_thermal_contrast (1.6e-4 at the default contrast; nothing for a constant image)
        bound = tol + (2e-3 * float(b.std()) if k == "thermal" and b.numel() > 1 else 0.0)
        assert err <= bound, f"{k}: max abs err {err:.3e} > {bound:.3e}"
        worst = max(worst, err)
    depth_rtol = depth_rtol if depth_rtol is not None else max(10 * tol, 1e-4)
    a, b = out["expected_depth"].detach().cpu(), ref["expected_depth"].detach().cpu()
    rel = ((a - b).abs() / b.abs().clamp_min(1e-3)).max().item()
    assert rel <= depth_rtol, f"expected_depth: max rel err {rel:.3e} > {depth_rtol}"
    # median depths pick a sample index: a last-ulp difference in the cumulative weight can flip to
    # the neighbouring sample, so bound the *fraction* of rays that differ noticeably.
    for k in ("depth", "prop_depth_0", "prop_depth_1"):
        a, b = out[k].detach().cpu(), ref[k].detach().cpu()
        assert a.shape == b.shape, (k, a.shape, b.shape)
        bad = (((a - b).abs() / b.abs().clamp_min(1e-3)) > depth_rtol).float().mean().item()
        assert bad <= median_bad_frac, f"{k}: {bad:.3%} rays differ"
    return worst
