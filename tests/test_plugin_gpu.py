"""KernelModelMixin over a FOREIGN module tree on the GPU: the plugin classes of nerfstudio_plugin.make_plugin_classes
built on a base whose modules are not this package's containers (here the oracle's nn.Modules, which carry
nerfstudio's ``implementation="torch"`` attribute layout and state_dict keys) - the situation of
``B200ThermalNerfModel(ThermalNerfModel)`` on a real nerfstudio install (tests/test_plugin_cpu.py runs that class
hierarchy itself, without a GPU).  Eval outputs, one training step (losses, gradients landing in the foreign modules'
``.grad``) and the field surface bound onto the foreign field are compared with the oracle."""

from __future__ import annotations

from types import SimpleNamespace

import numpy as np
import pytest
import torch
from torch import nn

pytestmark = pytest.mark.gpu

from tests.helpers import compare_outputs, make_trained_like, oracle_config  # noqa: E402


def _build(precision="fp32", num_images=6, seed=5):
    from oracle import OracleThermalNerf
    from thermo_nerf_b200.model import ThermalNerfModelConfig
    from thermo_nerf_b200.nerfstudio_plugin import make_plugin_classes

    ocfg = oracle_config(log2_field=14, log2_prop=11)
    oracle = OracleThermalNerf(ocfg, num_images, seed=seed)
    make_trained_like(oracle, seed)
    with torch.no_grad():  # temperature contrast
        oracle.field.mlp_thermal.layers[0].weight.mul_(6.0)
        oracle.field.mlp_thermal.layers[1].weight.mul_(4.0)
        oracle.field.field_head_thermal.net.weight.mul_(4.0)
        oracle.field.field_head_thermal.net.bias.fill_(0.45)

    class RefLike(nn.Module):  # nerfstudio Model protocol + the module tree of thermal_nerf_model.py:86-208
        def __init__(self, config, metadata, scene_box, num_train_data, **kw):
            if "thermal" not in metadata:
                raise ValueError("Thermal images not found in metadata.")
            super().__init__()
            self.config, self.scene_box, self.num_train_data = config, scene_box, num_train_data
            self.device_indicator_param = nn.Parameter(torch.empty(0))
            self.populate_modules()

        @property
        def device(self):
            return self.device_indicator_param.device

        def populate_modules(self):
            import copy

            o = copy.deepcopy(oracle)
            self.field, self.proposal_networks, self.camera_optimizer = o.field, o.proposal_networks, o.camera_optimizer
            self.field.use_contraction = True
            for p in self.proposal_networks:
                p.use_contraction = True
            cfg = self.config
            self.proposal_sampler = SimpleNamespace(
                _anneal=1.0, _steps_since_update=0, _step=0,
                update_sched=lambda s: np.clip(np.interp(s, [0, cfg.proposal_warmup], [0, cfg.proposal_update_every]), 1,
                                               cfg.proposal_update_every))
            self.density_fns = [p.density_fn for p in self.proposal_networks]

        def get_outputs(self, rb):
            raise AssertionError("the eager reference path must not run")

    M, C = make_plugin_classes(RefLike, ThermalNerfModelConfig, name="B200OverForeignModules")
    cfg = C(log2_hashmap_size=14, precision=precision, camera_optimizer_mode="off",
            proposal_net_args_list=[dict(a, use_linear=False) for a in ocfg.proposal_net_args_list])
    model = cfg.setup(metadata={"thermal": []}, scene_box=SimpleNamespace(aabb=torch.tensor([[-1.0, -1, -1], [1, 1, 1]])),
                      num_train_data=num_images).to("cuda:0")
    return oracle, model


def test_eval_and_field_surface_over_foreign_modules():
    from oracle import make_synthetic_rays
    from thermo_nerf_b200 import RayBundle, surface

    oracle, model = _build("fp32")
    model.eval()
    rays = make_synthetic_rays(384, num_images=6, seed=2)
    with torch.no_grad():
        ref = oracle.get_outputs(rays, training=False)
        out = model(RayBundle(origins=rays.origins.cuda(), directions=rays.directions.cuda(),
                              camera_indices=rays.camera_indices.cuda()))
    compare_outputs(out, ref, 2e-4)
    assert float(ref["thermal"].std()) > 0.02  # the temperature image has contrast
    # the surface functions were bound onto the foreign field / proposal networks
    assert model.field.get_density.__func__ is surface.field_get_density
    pos = (torch.rand(200, 3, generator=torch.Generator().manual_seed(0)) * 2 - 1).cuda()
    with torch.no_grad():
        d_ref, _ = oracle.field.get_density(pos.cpu())
        p_ref = oracle.proposal_networks[1].density_fn(pos.cpu())
    err = ((model.field.density_fn(pos).cpu() - d_ref).abs() - 2e-5 * d_ref.abs()).max()
    assert err <= 1e-6
    err = ((model.density_fns[1](pos).cpu() - p_ref).abs() - 2e-5 * p_ref.abs()).max()
    assert err <= 1e-6


def test_training_step_over_foreign_modules():
    from oracle import make_synthetic_rays
    from thermo_nerf_b200 import RayBundle

    oracle, model = _build("fp32")
    model.train()
    oracle.train()
    R = 256
    rays = make_synthetic_rays(R, num_images=6, seed=3)
    g = torch.Generator().manual_seed(1)
    gt_rgb, gt_th = torch.rand(R, 3, generator=g), torch.rand(R, 1, generator=g)
    jit = torch.rand(3, R, generator=g)
    # same stratified draws on both sides
    real_rand = torch.rand
    torch.rand = lambda *a, **k: jit.to(k.get("device", "cpu")) if tuple(a[0]) == (3, R) else real_rand(*a, **k)
    try:
        out = model(RayBundle(origins=rays.origins.cuda(), directions=rays.directions.cuda(),
                              camera_indices=rays.camera_indices.cuda()))
    finally:
        torch.rand = real_rand
    batch = {"image": gt_rgb.cuda(), "thermal": gt_th.cuda()}
    metrics = model.get_metrics_dict(out, batch)
    losses = model.get_loss_dict(out, batch, metrics)
    sum(losses.values()).backward()
    ref_out = oracle.get_outputs(rays, training=True, jitter=jit[..., None])
    ref_losses = oracle.get_loss_dict(ref_out, gt_rgb, gt_th, training=True)
    sum(ref_losses.values()).backward()
    for k in ("rgb_loss", "interlevel_loss", "distortion_loss", "thermal"):
        a, b = float(losses[k]), float(ref_losses[k])
        assert abs(a - b) <= 2e-4 * max(1.0, abs(b)), (k, a, b)
    # gradients arrive in the foreign modules' parameters
    pairs = [(model.field.mlp_base.encoder.hash_table, oracle.field.mlp_base.encoder.hash_table),
             (model.field.mlp_head.layers[1].weight, oracle.field.mlp_head.layers[1].weight),
             (model.field.field_head_thermal.net.weight, oracle.field.field_head_thermal.net.weight),
             (model.proposal_networks[0].mlp_base[1].layers[0].weight, oracle.proposal_networks[0].mlp_base[1].layers[0].weight)]
    # fp32 kernels; the accumulation order of the 256-ray sums differs from autograd's (atomics), and the sample
    # positions behind the proposal gradients sit on hash-cell boundaries now and then: 1e-2 relative L2 per tensor
    for i, (p, q) in enumerate(pairs):
        assert p.grad is not None
        rel = float((p.grad.cpu() - q.grad).norm() / q.grad.norm().clamp_min(1e-12))
        assert rel <= 1e-2, (i, rel)
    assert model.proposal_sampler._steps_since_update == 0  # an "updated" step (step < 10) reset the sampler state
