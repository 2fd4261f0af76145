"""The Field / Renderer plugin surface on the GPU against the oracle: ThermalNerfactoTField.get_density /
get_outputs / forward / density_fn (thermo_nerf/thermal_nerf/thermal_field.py:108-201), the proposal networks'
density_fn (thermal_nerf_model.py:127-148), ThermalRenderer.forward (thermal_renderer.py:113-149) and RGBTRenderer.forward
(rgb_concat/rgbt_renderer.py:134-174).  fp32 kernels: tolerance 2e-5 relative on densities (exp of an fp32 MLP output),
2e-5 absolute on colours / temperatures / geo features."""

from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import pytest
import torch

pytestmark = pytest.mark.gpu

from tests.helpers import make_pair  # noqa: E402


@dataclass
class _Frustums:
    origins: torch.Tensor
    directions: torch.Tensor
    starts: torch.Tensor
    ends: torch.Tensor

    def get_positions(self):
        return self.origins + self.directions * (self.starts + self.ends) / 2


@dataclass
class _RaySamples:
    frustums: _Frustums
    camera_indices: Optional[torch.Tensor] = None


def _samples(R=257, S=11, num_images=8, seed=0, device="cuda:0"):
    g = torch.Generator().manual_seed(seed)
    o = (torch.rand(R, 1, 3, generator=g) * 2 - 1) * 0.7
    d = torch.nn.functional.normalize(torch.randn(R, 1, 3, generator=g), dim=-1)
    starts = torch.sort(torch.rand(R, S, 1, generator=g) * 3.0, dim=1).values  # up to |x| ~ 3.7: contraction is exercised
    ends = starts + 0.05
    cam = torch.randint(0, num_images, (R, 1, 1), generator=g).expand(R, S, 1).contiguous()
    fr = _Frustums(o.expand(R, S, 3).contiguous().to(device), d.expand(R, S, 3).contiguous().to(device),
                   starts.to(device), ends.to(device))
    return _RaySamples(fr, cam.to(device))


def _close(a, b, atol, rtol=0.0, name=""):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    assert a.shape == b.shape, (name, a.shape, b.shape)
    err = ((a - b).abs() - rtol * b.abs()).max().item()
    assert err <= atol, f"{name}: |err| - rtol |ref| = {err:.3e} > {atol}"


@pytest.mark.parametrize("contraction", [True, False])
def test_field_surface_matches_oracle(contraction):
    from thermo_nerf_b200 import FieldHeadNames, FieldHeadNamesT

    oracle, model = make_pair(log2_field=14, log2_prop=11, num_images=8, seed=3, contraction=contraction)
    rs = _samples()
    pos = rs.frustums.get_positions()
    with torch.no_grad():
        ref_density, ref_geo = oracle.field.get_density(pos.cpu())
    density, geo = model.field.get_density(rs)
    assert density.shape == (*pos.shape[:-1], 1) and geo.shape == (*pos.shape[:-1], 15)
    _close(density, ref_density, 1e-6, 2e-5, "density")
    _close(geo, ref_geo, 2e-5, 2e-5, "geo")
    _close(model.field.density_fn(pos), ref_density, 1e-6, 2e-5, "density_fn")
    for training in (False, True):
        oracle.train(training)
        model.train(training)
        with torch.no_grad():
            ref_rgb, ref_th = oracle.field.get_outputs(rs.frustums.directions.cpu(), rs.camera_indices.cpu()[..., 0],
                                                       ref_geo, training=training)
            ref = {"rgb": ref_rgb, "thermal": ref_th}
            out = model.field.get_outputs(rs, density_embedding=ref_geo.to(pos.device))
            fwd = model.field(rs)
        _close(out[FieldHeadNames.RGB], ref["rgb"], 2e-5, name=f"rgb training={training}")
        _close(out[FieldHeadNamesT.THERMAL], ref["thermal"], 2e-5, name=f"thermal training={training}")
        assert set(fwd) == {FieldHeadNames.RGB, FieldHeadNamesT.THERMAL, FieldHeadNames.DENSITY}
        _close(fwd[FieldHeadNames.DENSITY], ref_density, 1e-6, 2e-5, "forward density")
        _close(fwd[FieldHeadNames.RGB], ref["rgb"], 5e-5, name="forward rgb")
    model.eval()
    # proposal networks: density_fn (what ProposalNetworkSampler calls) and the model-level list of them
    for i, net in enumerate(model.proposal_networks):
        with torch.no_grad():
            ref_p = oracle.proposal_networks[i].density_fn(pos.cpu())
        _close(net.density_fn(pos), ref_p, 1e-6, 2e-5, f"prop {i}")
        _close(model.density_fns[i](pos), ref_p, 1e-6, 2e-5, f"density_fns[{i}]")


def test_field_surface_error_behaviour():
    _, model = make_pair(log2_field=12, log2_prop=10, num_images=4, seed=0)
    rs = _samples(R=8, S=4, num_images=4)
    geo = torch.zeros(8, 4, 15, device="cuda:0")
    with pytest.raises(AssertionError):
        model.field.get_outputs(rs, density_embedding=None)                     # thermal_field.py:111
    with pytest.raises(AttributeError, match="Camera indices are not provided"):
        model.field.get_outputs(_RaySamples(rs.frustums, None), density_embedding=geo)  # thermal_field.py:113-114
    model.train()
    with pytest.raises(RuntimeError, match="inference only"):
        model.field.get_density(rs)                                             # no silent no-grad training
    with torch.no_grad():
        model.field.get_density(rs)
    with pytest.raises(ValueError):
        model.field(rs, compute_normals=True)


def test_renderers_match_the_reference_executed_vectors():
    """ThermalRenderer / RGBTRenderer against the vectors the reference's OWN modules produced
    (tests/golden/reference_renderers.pt, made by make_reference_renderer_golden.py from
    thermo_nerf/thermal_nerf/thermal_renderer.py and thermo_nerf/rgb_concat/rgbt_renderer.py): train / eval mode,
    NaN / inf and out-of-range samples, empty and opaque rays."""
    from pathlib import Path

    from thermo_nerf_b200.surface import RGBTRenderer, ThermalRenderer

    gold = torch.load(Path(__file__).parent / "golden" / "reference_renderers.pt", weights_only=False)
    checked = 0
    for kind, key, cls in (("thermal", "thermal", ThermalRenderer), ("rgbt", "rgbt", RGBTRenderer)):
        for case in gold[kind]:
            r = cls().train(bool(case["training"]))
            with torch.no_grad():
                out = r(case[key].cuda(), case["weights"].cuda())
            ref = case["out"]
            # training mode propagates NaN / inf samples in the reference: compare where the reference is finite
            finite = torch.isfinite(ref)
            assert torch.equal(torch.isfinite(out).cpu() | ~finite, torch.ones_like(finite)), (kind, case["training"])
            _close(torch.where(finite.cuda(), out, torch.zeros_like(out)),
                   torch.where(finite, ref, torch.zeros_like(ref)), 2e-6, 2e-6, f"{kind} training={case['training']}")
            checked += 1
    fb = gold["thermal_forced_background"]  # a background_color argument is ignored (thermal_renderer.py:49)
    with torch.no_grad():
        out = ThermalRenderer(background_color="black").train()(fb["thermal"].cuda(), fb["weights"].cuda(),
                                                               background_color="white")
    _close(out, fb["out"], 2e-6, 2e-6, "forced background")
    assert checked == 12
    with pytest.raises(NotImplementedError):
        ThermalRenderer().eval()(fb["thermal"].cuda(), fb["weights"].cuda(), ray_indices=torch.zeros(1), num_rays=1)
