"""The CUDA path (through the plugin surface and the C ABI) against the committed golden fixtures under
tests/golden/ - fixed (rays, weights seed) -> outputs vectors, so a drift of either side is caught without
running the oracle.  fp32 mode: |err| <= 2e-4 (rgb/thermal/accumulation), spacing bins <= 2e-5; the 16/8/16-sample
plumbing case gets 5e-4 (16 coarse samples: each carries a large delta*sigma, so a last-ulp difference in a bin
edge moves the weights more than at 256/96/48)."""

from pathlib import Path

import pytest
import torch

from tests.golden.make_golden import CASES, weights_checksum
from tests.helpers import compare_outputs, make_pair

pytestmark = pytest.mark.gpu
GOLDEN = Path(__file__).parent / "golden"


@pytest.mark.parametrize("name", sorted(CASES))
def test_cuda_path_reproduces_golden(name):
    from thermo_nerf_b200 import RayBundle

    R, ns, lf, lp, nimg, trained, training = CASES[name]
    blob = torch.load(GOLDEN / f"{name}.pt", weights_only=True)
    oracle, model = make_pair(log2_field=lf, log2_prop=lp, num_images=nimg, seed=0, trained_like=trained,
                              num_samples=ns, precision="fp32")
    if abs(weights_checksum(oracle) - blob["weights_checksum"]) > 1e-6 * abs(blob["weights_checksum"]):
        pytest.skip("torch CPU RNG stream differs from the one the golden was generated with")
    o, d, cam = blob["origins"].cuda(), blob["directions"].cuda(), blob["camera_indices"].cuda()
    if training:
        from thermo_nerf_b200 import _lib as L
        from thermo_nerf_b200 import functional as F

        model.train()
        # training-mode get_outputs first applies the camera optimiser (thermal_nerf_model.py:218-219); the golden's
        # jitter draws are then passed explicitly (get_outputs itself would draw fresh ones)
        rb = RayBundle(origins=o, directions=d, camera_indices=cam)
        with torch.no_grad():
            model.camera_optimizer.apply_to_raybundle(rb)
        o, d = rb.origins.contiguous(), rb.directions.contiguous()
        res = F.render_forward(model.tensors(), o, d, cam.reshape(-1), None, None, blob["jitter"].cuda().reshape(3, -1),
                               num_samples=ns, training=True, near_plane=model.config.near_plane,
                               far_plane=model.config.far_plane, anneal=blob["anneal"],
                               appearance_mode=L.APPEARANCE_LOOKUP, precision=L.PRECISION_FP32, return_samples=True)
        out = {k: (v.reshape(R, -1) if isinstance(v, torch.Tensor) else v) for k, v in res.items()}
    else:
        with torch.no_grad():
            out = model.get_outputs(RayBundle(origins=o, directions=d, camera_indices=cam))
    torch.cuda.synchronize()
    ref = blob["outputs"]
    compare_outputs(out, ref, 5e-4 if name == "e2e_mini_r32" else 2e-4)
    if training:
        for k in range(3):
            assert torch.allclose(out["sdist_list"][k].cpu(), ref["sdist_list"][k], atol=2e-5), k
            w, wr = out["weights_list"][k].cpu().reshape(R, -1), ref["weights_list"][k].reshape(R, -1)
            assert (w - wr).abs().max().item() <= 2e-4, k


@pytest.mark.parametrize("idx", [0, 1, 2])
def test_cuda_path_reproduces_the_reference_model_vectors(idx):
    """tests/golden/reference_model_wiring.pt: outputs of the REFERENCE'S OWN ThermalNerfModel / ThermalNerfactoTField /
    ThermalRenderer code executed over nerfstudio stand-ins (tests/golden/make_reference_wiring_golden.py) on a
    24/16/12-sample configuration, 48 rays, trained-like weights.  fp32 mode; 1e-3 abs (measured on B200: 2.6e-4 eval,
    8.9e-5 train): with 12 field samples per ray a last-ulp difference in a bin edge moves the accumulation by ~3e-4
    already between an fp32 and an fp64 evaluation of the oracle itself.  48 rays: up to 3 median-depth flips to the
    neighbouring sample are tolerated.  Case 2 has the geometry of case 0 and a temperature head with contrast
    (rendered temperatures 0.42-0.62 instead of an almost constant image that the eval clamp flattens to 0)."""
    from thermo_nerf_b200 import RayBundle, ThermalNerfModel, ThermalNerfModelConfig
    from thermo_nerf_b200 import _lib as L
    from thermo_nerf_b200 import functional as F

    blob = torch.load(GOLDEN / "reference_model_wiring.pt", weights_only=True)
    case, mini = blob["cases"][idx], blob["mini"]
    ns = (*mini["num_proposal_samples_per_ray"], mini["num_nerf_samples_per_ray"])
    cfg = ThermalNerfModelConfig(log2_hashmap_size=mini["log2_hashmap_size"], num_proposal_samples_per_ray=ns[:2],
                                 num_nerf_samples_per_ray=ns[2], proposal_net_args_list=mini["proposal_net_args_list"],
                                 pass_thermal_gradients=case["pass_thermal_gradients"], precision="fp32")
    model = ThermalNerfModel(cfg, {"thermal": []}, torch.tensor([[-1.0, -1, -1], [1, 1, 1]]), blob["num_images"])
    missing, unexpected = model.load_state_dict(case["state_dict"], strict=False)
    assert set(missing) <= {"device_indicator_param"} and not unexpected
    model = model.cuda()
    o, d, cam = case["origins"].cuda(), case["directions"].cuda(), case["camera_indices"].cuda()
    R = o.shape[0]
    # ---- eval mode through the plugin surface
    model.eval()
    with torch.no_grad():
        out = model.get_outputs(RayBundle(origins=o, directions=d, camera_indices=cam))
    torch.cuda.synchronize()
    worst_eval = compare_outputs(out, case["eval"]["outputs"], 1e-3, median_bad_frac=0.07)
    # ---- training mode: camera optimiser first, the vectors' jitter draws and anneal passed explicitly
    model.train()
    rb = RayBundle(origins=o.clone(), directions=d.clone(), camera_indices=cam)
    with torch.no_grad():
        model.camera_optimizer.apply_to_raybundle(rb)
    tr = case["train"]
    res = F.render_forward(model.tensors(), rb.origins.contiguous(), rb.directions.contiguous(), cam.reshape(-1), None, None,
                           case["jitter"].cuda().reshape(3, -1), num_samples=ns, training=True,
                           near_plane=model.config.near_plane, far_plane=model.config.far_plane, anneal=tr["anneal"],
                           appearance_mode=L.APPEARANCE_LOOKUP, precision=L.PRECISION_FP32, return_samples=True)
    torch.cuda.synchronize()
    out = {k: (v.reshape(R, -1) if isinstance(v, torch.Tensor) else v) for k, v in res.items()}
    worst_train = compare_outputs(out, tr["outputs"], 1e-3, median_bad_frac=0.07)
    for k in range(3):
        assert torch.allclose(out["sdist_list"][k].cpu(), tr["spacing_bins"][k], atol=5e-5), k
        w, wr = out["weights_list"][k].cpu().reshape(R, -1), tr["weights_list"][k].reshape(R, -1)
        assert (w - wr).abs().max().item() <= 1e-3, k
    print(f"[reference vectors] case {idx}: worst |err| eval {worst_eval:.3e}, train {worst_train:.3e}")
