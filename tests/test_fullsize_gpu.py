"""Parity at BASELINE.json's full sizes (2^19 x 16 field table, 2^17 x 5 proposal tables, 4096 - 65536 rays,
800x800 frames) through properties that do not need the oracle to run at that size:

* ray-permutation equivariance and call-splitting independence (bit-exact),
* the compositing invariants (weights >= 0, sum = accumulation <= 1, outputs inside [0,1], monotone bins),
* a closed-form gradient: d loss / d (thermal head output bias) = sum_r dL/d thermal_r, because the
  compositing weights plus the last-sample background weight add up to one on every ray,
* a 4096-ray slice of the full-size problem against the oracle itself.
"""

import pytest
import torch

from oracle import OracleRays, make_synthetic_rays
from tests.helpers import compare_outputs, make_pair

pytestmark = pytest.mark.gpu

OUT_KEYS = ("rgb", "thermal", "depth", "expected_depth", "accumulation", "prop_depth_0", "prop_depth_1")


@pytest.fixture(scope="module")
def full():
    oracle, model = make_pair(log2_field=19, log2_prop=17, num_images=100, trained_like=True, precision="tc_fp16")
    return oracle, model


def _fwd(model, o, d, **kw):
    from thermo_nerf_b200 import _lib as L
    from thermo_nerf_b200 import functional as F

    res = F.render_forward(model.tensors(), o, d, near_plane=0.0, far_plane=1000.0, appearance_mode=L.APPEARANCE_MEAN,
                           precision=L.PRECISION_TC_FP16, **kw)
    torch.cuda.synchronize()
    return res


def test_65536_rays_permutation_and_split_invariance(full):
    _, model = full
    rays = make_synthetic_rays(65536, num_images=100, seed=9)
    o, d = rays.origins.cuda(), rays.directions.cuda()
    a = _fwd(model, o, d)
    perm = torch.randperm(65536, generator=torch.Generator().manual_seed(1)).cuda()
    b = _fwd(model, o[perm].contiguous(), d[perm].contiguous())
    for k in OUT_KEYS:
        if k == "expected_depth":
            continue  # clipped to the call-global [min, max]: same set of rays -> same range
        assert torch.equal(a[k][perm], b[k]), k
    assert torch.equal(a["expected_depth"][perm], b["expected_depth"])
    # two half calls == one call (expected depth aside: its clip range is per call by definition)
    h1, h2 = _fwd(model, o[:32768].contiguous(), d[:32768].contiguous()), _fwd(model, o[32768:].contiguous(),
                                                                               d[32768:].contiguous())
    for k in OUT_KEYS:
        if k != "expected_depth":
            assert torch.equal(torch.cat([h1[k], h2[k]]), a[k]), k
    # invariants of the renderers
    assert float(a["rgb"].min()) >= 0.0 and float(a["rgb"].max()) <= 1.0
    assert float(a["thermal"].min()) >= 0.0 and float(a["thermal"].max()) <= 1.0
    assert float(a["accumulation"].min()) >= 0.0 and float(a["accumulation"].max()) <= 1.0 + 1e-5
    assert torch.isfinite(a["depth"]).all() and float(a["depth"].min()) >= 0.0


def test_full_frame_800x800_camera_vs_chunked_bundle(full):
    """One 640 000-ray launch from the camera == the reference's chunked loop over the generated rays."""
    from thermo_nerf_b200 import RayBundle, orbit_cameras
    from thermo_nerf_b200 import functional as F

    _, model = full
    cams = orbit_cameras(3, hw=800, focal=1111.1)
    out = model.get_outputs_for_camera(cams, 1)
    cam = F.pack_camera(cams.camera_to_worlds[1], cams.fx, cams.fy, cams.cx, cams.cy, 800, 800)
    o, d, _ = F.generate_rays(cam, "cuda:0")
    chunk = model.config.eval_num_rays_per_chunk
    parts = []
    for s in range(0, 640000, chunk):  # nerfstudio Model.get_outputs_for_camera_ray_bundle's loop
        with torch.no_grad():
            parts.append(model.get_outputs(RayBundle(origins=o[s:s + chunk], directions=d[s:s + chunk])))
    for k in OUT_KEYS:
        ref = torch.cat([p[k] for p in parts]).view(800, 800, -1)
        assert torch.equal(out[k], ref), k


def test_training_invariants_and_closed_form_gradient_8192(full):
    from thermo_nerf_b200 import _lib as L
    from thermo_nerf_b200 import functional as F

    _, model = full
    R = 8192
    rays = make_synthetic_rays(R, num_images=100, seed=12)
    o, d, cam = rays.origins.cuda(), rays.directions.cuda(), rays.camera_indices.cuda().reshape(-1)
    gen = torch.Generator().manual_seed(3)
    jitter = torch.rand((3, R), generator=gen).cuda()
    gt_rgb, gt_th = torch.rand((R, 3), generator=gen).cuda(), torch.rand((R,), generator=gen).cuda()
    t = model.tensors()
    res = F.render_forward(t, o, d, cam, None, None, jitter, training=True, return_samples=True, save_for_backward=True,
                           anneal=0.6, appearance_mode=L.APPEARANCE_LOOKUP, precision=L.PRECISION_TC_FP16)
    for k, S in enumerate((256, 96, 48)):
        w, sd = res["weights_list"][k].reshape(R, S), res["sdist_list"][k]
        assert float(w.min()) >= 0.0 and float(w.sum(1).max()) <= 1.0 + 1e-4
        assert bool((sd[:, 1:] >= sd[:, :-1]).all()) and float(sd.min()) >= 0.0 and float(sd.max()) <= 1.0
    assert torch.allclose(res["weights_list"][2].reshape(R, 48).sum(1), res["accumulation"].reshape(R), atol=1e-5)
    losses, g = F.losses_forward_backward(res["weights_list"], res["sdist_list"], res["rgb"], res["thermal"], gt_rgb, gt_th)
    # MSE gradients in closed form
    assert torch.allclose(g["thermal"], 2.0 * (res["thermal"].reshape(R) - gt_th) / R, atol=1e-9, rtol=1e-5)
    assert torch.allclose(g["rgb"], 2.0 * (res["rgb"] - gt_rgb) / (3 * R), atol=1e-9, rtol=1e-5)
    grads = [torch.zeros_like(p) for p in t.param_list()]
    res["_workspace"] = None
    F.render_backward(t, res["_model_struct"], o, d, cam, None, None, jitter, res,
                      {"rgb": g["rgb"], "thermal": g["thermal"], "weights_list": g["weights_list"]}, grads)
    torch.cuda.synchronize()
    names = ["p%d" % i for i in range(10)] + ["table"] + [f"{k}.{s}" for k in F.ModelTensors.FIELD_ORDER for s in "wb"] + ["app"]
    gmap = dict(zip(names, grads))
    # thermal = sum_s w_s tau_s + (1 - sum_s w_s) tau_last and tau = th2(.) + b  =>  d thermal_r / d b = 1
    want = float(g["thermal"].double().sum())
    got = float(gmap["th2.b"].double().sum())
    assert abs(got - want) <= 2e-3 * abs(want) + 1e-9, (got, want)
    for n, gr in gmap.items():
        assert torch.isfinite(gr).all(), n
    assert float(gmap["table"].abs().sum()) > 0 and float(gmap["p0"].abs().sum()) > 0


def test_4096_ray_slice_of_the_full_size_problem_matches_oracle(full):
    oracle, model = full
    from thermo_nerf_b200 import RayBundle

    rays = make_synthetic_rays(4096, num_images=100, seed=21)
    with torch.no_grad():
        ref = oracle.get_outputs(OracleRays(rays.origins, rays.directions, rays.camera_indices), training=False)
        out = model.get_outputs(RayBundle(origins=rays.origins.cuda(), directions=rays.directions.cuda(),
                                          camera_indices=rays.camera_indices.cuda()))
    torch.cuda.synchronize()
    compare_outputs(out, ref, 2e-2, median_bad_frac=0.05)


# ------------------------------------------------------------------------------------------------------------------
# the oracle itself at full size, run on the GPU inside the test (eager PyTorch: ~0.1 s per 65 536-ray chunk)
# ------------------------------------------------------------------------------------------------------------------
def _oracle_on_gpu(oracle):
    import copy

    return copy.deepcopy(oracle).to("cuda:0")


def test_65536_ray_chunk_matches_the_oracle_run_on_the_gpu(full):
    """One whole eval chunk (eval_num_rays_per_chunk = 1 << 16, config_thermal_nerf.py:30) of contiguous frame pixels,
    full-size tables, tensor-core mode against the fp32 oracle: 2e-2 abs, PSNR >= 40 dB, temperature MAE <= 4e-3.

    Measured on B200: the temperature image sits a constant 2.3e-3 above the fp32 oracle's (spread 1.7e-4).  That
    offset is the rounding of the field's weights to fp16 seen through the test weights' gain of a few hundred
    (helpers.add_thermal_contrast): with fp16-representable weights on both sides it is 5e-7, and rounding the
    activations alone (emulated in the oracle) gives none.  The reference's tcnn networks hold fp16 weights too."""
    oracle, model = full
    og = _oracle_on_gpu(oracle).eval()
    rays = make_synthetic_rays(1 << 16, num_images=100, seed=13, contiguous_pixels=True)
    o, d = rays.origins.cuda(), rays.directions.cuda()
    with torch.no_grad():
        ref = og.get_outputs(OracleRays(o, d, rays.camera_indices.cuda()), training=False)
    out = _fwd(model, o, d)
    compare_outputs({k: v.reshape(ref[k].shape) for k, v in out.items() if k in ref}, ref, 2e-2, median_bad_frac=0.05,
                    thermal_contrast=True)
    mse = torch.mean((out["rgb"] - ref["rgb"]) ** 2).item()
    assert -10 * torch.log10(torch.tensor(mse + 1e-20)).item() >= 40.0
    assert (out["thermal"].reshape(-1) - ref["thermal"].reshape(-1)).abs().mean().item() <= 4e-3


def test_800x800_frame_matches_the_oracle_run_on_the_gpu_chunk_by_chunk(full):
    """A whole 800x800 frame through get_outputs_for_camera (one launch, rays generated in the kernel) against the
    oracle driven the way nerfstudio's get_outputs_for_camera_ray_bundle drives the reference: 65 536-ray slices of the
    frame's rays, each with its own expected-depth clip range."""
    from thermo_nerf_b200 import orbit_cameras

    oracle, model = full
    og = _oracle_on_gpu(oracle).eval()
    cams = orbit_cameras(3, hw=800, focal=1111.1)
    model.eval()
    with torch.no_grad():
        out = model.get_outputs_for_camera(cams, 1)
        rb = cams.generate_rays(1).to("cuda:0").flatten()
        parts = []
        for s in range(0, 640000, 1 << 16):
            parts.append(og.get_outputs(OracleRays(rb.origins[s:s + (1 << 16)], rb.directions[s:s + (1 << 16)],
                                                   rb.camera_indices[s:s + (1 << 16)]), training=False))
    ref = {k: torch.cat([p[k] for p in parts]) for k in ("rgb", "thermal", "accumulation", "expected_depth", "depth",
                                                          "prop_depth_0", "prop_depth_1")}
    got = {k: out[k].reshape(ref[k].shape) for k in ref}
    compare_outputs(got, ref, 2e-2, median_bad_frac=0.05, thermal_contrast=True)
    mse = torch.mean((got["rgb"] - ref["rgb"]) ** 2).item()
    assert -10 * torch.log10(torch.tensor(mse + 1e-20)).item() >= 40.0


@pytest.mark.parametrize("R", [4096, 8192])
def test_fullsize_training_gradients_match_oracle_autograd_on_the_gpu(full, R):
    """BASELINE configs[1] / [2] shapes: a full training iteration's gradients (2^19 x 16 and 2^17 x 5 tables, 4096 and
    8192 rays, tensor-core mode) against the oracle's autograd run in fp32 on the GPU: relative L2 <= 5e-2 and cosine
    >= 0.998 per parameter tensor (fp16 forward / bf16 backward operands), losses to 3e-2 relative (the fp16 weight
    rounding moves the temperature image by ~2e-3, i.e. ~2% of a loss whose residual is 0.25).

    The temperature targets sit a fixed 0.25 above the rendered temperatures.  With targets drawn at random the
    residuals change sign from ray to ray, the temperature chain's gradient collapses by cancellation to the size of
    (forward rounding error) x (Jacobian), and the comparison then measures the tensor-core *forward's* 1-2e-3
    temperature error against a vanishing gradient (measured: 0.3 - 0.9 relative on the temperature MLP) instead of
    the backward kernel.  The colour targets stay random."""
    from thermo_nerf_b200 import _lib as L
    from thermo_nerf_b200 import functional as F

    oracle, model = full
    og = _oracle_on_gpu(oracle).train()
    og.camera_optimizer.mode = "off"  # F.render below takes the rays as they are
    og.anneal = 1.0
    rays = make_synthetic_rays(R, num_images=100, seed=17)
    g = torch.Generator().manual_seed(3)
    jitter = torch.rand((3, R, 1), generator=g).cuda()
    gt_rgb, gt_th = torch.rand((R, 3), generator=g).cuda(), torch.rand((R, 1), generator=g).cuda()
    o, d, cam = rays.origins.cuda(), rays.directions.cuda(), rays.camera_indices.cuda()
    og.zero_grad()
    ref_out = og.get_outputs(OracleRays(o, d, cam), training=True, jitter=jitter)
    gt_th = (ref_out["thermal"].detach() + 0.25).clamp(0.0, 1.0)
    ref_ld = og.get_loss_dict(ref_out, gt_rgb, gt_th, training=True)
    sum(ref_ld.values()).backward()
    model.train()
    model.zero_grad()
    out = F.render(model.tensors(), o, d, cam.reshape(-1), None, None, jitter.reshape(3, -1), num_samples=(256, 96, 48),
                   near_plane=0.05, far_plane=1000.0, anneal=1.0, appearance_mode=L.APPEARANCE_LOOKUP,
                   precision=L.PRECISION_TC_FP16)
    ld = F.losses(out, gt_rgb, gt_th.reshape(-1))
    sum(ld.values()).backward()
    torch.cuda.synchronize()
    for k in ref_ld:
        a, b = float(ld[k].detach()), float(ref_ld[k].detach())
        assert abs(a - b) <= 3e-2 * max(abs(b), 1e-3), (k, a, b)
    ref_params = dict(og.named_parameters())
    checked, bad = 0, []
    for name, p in model.named_parameters():
        q = ref_params.get(name)
        if q is None or q.grad is None or p.grad is None or name.startswith("camera_optimizer"):
            continue
        a, b = p.grad.flatten().double(), q.grad.flatten().double()
        if float(b.norm()) < 1e-12:
            continue
        rel = float((a - b).norm() / b.norm())
        cos = float(torch.dot(a, b) / (a.norm() * b.norm()).clamp_min(1e-30))
        print(f"{name}: rel-L2 {rel:.2e} cos {cos:.5f}")
        if rel > 5e-2 or cos < 0.998:
            bad.append((name, rel, cos))
        checked += 1
    assert not bad, bad
    assert checked >= 24
    model.eval()
    model.zero_grad()
