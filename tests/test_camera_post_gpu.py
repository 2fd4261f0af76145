"""GPU parity of the neighbours of the path, through the C ABI: in-kernel ray
generation, the frame post-processing kernel and the Renderer mirror, against the numpy oracle."""

import numpy as np
import pytest
import torch

import oracle
from tests.helpers import make_pair

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


def _cameras(n=3, hw=(40, 56), focal=70.0):
    from thermo_nerf_b200 import orbit_cameras

    cams = orbit_cameras(n, hw=hw[0], focal=focal)
    cams.width, cams.height, cams.cx, cams.cy = hw[1], hw[0], hw[1] / 2, hw[0] / 2
    return cams


def test_generate_rays_matches_oracle():
    from thermo_nerf_b200 import functional as F

    cams = _cameras()
    for i in range(cams.size):
        cam = F.pack_camera(cams.camera_to_worlds[i], cams.fx, cams.fy, cams.cx, cams.cy, cams.width, cams.height)
        o, d, n = F.generate_rays(cam, DEV)
        ro, rd, rn = oracle.generate_rays_np(cams.camera_to_worlds[i].numpy(), cams.fx, cams.fy, cams.cx, cams.cy,
                                             cams.height, cams.width)
        assert np.array_equal(o.cpu().numpy().reshape(ro.shape), ro)
        # same float32 operations in the same order: allow one ulp for the division/sqrt paths
        assert np.allclose(d.cpu().numpy().reshape(rd.shape), rd, rtol=0, atol=1.2e-7)
        assert np.allclose(n.cpu().numpy().reshape(rn.shape), rn, rtol=2e-7, atol=0)
    # a pixel sub-range (ragged tail) and the empty range
    cam = F.pack_camera(cams.camera_to_worlds[0], cams.fx, cams.fy, cams.cx, cams.cy, cams.width, cams.height)
    o_all, d_all, _ = F.generate_rays(cam, DEV)
    o, d, _ = F.generate_rays(cam, DEV, first_pixel=101, num_pixels=37)
    assert torch.equal(d, d_all[101:138]) and torch.equal(o, o_all[101:138])
    assert F.generate_rays(cam, DEV, first_pixel=5, num_pixels=0)[0].shape == (0, 3)


@pytest.mark.parametrize("precision", ["fp32", "tc_fp16"])
def test_camera_forward_equals_ray_forward(precision):
    """Rays generated inside the kernel give bit-identical outputs to the same rays passed as tensors."""
    from thermo_nerf_b200 import RayBundle
    from thermo_nerf_b200 import functional as F

    _, model = make_pair(trained_like=True, precision=precision)
    cams = _cameras(2)
    for i in range(cams.size):
        out_cam = model.get_outputs_for_camera(cams, i)
        cam = F.pack_camera(cams.camera_to_worlds[i], cams.fx, cams.fy, cams.cx, cams.cy, cams.width, cams.height)
        o, d, _ = F.generate_rays(cam, DEV)
        rb = RayBundle(origins=o.view(cams.height, cams.width, 3), directions=d.view(cams.height, cams.width, 3))
        out_rays = model.get_outputs_for_camera_ray_bundle(rb)
        for k in ("rgb", "thermal", "depth", "expected_depth", "accumulation", "prop_depth_0", "prop_depth_1", "img"):
            assert out_cam[k].shape == out_rays[k].shape == (cams.height, cams.width, out_cam[k].shape[-1])
            assert torch.equal(out_cam[k], out_rays[k]), k
    model.train()
    with pytest.raises(RuntimeError):
        model.get_outputs_for_camera(cams, 0)


def test_postprocess_frame_matches_oracle_exactly():
    from thermo_nerf_b200 import functional as F

    g = torch.Generator().manual_seed(0)
    H, W = 37, 53
    rgb = torch.rand((H, W, 3), generator=g)
    th = torch.rand((H, W, 1), generator=g)
    # values on and next to every bin edge of the colour map and of the uint8 grid
    edges = torch.arange(0, 257, dtype=torch.float32) / 256
    th.view(-1)[:257] = edges.clamp(0, 1)
    th.view(-1)[257:514] = torch.nextafter(edges, torch.zeros(())).clamp(0, 1)
    rgb.view(-1)[:256] = torch.arange(256, dtype=torch.float32) / 255
    rgb.view(-1)[256:512] = torch.nextafter(torch.arange(256, dtype=torch.float32) / 255, torch.ones(()))
    rgb.view(-1)[512:514] = torch.tensor([0.0, 1.0])
    cm = oracle.ListedColormapLike(np.random.default_rng(1).random((256, 3)))
    lut8 = torch.from_numpy(oracle.colormap_to_lut8(cm)).to(DEV)
    rgb8, th8 = F.postprocess_frame(rgb=rgb.to(DEV), scalar=th.to(DEV), lut8=lut8)
    assert rgb8.dtype == torch.uint8 and rgb8.shape == (H, W, 3) and th8.shape == (H, W, 3)
    assert np.array_equal(rgb8.cpu().numpy(), oracle.postprocess_np(rgb.numpy(), False))
    assert np.array_equal(th8.cpu().numpy(), oracle.postprocess_np(th.numpy(), True, cm))
    # grey path (depth / accumulation modalities) and a small odd-sized table
    _, g8 = F.postprocess_frame(scalar=th.to(DEV))
    assert np.array_equal(g8.cpu().numpy(), oracle.postprocess_np(th.numpy(), False))
    cm7 = oracle.ListedColormapLike(np.random.default_rng(2).random((7, 3)))
    _, t7 = F.postprocess_frame(scalar=th.to(DEV), lut8=torch.from_numpy(oracle.colormap_to_lut8(cm7)).to(DEV))
    assert np.array_equal(t7.cpu().numpy(), oracle.postprocess_np(th.numpy(), True, cm7))
    assert F.postprocess_frame() == (None, None)


def test_renderer_matches_reference_loop():
    """Renderer.render == the reference's loop (renderer.py:176-200) applied to this model's outputs:
    per modality, per camera: outputs[modality] -> numpy -> colour map / * 255 -> uint8."""
    from thermo_nerf_b200 import RenderedImageModality as M
    from thermo_nerf_b200 import Renderer

    _, model = make_pair(trained_like=True, precision="tc_fp16")
    cams = _cameras(3)
    cm = oracle.ListedColormapLike(np.random.default_rng(3).random((256, 3)))
    mods = [M.RGB, M.THERMAL, M.DEPTH, M.ACCUMULATION]
    r = Renderer(model)
    r.render(mods, cams, thermal_color_map=cm)
    assert set(r._rendered_images) == set(mods)
    for m in mods:
        assert len(r._rendered_images[m]) == cams.size
    from thermo_nerf_b200 import RayBundle
    from thermo_nerf_b200 import functional as F

    for i in range(cams.size):
        cam = F.pack_camera(cams.camera_to_worlds[i], cams.fx, cams.fy, cams.cx, cams.cy, cams.width, cams.height)
        o, d, _ = F.generate_rays(cam, DEV)
        rb = RayBundle(origins=o.view(cams.height, cams.width, 3), directions=d.view(cams.height, cams.width, 3))
        outputs = model.get_outputs_for_camera_ray_bundle(rb)
        for m in mods:
            img = outputs[m.value].cpu().numpy()
            want = oracle.postprocess_np(img, m == M.THERMAL, cm)
            got = r._rendered_images[m][i]
            assert got.dtype == np.uint8 and got.shape == (cams.height, cams.width, 3)
            assert np.array_equal(got, want), m
    with pytest.raises(Exception):
        r.render([M.THERMAL_COMBINED], cams, thermal_color_map=cm)


def test_evaluator_fused_and_pose_paths_agree():
    """Evaluator (evaluator.py:47-106) over the model: with the pose deltas at zero the path that generates the rays in
    the kernel (camera optimiser off) and the one that builds the bundle, applies the camera optimiser in PyTorch and
    renders it (SO3xR3) must give the same metrics; images come back as uint8 side-by-side panels."""
    from types import SimpleNamespace

    from thermo_nerf_b200 import Evaluator
    from thermo_nerf_b200 import RenderedImageModality as M

    cams = _cameras(2)
    g = torch.Generator().manual_seed(4)
    loader = []
    for i in range(cams.size):
        one = type(cams)(cams.camera_to_worlds[i:i + 1], cams.fx, cams.fy, cams.cx, cams.cy, cams.width, cams.height)
        loader.append((one, {"image": torch.rand((cams.height, cams.width, 3), generator=g),
                             "thermal": torch.rand((cams.height, cams.width, 1), generator=g)}))
    cfg = SimpleNamespace(experiment_name="exp", method_name="thermal-nerf")
    results = {}
    for mode in ("off", "SO3xR3"):
        _, model = make_pair(trained_like=True, precision="tc_fp16", camera_optimizer_mode=mode)
        with torch.no_grad():
            model.camera_optimizer.pose_adjustment.zero_()
        dm = SimpleNamespace(setup_eval=lambda: None, fixed_indices_eval_dataloader=loader)
        ev = Evaluator(SimpleNamespace(model=model, datamanager=dm), cfg, modalities_to_save=[M.RGB, M.THERMAL_COMBINED],
                       threshold=0.5)
        results[mode] = ev
        assert len(ev._evaluation_images[M.RGB]) == 2
        assert ev._evaluation_images[M.RGB][0].shape == (cams.height, 2 * cams.width, 3)
        assert ev._evaluation_images[M.THERMAL_COMBINED][0].dtype == np.uint8
    a, b = results["off"].metrics, results["SO3xR3"].metrics
    assert set(a) == set(b) and "mae_thermal_mean" in a and "psnr_thermal_std" in a
    for k in ("psnr", "ssim", "psnr_thermal", "mae_thermal", "mae_thermal_foreground"):
        assert a[k] == pytest.approx(b[k], rel=1e-6, abs=1e-9), k
