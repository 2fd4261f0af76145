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
