"""Device pixel sampler + ground-truth gather + ray generation (tnf_sample_batch) against
the numpy oracle on the same uniform draw: indices and gathered pixels bit-exact, rays within one ulp."""

import numpy as np
import pytest
import torch

import oracle

pytestmark = pytest.mark.gpu


def _dataset(n=5, h=24, w=36, u8=False, seed=0):
    from thermo_nerf_b200 import sphere_cameras

    g = torch.Generator().manual_seed(seed)
    if u8:
        images = torch.randint(0, 256, (n, h, w, 3), generator=g, dtype=torch.uint8)
        thermal = torch.randint(0, 256, (n, h, w, 1), generator=g, dtype=torch.uint8)
    else:
        images = torch.rand((n, h, w, 4), generator=g)  # RGBA: only the first three channels are colour
        thermal = torch.rand((n, h, w), generator=g)
    cams = sphere_cameras(n, hw=h, focal=40.0)
    fx = torch.linspace(38.0, 42.0, n)  # per-camera intrinsics
    return images, thermal, cams.camera_to_worlds, fx, fx * 1.01, torch.full((n,), w / 2), torch.full((n,), h / 2)


@pytest.mark.parametrize("u8", [False, True])
def test_sample_batch_matches_oracle(u8):
    from thermo_nerf_b200.data import DevicePixelSampler

    images, thermal, c2w, fx, fy, cx, cy = _dataset(u8=u8)
    s = DevicePixelSampler(images, thermal, c2w, fx, fy, cx, cy, device="cuda:0", seed=3)
    R = 4099
    rand = torch.rand((R, 3), generator=torch.Generator().manual_seed(7))
    rand[:6] = torch.tensor([[0.0, 0.0, 0.0], [0.999999, 0.999999, 0.999999], [0.2, 0.5, 0.5], [0.4, 1 / 24, 1 / 36],
                             [0.6, 23 / 24, 35 / 36], [0.8, 0.5 - 1e-7, 0.5 + 1e-7]])
    rb, batch = s.sample(R, rand.cuda())
    torch.cuda.synchronize()
    intr = torch.stack([fx, fy, cx, cy], 1).numpy()
    o, d, idx, rgb, th = oracle.sample_batch_np(rand.numpy(), images.numpy(), thermal.numpy(), c2w.numpy(), intr)
    assert np.array_equal(batch["indices"].cpu().numpy(), idx)
    assert np.array_equal(rb.camera_indices.cpu().numpy()[:, 0], idx[:, 0])
    assert np.array_equal(batch["image"].cpu().numpy(), rgb)
    assert np.array_equal(batch["thermal"].cpu().numpy()[:, 0], th)
    assert np.array_equal(rb.origins.cpu().numpy(), o)
    assert np.allclose(rb.directions.cpu().numpy(), d, rtol=0, atol=1.2e-7)
    assert batch["image"].shape == (R, 3) and batch["thermal"].shape == (R, 1) and rb.camera_indices.dtype == torch.int64


def test_sampler_draws_cover_the_dataset_and_feed_the_model():
    from tests.helpers import make_pair
    from thermo_nerf_b200.data import DevicePixelSampler

    images, thermal, c2w, fx, fy, cx, cy = _dataset(n=8)
    s = DevicePixelSampler(images, thermal, c2w, fx, fy, cx, cy, device="cuda:0", seed=11)
    rb, batch = s.next_train(0, num_rays=8192)
    idx = batch["indices"].cpu()
    assert int(idx[:, 0].min()) == 0 and int(idx[:, 0].max()) == 7
    assert int(idx[:, 1].max()) == 23 and int(idx[:, 2].max()) == 35 and int(idx.min()) == 0
    rb2, _ = s.next_train(1, num_rays=8192)
    assert not torch.equal(rb.origins, rb2.origins) or not torch.equal(rb.directions, rb2.directions)
    # a training step straight from the sampler
    _, model = make_pair(trained_like=False, precision="tc_fp16", log2_field=14, log2_prop=12)
    model.train()
    out = model(rb)
    ld = model.get_loss_dict(out, batch, model.get_metrics_dict(out, batch))
    sum(ld.values()).backward()
    torch.cuda.synchronize()
    assert torch.isfinite(model.field.mlp_base.encoder.hash_table.grad).all()
    # no thermal modality: the batch simply has no "thermal" entry
    s2 = DevicePixelSampler(images, None, c2w, 40.0, 40.0, 18.0, 12.0, device="cuda:0")
    assert "thermal" not in s2.sample(16)[1]
