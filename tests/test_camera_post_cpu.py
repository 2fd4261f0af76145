"""Oracle known-answer tests for the neighbours of the path: ray generation, the
matplotlib colour-map call and Renderer.render's uint8 conversion, plus the host side of the Renderer
mirror (camera-path loading against the reference's own fixture).  CPU only."""

import json
from pathlib import Path

import numpy as np
import pytest
import torch

import oracle
from thermo_nerf_b200 import PinholeCameras, RenderedImageModality, Renderer
from thermo_nerf_b200.render import lut8_from_colormap

GOLDEN = Path(__file__).parent / "golden"


def test_generate_rays_known_answers():
    # identity pose: the principal ray looks down -z; pixel centres sit at +0.5
    c2w = np.concatenate([np.eye(3), np.array([[1.0], [2.0], [3.0]])], 1)
    o, d, n = oracle.generate_rays_np(c2w, 100.0, 100.0, 4.0, 3.0, 6, 8)
    assert o.shape == (6, 8, 3) and np.all(o == np.array([1, 2, 3], np.float32))
    # pixel (y=2, x=3): centre (3.5, 2.5) -> camera dir (-0.005, +0.005, -1) normalised
    v = np.array([(3.5 - 4) / 100, -(2.5 - 3) / 100, -1.0])
    assert np.allclose(d[2, 3], v / np.linalg.norm(v), atol=1e-7)
    assert np.allclose(n[2, 3, 0], np.linalg.norm(v), atol=1e-7)
    assert np.allclose(np.linalg.norm(d, axis=-1), 1.0, atol=1e-6)
    # a rotated camera rotates every direction: R maps camera -z onto world +x
    R = np.array([[0.0, 0, -1], [0, 1, 0], [1, 0, 0]])
    _, d2, _ = oracle.generate_rays_np(np.concatenate([R, np.zeros((3, 1))], 1), 50.0, 50.0, 2.0, 2.0, 4, 4)
    assert np.allclose(d2, d2 @ np.eye(3)) and d2[..., 0].min() > 0.99  # all rays point along +x
    # the torch stand-in used to build benchmark rays agrees with the oracle
    cams = PinholeCameras(torch.tensor(c2w, dtype=torch.float32)[None], 100.0, 100.0, 4.0, 3.0, 8, 6)
    rb = cams.generate_rays(0)
    assert np.allclose(rb.directions.numpy(), d, atol=1e-6) and np.allclose(rb.origins.numpy(), o)


def test_colormap_call_semantics_and_uint8_conversion():
    rng = np.random.default_rng(0)
    colors = rng.random((7, 3))
    cm = oracle.ListedColormapLike(colors)
    x = np.array([0.0, 0.1428, 0.1429, 0.5, 0.999, 1.0], dtype=np.float32)
    idx = [0, 0, 1, 3, 6, 6]  # x*7 truncated; x == 1 -> last bin
    assert np.allclose(cm(x)[:, :3], colors[idx])
    lut8 = oracle.colormap_to_lut8(cm)
    assert lut8.dtype == np.uint8 and np.array_equal(lut8, (colors * 255).astype(np.uint8))
    assert np.array_equal(lut8_from_colormap(cm), lut8)             # product-side conversion = oracle's
    assert np.array_equal(lut8_from_colormap(colors), lut8)         # a float table is accepted as is
    assert np.array_equal(lut8_from_colormap(lut8), lut8)           # and a uint8 table
    # renderer.py:189-199: single-channel images are replicated, float -> uint8 truncates
    img = np.array([[[0.0], [0.5], [0.999], [1.0]]], dtype=np.float32)
    assert np.array_equal(oracle.postprocess_np(img, False)[0, :, 0], [0, 127, 254, 255])
    th = oracle.postprocess_np(img, True, cm)
    assert th.shape == (1, 4, 3) and np.array_equal(th[0, 3], lut8[6]) and np.array_equal(th[0, 0], lut8[0])


def test_load_cameras_reference_fixture():
    """tests/test_renderer.py:66-69 of the reference loads camera_path_facade_2.json and expects 96 poses;
    a copy of that fixture's header + the first/last pose is committed under tests/golden/."""
    path = GOLDEN / "camera_path_facade_2_excerpt.json"
    cams = Renderer.load_cameras(path)
    meta = json.load(open(path))
    assert cams.size == len(meta["camera_path"]) == meta["_num_poses_in_excerpt"]
    assert (cams.width, cams.height) == (1920, 1080)
    # three_js_perspective_camera_focal_length(fov=50, h=1080)
    assert cams.fx == pytest.approx(0.5 * 1080 / np.tan(np.deg2rad(25.0)))
    assert cams.cx == 960 and cams.cy == 540
    c2w, fx, fy, cx, cy, h, w = oracle.camera_path_to_cameras(meta)
    assert np.allclose(cams.camera_to_worlds.numpy(), c2w) and (h, w) == (1080, 1920) and fx == pytest.approx(cams.fx)
    half = Renderer.load_cameras(path, 0.5)
    assert (half.width, half.height) == (960, 540) and half.fx == pytest.approx(cams.fx / 2)


def test_renderer_surface():
    assert RenderedImageModality.RGB.value == "img" and RenderedImageModality.THERMAL.value == "thermal"
    r = Renderer(model=object())
    assert r.model is not None and r._rendered_images == {}
    with pytest.raises(ImportError):
        Renderer.from_pipeline_path(Path("."), Path("."))


def test_save_images_and_gif(tmp_path):
    # tests/test_renderer.py:44-64 of the reference: two frames -> two jpeg files and one gif
    r = Renderer(model=object())
    frame = (np.arange(64 * 48 * 3) % 255).astype(np.uint8).reshape(48, 64, 3)
    r._rendered_images = {RenderedImageModality.RGB: [frame, frame]}
    r.save_images([RenderedImageModality.RGB], tmp_path)
    r.save_gif([RenderedImageModality.RGB], 1, tmp_path)
    names = sorted(p.name for p in tmp_path.iterdir())
    assert names == ["img_00000.jpeg", "img_00001.jpeg", "synthesized_video_img.gif"]


# ---- the reference's own thermal fixture (tests/data/thermal/*, its tests/test_thermal_conversion.py) ----
def _thermal_kat():
    from pathlib import Path

    return torch.load(Path(__file__).parent / "golden" / "reference_thermal_image_kat.pt", weights_only=True)


def test_uint8_thermal_ground_truth_denormalises_to_the_flir_temperatures():
    """8-bit thermal PNG -> x / 255 (what the batch sampler hands to the loss) -> (max - min) x + min reproduces the
    FLIR temperatures of the reference fixture to the reference test's own tolerance (1 decimal,
    tests/test_thermal_conversion.py:46-52) and to within one 8-bit quantisation step."""
    from oracle.camera_post import sample_batch_np

    kat = _thermal_kat()
    px = kat["pixels_u8"].numpy()  # [80,60]
    temps = kat["temperature_c"].numpy()
    tmax, tmin = kat["absolute_max_temperature"], kat["absolute_min_temperature"]
    h, w = px.shape
    images = np.zeros((1, h, w, 3), np.uint8)
    c2w = np.eye(4, dtype=np.float32)[None, :3]
    intr = np.array([[50.0, 50.0, w / 2, h / 2]], np.float32)
    # one draw per pixel centre
    yy, xx = np.meshgrid(np.arange(h), np.arange(w), indexing="ij")
    rand = np.stack([np.zeros(h * w), (yy.ravel() + 0.5) / h, (xx.ravel() + 0.5) / w], -1).astype(np.float32)
    _, _, idx, _, gt_th = sample_batch_np(rand, images, px[None], c2w, intr)
    assert np.array_equal(idx[:, 1], yy.ravel()) and np.array_equal(idx[:, 2], xx.ravel())
    assert gt_th.dtype == np.float32 and np.array_equal(gt_th, px.ravel().astype(np.float32) / np.float32(255))
    celsius = gt_th.astype(np.float64) * (tmax - tmin) + tmin  # ThermalVisualiser.update_temperature
    err = np.abs(celsius - temps.ravel())
    step = (tmax - tmin) / 255
    assert err.max() < 0.15                      # the reference's assert_almost_equal(decimal=1)
    assert err.max() <= step * (1 + 1e-3), (err.max(), step)


def test_mae_thermal_on_the_reference_fixture():
    """mae_thermal between the 8-bit image and the exact (csv) temperatures, both normalised, is the mean absolute
    temperature error in degrees - oracle and product metric agree with the direct computation."""
    from oracle import nerfstudio_math as M
    from thermo_nerf_b200.model import ThermalNerfModel

    kat = _thermal_kat()
    tmax, tmin = kat["absolute_max_temperature"], kat["absolute_min_temperature"]
    pred = kat["pixels_u8"].float() / 255
    gt = ((kat["temperature_c"] - tmin) / (tmax - tmin)).float()
    want = (pred.double() * (tmax - tmin) + tmin - kat["temperature_c"]).abs().mean().item()
    got = M.mae_thermal(gt, pred, False, tmax, tmin).item()
    assert abs(got - want) < 1e-4 and 0.02 < got < 0.06
    from types import SimpleNamespace

    m = SimpleNamespace(max_temperature=tmax, min_temperature=tmin, config=SimpleNamespace(cold=False))
    assert abs(float(ThermalNerfModel.mae_thermal(m, gt, pred)) - want) < 1e-4


def test_frame_conversion_matches_the_reference_render_loop():
    """Renderer.render (renderer.py:160-201) executed from the reference (tests/golden/make_reference_render_frames_golden.py):
    the oracle's per-frame conversion reproduces its uint8 frames bit for bit, for the colour-mapped thermal modality and
    the replicated single-channel ones; the loop renders every frame once per modality and rejects the RGB modality."""
    from oracle.camera_post import ListedColormapLike, colormap_to_lut8, postprocess_np

    gold = torch.load(GOLDEN / "reference_render_frames.pt", weights_only=True)
    cmap = ListedColormapLike(gold["lut"].numpy())
    keys = {"THERMAL": "thermal", "DEPTH": "depth", "ACCUMULATION": "accumulation"}
    assert gold["modalities"] == list(keys)
    for name, key in keys.items():
        assert len(gold["rendered"][name]) == len(gold["frames"]) == 2
        for frame, ref in zip(gold["frames"], gold["rendered"][name]):
            ours = postprocess_np(frame[key].numpy(), name == "THERMAL", cmap)
            assert ours.dtype == np.uint8 and ours.shape == tuple(ref.shape) == (*gold["hw"], 3)
            assert np.array_equal(ours, ref.numpy()), name
    # the uint8 table the CUDA post-processing kernel indexes gives the same thermal pixels (x in [0,1])
    lut8 = colormap_to_lut8(cmap)
    x = gold["frames"][0]["thermal"].numpy()[..., 0].astype(np.float64)
    idx = np.minimum((x * 256).astype(np.int64), 255)
    assert np.array_equal(lut8[idx], gold["rendered"]["THERMAL"][0].numpy())
    # modality-outer / camera-inner: 3 modalities x 2 cameras, then the RGB attempt fails on its first frame
    assert gold["model_calls"] == [0, 1, 0, 1, 0, 1, 0]
    assert gold["rgb_modality_error"] == "img modality does not exist"  # RGB.value == "img", the model emits "rgb"


def test_uint8_ground_truth_conversion_conventions_agree_on_every_byte():
    """The reference converts thermal images as float64 `image / 255.0` then `.astype(float32)`
    (thermal_dataset.py:64-65) and colour images as `image.astype("float32") / 255.0` (nerfstudio InputDataset); the
    oracle's and the batch-sampling kernel's float32 `x / 255` give the same float32 for all 256 byte values."""
    x = np.arange(256, dtype=np.uint8)
    thermal_ref = (x / 255.0).astype(np.float32)
    colour_ref = x.astype("float32") / 255.0
    ours = x.astype(np.float32) / np.float32(255.0)
    assert colour_ref.dtype == np.float32
    assert np.array_equal(thermal_ref, ours) and np.array_equal(colour_ref, ours)


def test_saved_file_names_match_the_reference_renderer(tmp_path):
    """Renderer.save_images / save_gif of the reference (renderer.py:202-228) were executed with a recording imageio:
    the mirror writes the same files (names, one gif per modality over all frames) from the same frames."""
    from PIL import Image

    gold = torch.load(GOLDEN / "reference_render_frames.pt", weights_only=True)
    mods = [RenderedImageModality[n] for n in gold["modalities"]]
    r = Renderer(model=object())
    r._rendered_images = {m: [f.numpy() for f in gold["rendered"][m.name]] for m in mods}
    r.save_images(mods, tmp_path)
    r.save_gif(mods, 2.5, tmp_path)
    want = sorted(name for _, name, _, _ in gold["written"])
    assert sorted(p.name for p in tmp_path.iterdir()) == want
    for fn, name, frames, duration in gold["written"]:
        if fn == "mimsave":
            assert frames == 2 and duration == 2.5
            with Image.open(tmp_path / name) as gif:
                assert getattr(gif, "n_frames", 1) == frames
        else:
            with Image.open(tmp_path / name) as im:
                assert im.size == (gold["hw"][1], gold["hw"][0])
