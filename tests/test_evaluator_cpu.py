"""Evaluator mirror (thermo_nerf/evaluator/evaluator.py): aggregation, files and error behaviour with a stand-in
model (no GPU): per-key mean / std / list over the eval images, metrics.json, the psnr/ssim/lpips text files (the
thermal ones are always written, as in the reference) and one jpg per image and modality."""

import json
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from thermo_nerf_b200 import Evaluator, RenderedImageModality, sphere_cameras


class _FakeModel:
    device = torch.device("cpu")
    camera_optimizer = SimpleNamespace(mode="off", apply_to_raybundle=lambda rb: None)

    def __init__(self):
        self.calls = 0

    def get_outputs_for_camera_ray_bundle(self, bundle):
        self.calls += 1
        h, w = bundle.origins.shape[:2]
        return {"rgb": torch.full((h, w, 3), 0.1 * self.calls), "thermal": torch.full((h, w, 1), 0.5)}

    def get_image_metrics_and_images(self, outputs, batch, threshold=None):
        v = float(outputs["rgb"].mean())
        m = {k: v for k in ("psnr", "ssim", "lpips", "psnr_thermal", "ssim_thermal", "lpips_thermal", "mae_thermal")}
        m["mae_thermal_foreground"] = -1.0 if threshold is None else threshold
        return m, {"img": outputs["rgb"], "thermal": outputs["thermal"].repeat(1, 1, 3)}


def _pipeline(n_batches=3):
    cams = sphere_cameras(1, hw=8, focal=10.0)
    loader = [(cams, {"image": torch.zeros(8, 8, 3), "thermal": torch.zeros(8, 8, 1)}) for _ in range(n_batches)]
    dm = SimpleNamespace(setup_eval=lambda: None, fixed_indices_eval_dataloader=loader)
    return SimpleNamespace(model=_FakeModel(), datamanager=dm)


def test_metrics_aggregation_and_files(tmp_path):
    cfg = SimpleNamespace(experiment_name="exp", method_name="thermal-nerf")
    ev = Evaluator(_pipeline(), cfg, job_param_identifier="job7",
                   modalities_to_save=[RenderedImageModality.RGB, RenderedImageModality.THERMAL], threshold=0.3)
    m = ev.metrics
    assert m["psnr"] == pytest.approx([0.1, 0.2, 0.3]) and m["psnr_mean"] == pytest.approx(0.2)
    assert m["psnr_std"] == pytest.approx(float(torch.std(torch.tensor([0.1, 0.2, 0.3]))))
    assert m["mae_thermal_foreground"] == [0.3, 0.3, 0.3]
    ev.save_metrics(tmp_path)
    info = json.loads((tmp_path / "metrics.json").read_text())
    assert info["experiment_name"] == "exp" and info["method_name"] == "thermal-nerf" and info["job_param_identifier"] == "job7"
    assert info["results"]["ssim_mean"] == pytest.approx(0.2)
    for name in ("psnr", "ssim", "lpips"):
        assert json.loads((tmp_path / name / "job7.txt").read_text()) == pytest.approx([0.1, 0.2, 0.3])
        assert (tmp_path / name / "job7_thermal.txt").exists()
    ev.save_images([RenderedImageModality.RGB, RenderedImageModality.THERMAL], tmp_path)
    names = sorted(p.name for p in tmp_path.glob("*.jpg"))
    assert names == ["img_00000.jpg", "img_00001.jpg", "img_00002.jpg", "thermal_00000.jpg", "thermal_00001.jpg",
                     "thermal_00002.jpg"]
    assert ev._evaluation_images[RenderedImageModality.RGB][1].dtype == np.uint8
    assert int(ev._evaluation_images[RenderedImageModality.RGB][1][0, 0, 0]) == 51  # 0.2 * 255 truncated


def test_without_identifier_only_metrics_json(tmp_path):
    ev = Evaluator(_pipeline(1), SimpleNamespace(experiment_name="e", method_name="m"))
    ev.save_metrics(tmp_path)
    assert sorted(p.name for p in tmp_path.iterdir()) == ["metrics.json"]


def test_missing_eval_dataloader_raises():
    p = _pipeline()
    p.datamanager.fixed_indices_eval_dataloader = None
    with pytest.raises(RuntimeError):
        Evaluator(p, SimpleNamespace(experiment_name="e", method_name="m"))
