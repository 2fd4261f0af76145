"""``Evaluator``: the reference's thermo_nerf/evaluator/evaluator.py over the fused B200 path.

Same constructor, attributes and files (``metrics.json``, ``psnr/ ssim/ lpips/<identifier>[_thermal].txt``, one
``<modality>_<idx>.jpg`` per evaluation image).  ``pipeline`` is duck-typed on what the reference touches:
``pipeline.model``, ``pipeline.datamanager.setup_eval()`` and ``pipeline.datamanager.fixed_indices_eval_dataloader``
yielding ``(cameras, batch)`` pairs; ``config`` needs ``experiment_name`` and ``method_name``.

Per evaluation image (evaluator.py:64-87): rays of the camera -> ``camera_optimizer.apply_to_raybundle`` ->
``get_outputs_for_camera_ray_bundle`` -> ``get_image_metrics_and_images(outputs, batch, threshold=)``.  With plain
perspective cameras and no pose deltas to apply the rays are generated inside the forward kernel
(``get_outputs_for_camera``); otherwise ``tnf_generate_rays`` builds the bundle on the device and the pose deltas are
applied in PyTorch as in the reference.
"""

from __future__ import annotations

import json
from pathlib import Path
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import functional as F
from .rays import PinholeCameras, RayBundle
from .render import RenderedImageModality


class Evaluator:
    def __init__(self, pipeline, config, job_param_identifier: Optional[str] = None,
                 modalities_to_save: Sequence[RenderedImageModality] = (RenderedImageModality.RGB,),
                 threshold: Optional[float] = None) -> None:
        pipeline.datamanager.setup_eval()
        self._pipeline = pipeline
        self._run_names = (config.experiment_name, config.method_name)
        self.identifier = job_param_identifier
        self.modalities_to_save = list(modalities_to_save)
        self._evaluation_images: Dict[RenderedImageModality, List[np.ndarray]] = {m: [] for m in self.modalities_to_save}
        self._metrics = self._compute_metrics(threshold=threshold)

    # ------------------------------------------------------------------ results
    @property
    def metrics(self) -> dict:
        return self._metrics

    @property
    def _benchmark_info(self) -> dict:
        """What metrics.json holds (evaluator.py:38-43)."""
        experiment, method = self._run_names
        return {"experiment_name": experiment, "method_name": method, "job_param_identifier": self.identifier,
                "results": self._metrics}

    # ------------------------------------------------------------------ evaluation loop
    def _outputs_for(self, model, cameras, i: int):
        """One evaluation frame of camera ``i`` of ``cameras`` (evaluator.py:66-81)."""
        pose_off = getattr(model.camera_optimizer, "mode", "off") == "off"
        if isinstance(cameras, PinholeCameras) and model.device.type == "cuda":
            if pose_off:
                return model.get_outputs_for_camera(cameras, i)
            cam = F.pack_camera(cameras.camera_to_worlds[i], cameras.fx, cameras.fy, cameras.cx, cameras.cy,
                                cameras.width, cameras.height)
            o, d, _ = F.generate_rays(cam, model.device)
            idx = torch.full((o.shape[0], 1), i, dtype=torch.int64, device=model.device)
            flat = RayBundle(origins=o, directions=d, camera_indices=idx)
            shape = (cameras.height, cameras.width)
        else:
            bundle = cameras.generate_rays(torch.tensor([i]) if not isinstance(cameras, PinholeCameras) else i)
            shape = tuple(bundle.origins.shape[:-1])
            flat = bundle.flatten()
        with torch.no_grad():
            model.camera_optimizer.apply_to_raybundle(flat)
        return model.get_outputs_for_camera_ray_bundle(flat.reshape(shape))

    @staticmethod
    def _aggregate(per_image: List[Dict[str, float]]) -> dict:
        """<key>_mean, <key>_std (unbiased, torch.std_mean) and the per-image list of every metric."""
        summary: dict = {}
        for name in per_image[0]:
            column = [entry[name] for entry in per_image]
            std, mean = torch.std_mean(torch.tensor(column))
            summary[name + "_mean"], summary[name + "_std"], summary[name] = float(mean), float(std), column
        return summary

    def _compute_metrics(self, threshold: Optional[float]) -> dict:
        loader = self._pipeline.datamanager.fixed_indices_eval_dataloader
        if loader is None:
            raise RuntimeError("Cannot evaluate without a fixed indices eval dataloader")
        model = self._pipeline.model
        per_image: List[Dict[str, float]] = []
        for cameras, batch in loader:
            for i in range(int(cameras.camera_to_worlds.shape[0])):  # the reference's loader yields one camera at a time
                outputs = self._outputs_for(model, cameras, i)
                frame_metrics, frame_images = model.get_image_metrics_and_images(outputs, batch, threshold=threshold)
                per_image.append(frame_metrics)
                for modality, store in self._evaluation_images.items():
                    store.append((frame_images[modality.value] * 255).byte().cpu().numpy())
        if not per_image:
            raise RuntimeError("the eval dataloader is empty")
        return self._aggregate(per_image)

    # ------------------------------------------------------------------ files
    def save_images(self, modalities: Sequence[RenderedImageModality], output_path: Path) -> None:
        from PIL import Image

        for modality in modalities:
            for idx, frame in enumerate(self._evaluation_images[modality]):
                plane = frame[:, :, 3] if frame.shape[-1] == 4 else frame  # RGBA panels keep their alpha plane only
                Image.fromarray(plane).save(Path(output_path) / f"{modality.value}_{idx:05d}.jpg")

    def save_metrics(self, output_folder: Path) -> None:
        root = Path(output_folder)
        root.mkdir(parents=True, exist_ok=True)
        (root / "metrics.json").write_text(json.dumps(self._benchmark_info, indent=2), "utf8")
        if self.identifier is None:
            return
        # per-metric folders with the per-image lists; the reference's guard for the thermal files
        # (`if THERMAL or ...`, evaluator.py:155-158) is always true, so they are always written
        for name in ("psnr", "ssim", "lpips"):
            folder = root / name
            folder.mkdir(parents=True, exist_ok=True)
            for suffix in ("", "_thermal"):
                (folder / f"{self.identifier}{suffix}.txt").write_text(json.dumps(self._metrics[name + suffix], indent=2), "utf8")
