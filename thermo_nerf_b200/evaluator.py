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
        self._pipeline = pipeline
        self._pipeline.datamanager.setup_eval()
        self.identifier = job_param_identifier
        self._evaluation_images: Dict[RenderedImageModality, List[np.ndarray]] = {}
        self.modalities_to_save = list(modalities_to_save)
        self._metrics = self._compute_metrics(threshold=threshold)
        self._benchmark_info = {
            "experiment_name": config.experiment_name,
            "method_name": config.method_name,
            "job_param_identifier": self.identifier,
            "results": self._metrics,
        }

    @property
    def metrics(self) -> dict:
        return self._metrics

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

    def _compute_metrics(self, threshold: Optional[float]) -> dict:
        datamanager = self._pipeline.datamanager
        if datamanager.fixed_indices_eval_dataloader is None:
            raise RuntimeError("Cannot evaluate without a fixed indices eval dataloader")
        for modality in self.modalities_to_save:
            self._evaluation_images[modality] = []
        model = self._pipeline.model
        metrics_dict_list = []
        for cameras, batch in datamanager.fixed_indices_eval_dataloader:
            n = int(cameras.camera_to_worlds.shape[0])
            for i in range(n):  # the reference's dataloader yields one camera at a time
                outputs = self._outputs_for(model, cameras, i)
                metrics_dict, images_dict = model.get_image_metrics_and_images(outputs, batch, threshold=threshold)
                for modality in self.modalities_to_save:
                    self._evaluation_images[modality].append((images_dict[modality.value] * 255).byte().cpu().numpy())
                metrics_dict_list.append(metrics_dict)
        if not metrics_dict_list:
            raise RuntimeError("the eval dataloader is empty")
        out: dict = {}
        for key in metrics_dict_list[0].keys():
            vals = [m[key] for m in metrics_dict_list]
            key_std, key_mean = torch.std_mean(torch.tensor(vals))
            out[f"{key}_mean"] = float(key_mean)
            out[f"{key}_std"] = float(key_std)
            out[key] = vals
        return out

    def save_images(self, modalities: Sequence[RenderedImageModality], output_path: Path) -> None:
        from PIL import Image

        for modality in modalities:
            for idx, image in enumerate(self._evaluation_images[modality]):
                if image.shape[-1] == 4:
                    image = image[:, :, 3]
                Image.fromarray(image).save(Path(output_path) / f"{modality.value}_{idx:05d}.jpg")

    def save_metrics(self, output_folder: Path) -> None:
        output_folder = Path(output_folder)
        output_file = output_folder / "metrics.json"
        output_file.parent.mkdir(parents=True, exist_ok=True)
        output_file.write_text(json.dumps(self._benchmark_info, indent=2), "utf8")
        if self.identifier is None:
            return
        for name in ("psnr", "ssim", "lpips"):
            folder = output_folder / name
            folder.mkdir(parents=True, exist_ok=True)
            (folder / (self.identifier + ".txt")).write_text(json.dumps(self._metrics[name], indent=2), "utf8")
            # the reference's condition (`if THERMAL or ...`, evaluator.py:155-158) is always true: the thermal
            # files are always written
            (folder / (self.identifier + "_thermal.txt")).write_text(
                json.dumps(self._metrics[name + "_thermal"], indent=2), "utf8")
