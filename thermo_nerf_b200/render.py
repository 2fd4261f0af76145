"""``Renderer``: the reference's thermo_nerf/render/renderer.py on the fused B200 path.

Same class surface (``Renderer(model)``, ``model``, ``load_cameras``, ``render``, ``save_images``,
``save_gif``, ``from_pipeline_path``) and the same ``_rendered_images`` contents - lists of uint8
``[H, W, 3]`` numpy arrays per ``RenderedImageModality`` - but a frame goes through the device once:

* rays are generated inside the forward kernel from the camera (no ``generate_rays`` tensors),
* all requested modalities of a frame come from ONE forward pass (the reference renders every frame
  once per modality, renderer.py:180-182; the outputs are identical),
* the ``* 255 -> uint8`` conversion and the thermal colour map run in ``tnf_postprocess_frame`` and only
  uint8 pixels cross PCIe (renderer.py:189-199 does them on the host in numpy / matplotlib).

Cameras that are not plain perspective cameras fall back to ``cameras.generate_rays`` +
``model.get_outputs_for_camera_ray_bundle`` (still one pass per frame).
"""

from __future__ import annotations

import json
from enum import Enum
from pathlib import Path
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import functional as F
from .rays import PinholeCameras


class RenderedImageModality(Enum):  # thermo_nerf/rendered_image_modalities.py:4-9
    RGB = "img"
    DEPTH = "depth"
    ACCUMULATION = "accumulation"
    THERMAL = "thermal"
    THERMAL_COMBINED = "thermal_combined"


def lut8_from_colormap(cmap) -> np.ndarray:
    """uint8 [N,3] table of a matplotlib-style colour map (``cmap.N`` entries, ``cmap(x) -> RGBA``), converted
    the way renderer.py:195-197 converts the mapped image: ``(cmap(x)[..., :3] * 255).astype(uint8)``.
    A ``[N,3]`` / ``[N,4]`` float array in [0,1] or a uint8 ``[N,3]`` array is accepted as the table itself."""
    if isinstance(cmap, np.ndarray) or torch.is_tensor(cmap):
        arr = np.asarray(cmap.cpu() if torch.is_tensor(cmap) else cmap)
        if arr.ndim != 2 or arr.shape[1] not in (3, 4):
            raise ValueError("a colour table must be [N,3] or [N,4]")
        if arr.dtype == np.uint8:
            return np.ascontiguousarray(arr[:, :3])
        return (arr[:, :3].astype(np.float64) * 255).astype(np.uint8)
    n = int(cmap.N)
    x = (np.arange(n, dtype=np.float64) + 0.5) / n  # bin centres: x * N truncates to the bin index
    return (np.asarray(cmap(x))[:, :3] * 255).astype(np.uint8)


def _default_thermal_colormap():
    try:
        import matplotlib.pyplot as plt  # the reference's default: plt.colormaps["magma"] (renderer.py:163)
    except ImportError as e:  # pragma: no cover - depends on the environment
        raise RuntimeError("matplotlib is not installed: pass thermal_color_map (a Colormap-like object or an "
                           "[N,3] colour table) to Renderer.render") from e
    return plt.colormaps["magma"]


class Renderer:
    def __init__(self, model) -> None:
        self._rendered_images: Dict[RenderedImageModality, List[np.ndarray]] = {}
        self._model = model

    @property
    def model(self):
        return self._model

    # ------------------------------------------------------------------ loading
    @classmethod
    def from_pipeline_path(cls, model_path: Path, transforms_path: Path,
                           eval_num_rays_per_chunk: Optional[int] = None) -> "Renderer":
        """renderer.py:116-141: ``config.yml`` + the last ``*.ckpt`` of a training run -> a renderer on its model.
        The pipeline (data manager, TrainerConfig unpickling, checkpoint layout) is nerfstudio's and the
        reference's own loader builds it (``Renderer.extract_pipeline``, renderer.py:70-115);
        ``nerfstudio_plugin.install()`` makes that loader construct the B200 model - also for runs trained with the
        stock model, whose state_dict keys are the same.  Raises ImportError without nerfstudio + thermo_nerf."""
        from . import nerfstudio_plugin

        nerfstudio_plugin.install()  # raises ImportError naming what is missing
        from thermo_nerf.render.renderer import Renderer as ReferenceRenderer  # type: ignore[import-not-found]

        pipeline, _, _ = ReferenceRenderer.extract_pipeline(model_path=model_path, transforms_path=transforms_path,
                                                            eval_num_rays_per_chunk=eval_num_rays_per_chunk)
        return cls(pipeline.model)

    @staticmethod
    def load_cameras(load_camera_trajectory: Path, rendered_resolution_scaling_factor: float = 1.0) -> PinholeCameras:
        """renderer.py:144-157: a nerfstudio camera-path JSON as cameras (perspective paths)."""
        with open(load_camera_trajectory, "r", encoding="utf-8") as f:
            camera_path = json.load(f)
        if camera_path.get("camera_type", "perspective") != "perspective":
            raise ValueError("only perspective camera paths are supported")
        h, w = int(camera_path["render_height"]), int(camera_path["render_width"])
        c2ws, focals = [], []
        for cam in camera_path["camera_path"]:
            c2ws.append(torch.tensor(cam["camera_to_world"], dtype=torch.float32).view(4, 4)[:3])
            focals.append(0.5 * h / float(np.tan(0.5 * float(cam["fov"]) * np.pi / 180.0)))
        if len(set(focals)) != 1:
            raise ValueError("camera paths with a varying field of view are not supported")
        s = float(rendered_resolution_scaling_factor)
        return PinholeCameras(torch.stack(c2ws), focals[0] * s, focals[0] * s, w / 2 * s, h / 2 * s, int(w * s),
                              int(h * s))

    # ------------------------------------------------------------------ rendering
    _OUTPUT_KEY = {RenderedImageModality.RGB: "rgb"}  # "img" is an alias of "rgb" in the output dict

    def render(self, rendered_image_modalities: Sequence[RenderedImageModality], cameras,
               thermal_color_map=None) -> None:
        """renderer.py:160-200.  Fills ``_rendered_images[modality]`` with one uint8 [H,W,3] array per camera."""
        model = self._model
        device = model.device
        if device.type != "cuda":
            raise RuntimeError("Renderer.render needs the model on a CUDA device; there is no CPU path")
        cameras = cameras.to(device)
        modalities = list(rendered_image_modalities)
        lut8 = None
        if RenderedImageModality.THERMAL in modalities:
            cmap = thermal_color_map if thermal_color_map is not None else _default_thermal_colormap()
            lut8 = torch.from_numpy(lut8_from_colormap(cmap)).to(device)
        self._rendered_images = {m: [] for m in modalities}
        was_training = model.training
        model.eval()
        fused = isinstance(cameras, PinholeCameras)
        staging: Dict[RenderedImageModality, torch.Tensor] = {}
        try:
            with torch.no_grad():
                for camera_idx in range(cameras.size):
                    if fused:
                        outputs = model.get_outputs_for_camera(cameras, camera_idx)
                    else:
                        outputs = model.get_outputs_for_camera_ray_bundle(cameras.generate_rays(camera_indices=camera_idx))
                    images = {}
                    for modality in modalities:
                        key = self._OUTPUT_KEY.get(modality, modality.value)
                        if key not in outputs:
                            raise Exception(f"{modality.value} modality does not exist")
                        img = outputs[key]
                        if img.shape[-1] == 3:
                            images[modality] = F.postprocess_frame(rgb=img.contiguous())[0]
                        else:
                            images[modality] = F.postprocess_frame(
                                scalar=img.contiguous(),
                                lut8=lut8 if modality == RenderedImageModality.THERMAL else None)[1]
                    # uint8 frames -> pinned host staging (async), one sync per frame
                    for modality, img8 in images.items():
                        st = staging.get(modality)
                        if st is None or st.shape != img8.shape:
                            st = torch.empty(img8.shape, dtype=torch.uint8).pin_memory()
                            staging[modality] = st
                        st.copy_(img8, non_blocking=True)
                    torch.cuda.current_stream(device).synchronize()
                    for modality in modalities:
                        self._rendered_images[modality].append(staging[modality].numpy().copy())
        finally:
            model.train(was_training)

    # ------------------------------------------------------------------ export
    def save_images(self, modalities: Sequence[RenderedImageModality], output_dir: Path) -> None:
        """renderer.py:202-213 (imageio when present, PIL otherwise)."""
        for modality in modalities:
            for idx, image in enumerate(self._rendered_images[modality]):
                _imwrite(Path(output_dir) / f"{modality.value}_{idx:05d}.jpeg", image)

    def save_gif(self, modalities: Sequence[RenderedImageModality], seconds: float, output_dir: Path) -> None:
        """renderer.py:215-228."""
        for modality in modalities:
            path = Path(output_dir) / f"synthesized_video_{modality.value}.gif"
            frames = self._rendered_images[modality]
            try:
                import imageio

                imageio.mimsave(path, frames, duration=seconds)
            except ImportError:
                from PIL import Image

                ims = [Image.fromarray(np.asarray(f)) for f in frames]
                ims[0].save(path, save_all=True, append_images=ims[1:], duration=int(seconds * 1000), loop=0)


def _imwrite(path: Path, image: np.ndarray) -> None:
    try:
        import imageio

        imageio.imwrite(path, image)
    except ImportError:
        from PIL import Image

        Image.fromarray(np.asarray(image)).save(path)
