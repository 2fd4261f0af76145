"""CPU oracle (numpy) for the two neighbours of the render path: ray generation
and the per-frame uint8 conversion of ``Renderer.render``.  TEST INFRASTRUCTURE ONLY (see __init__).

PARITY: ``generate_rays_np`` restates nerfstudio-1.1.5 ``Cameras._generate_rays_from_coords`` for
perspective cameras without distortion (the call sites are thermo_nerf/render/renderer.py:183 and
thermo_nerf/evaluator/evaluator.py:69; nerfstudio itself is not available offline, so this part is
unpinned like the rest of the oracle).  ``postprocess_np`` follows thermo_nerf/render/renderer.py:189-199
line by line; ``ListedColormapLike`` restates matplotlib's ``Colormap.__call__`` for float input (matplotlib
is not installed here), and ``camera_path_to_cameras`` restates nerfstudio's ``get_path_from_json`` for the
fields the reference fixture tests/data/trajectories/camera_path_facade_2.json carries.
"""

from __future__ import annotations

import math
from typing import Optional, Sequence

import numpy as np


def generate_rays_np(c2w: np.ndarray, fx: float, fy: float, cx: float, cy: float, height: int, width: int):
    """(origins [H,W,3], directions [H,W,3], directions_norm [H,W,1]) in float32.
    coords = pixel centres (+0.5); camera-frame direction ((x-cx)/fx, -(y-cy)/fy, -1); rotated with
    ``sum(dir[..., None, :] * R, -1)``; ``normalize_with_norm``; origins = c2w[:3, 3]."""
    c2w = np.asarray(c2w, dtype=np.float32).reshape(3, 4)
    ys, xs = np.meshgrid(np.arange(height, dtype=np.float32) + np.float32(0.5),
                         np.arange(width, dtype=np.float32) + np.float32(0.5), indexing="ij")
    x = (xs - np.float32(cx)) / np.float32(fx)
    y = (ys - np.float32(cy)) / np.float32(fy)
    dirs = np.stack([x, -y, -np.ones_like(x)], -1).astype(np.float32)
    prod = dirs[..., None, :] * c2w[:3, :3]  # [H,W,3,3]
    d = (prod[..., 0] + prod[..., 1]) + prod[..., 2]
    norm = np.sqrt((d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1]) + d[..., 2] * d[..., 2])[..., None]
    d = d / norm
    o = np.broadcast_to(c2w[:3, 3], d.shape).copy()
    return o.astype(np.float32), d.astype(np.float32), norm.astype(np.float32)


class ListedColormapLike:
    """matplotlib ``Colormap.__call__`` for a float array in [0,1] (float input path): ``xa = X * N``,
    ``xa == N -> N-1``, under/over/bad handled with the default end colours / transparent black,
    ``xa.astype(int)`` (truncation), ``lut.take(xa, mode='clip')``; returns RGBA float64."""

    def __init__(self, colors: np.ndarray) -> None:
        colors = np.asarray(colors, dtype=np.float64)
        if colors.shape[1] == 3:
            colors = np.concatenate([colors, np.ones((colors.shape[0], 1))], 1)
        self.N = colors.shape[0]
        self._lut = np.concatenate([colors, colors[:1], colors[-1:], np.zeros((1, 4))], 0)  # under, over, bad

    def __call__(self, X):
        xa = np.array(X, copy=True)
        if xa.dtype.kind == "f":
            xa *= self.N
            xa[xa == self.N] = self.N - 1
        mask_under, mask_over, mask_bad = xa < 0, xa >= self.N, np.isnan(xa)
        with np.errstate(invalid="ignore"):
            xa = xa.astype(int)
        xa[mask_under], xa[mask_over], xa[mask_bad] = self.N, self.N + 1, self.N + 2
        return self._lut.take(xa, axis=0, mode="clip")


def colormap_to_lut8(cmap) -> np.ndarray:
    """uint8 table of a matplotlib-style colour map, converted exactly as renderer.py:195-197 converts its
    output: ``(cmap(x)[..., :3] * 255).astype(uint8)`` evaluated on the N bin centres."""
    n = int(cmap.N)
    x = (np.arange(n, dtype=np.float64) + 0.5) / n
    return (np.asarray(cmap(x))[:, :3] * 255).astype(np.uint8)


def postprocess_np(image: np.ndarray, is_thermal: bool, cmap=None) -> np.ndarray:
    """thermo_nerf/render/renderer.py:189-199 for one output image ([H,W,3] or [H,W,1] float32)."""
    output_image = np.asarray(image)
    if output_image.shape[-1] == 1:
        output_image = np.concatenate((output_image,) * 3, axis=-1)
    if is_thermal:
        return (cmap(output_image[:, :, 0])[:, :, :3] * 255).astype(np.uint8)
    return (output_image * 255).astype(np.uint8)


def camera_path_to_cameras(camera_path: dict, scaling: float = 1.0):
    """nerfstudio ``get_path_from_json`` + ``rescale_output_resolution`` (renderer.py:144-157) for a
    perspective path: returns (c2w [N,3,4] float32, fx, fy, cx, cy, height, width) with per-path constant
    intrinsics (every camera of the reference fixture shares one fov)."""
    h, w = int(camera_path["render_height"]), int(camera_path["render_width"])
    c2ws, fxs = [], []
    for cam in camera_path["camera_path"]:
        c2ws.append(np.asarray(cam["camera_to_world"], dtype=np.float32).reshape(4, 4)[:3])
        fov = float(cam["fov"])
        fxs.append(0.5 * h / math.tan(0.5 * fov * math.pi / 180.0))  # three_js_perspective_camera_focal_length
    if len(set(fxs)) != 1:
        raise ValueError("per-camera focal lengths differ")
    fx = fy = fxs[0] * scaling
    cx, cy = w / 2 * scaling, h / 2 * scaling
    # rescale_output_resolution: height/width scaled and truncated to int (scaling_factor float path)
    return np.stack(c2ws), fx, fy, cx, cy, int(h * scaling), int(w * scaling)


def sample_batch_np(rand: np.ndarray, images: np.ndarray, thermal: Optional[np.ndarray], c2w: np.ndarray,
                    intrinsics: np.ndarray):
    """nerfstudio PixelSampler.sample_method (no mask) + collate_image_dataset_batch + RayGenerator for a given
    uniform draw ``rand`` [R,3] (float32): indices = floor(rand * [N,H,W]).long(); batch[key] = value[c, y, x];
    rays through the pixel centres (y + 0.5, x + 0.5) of camera c.  uint8 images are converted as x / 255
    (InputDataset.get_image_float32).  Returns (origins, directions, indices, gt_rgb, gt_thermal)."""
    n, h, w = images.shape[:3]
    dims = np.array([n, h, w], dtype=np.float32)
    idx = np.floor(rand.astype(np.float32) * dims).astype(np.int64)
    c, y, x = idx[:, 0], idx[:, 1], idx[:, 2]
    img = images[c, y, x, :3]
    gt_rgb = img.astype(np.float32) / np.float32(255.0) if images.dtype == np.uint8 else img.astype(np.float32)
    gt_th = None
    if thermal is not None:
        t = thermal.reshape(n, h, w)[c, y, x]
        gt_th = t.astype(np.float32) / np.float32(255.0) if thermal.dtype == np.uint8 else t.astype(np.float32)
    o = np.empty((len(c), 3), np.float32)
    d = np.empty((len(c), 3), np.float32)
    for cam in np.unique(c):
        sel = np.nonzero(c == cam)[0]
        fx, fy, cx, cy = (float(v) for v in intrinsics[cam])
        oo, dd, _ = generate_rays_np(c2w[cam], fx, fy, cx, cy, h, w)
        o[sel], d[sel] = oo[y[sel], x[sel]], dd[y[sel], x[sel]]
    return o, d, idx, gt_rgb, gt_th
