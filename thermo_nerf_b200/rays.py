"""Minimal stand-ins for nerfstudio's ``RayBundle`` / ``Cameras`` used when nerfstudio is
not importable.  The model only touches the attributes nerfstudio's own objects expose
(``origins``, ``directions``, ``camera_indices``, ``nears``, ``fars``, ``shape``,
``flatten()``, ``get_row_major_sliced_ray_bundle``), so a real nerfstudio RayBundle works
unchanged (thermo_nerf/render/renderer.py:183-185, evaluator/evaluator.py:69-79).
"""

from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, Optional, Tuple

import torch
from torch import Tensor


@dataclass
class RayBundle:
    origins: Tensor  # [*bs, 3]
    directions: Tensor  # [*bs, 3]
    pixel_area: Optional[Tensor] = None  # [*bs, 1]
    camera_indices: Optional[Tensor] = None  # [*bs, 1] int64
    nears: Optional[Tensor] = None  # [*bs, 1]
    fars: Optional[Tensor] = None  # [*bs, 1]
    metadata: Dict[str, Tensor] = field(default_factory=dict)

    _TENSOR_FIELDS = ("origins", "directions", "pixel_area", "camera_indices", "nears", "fars")

    @property
    def shape(self) -> Tuple[int, ...]:
        return tuple(self.origins.shape[:-1])

    def __len__(self) -> int:
        return int(self.origins.shape[0])

    def _map(self, fn) -> "RayBundle":
        kw = {k: (fn(getattr(self, k)) if getattr(self, k) is not None else None) for k in self._TENSOR_FIELDS}
        kw["metadata"] = {k: fn(v) for k, v in self.metadata.items()}
        return RayBundle(**kw)

    def flatten(self) -> "RayBundle":
        return self._map(lambda t: t.reshape(-1, t.shape[-1]))

    def reshape(self, shape) -> "RayBundle":
        return self._map(lambda t: t.reshape(*shape, t.shape[-1]))

    def to(self, device) -> "RayBundle":
        return self._map(lambda t: t.to(device))

    def get_row_major_sliced_ray_bundle(self, start_idx: int, end_idx: int) -> "RayBundle":
        return self.flatten()._map(lambda t: t[start_idx:end_idx])


@dataclass
class PinholeCameras:
    """Pinhole cameras, nerfstudio conventions: camera looks down -z, +x right, +y up;
    pixel centres at +0.5; directions normalised (``Cameras.generate_rays``)."""

    camera_to_worlds: Tensor  # [N,3,4]
    fx: float
    fy: float
    cx: float
    cy: float
    width: int
    height: int

    @property
    def size(self) -> int:
        return int(self.camera_to_worlds.shape[0])

    def to(self, device) -> "PinholeCameras":
        return PinholeCameras(self.camera_to_worlds.to(device), self.fx, self.fy, self.cx, self.cy, self.width,
                              self.height)

    def generate_rays(self, camera_indices: int) -> RayBundle:
        c2w = self.camera_to_worlds[camera_indices]
        dev = c2w.device
        ys, xs = torch.meshgrid(torch.arange(self.height, device=dev, dtype=torch.float32) + 0.5,
                                torch.arange(self.width, device=dev, dtype=torch.float32) + 0.5, indexing="ij")
        dirs = torch.stack([(xs - self.cx) / self.fx, -(ys - self.cy) / self.fy, -torch.ones_like(xs)], -1)
        d = (dirs[..., None, :] * c2w[:3, :3]).sum(-1)
        norm = d.norm(dim=-1, keepdim=True)
        d = d / norm
        o = c2w[:3, 3].expand_as(d).contiguous()
        cam = torch.full((*d.shape[:-1], 1), int(camera_indices), dtype=torch.int64, device=dev)
        return RayBundle(origins=o, directions=d.contiguous(), pixel_area=torch.ones_like(norm), camera_indices=cam,
                         metadata={"directions_norm": norm})


    def generate_pixel_rays(self, camera_indices: Tensor, ys: Tensor, xs: Tensor) -> RayBundle:
        """Rays through pixels (ys, xs) of cameras ``camera_indices`` ([R] each): the output of nerfstudio's
        pixel sampler + ray generator (``datamanager.next_train``) for one training batch."""
        c2w = self.camera_to_worlds[camera_indices]  # [R,3,4]
        x = (xs.to(torch.float32) + 0.5 - self.cx) / self.fx
        y = -(ys.to(torch.float32) + 0.5 - self.cy) / self.fy
        dirs = torch.stack([x, y, -torch.ones_like(x)], -1)
        d = (dirs[:, None, :] * c2w[:, :3, :3]).sum(-1)
        norm = d.norm(dim=-1, keepdim=True)
        d = d / norm
        return RayBundle(origins=c2w[:, :3, 3].contiguous(), directions=d.contiguous(), pixel_area=torch.ones_like(norm),
                         camera_indices=camera_indices.reshape(-1, 1).to(torch.int64),
                         metadata={"directions_norm": norm})


def sphere_cameras(n: int, radius: float = 0.8, hw: int = 800, focal: float = 1111.1, device="cpu") -> PinholeCameras:
    """n cameras on a Fibonacci sphere looking at the origin: the shape of a ThermoScenes capture after the
    dataparser's auto-orient / auto-scale to +-1 (thermal_dataparser.py:219-225)."""
    k = torch.arange(n, dtype=torch.float32) + 0.5
    phi = torch.acos(1 - 2 * k / n)
    theta = torch.pi * (1 + 5**0.5) * k
    pos = radius * torch.stack([torch.cos(theta) * torch.sin(phi), torch.sin(theta) * torch.sin(phi), torch.cos(phi)], -1)
    back = pos / pos.norm(dim=-1, keepdim=True)
    up = torch.tensor([0.0, 0.0, 1.0]).expand_as(back)
    right = torch.linalg.cross(up, back)
    right = right / right.norm(dim=-1, keepdim=True).clamp_min(1e-6)
    true_up = torch.linalg.cross(back, right)
    c2w = torch.stack([right, true_up, back, pos], dim=-1)
    return PinholeCameras(c2w.to(device), focal, focal, hw / 2, hw / 2, hw, hw)


def orbit_cameras(n: int, radius: float = 0.8, height: float = 0.25, hw: int = 800, focal: float = 1111.1,
                  device="cpu") -> PinholeCameras:
    """n cameras on a circle looking at the origin (synthetic ThermoScenes-shaped path)."""
    ang = torch.arange(n, dtype=torch.float32) * (2 * torch.pi / max(n, 1))
    pos = torch.stack([radius * torch.cos(ang), radius * torch.sin(ang), torch.full_like(ang, height)], -1)
    back = pos / pos.norm(dim=-1, keepdim=True)  # camera +z points away from the target
    up = torch.tensor([0.0, 0.0, 1.0]).expand_as(back)
    right = torch.linalg.cross(up, back)
    right = right / right.norm(dim=-1, keepdim=True)
    true_up = torch.linalg.cross(back, right)
    c2w = torch.stack([right, true_up, back, pos], dim=-1)  # [n,3,4]
    return PinholeCameras(c2w.to(device), focal, focal, hw / 2, hw / 2, hw, hw)
