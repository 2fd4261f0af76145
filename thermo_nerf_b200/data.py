"""Device-resident training data: nerfstudio's pixel sampler, batch collation and ray
generator (``VanillaDataManager.next_train``; driven from thermo_nerf/nerfstudio_config/
pipeline_tracking.py:47-59) as one ``tnf_sample_batch`` launch.

``DevicePixelSampler.next_train(step)`` returns what ``datamanager.next_train(step)`` returns - a ``RayBundle``
and a batch dict with ``"image"`` [R,3], ``"thermal"`` [R,1] and ``"indices"`` [R,3] - but nothing crosses PCIe:
the RGB and the thermal images (which the reference keeps on the host, thermal_dataset.py:18-20) both live in HBM.
"""

from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Tuple

import torch
from torch import Tensor

from . import _lib as L
from .rays import RayBundle


class DevicePixelSampler:
    def __init__(self, images: Tensor, thermal: Optional[Tensor], camera_to_worlds: Tensor, fx, fy, cx, cy,
                 device="cuda", seed: Optional[int] = None) -> None:
        """``images`` [N,H,W,C>=3] float32 in [0,1] or uint8; ``thermal`` [N,H,W] / [N,H,W,1] float32 or uint8;
        ``camera_to_worlds`` [N,3,4]; intrinsics scalars or [N] tensors (nerfstudio ``Cameras`` fields)."""
        dev = torch.device(device)
        if dev.type != "cuda":
            raise RuntimeError("DevicePixelSampler keeps the dataset in GPU memory; there is no CPU path")
        if dev.index is None:
            dev = torch.device("cuda", torch.cuda.current_device())
        if images.dim() != 4 or images.shape[-1] < 3 or images.dtype not in (torch.float32, torch.uint8):
            raise ValueError("images must be [N,H,W,C>=3] float32 or uint8")
        self.images = images.to(dev).contiguous()
        n, h, w, c = self.images.shape
        self.thermal = None
        if thermal is not None:
            if thermal.dtype not in (torch.float32, torch.uint8):
                raise ValueError("thermal must be float32 or uint8")
            t = thermal.reshape(n, h, w) if thermal.numel() == n * h * w else None
            if t is None:
                raise ValueError("thermal must hold one value per pixel of every image")
            self.thermal = t.to(dev).contiguous()
        self.c2w = camera_to_worlds.to(dev, torch.float32).reshape(n, 3, 4).contiguous()

        def per_cam(v):
            t = torch.as_tensor(v, dtype=torch.float32).reshape(-1)
            return t.expand(n) if t.numel() == 1 else t.reshape(n)

        self.intrinsics = torch.stack([per_cam(fx), per_cam(fy), per_cam(cx), per_cam(cy)], 1).to(dev).contiguous()
        self.device = dev
        self.num_images, self.height, self.width, self.channels = n, h, w, c
        self.generator = torch.Generator(device=dev)
        if seed is not None:
            self.generator.manual_seed(seed)
        ds = L.TnfDataset()
        ds.images, ds.thermal = self.images.data_ptr(), (self.thermal.data_ptr() if self.thermal is not None else 0)
        ds.camera_to_worlds, ds.intrinsics = self.c2w.data_ptr(), self.intrinsics.data_ptr()
        ds.num_images, ds.height, ds.width, ds.channels = n, h, w, c
        ds.images_uint8 = int(self.images.dtype == torch.uint8)
        ds.thermal_uint8 = int(self.thermal is not None and self.thermal.dtype == torch.uint8)
        self._struct = ds

    def sample(self, num_rays: int, rand: Optional[Tensor] = None) -> Tuple[RayBundle, Dict[str, Tensor]]:
        """One training batch.  ``rand`` [R,3] uniform [0,1) reproduces a given draw (PixelSampler uses
        ``torch.rand((R,3))``); by default it is drawn from this sampler's device generator."""
        lib = L.load()
        dev = self.device
        R = int(num_rays)
        if rand is None:
            rand = torch.rand((R, 3), device=dev, generator=self.generator)
        if rand.shape != (R, 3) or rand.dtype != torch.float32 or rand.device != dev:
            raise ValueError("rand must be a float32 [R,3] tensor on the sampler's device")
        rand = rand.contiguous()
        o = torch.empty((R, 3), dtype=torch.float32, device=dev)
        d = torch.empty((R, 3), dtype=torch.float32, device=dev)
        cam = torch.empty((R, 1), dtype=torch.int64, device=dev)
        idx = torch.empty((R, 3), dtype=torch.int64, device=dev)
        rgb = torch.empty((R, 3), dtype=torch.float32, device=dev)
        th = torch.empty((R, 1), dtype=torch.float32, device=dev) if self.thermal is not None else None
        stream = torch.cuda.current_stream(dev).cuda_stream
        with torch.cuda.device(dev):
            rc = lib.tnf_sample_batch(C.byref(self._struct), rand.data_ptr(), R, o.data_ptr(), d.data_ptr(), cam.data_ptr(),
                                      idx.data_ptr(), rgb.data_ptr(), th.data_ptr() if th is not None else None,
                                      C.c_void_p(stream))
        L.check(rc)
        batch = {"image": rgb, "indices": idx}
        if th is not None:
            batch["thermal"] = th
        return RayBundle(origins=o, directions=d, camera_indices=cam), batch

    def next_train(self, step: int, num_rays: int = 4096) -> Tuple[RayBundle, Dict[str, Tensor]]:
        """``VanillaDataManager.next_train`` shape: (ray_bundle, batch)."""
        return self.sample(num_rays)
