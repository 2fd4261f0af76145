"""Parameter containers of the B200 ThermoNeRF model.

These modules own the tensors only; all arithmetic happens in libtnf_b200.so.  Attribute
names follow the nerfstudio/ThermoNeRF module tree so that ``state_dict()`` keys line up
with the ones a reference checkpoint carries after its ``_model.`` prefix:

* ``ThermalNerfactoTField``   <- thermo_nerf/thermal_nerf/thermal_field.py:33-106
* ``ThermalFieldHead``        <- thermo_nerf/thermal_nerf/thermal_field.py:18-30,
                                 thermal_field_head.py:15-71
* ``HashMLPDensityField``     <- built at thermo_nerf/thermal_nerf/thermal_nerf_model.py:127-148
* ``CameraOptimizer``         <- built at thermal_nerf_model.py:118-120 (nerfstudio SO3xR3)

The leaf containers (``HashEncoding``, ``MLP`` ...) raise from ``forward``: there is deliberately no PyTorch
fallback.  The two field classes carry the reference's Field surface - ``get_density`` / ``get_outputs`` /
``forward`` / ``density_fn`` (thermal_field.py:108-201) - backed by the stand-alone kernels of ``surface.py``
(inference only; training and rendering go through the fused ``Model.get_outputs``).
"""

from __future__ import annotations

from enum import Enum
from typing import Optional

import numpy as np
import torch
from torch import Tensor, nn


class FieldHeadNamesT(Enum):
    """thermal_field_head.py:9-12."""

    THERMAL = "thermal"


class FieldHeadNames(Enum):
    """Subset of nerfstudio FieldHeadNames used on the path."""

    RGB = "rgb"
    DENSITY = "density"


def _no_torch_path(name: str):
    raise RuntimeError(
        f"{name} is a parameter container; its arithmetic runs inside libtnf_b200.so "
        "(thermo_nerf_b200.functional).  There is no PyTorch fallback."
    )


class HashEncoding(nn.Module):
    """nerfstudio HashEncoding (torch layout): ``hash_table`` [L*2^T, F], buffer ``scalings`` [L]."""

    def __init__(self, num_levels: int, min_res: int, max_res: int, log2_hashmap_size: int,
                 features_per_level: int = 2, hash_init_scale: float = 1e-3) -> None:
        super().__init__()
        if features_per_level != 2:
            raise ValueError("libtnf_b200 kernels are built for features_per_level == 2")
        if max_res >= 2048 + 1:  # the scatter's cell key packs grid coordinates into 11 bits (tnf_device.cuh)
            raise ValueError(f"libtnf_b200 kernels are built for max_res <= 2048 (got {max_res}): nerfacto-big / -huge "
                             "grids would alias cells in the hash-table gradient scatter")
        self.num_levels = num_levels
        self.min_res = min_res
        self.max_res = max_res
        self.features_per_level = features_per_level
        self.log2_hashmap_size = log2_hashmap_size
        self.hash_table_size = 2**log2_hashmap_size
        levels = torch.arange(num_levels)
        growth = np.exp((np.log(max_res) - np.log(min_res)) / (num_levels - 1)) if num_levels > 1 else 1
        # same expression as nerfstudio so the float32 rounding (2047 at the top level) agrees
        self.register_buffer("scalings", torch.floor(min_res * growth**levels).to(torch.float32))
        table = torch.rand(size=(self.hash_table_size * num_levels, features_per_level)) * 2 - 1
        self.hash_table = nn.Parameter(table * hash_init_scale)

    def get_out_dim(self) -> int:
        return self.num_levels * self.features_per_level

    def forward(self, *_):
        _no_torch_path("HashEncoding")


class MLP(nn.Module):
    """nerfstudio MLP (torch layout): ``layers`` = ModuleList of nn.Linear (with bias)."""

    def __init__(self, in_dim: int, num_layers: int, layer_width: int, out_dim: int) -> None:
        super().__init__()
        dims = [in_dim] + [layer_width] * (num_layers - 1) + [out_dim]
        self.in_dim, self.out_dim = in_dim, out_dim
        self.layers = nn.ModuleList([nn.Linear(dims[i], dims[i + 1]) for i in range(num_layers)])

    def get_out_dim(self) -> int:
        return self.out_dim

    def forward(self, *_):
        _no_torch_path("MLP")


class MLPWithHashEncoding(nn.Module):
    def __init__(self, num_levels, min_res, max_res, log2_hashmap_size, num_layers, layer_width, out_dim) -> None:
        super().__init__()
        self.encoder = HashEncoding(num_levels, min_res, max_res, log2_hashmap_size)
        self.mlp = MLP(self.encoder.get_out_dim(), num_layers, layer_width, out_dim)

    def forward(self, *_):
        _no_torch_path("MLPWithHashEncoding")


class Embedding(nn.Module):
    def __init__(self, in_dim: int, out_dim: int) -> None:
        super().__init__()
        self.embedding = nn.Embedding(in_dim, out_dim)

    def mean(self, dim: int = 0) -> Tensor:
        return self.embedding.weight.mean(dim)

    def forward(self, *_):
        _no_torch_path("Embedding")


class ThermalFieldHead(nn.Module):
    """Linear(in_dim -> 1), no activation (thermal_field.py:18-30)."""

    def __init__(self, in_dim: int) -> None:
        super().__init__()
        self.in_dim, self.out_dim = in_dim, 1
        self.field_head_name = FieldHeadNamesT.THERMAL
        self.net = nn.Linear(in_dim, 1)

    def forward(self, *_):
        _no_torch_path("ThermalFieldHead")


class _ZeroLinear(nn.Module):
    """A Linear whose weight and bias are constant zeros kept as NON-persistent buffers: present for the kernels
    (which always evaluate the thermal head) but neither a parameter nor a state_dict entry - a plain nerfacto
    field (thermo_nerf/nerfacto_config/thermal_nerfacto.py) has no thermal head."""

    def __init__(self, in_dim: int, out_dim: int) -> None:
        super().__init__()
        self.register_buffer("weight", torch.zeros(out_dim, in_dim), persistent=False)
        self.register_buffer("bias", torch.zeros(out_dim), persistent=False)

    def forward(self, *_):
        _no_torch_path("_ZeroLinear")


class _Seq(nn.Sequential):
    def forward(self, *_):  # type: ignore[override]
        _no_torch_path("mlp_base")


class HashMLPDensityField(nn.Module):
    """Proposal density network: hash grid -> 16 -> 1 (thermal_nerf_model.py:127-148)."""

    def __init__(self, aabb: Tensor, num_layers: int = 2, hidden_dim: int = 64, num_levels: int = 8,
                 max_res: int = 1024, base_res: int = 16, log2_hashmap_size: int = 18,
                 features_per_level: int = 2, use_contraction: bool = True) -> None:
        super().__init__()
        self.use_contraction = bool(use_contraction)  # nerfstudio: spatial_distortion=SceneContraction(inf) or None
        if num_layers != 2 or hidden_dim != 16:
            raise ValueError("libtnf_b200 proposal kernels are built for num_layers=2, hidden_dim=16")
        if num_levels > 8:
            raise ValueError("libtnf_b200 proposal kernels support at most 8 hash levels")
        self.register_buffer("aabb", aabb)
        self.encoding = HashEncoding(num_levels, base_res, max_res, log2_hashmap_size, features_per_level)
        network = MLP(self.encoding.get_out_dim(), num_layers, hidden_dim, 1)
        self.mlp_base = _Seq(self.encoding, network)

    # nerfstudio DensityField surface (HashMLPDensityField.get_density / Field.density_fn / Field.forward)
    def get_density(self, ray_samples):
        from . import surface

        surface._no_grad_surface(self, "get_density")
        return surface.density_at(self, ray_samples.frustums.get_positions())[0], None

    def density_fn(self, positions: Tensor, times: Optional[Tensor] = None) -> Tensor:
        from . import surface

        return surface.density_fn(self, positions, times)

    def get_outputs(self, ray_samples, density_embedding: Optional[Tensor] = None) -> dict:
        return {}

    def forward(self, ray_samples, compute_normals: bool = False) -> dict:
        return {FieldHeadNames.DENSITY: self.get_density(ray_samples)[0]}


class ThermalNerfactoTField(nn.Module):
    """ThermoNeRF field parameters (thermal_field.py:33-106).  Fixed architecture: 16x2 hash
    grid, base MLP 32->64->16, colour head 63->64->64->3, thermal MLP 15->64->64 + Linear 64->1."""

    def __init__(self, aabb: Tensor, num_images: int, num_layers: int = 2, hidden_dim: int = 64,
                 geo_feat_dim: int = 15, num_levels: int = 16, base_res: int = 16, max_res: int = 2048,
                 log2_hashmap_size: int = 19, num_layers_color: int = 3, features_per_level: int = 2,
                 hidden_dim_color: int = 64, hidden_dim_transient: int = 64, appearance_embedding_dim: int = 32,
                 use_average_appearance_embedding: bool = False, pass_thermal_gradients: bool = False,
                 thermal_head: bool = True, use_contraction: bool = True, rgb_out_dim: int = 3) -> None:
        super().__init__()
        if rgb_out_dim not in (3, 4):
            raise ValueError("the colour head has 3 outputs, or 4 for the RGBT head of concat_nerf")
        if rgb_out_dim == 4 and thermal_head:
            raise ValueError("the RGBT field (rgb_concat/concat_field.py) has no separate thermal head")
        self.use_contraction = bool(use_contraction)  # nerfstudio: spatial_distortion=SceneContraction(inf) or None
        fixed = dict(num_layers=(num_layers, 2), hidden_dim=(hidden_dim, 64), geo_feat_dim=(geo_feat_dim, 15),
                     num_levels=(num_levels, 16), num_layers_color=(num_layers_color, 3),
                     hidden_dim_color=(hidden_dim_color, 64), hidden_dim_transient=(hidden_dim_transient, 64),
                     appearance_embedding_dim=(appearance_embedding_dim, 32))
        for k, (got, want) in fixed.items():
            if got != want:
                raise ValueError(f"libtnf_b200 field kernels are built for {k}={want}, got {got}")
        self.register_buffer("aabb", aabb)
        self.geo_feat_dim = geo_feat_dim
        self.appearance_embedding_dim = appearance_embedding_dim
        self.use_average_appearance_embedding = use_average_appearance_embedding
        self.mlp_base = MLPWithHashEncoding(num_levels, base_res, max_res, log2_hashmap_size, num_layers,
                                            hidden_dim, 1 + geo_feat_dim)
        self.embedding_appearance = Embedding(num_images, appearance_embedding_dim)
        # rgb_out_dim = 4: ConcatNerfactoTField (rgb_concat/concat_field.py:65-75), temperature = colour channel 3
        self.mlp_head = MLP(16 + geo_feat_dim + appearance_embedding_dim, num_layers_color, hidden_dim_color,
                            rgb_out_dim)
        self.mlp_thermal = MLP(geo_feat_dim, 2, 64, hidden_dim_transient)
        self.field_head_thermal = ThermalFieldHead(in_dim=self.mlp_thermal.get_out_dim())
        self.thermal_head = bool(thermal_head)
        if not self.thermal_head:  # nerfacto field: the head exists only as constant zeros for the kernels
            self.mlp_thermal.layers = nn.ModuleList([_ZeroLinear(geo_feat_dim, 64), _ZeroLinear(64, hidden_dim_transient)])
            self.field_head_thermal.net = _ZeroLinear(hidden_dim_transient, 1)
            pass_thermal_gradients = False
        self.pass_thermal_gradients = pass_thermal_gradients  # thermal_field.py:103
        self.training_iteration = 0
        self.pass_rgb_gradients = True  # thermal_field.py:106

    # the reference's Field surface (thermal_field.py:108-201), kernel backed (surface.py)
    def get_density(self, ray_samples):
        from . import surface

        return surface.field_get_density(self, ray_samples)

    def get_outputs(self, ray_samples, density_embedding: Optional[Tensor] = None) -> dict:
        from . import surface

        return surface.field_get_outputs(self, ray_samples, density_embedding)

    def forward(self, ray_samples, compute_normals: bool = False) -> dict:
        from . import surface

        return surface.field_forward(self, ray_samples, compute_normals)

    def density_fn(self, positions: Tensor, times: Optional[Tensor] = None) -> Tensor:
        from . import surface

        return surface.density_fn(self, positions, times)


def exp_map_so3xr3(tangent: Tensor) -> Tensor:
    """[N,6] (translation | so3 log) -> [N,3,4]; nerfstudio's SO3xR3 pose delta."""
    log_rot = tangent[:, 3:]
    nrms = (log_rot * log_rot).sum(1)
    angle = torch.clamp(nrms, 1e-4).sqrt()
    inv = 1.0 / angle
    fac1 = inv * angle.sin()
    fac2 = inv * inv * (1.0 - angle.cos())
    zeros = torch.zeros_like(log_rot[:, 0])
    skew = torch.stack([
        torch.stack([zeros, -log_rot[:, 2], log_rot[:, 1]], -1),
        torch.stack([log_rot[:, 2], zeros, -log_rot[:, 0]], -1),
        torch.stack([-log_rot[:, 1], log_rot[:, 0], zeros], -1),
    ], 1)
    rot = fac1[:, None, None] * skew + fac2[:, None, None] * torch.bmm(skew, skew)
    rot = rot + torch.eye(3, dtype=tangent.dtype, device=tangent.device)[None]
    return torch.cat([rot, tangent[:, :3, None]], dim=-1)


class CameraOptimizer(nn.Module):
    """Per-camera pose deltas (row a2 of SURVEY 8a): kept in PyTorch, upstream of the kernel."""

    def __init__(self, num_cameras: int, mode: str = "SO3xR3") -> None:
        super().__init__()
        if mode not in ("off", "SO3xR3"):
            raise ValueError(f"camera optimizer mode {mode!r} is not supported (off | SO3xR3)")
        self.mode = mode
        self.num_cameras = num_cameras
        self.pose_adjustment = nn.Parameter(torch.zeros((num_cameras, 6)))

    def forward(self, indices: Tensor) -> Tensor:
        if self.mode == "off":
            return torch.eye(4, device=indices.device)[None, :3, :4].tile(indices.shape[0], 1, 1)
        return exp_map_so3xr3(self.pose_adjustment[indices, :])

    def apply_to_raybundle(self, raybundle) -> None:
        """thermal_nerf_model.py:218-219 / evaluator.py:71-73: in-place origin/direction update."""
        if self.mode == "off":
            return
        c = self(raybundle.camera_indices.squeeze(-1) if raybundle.camera_indices.dim() > 1
                 else raybundle.camera_indices)
        raybundle.origins = raybundle.origins + c[:, :3, 3]
        raybundle.directions = torch.bmm(c[:, :3, :3], raybundle.directions[..., None]).squeeze(-1)

    def get_param_groups(self, param_groups: dict) -> None:
        if self.mode != "off":
            param_groups["camera_opt"] = list(self.parameters())


def hash_grid_of(module: nn.Module) -> Optional[nn.Module]:
    """Find the HashEncoding-like submodule (``hash_table`` + ``scalings``) of a field part."""
    for m in module.modules():
        if hasattr(m, "hash_table") and hasattr(m, "scalings"):
            return m
    return None
