"""Kernel-backed Field / Renderer plugin surface.

The per-module entry points of the reference's field and renderer classes, for callers that compose the modules
by hand instead of going through ``Model.get_outputs`` (viewer density queries, export scripts, user code):

* ``field_get_density``   <- nerfstudio ``NerfactoField.get_density`` as reached from
                             ``ThermalNerfactoTField.forward`` (thermo_nerf/thermal_nerf/thermal_field.py:183-201)
* ``field_get_outputs``   <- ``ThermalNerfactoTField.get_outputs`` (thermal_field.py:108-181)
* ``field_forward``       <- ``ThermalNerfactoTField.forward`` (thermal_field.py:183-201)
* ``density_fn``          <- ``Field.density_fn`` / ``HashMLPDensityField.density_fn`` (thermal_nerf_model.py:127-148)
* ``ThermalRenderer``     <- thermo_nerf/thermal_nerf/thermal_renderer.py:14-149
* ``RGBTRenderer``        <- thermo_nerf/rgb_concat/rgbt_renderer.py:17-174 (concat baseline)

The functions take any module whose attribute tree follows nerfstudio's ``implementation="torch"`` layout: this
package's parameter containers (``modules.py``) and the stock nerfstudio modules the reference builds
(``nerfstudio_plugin.py`` binds them as methods).  They run ``tnf_field_density`` / ``tnf_field_heads`` /
``tnf_composite`` (fp32) on the current CUDA stream.  Inference only: outputs carry no autograd graph - training
goes through the fused ``Model.get_outputs``.  There is no PyTorch fallback.
"""

from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Tuple

import torch
from torch import Tensor, nn

from . import _lib as L
from .functional import ModelTensors, _dev_f32
from .modules import FieldHeadNames, FieldHeadNamesT, hash_grid_of


def _no_grad_surface(module: nn.Module, what: str) -> None:
    if module.training and torch.is_grad_enabled():
        raise RuntimeError(
            f"{what}: the kernel-backed field surface is inference only (outputs carry no autograd graph); "
            "training runs through the fused Model.get_outputs.  Call it under torch.no_grad() or in eval mode.")


def _geometry(module, m: L.TnfModel) -> None:
    """Scene contraction vs. aabb normalisation, from the module (nerfstudio: ``spatial_distortion``)."""
    if hasattr(module, "use_contraction"):
        contraction = bool(module.use_contraction)
    else:
        contraction = getattr(module, "spatial_distortion", None) is not None
    m.use_contraction = int(contraction)
    aabb = getattr(module, "aabb", None)
    box = [-1.0, -1.0, -1.0, 1.0, 1.0, 1.0] if aabb is None else [float(v) for v in aabb.reshape(-1).tolist()]
    for i in range(6):
        m.aabb[i] = box[i]


def _pack_field(field, *, heads: bool, appearance_mode: int) -> L.TnfModel:
    m = L.TnfModel()
    mlp_base = field.mlp_base
    enc = getattr(mlp_base, "encoder", None) or hash_grid_of(mlp_base)
    ModelTensors._fill_grid(m.field.grid, ModelTensors._grid_of(enc), "field.mlp_base.encoder")
    layers = mlp_base.mlp.layers
    ModelTensors._fill_lin(m.field.base0, ModelTensors._lin(layers[0]), "field.mlp_base.0")
    ModelTensors._fill_lin(m.field.base1, ModelTensors._lin(layers[1]), "field.mlp_base.1")
    if heads:
        for k, layer in (("rgb0", field.mlp_head.layers[0]), ("rgb1", field.mlp_head.layers[1]),
                         ("rgb2", field.mlp_head.layers[2]), ("th0", field.mlp_thermal.layers[0]),
                         ("th1", field.mlp_thermal.layers[1]), ("th2", field.field_head_thermal.net)):
            ModelTensors._fill_lin(getattr(m.field, k), ModelTensors._lin(layer), "field." + k)
        app = field.embedding_appearance.embedding.weight
        m.field.appearance = _dev_f32(app, "field.embedding_appearance").data_ptr()
        m.field.num_images = int(app.shape[0])
    m.appearance_mode = int(appearance_mode)
    for i, s in enumerate((256, 96, 48)):
        m.num_samples[i] = s
    _geometry(field, m)
    return m


def _pack_density_net(net, which: int = 0) -> L.TnfModel:
    m = L.TnfModel()
    ModelTensors._fill_grid(m.prop[which].grid, ModelTensors._grid_of(net.encoding), "proposal.encoding")
    mlp = net.mlp_base[1]
    ModelTensors._fill_lin(m.prop[which].l0, ModelTensors._lin(mlp.layers[0]), "proposal.mlp.0")
    ModelTensors._fill_lin(m.prop[which].l1, ModelTensors._lin(mlp.layers[1]), "proposal.mlp.1")
    _geometry(net, m)
    return m


def _stream(dev) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def _flat_f32(t: Tensor, width: int, name: str) -> Tensor:
    return _dev_f32(t.detach().reshape(-1, width).float().contiguous(), name)


# --------------------------------------------------------------------------------------------- field
def density_at(module, positions: Tensor, *, want_geo: bool = False) -> Tuple[Tensor, Optional[Tensor]]:
    """density [*bs,1] (and the 15 geo features [*bs,15]) at world-space ``positions`` [*bs,3].  ``module`` is a
    field (``mlp_base.encoder`` + ``mlp_base.mlp``) or a proposal density network (``encoding`` + ``mlp_base[1]``)."""
    shape = tuple(positions.shape[:-1])
    pos = _flat_f32(positions, 3, "positions")
    n, dev = int(pos.shape[0]), pos.device
    is_prop = hasattr(module, "encoding")
    if is_prop and want_geo:
        raise ValueError("a proposal density network has no geometry features")
    m = _pack_density_net(module) if is_prop else _pack_field(module, heads=False, appearance_mode=L.APPEARANCE_ZEROS)
    density = torch.empty((max(n, 1),), dtype=torch.float32, device=dev)[:n]
    geo = torch.empty((max(n, 1), 15), dtype=torch.float32, device=dev)[:n] if want_geo else None
    with torch.cuda.device(dev):
        rc = L.load().tnf_field_density(C.byref(m), 0 if is_prop else -1, pos.data_ptr(), n, density.data_ptr(),
                                        geo.data_ptr() if geo is not None else None, _stream(dev))
    L.check(rc)
    return density.view(*shape, 1), (geo.view(*shape, 15) if geo is not None else None)


def field_get_density(field, ray_samples) -> Tuple[Tensor, Tensor]:
    """``Field.get_density(ray_samples) -> (density [*bs,1], density_embedding [*bs,15])``."""
    _no_grad_surface(field, "get_density")
    return density_at(field, ray_samples.frustums.get_positions(), want_geo=True)


def density_fn(module, positions: Tensor, times: Optional[Tensor] = None) -> Tensor:
    """``Field.density_fn(positions) -> density [*bs,1]`` (nerfstudio wraps the points in dummy RaySamples)."""
    del times
    _no_grad_surface(module, "density_fn")
    return density_at(module, positions)[0]


def field_get_outputs(field, ray_samples, density_embedding: Optional[Tensor] = None) -> Dict[object, Tensor]:
    """``ThermalNerfactoTField.get_outputs`` (thermal_field.py:108-181): keys ``FieldHeadNames.RGB`` [*bs,3] and
    ``FieldHeadNamesT.THERMAL`` [*bs,1] from the sample directions, the camera indices and the 15 geo features."""
    assert density_embedding is not None  # thermal_field.py:111
    if ray_samples.camera_indices is None:
        raise AttributeError("Camera indices are not provided.")  # thermal_field.py:113-114
    _no_grad_surface(field, "get_outputs")
    directions = ray_samples.frustums.directions
    shape = tuple(directions.shape[:-1])
    d = _flat_f32(directions, 3, "directions")
    geo = _flat_f32(density_embedding, 15, "density_embedding")
    n, dev = int(d.shape[0]), d.device
    if geo.shape[0] != n:
        raise ValueError(f"density_embedding has {geo.shape[0]} rows for {n} samples")
    if field.training:
        mode = L.APPEARANCE_LOOKUP
    else:
        mode = L.APPEARANCE_MEAN if getattr(field, "use_average_appearance_embedding", False) else L.APPEARANCE_ZEROS
    cam = None
    if mode == L.APPEARANCE_LOOKUP:
        cam = ray_samples.camera_indices.reshape(-1).to(torch.int64).contiguous()
        if cam.numel() != n or not cam.is_cuda:
            raise ValueError("camera_indices must be a CUDA tensor with one index per sample")
    m = _pack_field(field, heads=True, appearance_mode=mode)
    rgb = torch.empty((max(n, 1), 3), dtype=torch.float32, device=dev)[:n]
    thermal = torch.empty((max(n, 1),), dtype=torch.float32, device=dev)[:n]
    with torch.cuda.device(dev):
        rc = L.load().tnf_field_heads(C.byref(m), d.data_ptr(), cam.data_ptr() if cam is not None else None,
                                      geo.data_ptr(), n, rgb.data_ptr(), thermal.data_ptr(), _stream(dev))
    L.check(rc)
    out: Dict[object, Tensor] = {_rgb_key(field): rgb.view(*shape, 3)}
    if getattr(field, "thermal_head", True):
        out[_thermal_key(field)] = thermal.view(*shape, 1)
    return out


def _rgb_key(field):
    return getattr(field, "_field_head_names", FieldHeadNames).RGB


def _density_key(field):
    return getattr(field, "_field_head_names", FieldHeadNames).DENSITY


def _thermal_key(field):
    return getattr(field, "_field_head_names_t", FieldHeadNamesT).THERMAL


def field_forward(field, ray_samples, compute_normals: bool = False) -> Dict[object, Tensor]:
    """``ThermalNerfactoTField.forward`` (thermal_field.py:183-201): get_density, get_outputs, + DENSITY."""
    if compute_normals:
        raise ValueError("compute_normals needs d density / d position through autograd: not on the thermal-nerf "
                         "hot path (predict_normals=False) and not built into libtnf_b200")
    density, embedding = field_get_density(field, ray_samples)
    outputs = field_get_outputs(field, ray_samples, density_embedding=embedding)
    outputs[_density_key(field)] = density
    return outputs


# --------------------------------------------------------------------------------------------- renderers
def composite(values: Tensor, weights: Tensor, *, last_sample_background: bool, eval_mode: bool) -> Tensor:
    """``tnf_composite``: values [*bs,S,C], weights [*bs,S,1] -> [*bs,C]."""
    if values.dim() < 2 or weights.shape[:-1] != values.shape[:-1] or weights.shape[-1] != 1:
        raise ValueError(f"values {tuple(values.shape)} / weights {tuple(weights.shape)}: expected [*bs,S,C] and [*bs,S,1]")
    bs, S, Cc = tuple(values.shape[:-2]), int(values.shape[-2]), int(values.shape[-1])
    v = _dev_f32(values.detach().reshape(-1, S, Cc).float().contiguous(), "values")
    w = _dev_f32(weights.detach().reshape(-1, S).float().contiguous(), "weights")
    R, dev = int(v.shape[0]), v.device
    out = torch.empty((max(R, 1), Cc), dtype=torch.float32, device=dev)[:R]
    with torch.cuda.device(dev):
        rc = L.load().tnf_composite(v.data_ptr(), w.data_ptr(), R, S, Cc, int(last_sample_background), int(eval_mode),
                                    out.data_ptr(), _stream(dev))
    L.check(rc)
    return out.view(*bs, Cc)


class ThermalRenderer(nn.Module):
    """thermo_nerf/thermal_nerf/thermal_renderer.py:14-149.  The reference forces the background to the last
    sample whatever ``background_color`` says (:49) and rejects packed samples (:50-53); eval mode applies
    ``nan_to_num`` to the samples and clamps the result to [0,1] (:136-137,146-147)."""

    def __init__(self, background_color="random") -> None:
        super().__init__()
        self.background_color = background_color

    def forward(self, thermal: Tensor, weights: Tensor, ray_indices: Optional[Tensor] = None,
                num_rays: Optional[int] = None, background_color=None) -> Tensor:
        if ray_indices is not None and num_rays is not None:
            raise NotImplementedError("Background color 'last_sample' not implemented for packed samples.")
        _no_grad_surface(self, "ThermalRenderer.forward")
        return composite(thermal, weights, last_sample_background=True, eval_mode=not self.training)


class RGBTRenderer(nn.Module):
    """thermo_nerf/rgb_concat/rgbt_renderer.py: the concat baseline's 4-channel renderer.  With its default
    "random" background the composite carries no background term (:63-71); "last_sample" adds the last sample."""

    def __init__(self, background_color="random") -> None:
        super().__init__()
        if background_color not in ("random", "last_sample"):
            raise ValueError("libtnf_b200 composites with background 'random' (none) or 'last_sample'")
        self.background_color = background_color

    def forward(self, rgb: Tensor, weights: Tensor, ray_indices: Optional[Tensor] = None,
                num_rays: Optional[int] = None, background_color=None) -> Tensor:
        if ray_indices is not None and num_rays is not None:
            raise NotImplementedError("packed samples (nerfacc) are not reachable from the proposal sampler")
        _no_grad_surface(self, "RGBTRenderer.forward")
        bg = background_color if background_color is not None else self.background_color
        return composite(rgb, weights, last_sample_background=(bg == "last_sample"), eval_mode=not self.training)
