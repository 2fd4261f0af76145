"""Host-side entry points over the C ABI: pack a module tree into ``TnfModel`` and run
``tnf_render_forward`` on the current CUDA stream.

``ModelTensors.from_module`` walks the *reference's* attribute paths
(``field.mlp_base.encoder.hash_table``, ``proposal_networks[i].mlp_base[1].layers`` ...,
i.e. the nerfstudio ``implementation="torch"`` module tree that
thermo_nerf/thermal_nerf/thermal_nerf_model.py:86-208 builds), so the same packing works
on a stock ThermoNeRF model and on ``thermo_nerf_b200.model.ThermalNerfModel``.
"""

from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import torch
from torch import Tensor

from . import _lib as L


def _dev_f32(t: Tensor, name: str) -> Tensor:
    if not isinstance(t, Tensor):
        raise TypeError(f"{name}: expected a tensor, got {type(t)}")
    if not t.is_cuda:
        raise RuntimeError(f"{name}: libtnf_b200 needs CUDA tensors (got device {t.device}); there is no CPU path")
    if t.dtype != torch.float32:
        raise TypeError(f"{name}: expected float32, got {t.dtype}")
    if not t.is_contiguous():
        raise ValueError(f"{name}: tensor must be contiguous")
    return t


@dataclass
class _Grid:
    table: Tensor
    scalings: List[float]
    num_levels: int
    log2_size: int


@dataclass
class _Linear:
    weight: Tensor
    bias: Tensor


@dataclass
class ModelTensors:
    """References to every tensor of the path plus the host-side constants."""

    prop_grids: List[_Grid]
    prop_l0: List[_Linear]
    prop_l1: List[_Linear]
    field_grid: _Grid
    field_linears: Dict[str, _Linear]
    appearance: Tensor
    extra: dict = field(default_factory=dict)

    # ------------------------------------------------------------------ construction
    @staticmethod
    def _grid_of(enc) -> _Grid:
        table = enc.hash_table
        scal = enc.scalings
        num_levels = int(scal.shape[0])
        if table.dim() != 2 or table.shape[1] != 2:
            raise ValueError(f"hash_table must be [L*2^T, 2], got {tuple(table.shape)}")
        rows = table.shape[0] // num_levels
        log2 = rows.bit_length() - 1
        if (1 << log2) != rows or rows * num_levels != table.shape[0]:
            raise ValueError(f"hash_table rows {table.shape[0]} are not num_levels * 2^T")
        return _Grid(table, [float(x) for x in scal.detach().cpu().tolist()], num_levels, log2)

    @staticmethod
    def _lin(layer) -> _Linear:
        return _Linear(layer.weight, layer.bias)

    @classmethod
    def from_module(cls, model) -> "ModelTensors":
        f = model.field
        props = list(model.proposal_networks)
        if len(props) != L.TNF_NUM_PROP:
            raise ValueError(f"libtnf_b200 is built for {L.TNF_NUM_PROP} proposal networks, got {len(props)}")
        prop_grids, l0, l1 = [], [], []
        for p in props:
            prop_grids.append(cls._grid_of(p.encoding))
            mlp = p.mlp_base[1]
            if len(mlp.layers) != 2:
                raise ValueError("proposal MLP must have 2 layers")
            l0.append(cls._lin(mlp.layers[0]))
            l1.append(cls._lin(mlp.layers[1]))
        lin = {
            "base0": cls._lin(f.mlp_base.mlp.layers[0]),
            "base1": cls._lin(f.mlp_base.mlp.layers[1]),
            "rgb0": cls._lin(f.mlp_head.layers[0]),
            "rgb1": cls._lin(f.mlp_head.layers[1]),
            "rgb2": cls._lin(f.mlp_head.layers[2]),
            "th0": cls._lin(f.mlp_thermal.layers[0]),
            "th1": cls._lin(f.mlp_thermal.layers[1]),
            "th2": cls._lin(f.field_head_thermal.net),
        }
        expect = {"base0": (64, 32), "base1": (16, 64), "rgb0": (64, 63), "rgb1": (64, 64), "rgb2": (3, 64),
                  "th0": (64, 15), "th1": (64, 64), "th2": (1, 64)}
        for k, shp in expect.items():
            if tuple(lin[k].weight.shape) != shp:
                raise ValueError(f"{k}: weight shape {tuple(lin[k].weight.shape)} != {shp} (fixed architecture)")
        for i in range(len(props)):
            if tuple(l0[i].weight.shape) != (16, 2 * prop_grids[i].num_levels) or tuple(l1[i].weight.shape) != (1, 16):
                raise ValueError("proposal MLP must be grid -> 16 -> 1")
        return cls(prop_grids, l0, l1, cls._grid_of(f.mlp_base.encoder), lin, f.embedding_appearance.embedding.weight)

    # ------------------------------------------------------------------ packing
    @staticmethod
    def _fill_grid(dst: L.TnfHashGrid, g: _Grid, name: str) -> None:
        dst.table = _dev_f32(g.table, name + ".hash_table").data_ptr()
        for i, s in enumerate(g.scalings):
            dst.scalings[i] = s
        dst.num_levels = g.num_levels
        dst.log2_size = g.log2_size

    @staticmethod
    def _fill_lin(dst: L.TnfLinear, l: _Linear, name: str) -> None:
        dst.weight = _dev_f32(l.weight, name + ".weight").data_ptr()
        dst.bias = _dev_f32(l.bias, name + ".bias").data_ptr()

    def pack(self, *, num_samples: Sequence[int], training: bool, near_plane: float, far_plane: float,
             anneal: float, use_contraction: bool, aabb: Optional[Sequence[float]], appearance_mode: int,
             precision: int) -> L.TnfModel:
        m = L.TnfModel()
        for i in range(L.TNF_NUM_PROP):
            self._fill_grid(m.prop[i].grid, self.prop_grids[i], f"proposal_networks.{i}.encoding")
            self._fill_lin(m.prop[i].l0, self.prop_l0[i], f"proposal_networks.{i}.mlp.0")
            self._fill_lin(m.prop[i].l1, self.prop_l1[i], f"proposal_networks.{i}.mlp.1")
        self._fill_grid(m.field.grid, self.field_grid, "field.mlp_base.encoder")
        for k, l in self.field_linears.items():
            self._fill_lin(getattr(m.field, k), l, "field." + k)
        m.field.appearance = _dev_f32(self.appearance, "field.embedding_appearance").data_ptr()
        m.field.num_images = int(self.appearance.shape[0])
        if len(num_samples) != L.TNF_NUM_PROP + 1:
            raise ValueError(f"num_samples must have {L.TNF_NUM_PROP + 1} entries")
        for i, s in enumerate(num_samples):
            m.num_samples[i] = int(s)
        m.training = int(bool(training))
        m.near_plane, m.far_plane, m.anneal = float(near_plane), float(far_plane), float(anneal)
        m.use_contraction = int(bool(use_contraction))
        box = list(aabb) if aabb is not None else [-1.0, -1.0, -1.0, 1.0, 1.0, 1.0]
        for i in range(6):
            m.aabb[i] = float(box[i])
        m.appearance_mode = int(appearance_mode)
        m.precision = int(precision)
        return m


_OUT_KEYS = ("rgb", "thermal", "depth", "expected_depth", "accumulation", "prop_depth_0", "prop_depth_1")


def render_forward(
    tensors: ModelTensors,
    origins: Tensor,
    directions: Tensor,
    camera_indices: Optional[Tensor] = None,
    nears: Optional[Tensor] = None,
    fars: Optional[Tensor] = None,
    jitter: Optional[Tensor] = None,
    *,
    num_samples: Sequence[int] = (256, 96, 48),
    training: bool = False,
    near_plane: float = 0.05,
    far_plane: float = 1000.0,
    anneal: float = 1.0,
    use_contraction: bool = True,
    aabb: Optional[Sequence[float]] = None,
    appearance_mode: int = L.APPEARANCE_MEAN,
    precision: int = L.PRECISION_TC_FP16,
    depth_clip_chunk: int = 0,
    return_samples: bool = False,
    out: Optional[Dict[str, Tensor]] = None,
) -> Dict[str, object]:
    """One call of ``tnf_render_forward`` over R rays (flat).  Returns the output dict of
    ThermalNerfModel.get_outputs (thermal_nerf_model.py:245-275): rgb [R,3], thermal,
    depth, expected_depth, accumulation, prop_depth_0/1 [R,1]; with ``return_samples``
    also ``weights_list`` ([R,S_k,1]) and ``sdist_list`` ([R,S_k+1])."""
    lib = L.load()
    o = _dev_f32(origins, "origins")
    d = _dev_f32(directions, "directions")
    if o.dim() != 2 or o.shape[1] != 3 or d.shape != o.shape:
        raise ValueError(f"origins/directions must both be [R,3], got {tuple(o.shape)} / {tuple(d.shape)}")
    R = int(o.shape[0])
    dev = o.device
    rays = L.TnfRays()
    rays.origins, rays.directions, rays.num_rays = o.data_ptr(), d.data_ptr(), R
    keep = [o, d]
    if camera_indices is not None:
        ci = camera_indices.reshape(-1)
        if ci.dtype != torch.int64 or not ci.is_cuda or ci.numel() != R:
            raise ValueError("camera_indices must be a CUDA int64 tensor with R elements")
        ci = ci.contiguous()
        rays.camera_indices = ci.data_ptr()
        keep.append(ci)
    for name, t in (("nears", nears), ("fars", fars)):
        if t is not None:
            t = _dev_f32(t.reshape(-1), name)
            if t.numel() != R:
                raise ValueError(f"{name} must have R elements")
            setattr(rays, name, t.data_ptr())
            keep.append(t)
    if jitter is not None:
        j = _dev_f32(jitter.reshape(L.TNF_NUM_PROP + 1, -1), "jitter")
        if j.shape[1] != R:
            raise ValueError("jitter must be [3, R] (or [3, R, 1])")
        rays.jitter = j.data_ptr()
        keep.append(j)

    model = tensors.pack(num_samples=num_samples, training=training, near_plane=near_plane, far_plane=far_plane,
                         anneal=anneal, use_contraction=use_contraction, aabb=aabb,
                         appearance_mode=appearance_mode, precision=precision)

    res: Dict[str, object] = {}
    outs = L.TnfOutputs()
    if out is None:
        # one allocation for the seven per-ray outputs: rgb(3) + 6 scalars
        buf = torch.empty((9, max(R, 1)), dtype=torch.float32, device=dev)
        res["rgb"] = buf[0:3].view(-1)[: 3 * R].view(R, 3)
        for i, k in enumerate(_OUT_KEYS[1:]):
            res[k] = buf[3 + i, :R].view(R, 1)
    else:
        for k in _OUT_KEYS:
            res[k] = _dev_f32(out[k], "out." + k)
    outs.rgb = res["rgb"].data_ptr()
    outs.thermal = res["thermal"].data_ptr()
    outs.depth = res["depth"].data_ptr()
    outs.expected_depth = res["expected_depth"].data_ptr()
    outs.accumulation = res["accumulation"].data_ptr()
    outs.prop_depth[0] = res["prop_depth_0"].data_ptr()
    outs.prop_depth[1] = res["prop_depth_1"].data_ptr()
    if return_samples:
        wl, sl = [], []
        for k, s in enumerate(num_samples):
            w = torch.empty((R, int(s), 1), dtype=torch.float32, device=dev)
            sd = torch.empty((R, int(s) + 1), dtype=torch.float32, device=dev)
            outs.weights[k], outs.sdist[k] = w.data_ptr(), sd.data_ptr()
            wl.append(w)
            sl.append(sd)
        res["weights_list"], res["sdist_list"] = wl, sl

    chunk = int(depth_clip_chunk)
    ws_bytes = int(lib.tnf_forward_workspace_bytes(R, chunk))
    ws = torch.empty((ws_bytes + 3) // 4, dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream(dev).cuda_stream
    with torch.cuda.device(dev):
        rc = lib.tnf_render_forward(C.byref(model), C.byref(rays), C.byref(outs), chunk, ws.data_ptr(), ws_bytes,
                                    C.c_void_p(stream))
    L.check(rc)
    del keep  # inputs stay alive until the launch is enqueued; stream order protects them afterwards
    return res
