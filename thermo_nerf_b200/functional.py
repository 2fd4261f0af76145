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
import os
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
        }
        if hasattr(f, "mlp_thermal"):
            lin.update(th0=cls._lin(f.mlp_thermal.layers[0]), th1=cls._lin(f.mlp_thermal.layers[1]),
                       th2=cls._lin(f.field_head_thermal.net))
        else:
            # a stock nerfstudio NerfactoField (nerfacto / thermal-nerfacto model types): the kernels always
            # evaluate the thermal head, so it gets constant zeros that are neither parameters nor state
            ref = lin["rgb0"].weight
            z = lambda *shape: torch.zeros(shape, dtype=torch.float32, device=ref.device)  # noqa: E731
            lin.update(th0=_Linear(z(64, 15), z(64)), th1=_Linear(z(64, 64), z(64)), th2=_Linear(z(1, 64), z(1)))
        expect = {"base0": (64, 32), "base1": (16, 64), "rgb0": (64, 63), "rgb1": (64, 64), "rgb2": (3, 64),
                  "th0": (64, 15), "th1": (64, 64), "th2": (1, 64)}
        for k, shp in expect.items():
            got = tuple(lin[k].weight.shape)
            if k == "rgb2" and got == (4, 64):
                continue  # the RGBT colour head of the concat_nerf model type (rgb_concat/concat_field.py:65-75)
            if got != shp:
                raise ValueError(f"{k}: weight shape {got} != {shp} (fixed architecture)")
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
             precision: int, detach_thermal_geo: bool = False, head_mode: int = L.HEAD_THERMAL) -> L.TnfModel:
        """Fresh ``TnfModel`` for one call.  The pointer part is validated and filled once per set of
        tensor addresses and memcpy'd afterwards (this runs every training step)."""
        plist = self.param_list()
        key = tuple(t.data_ptr() for t in plist)
        cached = self.extra.get("_packed")
        if cached is None or cached[0] != key:
            base = L.TnfModel()
            for i in range(L.TNF_NUM_PROP):
                self._fill_grid(base.prop[i].grid, self.prop_grids[i], f"proposal_networks.{i}.encoding")
                self._fill_lin(base.prop[i].l0, self.prop_l0[i], f"proposal_networks.{i}.mlp.0")
                self._fill_lin(base.prop[i].l1, self.prop_l1[i], f"proposal_networks.{i}.mlp.1")
            self._fill_grid(base.field.grid, self.field_grid, "field.mlp_base.encoder")
            for k, l in self.field_linears.items():
                self._fill_lin(getattr(base.field, k), l, "field." + k)
            base.field.appearance = _dev_f32(self.appearance, "field.embedding_appearance").data_ptr()
            base.field.num_images = int(self.appearance.shape[0])
            cached = (key, base)
            self.extra["_packed"] = cached
        m = L.TnfModel.from_buffer_copy(cached[1])
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
        m.detach_thermal_geo = int(bool(detach_thermal_geo))
        m.head_mode = int(head_mode)
        nout = int(self.field_linears["rgb2"].weight.shape[0])
        if nout != (4 if m.head_mode == L.HEAD_CONCAT else 3):
            raise ValueError(f"head_mode={m.head_mode} does not match a colour head with {nout} outputs")
        return m

    # ------------------------------------------------------------------ autograd plumbing
    FIELD_ORDER = ("base0", "base1", "rgb0", "rgb1", "rgb2", "th0", "th1", "th2")

    def param_list(self) -> List[Tensor]:
        """Every parameter of the path in a fixed order: per proposal net (table, l0.w, l0.b, l1.w, l1.b),
        then the field (table, 8 x (w, b) in FIELD_ORDER, appearance embedding) - 28 tensors."""
        out = self.extra.get("_plist")
        if out is not None:
            return out
        out = []
        for i in range(L.TNF_NUM_PROP):
            out += [self.prop_grids[i].table, self.prop_l0[i].weight, self.prop_l0[i].bias,
                    self.prop_l1[i].weight, self.prop_l1[i].bias]
        out.append(self.field_grid.table)
        for k in self.FIELD_ORDER:
            out += [self.field_linears[k].weight, self.field_linears[k].bias]
        out.append(self.appearance)
        self.extra["_plist"] = out
        return out

    def with_params(self, params: Sequence[Tensor]) -> "ModelTensors":
        """Same constants, tensors replaced by ``params`` (in ``param_list`` order)."""
        it = iter(params)
        grids, l0, l1 = [], [], []
        for i in range(L.TNF_NUM_PROP):
            g = self.prop_grids[i]
            grids.append(_Grid(next(it), g.scalings, g.num_levels, g.log2_size))
            l0.append(_Linear(next(it), next(it)))
            l1.append(_Linear(next(it), next(it)))
        fg = _Grid(next(it), self.field_grid.scalings, self.field_grid.num_levels, self.field_grid.log2_size)
        lin = {k: _Linear(next(it), next(it)) for k in self.FIELD_ORDER}
        extra = {k: v for k, v in self.extra.items() if not k.startswith("_")}
        return ModelTensors(grids, l0, l1, fg, lin, next(it), extra)

    @staticmethod
    def pack_grads(grads: Sequence[Optional[Tensor]], ray_grads: Optional[Tuple[Tensor, Tensor]] = None
                   ) -> L.TnfModelGrad:
        """``grads`` in ``param_list`` order (None = not wanted; a proposal net is skipped as a whole
        when its table gradient is None); ``ray_grads`` = (d origins, d directions) [R,3] buffers or None."""
        g = L.TnfModelGrad()
        if ray_grads is not None:
            g.ray_origins = _dev_f32(ray_grads[0], "grad.origins").data_ptr()
            g.ray_directions = _dev_f32(ray_grads[1], "grad.directions").data_ptr()
        it = iter(grads)

        def ptr(t, name):
            return 0 if t is None else _dev_f32(t, name).data_ptr()

        for i in range(L.TNF_NUM_PROP):
            five = [next(it) for _ in range(5)]
            if five[0] is not None and any(t is None for t in five):
                raise ValueError(f"proposal net {i}: gradients must be requested for all five tensors or none")
            g.prop[i].table = ptr(five[0], "grad.table")
            g.prop[i].l0.weight, g.prop[i].l0.bias = ptr(five[1], "grad"), ptr(five[2], "grad")
            g.prop[i].l1.weight, g.prop[i].l1.bias = ptr(five[3], "grad"), ptr(five[4], "grad")
        g.field.table = ptr(next(it), "grad.field.table")
        for k in ModelTensors.FIELD_ORDER:
            lg = getattr(g.field, k)
            lg.weight, lg.bias = ptr(next(it), "grad." + k), ptr(next(it), "grad." + k)
        g.field.appearance = ptr(next(it), "grad.appearance")
        return g


_OUT_KEYS = ("rgb", "thermal", "depth", "expected_depth", "accumulation", "prop_depth_0", "prop_depth_1")


# TNF_EVAL_SPLIT=0: eval calls as one fused launch whatever their size
_EVAL_SPLIT = os.environ.get("TNF_EVAL_SPLIT", "1") != "0"
_EVAL_SPLIT_MIN_RAYS = 16384


def render_forward(
    tensors: ModelTensors,
    origins: Optional[Tensor],
    directions: Optional[Tensor],
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
    save_for_backward: bool = False,
    detach_thermal_geo: bool = False,
    head_mode: int = L.HEAD_THERMAL,
    camera: Optional[L.TnfCamera] = None,
    first_pixel: int = 0,
    num_pixels: Optional[int] = None,
    field_ready_event: Optional["torch.cuda.Event"] = None,
) -> Dict[str, object]:
    """One call of ``tnf_render_forward`` over R rays (flat).  Returns the output dict of
    ThermalNerfModel.get_outputs (thermal_nerf_model.py:245-275): rgb [R,3], thermal,
    depth, expected_depth, accumulation, prop_depth_0/1 [R,1]; with ``return_samples``
    also ``weights_list`` ([R,S_k,1]) and ``sdist_list`` ([R,S_k+1])."""
    lib = L.load()
    rays = L.TnfRays()
    if camera is not None:
        # rays generated inside the kernel from the camera (eval only): no [R,3] tensors cross HBM or PCIe
        if origins is not None or directions is not None:
            raise ValueError("pass either origins/directions or camera, not both")
        if training:
            raise ValueError("camera rays are an eval-mode input")
        total = int(camera.width) * int(camera.height)
        R = total - int(first_pixel) if num_pixels is None else int(num_pixels)
        if first_pixel < 0 or R < 0 or first_pixel + R > total:
            raise ValueError(f"pixels [{first_pixel}, {first_pixel + R}) outside the {camera.width}x{camera.height} image")
        dev = tensors.field_grid.table.device
        rays.from_camera, rays.first_pixel, rays.camera, rays.num_rays = 1, int(first_pixel), camera, R
        keep = []
    else:
        o = _dev_f32(origins, "origins")
        d = _dev_f32(directions, "directions")
        if o.dim() != 2 or o.shape[1] != 3 or d.shape != o.shape:
            raise ValueError(f"origins/directions must both be [R,3], got {tuple(o.shape)} / {tuple(d.shape)}")
        R = int(o.shape[0])
        dev = o.device
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
                         appearance_mode=appearance_mode, precision=precision,
                         detach_thermal_geo=detach_thermal_geo, head_mode=head_mode)

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
    elif precision == L.PRECISION_TC_FP16 and _EVAL_SPLIT and R >= _EVAL_SPLIT_MIN_RAYS:
        # two-launch form for large eval calls as well (proposal levels at twice the occupancy, then the field level):
        # the 49 spacing bins of the last level cross HBM once (196 B per ray) in a scratch buffer of this call
        scratch = torch.empty((R, int(num_samples[-1]) + 1), dtype=torch.float32, device=dev)
        outs.sdist[L.TNF_NUM_PROP] = scratch.data_ptr()
        keep.append(scratch)
    if save_for_backward:
        if not (return_samples and training):
            raise ValueError("save_for_backward needs training=True and return_samples=True")
        ns = R * int(num_samples[-1])
        fdt = torch.float32 if precision == L.PRECISION_FP32 else torch.float16
        res["field_features"] = torch.empty((max(ns, 1), 32), dtype=fdt, device=dev)
        res["field_samples"] = torch.empty((max(ns, 1), 5), dtype=torch.float32, device=dev)
        outs.field_features = res["field_features"].data_ptr()
        outs.field_samples = res["field_samples"].data_ptr()
        res["_model_struct"] = model

    chunk = int(depth_clip_chunk)
    ws_bytes = int(lib.tnf_forward_workspace_bytes(R, chunk))
    ws = torch.empty((ws_bytes + 3) // 4, dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream(dev).cuda_stream
    with torch.cuda.device(dev):
        if field_ready_event is not None:
            # tnf_render_forward_staged: the field level waits for the event (a parameter exchange of the field
            # network still in flight on another stream); the proposal levels start right away
            rc = lib.tnf_render_forward_staged(C.byref(model), C.byref(rays), C.byref(outs), chunk, ws.data_ptr(),
                                               ws_bytes, C.c_void_p(stream), C.c_void_p(field_ready_event.cuda_event))
        else:
            rc = lib.tnf_render_forward(C.byref(model), C.byref(rays), C.byref(outs), chunk, ws.data_ptr(), ws_bytes,
                                        C.c_void_p(stream))
    L.check(rc)
    del keep  # inputs stay alive until the launch is enqueued; stream order protects them afterwards
    return res


# ---------------------------------------------------------------------------------------------
# training: autograd node over tnf_render_forward / tnf_render_backward
# ---------------------------------------------------------------------------------------------
def _ptr(t: Optional[Tensor]) -> int:
    return 0 if t is None else t.data_ptr()


_WORKSPACES: Dict[torch.device, Tensor] = {}


def _workspace_for(model_struct: L.TnfModel, R: int, dev: torch.device) -> Tensor:
    """Per-device scratch of tnf_render_backward, kept across steps (it is 66 KB per ray: allocating it per
    call would churn the caching allocator with a quarter-gigabyte block every iteration).  Stream order
    makes reuse safe: every user is enqueued on the current stream."""
    nbytes = int(L.load().tnf_backward_workspace_bytes(C.byref(model_struct), R))
    ws = _WORKSPACES.get(dev)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        _WORKSPACES[dev] = ws
    return ws


def render_backward(tensors: ModelTensors, model_struct: L.TnfModel, origins: Tensor, directions: Tensor,
                    camera_indices: Optional[Tensor], nears: Optional[Tensor], fars: Optional[Tensor],
                    jitter: Optional[Tensor], saved: Dict[str, object], grad_outputs: Dict[str, Optional[Tensor]],
                    grads: Sequence[Optional[Tensor]], ray_grads: Optional[Tuple[Tensor, Tensor]] = None,
                    field_grads_event: Optional["torch.cuda.Event"] = None, reserve_ctas: int = 0) -> None:
    """``tnf_render_backward``: accumulates into ``grads`` (``ModelTensors.param_list`` order) and, when
    ``ray_grads`` = (d origins, d directions) is given, into those [R,3] buffers (camera-optimiser path).
    ``field_grads_event`` (an event that has been recorded once, so that its handle exists): the field level runs
    first and the event is re-recorded behind it, before the proposal levels, whose kernel leaves ``reserve_ctas``
    CTA slots free (``tnf_render_backward_staged``)."""
    lib = L.load()
    R = int(origins.shape[0])
    dev = origins.device
    rays = L.TnfRays()
    rays.origins, rays.directions, rays.num_rays = origins.data_ptr(), directions.data_ptr(), R
    rays.camera_indices, rays.nears, rays.fars, rays.jitter = (_ptr(camera_indices), _ptr(nears), _ptr(fars),
                                                               _ptr(jitter))
    sv = L.TnfSaved()
    for k in range(L.TNF_NUM_PROP + 1):
        sv.sdist[k] = saved["sdist_list"][k].data_ptr()
        sv.weights[k] = saved["weights_list"][k].data_ptr()
    sv.field_features = saved["field_features"].data_ptr()
    sv.field_samples = saved["field_samples"].data_ptr()
    go = L.TnfOutputGrads()
    keep = []

    def gptr(t, n):
        if t is None:
            return 0
        t = _dev_f32(t.contiguous(), n)
        keep.append(t)
        return t.data_ptr()

    go.rgb = gptr(grad_outputs.get("rgb"), "grad rgb")
    go.thermal = gptr(grad_outputs.get("thermal"), "grad thermal")
    go.accumulation = gptr(grad_outputs.get("accumulation"), "grad accumulation")
    gw = grad_outputs.get("weights_list") or [None] * (L.TNF_NUM_PROP + 1)
    for k in range(L.TNF_NUM_PROP + 1):
        go.weights[k] = gptr(gw[k], f"grad weights[{k}]")
    gstruct = ModelTensors.pack_grads(grads, ray_grads)
    nbytes = int(lib.tnf_backward_workspace_bytes(C.byref(model_struct), R))
    ws = saved.get("_workspace")
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream(dev).cuda_stream
    with torch.cuda.device(dev):
        if field_grads_event is not None:
            rc = lib.tnf_render_backward_staged(C.byref(model_struct), C.byref(rays), C.byref(sv), C.byref(go),
                                                C.byref(gstruct), ws.data_ptr(), ws.numel(), C.c_void_p(stream),
                                                C.c_void_p(field_grads_event.cuda_event), int(reserve_ctas))
        else:
            rc = lib.tnf_render_backward(C.byref(model_struct), C.byref(rays), C.byref(sv), C.byref(go),
                                         C.byref(gstruct), ws.data_ptr(), ws.numel(), C.c_void_p(stream))
    L.check(rc)
    del keep


class _RenderFn(torch.autograd.Function):
    """get_outputs in training mode.  Differentiable outputs: rgb, thermal, accumulation and the three
    ``weights_list`` entries; depth-type outputs and the spacing bins are constants (as in the
    reference: median depth is computed under no_grad, PDFSampler detaches its bins)."""

    NUM_FIXED = 9  # inputs before *params

    @staticmethod
    def forward(ctx, tensors: ModelTensors, kw: dict, prop_grad: bool, origins, directions, camera_indices, nears,
                fars, jitter, *params):
        own = tensors.param_list()
        same = len(own) == len(params) and all(a is b for a, b in zip(own, params))
        t = tensors if same else tensors.with_params([p.detach() for p in params])
        res = render_forward(t, origins, directions, camera_indices, nears, fars, jitter, training=True,
                             return_samples=True, save_for_backward=True, **kw)
        ctx.tensors = t
        ctx.model_struct = res.pop("_model_struct")
        ctx.prop_grad = bool(prop_grad)
        ctx.set_materialize_grads(False)
        w = res["weights_list"]
        sd = res["sdist_list"]
        # save_for_backward (not attributes): weights_list / sdist_list are OUTPUTS of this node, and an
        # output kept on ctx directly forms a reference cycle (tensor -> grad_fn -> ctx -> tensor) that
        # only the cyclic GC frees - i.e. every step would leak its saved activations until a GC pass.
        ctx.save_for_backward(origins, directions, camera_indices, nears, fars, jitter, *sd, *w,
                              res["field_features"], res["field_samples"])
        nondiff = (res["depth"], res["expected_depth"], res["prop_depth_0"], res["prop_depth_1"], *sd)
        ctx.mark_non_differentiable(*nondiff)
        return (res["rgb"], res["thermal"], res["accumulation"], w[0], w[1], w[2], *nondiff)

    @staticmethod
    def backward(ctx, g_rgb, g_th, g_acc, g_w0, g_w1, g_w2, *_):
        params = ctx.tensors.param_list()
        need = list(ctx.needs_input_grad[_RenderFn.NUM_FIXED:])
        # a proposal net takes part only if its upstream weights carry a gradient (the reference's
        # no_grad steps) and its tensors want one
        gws = [g_w0, g_w1]
        on = [ctx.prop_grad and gws[i] is not None and all(need[5 * i:5 * i + 5]) for i in range(L.TNF_NUM_PROP)]
        # one zero-filled arena (a single memset), carved into per-parameter gradients (16-byte aligned)
        want = [on[j // 5] if j < 5 * L.TNF_NUM_PROP else True for j in range(len(params))]
        offs, total = [], 0
        for p_, w_ in zip(params, want):
            offs.append(total)
            if w_:
                total += (p_.numel() + 3) // 4 * 4
        arena = torch.zeros(total, dtype=torch.float32, device=params[-1].device)
        grads: List[Optional[Tensor]] = [arena[o_:o_ + p_.numel()].view(p_.shape) if w_ else None
                                         for p_, w_, o_ in zip(params, want, offs)]
        sv = ctx.saved_tensors
        o, d, cam, nears, fars, jitter = sv[:6]
        n = L.TNF_NUM_PROP + 1
        saved = {"sdist_list": list(sv[6:6 + n]), "weights_list": list(sv[6 + n:6 + 2 * n]),
                 "field_features": sv[6 + 2 * n], "field_samples": sv[7 + 2 * n],
                 "_workspace": _workspace_for(ctx.model_struct, int(o.shape[0]), o.device)}

        def flat(g):
            return None if g is None else g.reshape(g.shape[0], -1)

        # pose path: the camera optimiser made origins / directions part of the graph (thermal_nerf_model.py:218-219)
        ray_grads = None
        if ctx.needs_input_grad[3] or ctx.needs_input_grad[4]:
            ray_grads = (torch.zeros_like(o), torch.zeros_like(d))
        render_backward(ctx.tensors, ctx.model_struct, o, d, cam, nears, fars, jitter, saved,
                        {"rgb": g_rgb, "thermal": None if g_th is None else g_th.reshape(-1),
                         "accumulation": None if g_acc is None else g_acc.reshape(-1),
                         "weights_list": [flat(g_w0) if ctx.prop_grad else None,
                                          flat(g_w1) if ctx.prop_grad else None, flat(g_w2)]},
                        grads, ray_grads)
        out = [g if n_ else None for g, n_ in zip(grads, need)]
        fixed = [None] * _RenderFn.NUM_FIXED
        if ray_grads is not None:
            fixed[3] = ray_grads[0] if ctx.needs_input_grad[3] else None
            fixed[4] = ray_grads[1] if ctx.needs_input_grad[4] else None
        return tuple(fixed) + tuple(out)


def render(tensors: ModelTensors, origins: Tensor, directions: Tensor, camera_indices: Optional[Tensor] = None,
           nears: Optional[Tensor] = None, fars: Optional[Tensor] = None, jitter: Optional[Tensor] = None,
           *, prop_grad: bool = True, **kw) -> Dict[str, object]:
    """Training-mode ``get_outputs`` with autograd: the dict of :func:`render_forward` (with
    ``weights_list`` / ``sdist_list``), differentiable w.r.t. every parameter of ``tensors``.
    ``prop_grad=False`` reproduces the reference's no_grad proposal steps
    (ProposalNetworkSampler ``updated == False``)."""
    o = _dev_f32(origins, "origins")
    d = _dev_f32(directions, "directions")
    R = int(o.shape[0])
    cam = camera_indices.reshape(-1).contiguous() if camera_indices is not None else None
    nears = _dev_f32(nears.reshape(-1), "nears") if nears is not None else None
    fars = _dev_f32(fars.reshape(-1), "fars") if fars is not None else None
    if jitter is not None:
        jitter = _dev_f32(jitter.reshape(L.TNF_NUM_PROP + 1, -1), "jitter")
    for k in ("training", "return_samples", "save_for_backward", "out"):
        if k in kw:
            raise TypeError(f"render() fixes {k!r}; use render_forward() for the non-differentiable call")
    outs = _RenderFn.apply(tensors, dict(kw), bool(prop_grad), o, d, cam, nears, fars, jitter,
                           *tensors.param_list())
    rgb, th, acc, w0, w1, w2, depth, ed, pd0, pd1, sd0, sd1, sd2 = outs
    if not prop_grad:
        w0, w1 = w0.detach(), w1.detach()
    return {"rgb": rgb, "thermal": th.view(R, 1), "accumulation": acc.view(R, 1), "depth": depth,
            "expected_depth": ed, "prop_depth_0": pd0, "prop_depth_1": pd1,
            "weights_list": [w0, w1, w2], "sdist_list": [sd0, sd1, sd2]}


# ---------------------------------------------------------------------------------------------
# losses (get_loss_dict) and Adam
# ---------------------------------------------------------------------------------------------
LOSS_NAMES = ("rgb_loss", "interlevel_loss", "distortion_loss", "thermal")


def losses_forward_backward(weights_list: Sequence[Tensor], sdist_list: Sequence[Tensor], rgb: Tensor,
                            thermal: Tensor, gt_rgb: Tensor, gt_thermal: Tensor, *, interlevel_mult: float = 1.0,
                            distortion_mult: float = 0.002, use_rgb_loss: bool = True,
                            use_thermal_loss: bool = True, grad_scale: float = 1.0, want_grads: bool = True,
                            prop_grad: bool = True, concat_accumulation: Optional[Tensor] = None,
                            concat_noise: Optional[Tensor] = None):
    """``tnf_losses``: returns (losses[4] device tensor in LOSS_NAMES order, grads dict or None).
    grads: d(loss)/d(input) * grad_scale for rgb [R,3], thermal [R], weights_list[k] [R,S_k].
    With ``concat_accumulation`` [R] and ``concat_noise`` [R,4] the colour term is the concat_nerf loss
    (rgb_concat/concat_nerfacto_model.py:197-233): MSE over the 4 RGBT channels of pred + noise (1 - accumulation);
    grads then also carry "accumulation" [R]."""
    lib = L.load()
    R = int(rgb.shape[0])
    dev = rgb.device
    a = L.TnfLossArgs()
    keep = []

    def f32(t, n):
        t = _dev_f32(t.detach().contiguous(), n)
        keep.append(t)
        return t

    for k in range(L.TNF_NUM_PROP + 1):
        w = f32(weights_list[k].reshape(R, -1), f"weights[{k}]")
        sd = f32(sdist_list[k], f"sdist[{k}]")
        if sd.shape != (R, w.shape[1] + 1):
            raise ValueError(f"sdist[{k}] must be [R, S+1], got {tuple(sd.shape)} for S={w.shape[1]}")
        a.weights[k], a.sdist[k], a.num_samples[k] = w.data_ptr(), sd.data_ptr(), int(w.shape[1])
    a.rgb = f32(rgb.reshape(R, 3), "rgb").data_ptr()
    a.thermal = f32(thermal.reshape(R), "thermal").data_ptr()
    a.gt_rgb = f32(gt_rgb.reshape(R, 3), "gt_rgb").data_ptr()
    a.gt_thermal = f32(gt_thermal.reshape(R), "gt_thermal").data_ptr()
    a.num_rays = R
    a.interlevel_mult, a.distortion_mult = float(interlevel_mult), float(distortion_mult)
    a.use_rgb_loss, a.use_thermal_loss = int(bool(use_rgb_loss)), int(bool(use_thermal_loss))
    a.grad_scale = float(grad_scale)
    concat = concat_accumulation is not None
    if concat:
        if concat_noise is None:
            raise ValueError("the concat loss needs the random-background draw (torch.rand_like(pred), [R,4])")
        a.concat = 1
        a.accumulation = f32(concat_accumulation.reshape(R), "accumulation").data_ptr()
        a.noise = f32(concat_noise.reshape(R, 4), "noise").data_ptr()
    losses = torch.empty(4, dtype=torch.float32, device=dev)
    a.losses = losses.data_ptr()
    grads = None
    if want_grads:
        grads = {"rgb": torch.empty((R, 3), dtype=torch.float32, device=dev),
                 "thermal": torch.empty((R,), dtype=torch.float32, device=dev), "weights_list": []}
        a.g_rgb, a.g_thermal = grads["rgb"].data_ptr(), grads["thermal"].data_ptr()
        if concat:
            grads["accumulation"] = torch.empty((R,), dtype=torch.float32, device=dev)
            a.g_accumulation = grads["accumulation"].data_ptr()
        for k in range(L.TNF_NUM_PROP + 1):
            if k < L.TNF_NUM_PROP and not prop_grad:
                grads["weights_list"].append(None)
                continue
            g = torch.empty((R, a.num_samples[k]), dtype=torch.float32, device=dev)
            a.g_weights[k] = g.data_ptr()
            grads["weights_list"].append(g)
    stream = torch.cuda.current_stream(dev).cuda_stream
    with torch.cuda.device(dev):
        rc = lib.tnf_losses(C.byref(a), C.c_void_p(stream))
    L.check(rc)
    del keep
    return losses, grads


class _LossFn(torch.autograd.Function):
    """The four losses as one node: forward evaluates losses and their gradients in one kernel,
    backward scales the stashed gradients by the upstream scalars."""

    @staticmethod
    def forward(ctx, cfg: dict, rgb, thermal, w0, w1, w2, sd0, sd1, sd2, gt_rgb, gt_thermal, accumulation=None,
                noise=None):
        prop_grad = bool(ctx.needs_input_grad[3] or ctx.needs_input_grad[4])
        losses, g = losses_forward_backward([w0, w1, w2], [sd0, sd1, sd2], rgb, thermal, gt_rgb, gt_thermal,
                                            interlevel_mult=cfg["interlevel_mult"],
                                            distortion_mult=cfg["distortion_mult"], use_rgb_loss=cfg["use_rgb_loss"],
                                            use_thermal_loss=cfg["use_thermal_loss"], prop_grad=prop_grad,
                                            concat_accumulation=accumulation, concat_noise=noise)
        ctx.g = g
        ctx.set_materialize_grads(False)
        ctx.shapes = (rgb.shape, thermal.shape, w0.shape, w1.shape, w2.shape,
                      None if accumulation is None else accumulation.shape)
        return losses[0], losses[1], losses[2], losses[3]

    @staticmethod
    def backward(ctx, u_rgb, u_inter, u_dist, u_th):
        g = ctx.g
        s = ctx.shapes
        gw = g["weights_list"]
        concat = s[5] is not None

        def sc(t, u, shape):
            return None if (t is None or u is None) else (t * u).view(shape)

        # concat_nerf: the temperature channel and the accumulation feed losses[0] (the 4-channel colour term)
        return (None, sc(g["rgb"], u_rgb, s[0]), sc(g["thermal"], u_rgb if concat else u_th, s[1]),
                sc(gw[0], u_inter, s[2]), sc(gw[1], u_inter, s[3]), sc(gw[2], u_dist, s[4]), None, None, None, None,
                None, sc(g.get("accumulation"), u_rgb, s[5]) if concat else None, None)


def losses(outputs: Dict[str, object], gt_rgb: Tensor, gt_thermal: Tensor, *, interlevel_mult: float = 1.0,
           distortion_mult: float = 0.002, use_rgb_loss: bool = True, use_thermal_loss: bool = True,
           concat_noise: Optional[Tensor] = None) -> Dict[str, Tensor]:
    """Differentiable loss dict of get_loss_dict (thermal_nerf_model.py:277-326) from the training outputs
    of :func:`render`.  ``concat_noise`` [R,4] selects the concat_nerf loss (one 4-channel colour term with the
    random background blended into the prediction, rgb_concat/concat_nerfacto_model.py:197-233; no "thermal" entry);
    ``outputs`` must then carry "accumulation"."""
    w, sd = outputs["weights_list"], outputs["sdist_list"]
    cfg = dict(interlevel_mult=interlevel_mult, distortion_mult=distortion_mult, use_rgb_loss=use_rgb_loss,
               use_thermal_loss=use_thermal_loss)
    extra = () if concat_noise is None else (outputs["accumulation"], concat_noise)
    vals = _LossFn.apply(cfg, outputs["rgb"], outputs["thermal"], w[0], w[1], w[2], sd[0], sd[1], sd[2], gt_rgb,
                         gt_thermal, *extra)
    out = dict(zip(LOSS_NAMES, vals))
    if concat_noise is not None:
        out.pop("thermal")
        return out
    if not use_rgb_loss:
        out.pop("rgb_loss")
    if not use_thermal_loss:
        out.pop("thermal")
    return out


def adam_step(params: Sequence[Tensor], grads: Sequence[Tensor], exp_avgs: Sequence[Tensor],
              exp_avg_sqs: Sequence[Tensor], lrs: Sequence[float], *, step: int, beta1: float = 0.9,
              beta2: float = 0.999, eps: float = 1e-8, inv_grad_scale: float = 1.0,
              grad_scale: Optional[Tensor] = None, found_inf: Optional[Tensor] = None,
              zero_grads: bool = False) -> None:
    """``tnf_adam_step`` over a list of fp32 CUDA tensors (one launch per 48 tensors)."""
    lib = L.load()
    n = len(params)
    if not (len(grads) == len(exp_avgs) == len(exp_avg_sqs) == len(lrs) == n):
        raise ValueError("params/grads/exp_avgs/exp_avg_sqs/lrs must have the same length")
    if n == 0:
        return
    dev = params[0].device
    stream = torch.cuda.current_stream(dev).cuda_stream
    fi = gs = 0
    if found_inf is not None:
        fi = _dev_f32(found_inf.reshape(-1), "found_inf").data_ptr()
    if grad_scale is not None:
        gs = _dev_f32(grad_scale.reshape(-1), "grad_scale").data_ptr()
    for s0 in range(0, n, L.TNF_ADAM_MAX_TENSORS):
        m = min(L.TNF_ADAM_MAX_TENSORS, n - s0)
        arr = (L.TnfAdamTensor * m)()
        for j in range(m):
            i = s0 + j
            p, g, ea, es = params[i], grads[i], exp_avgs[i], exp_avg_sqs[i]
            for t, nm in ((p, "param"), (g, "grad"), (ea, "exp_avg"), (es, "exp_avg_sq")):
                _dev_f32(t, nm)
                if t.numel() != p.numel():
                    raise ValueError(f"{nm} numel {t.numel()} != param numel {p.numel()}")
            arr[j].param, arr[j].grad = p.data_ptr(), g.data_ptr()
            arr[j].exp_avg, arr[j].exp_avg_sq = ea.data_ptr(), es.data_ptr()
            arr[j].numel, arr[j].lr = p.numel(), float(lrs[i])
        with torch.cuda.device(dev):
            rc = lib.tnf_adam_step(arr, m, float(beta1), float(beta2), float(eps), int(step), float(inv_grad_scale),
                                   C.c_void_p(gs), C.c_void_p(fi), int(bool(zero_grads)), C.c_void_p(stream))
        L.check(rc)


# ---------------------------------------------------------------------------------------------
# either side of the path: ray generation and frame post-processing
# ---------------------------------------------------------------------------------------------
def pack_camera(c2w: Tensor, fx: float, fy: float, cx: float, cy: float, width: int, height: int) -> L.TnfCamera:
    """``TnfCamera`` of one perspective camera (``c2w`` = camera_to_worlds[i], [3,4]; host values)."""
    cam = L.TnfCamera()
    vals = [float(x) for x in torch.as_tensor(c2w, dtype=torch.float32).reshape(-1).tolist()]
    if len(vals) != 12:
        raise ValueError("c2w must be [3,4]")
    for i, v in enumerate(vals):
        cam.c2w[i] = v
    cam.fx, cam.fy, cam.cx, cam.cy = float(fx), float(fy), float(cx), float(cy)
    cam.width, cam.height = int(width), int(height)
    return cam


def generate_rays(camera: L.TnfCamera, device, first_pixel: int = 0, num_pixels: Optional[int] = None
                  ) -> Tuple[Tensor, Tensor, Tensor]:
    """``tnf_generate_rays``: (origins [n,3], directions [n,3], directions_norm [n,1]) of the pixels
    [first_pixel, first_pixel + n) of ``camera`` - Cameras.generate_rays (renderer.py:183, evaluator.py:69)."""
    lib = L.load()
    dev = torch.device(device)
    if dev.type != "cuda":
        raise RuntimeError("generate_rays runs on a CUDA device; there is no CPU path")
    n = int(camera.width) * int(camera.height) - int(first_pixel) if num_pixels is None else int(num_pixels)
    o = torch.empty((max(n, 0), 3), dtype=torch.float32, device=dev)
    d = torch.empty((max(n, 0), 3), dtype=torch.float32, device=dev)
    nrm = torch.empty((max(n, 0), 1), dtype=torch.float32, device=dev)
    stream = torch.cuda.current_stream(dev).cuda_stream
    with torch.cuda.device(dev):
        rc = lib.tnf_generate_rays(C.byref(camera), int(first_pixel), n, o.data_ptr(), d.data_ptr(), nrm.data_ptr(),
                                   C.c_void_p(stream))
    L.check(rc)
    return o, d, nrm


def postprocess_frame(rgb: Optional[Tensor] = None, scalar: Optional[Tensor] = None, lut8: Optional[Tensor] = None
                      ) -> Tuple[Optional[Tensor], Optional[Tensor]]:
    """``tnf_postprocess_frame``: the uint8 conversion of Renderer.render (renderer.py:189-199).
    ``rgb`` [...,3] float in [0,1] -> uint8 [...,3]; ``scalar`` [...,1] or [...] -> uint8 [...,3], through the
    uint8 colour table ``lut8`` [N,3] when given, else grey.  Returns (rgb8, scalar8)."""
    lib = L.load()
    ref = rgb if rgb is not None else scalar
    if ref is None:
        return None, None
    dev = ref.device
    rgb8 = scalar8 = None
    n = 0
    if rgb is not None:
        r = _dev_f32(rgb.reshape(-1, 3), "rgb")
        n = int(r.shape[0])
        rgb8 = torch.empty((*rgb.shape[:-1], 3), dtype=torch.uint8, device=dev)
    if scalar is not None:
        sc = _dev_f32(scalar.reshape(-1), "scalar")
        if rgb is not None and sc.numel() != n:
            raise ValueError("rgb and scalar must cover the same pixels")
        n = int(sc.numel())
        shape = scalar.shape[:-1] if (scalar.dim() > 1 and scalar.shape[-1] == 1) else scalar.shape
        scalar8 = torch.empty((*shape, 3), dtype=torch.uint8, device=dev)
    lut_ptr, lut_n = 0, 0
    if lut8 is not None:
        if lut8.dtype != torch.uint8 or lut8.dim() != 2 or lut8.shape[1] != 3 or not lut8.is_cuda:
            raise ValueError("lut8 must be a CUDA uint8 tensor [N,3]")
        lut8 = lut8.contiguous()
        lut_ptr, lut_n = lut8.data_ptr(), int(lut8.shape[0])
    stream = torch.cuda.current_stream(dev).cuda_stream
    with torch.cuda.device(dev):
        rc = lib.tnf_postprocess_frame(C.c_void_p(r.data_ptr() if rgb is not None else 0),
                                       C.c_void_p(sc.data_ptr() if scalar is not None else 0), n,
                                       C.c_void_p(lut_ptr), lut_n,
                                       C.c_void_p(rgb8.data_ptr() if rgb8 is not None else 0),
                                       C.c_void_p(scalar8.data_ptr() if scalar8 is not None else 0),
                                       C.c_void_p(stream))
    L.check(rc)
    return rgb8, scalar8
