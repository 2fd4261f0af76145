"""Fused Adam over libtnf_b200 (``tnf_adam_step``): the optimisers of
thermo_nerf/thermal_nerf/config_thermal_nerf.py:32-45 (Adam lr=1e-2 eps=1e-15 for the
``proposal_networks`` and ``fields`` groups) as ONE launch per step instead of one
multi-tensor chain per group.

Same update rule and state layout as ``torch.optim.Adam`` (amsgrad=False, weight_decay=0):
``state[p] = {"step", "exp_avg", "exp_avg_sq"}``, so optimiser state saved by either loads into
the other.  Implements the ``_step_supports_amp_scaling`` protocol of ``torch.amp.GradScaler``
(the reference trains with ``mixed_precision=True``, config_thermal_nerf.py:22): the unscale and
the found-inf skip happen inside the kernel, without a host sync.

The host side of a step is a dictionary lookup: the argument block of the launch is cached per set
of (parameter, gradient) addresses, and the per-parameter step counters are Python ints that are
written back into ``state[p]["step"]`` tensors only when the state is exported.
"""

from __future__ import annotations

import ctypes as C
from typing import Dict, List, Tuple

import torch
from torch import Tensor

from . import _lib as L
from . import functional as F


class FusedAdam(torch.optim.Optimizer):
    _step_supports_amp_scaling = True

    def __init__(self, params, lr: float = 1e-3, betas: Tuple[float, float] = (0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 0.0, zero_grads: bool = False) -> None:
        if weight_decay != 0.0:
            raise ValueError("FusedAdam implements weight_decay=0 (what the reference configures)")
        if lr < 0 or eps < 0 or not (0 <= betas[0] < 1 and 0 <= betas[1] < 1):
            raise ValueError("invalid Adam hyper-parameters")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, zero_grads=zero_grads))
        self._steps: Dict[Tensor, int] = {}   # host step counters (authoritative between exports)
        self._plans: Dict[tuple, tuple] = {}  # (param/grad addresses) -> prepared TnfAdamTensor array

    # ---- state export / import keep torch.optim.Adam's layout ---------------------------------
    def _export_steps(self) -> None:
        for p, n in self._steps.items():
            st = self.state.get(p)
            if st is not None and "step" in st:
                st["step"] = torch.tensor(float(n))

    def state_dict(self):
        self._export_steps()
        return super().state_dict()

    def load_state_dict(self, state_dict) -> None:
        super().load_state_dict(state_dict)
        self._steps = {p: int(float(st["step"])) for p, st in self.state.items() if "step" in st}
        self._plans.clear()

    # ---- one launch per (group, step count) ---------------------------------------------------
    def _launch(self, ps: List[Tensor], lr: float, step: int, beta1: float, beta2: float, eps: float,
                zero_grads: bool, grad_scale, found_inf) -> None:
        key = tuple(p.data_ptr() for p in ps) + tuple(p.grad.data_ptr() for p in ps)
        plan = self._plans.get(key)
        if plan is None:
            if len(self._plans) > 64:
                self._plans.clear()
            chunks = []
            for s0 in range(0, len(ps), L.TNF_ADAM_MAX_TENSORS):
                sub = ps[s0:s0 + L.TNF_ADAM_MAX_TENSORS]
                arr = (L.TnfAdamTensor * len(sub))()
                for j, p in enumerate(sub):
                    st = self.state[p]
                    for t, nm in ((p, "param"), (p.grad, "grad"), (st["exp_avg"], "exp_avg"),
                                  (st["exp_avg_sq"], "exp_avg_sq")):
                        F._dev_f32(t, nm)
                        if t.numel() != p.numel():
                            raise ValueError(f"{nm} numel {t.numel()} != param numel {p.numel()}")
                    arr[j].param, arr[j].grad = p.data_ptr(), p.grad.data_ptr()
                    arr[j].exp_avg, arr[j].exp_avg_sq = st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr()
                    arr[j].numel, arr[j].lr = p.numel(), float(lr)
                chunks.append((arr, len(sub)))
            plan = (chunks, [float(lr)])
            self._plans[key] = plan
        chunks, cur_lr = plan
        if cur_lr[0] != lr:
            for arr, m in chunks:
                for j in range(m):
                    arr[j].lr = lr
            cur_lr[0] = float(lr)
        lib = L.load()
        dev = ps[0].device
        fi = 0 if found_inf is None else F._dev_f32(found_inf.reshape(-1), "found_inf").data_ptr()
        gs = 0 if grad_scale is None else F._dev_f32(grad_scale.reshape(-1), "grad_scale").data_ptr()
        stream = torch.cuda.current_stream(dev).cuda_stream
        with torch.cuda.device(dev):
            for arr, m in chunks:
                L.check(lib.tnf_adam_step(arr, m, float(beta1), float(beta2), float(eps), int(step), 1.0,
                                          C.c_void_p(gs), C.c_void_p(fi), int(bool(zero_grads)), C.c_void_p(stream)))

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        grad_scale = getattr(self, "grad_scale", None)
        found_inf = getattr(self, "found_inf", None)
        steps = self._steps
        for group in self.param_groups:
            beta1, beta2 = group["betas"]
            # tensors whose step counters agree go into one launch
            buckets: Dict[int, List[Tensor]] = {}
            for p in group["params"]:
                g = p.grad
                if g is None:
                    continue
                n = steps.get(p)
                if n is None:
                    if g.is_sparse:
                        raise RuntimeError("FusedAdam does not support sparse gradients")
                    st = self.state[p]
                    if len(st) == 0:
                        st["step"] = torch.tensor(0.0)  # host-side counter, as torch.optim.Adam (capturable=False)
                        st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                        st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    n = int(float(st["step"]))
                n += 1
                steps[p] = n
                b = buckets.get(n)
                if b is None:
                    buckets[n] = [p]
                else:
                    b.append(p)
            for step, ps in buckets.items():
                self._launch(ps, group["lr"], step, beta1, beta2, group["eps"], group["zero_grads"], grad_scale,
                             found_inf)
        return loss
