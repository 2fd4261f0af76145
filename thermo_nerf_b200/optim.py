"""Fused Adam over libtnf_b200 (``tnf_adam_step``): the optimisers of
thermo_nerf/thermal_nerf/config_thermal_nerf.py:32-45 (Adam lr=1e-2 eps=1e-15 for the
``proposal_networks`` and ``fields`` groups) as ONE launch per step instead of one
multi-tensor chain per group.

Same update rule and state layout as ``torch.optim.Adam`` (amsgrad=False, weight_decay=0):
``state[p] = {"step", "exp_avg", "exp_avg_sq"}``, so optimiser state saved by either loads into
the other.  Implements the ``_step_supports_amp_scaling`` protocol of ``torch.amp.GradScaler``
(the reference trains with ``mixed_precision=True``, config_thermal_nerf.py:22): the unscale and
the found-inf skip happen inside the kernel, without a host sync.
"""

from __future__ import annotations

from collections import defaultdict
from typing import Dict, List, Tuple

import torch
from torch import Tensor

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

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        grad_scale = getattr(self, "grad_scale", None)
        found_inf = getattr(self, "found_inf", None)
        for group in self.param_groups:
            beta1, beta2 = group["betas"]
            # tensors whose step counters agree go into one launch
            buckets: Dict[int, List[Tensor]] = defaultdict(list)
            for p in group["params"]:
                if p.grad is None:
                    continue
                if p.grad.is_sparse:
                    raise RuntimeError("FusedAdam does not support sparse gradients")
                st = self.state[p]
                if len(st) == 0:
                    st["step"] = torch.tensor(0.0)  # host-side counter, as torch.optim.Adam (capturable=False)
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["step"] += 1
                buckets[int(st["step"].item())].append(p)
            for step, ps in buckets.items():
                F.adam_step(ps, [p.grad for p in ps], [self.state[p]["exp_avg"] for p in ps],
                            [self.state[p]["exp_avg_sq"] for p in ps], [group["lr"]] * len(ps), step=step,
                            beta1=beta1, beta2=beta2, eps=group["eps"], grad_scale=grad_scale, found_inf=found_inf,
                            zero_grads=group["zero_grads"])
        return loss
