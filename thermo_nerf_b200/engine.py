"""One ThermoNeRF training iteration on persistent device buffers, without autograd:

    forward (tnf_render_forward, training) -> losses + their gradients (tnf_losses)
    -> backward (tnf_render_backward, into one flat gradient arena)
    -> Adam (tnf_adam_step, zeroes the arena)

With several GPUs (``peer_fused=True``) the last two lines become the pipelined exchange of DESIGN.md section 6: the
field level of the backward first, its gradients reduced inside the NVSwitch, Adam on the shard this rank owns and the
result multicast to every rank (tnf_peer_adam_range) on a side stream, under the proposal backward and the next
iteration's proposal pass; the proposal networks' slice on the critical path.  ``peer_fused=False`` keeps the
baseline: one NCCL all-reduce(mean) of the arena, then tnf_adam_step.

It reproduces what nerfstudio's ``Trainer.train_iteration`` does around the reference model
(SURVEY 3.1): the proposal-weight anneal and the proposal update schedule
(thermal_nerf_model.py:152-161, ProposalNetworkSampler), single-jitter stratified sampling, the
loss multipliers of get_loss_dict (:277-326), Adam(lr=1e-2, eps=1e-15) with exponential decay to
1e-4 over 200k steps for the ``fields`` and ``proposal_networks`` groups
(config_thermal_nerf.py:32-45), and DDP semantics for world_size > 1 (each rank draws its own rays;
gradients are averaged).  The autograd route (``functional.render`` + ``FusedAdam``) computes the
same numbers through the plugin API; this class is the launch-lean fast path the benchmark times.
"""

from __future__ import annotations

import os
from typing import Dict, List, Optional

import numpy as np
import torch
from torch import Tensor

from . import _lib as L
from . import functional as F
from .dist import PeerArena, allreduce_mean_, peer_padding, world_info

NUM_PROP_TENSORS = 5 * L.TNF_NUM_PROP  # leading entries of ModelTensors.param_list()


def exponential_decay_lr(step: int, lr_init: float = 1e-2, lr_final: float = 1e-4, max_steps: int = 200000) -> float:
    """nerfstudio ExponentialDecayScheduler (no warm-up, ramp 'cosine' unused): log-linear interpolation."""
    t = float(np.clip(step / max_steps, 0.0, 1.0))
    return float(np.exp(np.log(lr_init) * (1 - t) + np.log(lr_final) * t))


def arena_layout(numels, num_head_tensors: int, world_size: int = 1):
    """Offsets (in floats) of the tensors in the flat gradient / parameter / Adam-state arenas and the arena's length.
    Every tensor starts on a multiple of 4 floats (16-byte vector accesses).  The first ``num_head_tensors`` tensors
    (the proposal networks) and the rest (the field) form two slices that the multi-GPU exchange shards over the ranks
    separately: the second slice starts on a multiple of 4 * world_size, and PeerArena pads the total to one."""
    pad = peer_padding(world_size)
    offs, total = [], 0
    for i, n in enumerate(numels):
        if i == num_head_tensors:
            total = (total + pad - 1) // pad * pad
        offs.append(total)
        total += (int(n) + 3) // 4 * 4
    return offs, total


class TrainEngine:
    def __init__(self, model, *, lr: float = 1e-2, lr_final: float = 1e-4, lr_max_steps: int = 200000,
                 betas=(0.9, 0.999), eps: float = 1e-15, process_group=None, world_size: int = 1,
                 peer_fused: bool = False) -> None:
        self.model = model
        self.cfg = model.config
        self.tensors = model.tensors()
        self.params: List[Tensor] = self.tensors.param_list()
        dev = self.params[0].device
        if dev.type != "cuda":
            raise RuntimeError("TrainEngine needs the model on a CUDA device; there is no CPU path")
        self.device = dev
        # flat arenas (each tensor 16-byte aligned): gradients, exp_avg, exp_avg_sq
        world = world_info(process_group)[1] if peer_fused else 1
        offs, total = arena_layout([p.numel() for p in self.params], NUM_PROP_TENSORS, world if peer_fused else 1)
        self.offsets, self.total = offs, total
        self.prop_end = offs[NUM_PROP_TENSORS]  # [0, prop_end): proposal networks, [prop_end, total): field
        self.arena: Optional[PeerArena] = None
        # TNF_PEER_PIPELINE=0: the exchange as one block on the critical path (the two-barrier form of round 1)
        self.pipelined = os.environ.get("TNF_PEER_PIPELINE", "1") != "0"

        def views(arena):
            return [arena[o:o + p.numel()].view_as(p) for o, p in zip(offs, self.params)]

        if peer_fused:
            # multi-GPU exchange fused with Adam over NVLink peer memory (tnf_peer_adam_step): parameters move
            # into a peer-mapped arena (every rank writes its updated shard into everybody's copy), the gradient
            # arena is peer-mapped too, and the Adam state exists only for the shard this rank owns
            self.arena = PeerArena(total, dev, process_group, split=self.prop_end)
            self.pipelined = self.pipelined and self.arena.world > 1 and self.arena.gather in ("push", "multimem")
            model._peer_arena = self.arena  # model.tensors() orders other readers after a pending exchange
            with torch.no_grad():
                for v, p in zip(views(self.arena.params), self.params):
                    v.copy_(p)
                    p.data = v
            model._tensors = None
            self.tensors = model.tensors()
            self.params = self.tensors.param_list()
            self.grad_arena = self.arena.grads
            self.m_arena, self.v_arena = self.arena.exp_avg, self.arena.exp_avg_sq  # this rank's shard
            self.grads = views(self.grad_arena)
            self.exp_avg = self.exp_avg_sq = None
        else:
            self.grad_arena = torch.zeros(total, dtype=torch.float32, device=dev)
            self.m_arena = torch.zeros(total, dtype=torch.float32, device=dev)
            self.v_arena = torch.zeros(total, dtype=torch.float32, device=dev)
            self.grads, self.exp_avg, self.exp_avg_sq = views(self.grad_arena), views(self.m_arena), views(self.v_arena)
        self.lr, self.lr_final, self.lr_max_steps = lr, lr_final, lr_max_steps
        self.betas, self.eps = betas, eps
        self.pg, self.world_size = process_group, world_size
        self.step_count = 0          # trainer step
        self.field_steps = 0         # Adam step counters (a tensor without gradient does not advance)
        self.prop_steps = 0
        self.steps_since_update = 0
        self.sampler_step = 0        # ProposalNetworkSampler._step: the step of the PREVIOUS iteration (step_cb)
        self._ws: Optional[Tensor] = None
        self._field_done: Optional[torch.cuda.Event] = None
        # CTAs (256 threads, <= 64 registers) of the field-slice exchange, which shares the SMs with the proposal
        # backward (told to leave that many slots free) and with the next proposal pass (512 CTAs at 4096 rays in
        # 4 x 148 slots).  64 CTAs keep 1 MB of switch reductions in flight; two ranks move four times the bytes per
        # rank with plain peer loads and want more
        # Measured (profiles/r2_exchange_pipeline.json): 8 GPUs, in-switch reduction - 64 CTAs started under the
        # proposal backward: 0.717 ms / step (0.796 started after it, 0.830 with 296 CTAs, 0.846 unpipelined);
        # 2 GPUs, peer loads / stores, 34 MB per rank and direction - 296 CTAs started after the backward: 0.765 ms
        # (0.795 started under it: the proposal kernel then runs on half its CTAs; 0.799 unpipelined)
        switch = self.arena is not None and self.arena.gather == "multimem"  # in-switch reduction: 1/N of the bytes per rank
        self._side_ctas = int(os.environ.get("TNF_PEER_SIDE_CTAS", 64 if switch else 296))
        # start the field exchange under the proposal backward (1) or after the whole backward (0)
        self._early = os.environ.get("TNF_PEER_EARLY", "1" if switch else "0") != "0"
        mode = getattr(getattr(model, "camera_optimizer", None), "mode", "off")
        if mode != "off":
            # the engine feeds origins / directions straight to the kernels: pose refinement lives in the
            # autograd route (model(ray_bundle) ... FusedAdam), which applies the deltas and trains pose_adjustment
            raise ValueError(f"TrainEngine does not train the camera optimiser (camera_optimizer_mode={mode!r}): build "
                             "the model with camera_optimizer_mode='off' or use the plugin route, whose backward "
                             "returns dL/d origins and dL/d directions")

    # ---- reference schedules -------------------------------------------------------------
    def anneal(self, step: int) -> float:
        if not self.cfg.use_proposal_weight_anneal:
            return 1.0
        n, b = self.cfg.proposal_weights_anneal_max_num_iters, self.cfg.proposal_weights_anneal_slope
        f = float(np.clip(step / n, 0, 1))
        return b * f / ((b - 1) * f + 1)

    def prop_updated(self, step: Optional[int] = None) -> bool:
        """ProposalNetworkSampler.generate_ray_samples: the schedule and the `< 10` warm-up are evaluated with the
        sampler's `_step`, which step_cb sets AFTER each iteration - i.e. with the previous iteration's step."""
        s = self.sampler_step if step is None else step
        return self.steps_since_update > self.model._update_schedule(s) or s < 10

    # ---- one iteration -------------------------------------------------------------------
    def step(self, origins: Tensor, directions: Tensor, camera_indices: Tensor, gt_rgb: Tensor, gt_thermal: Tensor,
             jitter: Optional[Tensor] = None) -> Tensor:
        """Runs one full iteration on the current stream; returns the 4 losses (device tensor, order
        ``functional.LOSS_NAMES``).  Nothing synchronises the host."""
        cfg, step = self.cfg, self.step_count
        R = int(origins.shape[0])
        if jitter is None:
            jitter = torch.rand((L.TNF_NUM_PROP + 1, R), device=self.device)
        updated = self.prop_updated()
        kw = dict(num_samples=(*cfg.num_proposal_samples_per_ray, cfg.num_nerf_samples_per_ray),
                  near_plane=cfg.near_plane, far_plane=cfg.far_plane, anneal=self.anneal(step),
                  use_contraction=not cfg.disable_scene_contraction,
                  aabb=self.model._aabb_list(),
                  appearance_mode=L.APPEARANCE_LOOKUP, precision=self.model._precision(),
                  detach_thermal_geo=not self.model.field.pass_thermal_gradients,
                  head_mode=L.HEAD_CONCAT if self.model._is_concat() else L.HEAD_THERMAL)
        cam = camera_indices.reshape(-1)
        # pipelined exchange: the previous iteration's field slice may still be in flight - the proposal levels
        # start now, the field level waits for the event
        ready = self.arena.field_ready_event() if self.arena is not None else None
        res = F.render_forward(self.tensors, origins, directions, cam, None, None, jitter, training=True,
                               return_samples=True, save_for_backward=True, field_ready_event=ready, **kw)
        losses, g = F.losses_forward_backward(
            res["weights_list"], res["sdist_list"], res["rgb"], res["thermal"], gt_rgb, gt_thermal,
            interlevel_mult=cfg.interlevel_loss_mult, distortion_mult=cfg.distortion_loss_mult,
            use_rgb_loss=self.model.field.pass_rgb_gradients, use_thermal_loss=self.model.field.pass_thermal_gradients,
            prop_grad=updated,
            # concat_nerf: gt_thermal is channel 3 of the RGBT image, the random background is drawn here
            concat_accumulation=res["accumulation"] if self.model._is_concat() else None,
            concat_noise=torch.rand((R, 4), device=self.device) if self.model._is_concat() else None)
        grads = list(self.grads)
        if not updated:
            grads[:NUM_PROP_TENSORS] = [None] * NUM_PROP_TENSORS
        nbytes = int(L.load().tnf_backward_workspace_bytes(res["_model_struct"], R))
        if self._ws is None or self._ws.numel() < nbytes:
            self._ws = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        res["_workspace"] = self._ws
        field_done = None
        if self.arena is not None:
            self.arena.wait_zeroed()  # the previous step's memset of the gradient arena ran beside this forward
            if self.pipelined and self._early:
                # field level first, event, proposal levels: the field gradients start travelling under the latter
                if self._field_done is None:
                    self._field_done = torch.cuda.Event()
                    self._field_done.record()  # creates the handle the kernel launcher re-records
                field_done = self._field_done
        F.render_backward(self.tensors, res["_model_struct"], origins, directions, cam, None, None, jitter, res,
                          {"rgb": g["rgb"], "thermal": g["thermal"], "accumulation": g.get("accumulation"),
                           "weights_list": g["weights_list"]}, grads, field_grads_event=field_done,
                          reserve_ctas=self._side_ctas if field_done is not None else 0)
        lr = exponential_decay_lr(step, self.lr, self.lr_final, self.lr_max_steps)
        self.field_steps += 1
        if updated:
            self.prop_steps += 1
            self.steps_since_update = 0
        if self.arena is not None:
            # one kernel per slice: mean over ranks + Adam on the owned shard + parameter broadcast
            segs = [(0, self.prop_end, lr, max(self.prop_steps, 1), updated),
                    (self.prop_end, self.arena.numel, lr, self.field_steps, True)]
            if self.pipelined:
                self.arena.adam_step_pipelined(segs, beta1=self.betas[0], beta2=self.betas[1], eps=self.eps,
                                               side_ctas=self._side_ctas, tail_grads_event=field_done)
            else:
                self.arena.adam_step(segs, beta1=self.betas[0], beta2=self.betas[1], eps=self.eps, async_zero=True)
        else:
            if self.world_size > 1:
                # DDP semantics: average over ranks.  The proposal part of the arena is all zeros on
                # non-updated steps (the schedule is identical on every rank), so one collective covers both.
                allreduce_mean_(self.grad_arena, self.pg, self.world_size)
            self._adam(slice(NUM_PROP_TENSORS, len(self.params)), lr, self.field_steps)
            if updated:
                self._adam(slice(0, NUM_PROP_TENSORS), lr, self.prop_steps)
        self.sampler_step = step      # ProposalNetworkSampler.step_cb (AFTER_TRAIN_ITERATION)
        self.steps_since_update += 1
        self.step_count += 1
        return losses

    def step_host(self, origins: Tensor, directions: Tensor, camera_indices: Tensor, gt_rgb: Tensor,
                  gt_thermal: Tensor) -> Tensor:
        """One iteration from a HOST batch (what ``datamanager.next_train`` hands the trainer, pipeline_tracking.py:47-59):
        the five tensors - pinned for the copies to be asynchronous - go into persistent device buffers on the current
        stream, then :meth:`step` runs.  Returns the 4 losses on the device; the caller decides when to read them."""
        R = int(origins.shape[0])
        bufs = getattr(self, "_host_batch", None)
        if bufs is None or bufs[0].shape[0] != R:
            dev = self.device
            bufs = (torch.empty((R, 3), dtype=torch.float32, device=dev), torch.empty((R, 3), dtype=torch.float32, device=dev),
                    torch.empty((R,), dtype=torch.int64, device=dev), torch.empty((R, 3), dtype=torch.float32, device=dev),
                    torch.empty((R,), dtype=torch.float32, device=dev))
            self._host_batch = bufs
        for d, h in zip(bufs, (origins, directions, camera_indices.reshape(-1), gt_rgb, gt_thermal.reshape(-1))):
            d.copy_(h, non_blocking=True)
        return self.step(*bufs)

    def _adam(self, sl: slice, lr: float, step: int) -> None:
        n = len(self.params[sl])
        F.adam_step(self.params[sl], self.grads[sl], self.exp_avg[sl], self.exp_avg_sq[sl], [lr] * n, step=step,
                    beta1=self.betas[0], beta2=self.betas[1], eps=self.eps, zero_grads=True)

    def state_dict(self) -> Dict[str, object]:
        if self.arena is not None:
            self.arena.wait_params()
        return {"step": self.step_count, "field_steps": self.field_steps, "prop_steps": self.prop_steps,
                "steps_since_update": self.steps_since_update, "sampler_step": self.sampler_step,
                "exp_avg": self.m_arena, "exp_avg_sq": self.v_arena}

    def load_state_dict(self, sd: Dict[str, object]) -> None:
        self.step_count, self.field_steps = int(sd["step"]), int(sd["field_steps"])
        self.prop_steps, self.steps_since_update = int(sd["prop_steps"]), int(sd["steps_since_update"])
        self.sampler_step = int(sd.get("sampler_step", max(self.step_count - 1, 0)))
        self.m_arena.copy_(sd["exp_avg"])
        self.v_arena.copy_(sd["exp_avg_sq"])
