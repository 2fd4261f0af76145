"""B200 ThermoNeRF model: the reference's ``ThermalNerfModel`` plugin surface over the fused
sm_100a kernels.

Mirrors thermo_nerf/thermal_nerf/thermal_nerf_model.py (config :46-56, ctor :67-84,
populate_modules :86-208, get_outputs :210-275, get_loss_dict :277-326) and the pieces of
nerfstudio's ``Model`` / ``NerfactoModel`` the reference's callers use
(``forward``, ``get_outputs_for_camera_ray_bundle``, ``get_param_groups``,
``get_training_callbacks``, ``get_metrics_dict``).  Same names, argument meaning, output
keys/shapes and error behaviour; the arithmetic is one call into libtnf_b200.so.
"""

from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any, Callable, Dict, List, Literal, Optional, Tuple, Type

import numpy as np
import torch
from torch import Tensor, nn

from . import _lib as L
from . import functional as F
from .modules import CameraOptimizer, HashMLPDensityField, ThermalNerfactoTField
from .rays import RayBundle


@dataclass
class ThermalNerfactoModelConfig:
    """ThermalNerfactoModelConfig (nerfacto_config/thermal_nerfacto.py:13-25) + the NerfactoModelConfig defaults it
    inherits: nerfacto on one image modality (the ``nerfacto`` / ``thermal-nerfacto`` ModelTypes of
    train_eval_script.py:66-73) - sampler / field / renderers without the thermal head."""

    _target: Type = field(default_factory=lambda: ThermalNerfactoModel)
    # ThermalNerfactoModelConfig
    max_temperature: float = 1.0
    min_temperature: float = 0.0
    cold: bool = False
    camera_optimizer_mode: Literal["off", "SO3xR3"] = "SO3xR3"
    # NerfactoModelConfig
    near_plane: float = 0.05
    far_plane: float = 1000.0
    background_color: str = "last_sample"
    hidden_dim: int = 64
    hidden_dim_color: int = 64
    hidden_dim_transient: int = 64
    num_levels: int = 16
    base_res: int = 16
    max_res: int = 2048
    log2_hashmap_size: int = 19
    features_per_level: int = 2
    num_proposal_samples_per_ray: Tuple[int, ...] = (256, 96)
    num_nerf_samples_per_ray: int = 48
    proposal_update_every: int = 5
    proposal_warmup: int = 5000
    num_proposal_iterations: int = 2
    use_same_proposal_network: bool = False
    proposal_net_args_list: List[Dict] = field(
        default_factory=lambda: [
            {"hidden_dim": 16, "log2_hashmap_size": 17, "num_levels": 5, "max_res": 128, "use_linear": False},
            {"hidden_dim": 16, "log2_hashmap_size": 17, "num_levels": 5, "max_res": 256, "use_linear": False},
        ]
    )
    proposal_initial_sampler: Literal["piecewise", "uniform"] = "piecewise"
    interlevel_loss_mult: float = 1.0
    distortion_loss_mult: float = 0.002
    use_proposal_weight_anneal: bool = True
    use_average_appearance_embedding: bool = True
    proposal_weights_anneal_slope: float = 10.0
    proposal_weights_anneal_max_num_iters: int = 1000
    use_single_jitter: bool = True
    predict_normals: bool = False
    disable_scene_contraction: bool = False
    use_gradient_scaling: bool = False
    appearance_embed_dim: int = 32
    implementation: str = "b200"
    eval_num_rays_per_chunk: int = 1 << 16  # config_thermal_nerf.py:30
    # B200 specific
    precision: Literal["fp32", "tc_fp16"] = "tc_fp16"
    thermal_head: bool = False  # the thermal head belongs to ThermalNerfModelConfig
    concat_head: bool = False   # the 4-channel RGBT colour head belongs to ConcatNerfModelConfig

    def setup(self, **kwargs) -> Any:
        return self._target(self, **kwargs)


@dataclass
class ThermalNerfModelConfig(ThermalNerfactoModelConfig):
    """ThermalNerfModelConfig (thermal_nerf_model.py:46-56): a subclass of ThermalNerfactoModelConfig, as
    train_eval_script.py:94 requires."""

    _target: Type = field(default_factory=lambda: ThermalNerfModel)
    use_transient_embedding: bool = False
    thermal_loss_weight: float = 1.0  # declared but unused by the reference (thermal_nerf_model.py:53 vs :321-324)
    pass_thermal_gradients: bool = True
    thermal_head: bool = True


@dataclass
class TrainingCallback:
    """Shape of nerfstudio's TrainingCallback (where_to_run / update_every_num_iters / func)."""

    where_to_run: List[str]
    update_every_num_iters: int
    func: Callable

    def run_callback(self, step: int) -> None:
        if step % self.update_every_num_iters == 0:
            self.func(step)


BEFORE_TRAIN_ITERATION = "BEFORE_TRAIN_ITERATION"
AFTER_TRAIN_ITERATION = "AFTER_TRAIN_ITERATION"


class _SceneBox:
    def __init__(self, aabb: Tensor) -> None:
        self.aabb = aabb


class KernelModelMixin:
    """The kernel-backed part of the Model surface: ``forward`` / ``get_outputs`` / ``get_outputs_for_camera_ray_bundle``
    / ``get_metrics_dict`` / ``get_loss_dict`` over libtnf_b200.  It only touches what nerfstudio's ``NerfactoModel``
    and the reference's ``ThermalNerfModel`` expose (``config``, ``field``, ``proposal_networks``, ``camera_optimizer``,
    ``scene_box``, ``device``, ``training`` and the ProposalNetworkSampler state), so the same code serves this
    package's stand-alone model classes and the subclass of the reference's own class in ``nerfstudio_plugin.py``."""

    _tensors: Optional[F.ModelTensors] = None

    # ------------------------------------------------------------------ tensors of the path
    def _apply(self, fn, *a, **k):  # parameters may move: re-resolve tensor references
        self._tensors = None
        return super()._apply(fn, *a, **k)

    def load_state_dict(self, state_dict, strict: bool = True, **k):  # type: ignore[override]
        self._tensors = None
        return super().load_state_dict(state_dict, strict=strict, **k)

    def tensors(self) -> F.ModelTensors:
        arena = self.__dict__.get("_peer_arena")
        if arena is not None:
            # a TrainEngine with the pipelined multi-GPU exchange owns the parameters: a field slice may still be in
            # flight on its side stream - readers that come through here (eval renders, the autograd route) wait for it
            arena.wait_params()
        if self._tensors is None:
            self._tensors = F.ModelTensors.from_module(self)
        return self._tensors

    # ------------------------------------------------------------------ ProposalNetworkSampler state
    def _sampler_state(self):
        """The object carrying ``_anneal`` / ``_steps_since_update`` / ``_step``: nerfstudio's ProposalNetworkSampler
        when the model has one (the reference's populate_modules builds it, thermal_nerf_model.py:172-179), else
        the model itself."""
        return getattr(self, "proposal_sampler", None) or self

    def _update_schedule(self, step: int) -> float:
        sched = getattr(self._sampler_state(), "update_sched", None)
        return float(sched(step)) if sched is not None else self.update_schedule(step)

    def _precision(self) -> int:
        return L.PRECISION_FP32 if getattr(self.config, "precision", "tc_fp16") == "fp32" else L.PRECISION_TC_FP16

    def _has_thermal_head(self) -> bool:
        return bool(getattr(self.config, "thermal_head", True))

    def _is_concat(self) -> bool:
        """ConcatNerfModel (rgb_concat/concat_nerfacto_model.py): 4-channel RGBT colour head, no thermal head."""
        return bool(getattr(self.config, "concat_head", False))

    def _appearance_mode(self) -> int:
        if self.training:
            return L.APPEARANCE_LOOKUP
        return L.APPEARANCE_MEAN if self.config.use_average_appearance_embedding else L.APPEARANCE_ZEROS

    def _collider_near(self) -> float:
        # NearFarCollider(reset_near_plane=True): eval renders from t=0
        return self.config.near_plane if self.training else 0.0

    def forward(self, ray_bundle) -> Dict[str, Any]:
        """Model.forward: collider (thermal_nerf_model.py:182-184) then get_outputs.  The
        NearFarCollider only writes two constants per ray, so they travel as kernel
        arguments instead of [R,1] tensors (a caller-supplied nears/fars still wins)."""
        return self.get_outputs(ray_bundle)

    def _aabb_list(self) -> List[float]:
        box = getattr(self, "_aabb_cache", None)
        if box is None:  # host copy made once: a CUDA-resident aabb would otherwise cost a sync per call
            box = [float(x) for x in torch.as_tensor(self.scene_box.aabb).reshape(-1).tolist()]
            self._aabb_cache = box
        return box

    def _render_kwargs(self) -> Dict[str, Any]:
        cfg = self.config
        return dict(
            num_samples=(*cfg.num_proposal_samples_per_ray, cfg.num_nerf_samples_per_ray),
            near_plane=self._collider_near(), far_plane=cfg.far_plane, anneal=float(self._sampler_state()._anneal),
            use_contraction=not cfg.disable_scene_contraction,
            aabb=self._aabb_list(), appearance_mode=self._appearance_mode(), precision=self._precision(),
            head_mode=L.HEAD_CONCAT if self._is_concat() else L.HEAD_THERMAL)

    def get_outputs(self, ray_bundle, depth_clip_chunk: int = 0) -> Dict[str, Any]:
        """thermal_nerf_model.py:210-275 as one fused kernel launch (eval) or one autograd node over
        tnf_render_forward / tnf_render_backward (training)."""
        cfg = self.config
        if self.training:
            self.camera_optimizer.apply_to_raybundle(ray_bundle)
        shape = tuple(ray_bundle.origins.shape[:-1])
        o = ray_bundle.origins.reshape(-1, 3).contiguous().float()
        d = ray_bundle.directions.reshape(-1, 3).contiguous().float()
        R = o.shape[0]
        cam = ray_bundle.camera_indices
        if self.training and cam is None:
            raise AttributeError("Camera indices are not provided.")  # thermal_field.py:113-114
        nears = ray_bundle.nears.reshape(-1).contiguous().float() if ray_bundle.nears is not None else None
        fars = ray_bundle.fars.reshape(-1).contiguous().float() if ray_bundle.fars is not None else None
        cam_flat = cam.reshape(-1) if cam is not None else None
        if self.training and torch.is_grad_enabled():
            # ProposalNetworkSampler: the proposal densities only carry gradients on "updated" steps
            st = self._sampler_state()
            updated = st._steps_since_update > self._update_schedule(st._step) or st._step < 10
            jitter = torch.rand((L.TNF_NUM_PROP + 1, R), device=o.device)
            # origins / directions stay in the graph when the camera optimiser produced them: the backward then
            # also returns dL/d origins, dL/d directions and autograd carries them into the pose deltas
            res = F.render(self.tensors(), o, d, cam_flat, nears, fars, jitter, prop_grad=updated,
                           detach_thermal_geo=not self.field.pass_thermal_gradients, **self._render_kwargs())
            if updated:
                st._steps_since_update = 0
        else:
            jitter = torch.rand((L.TNF_NUM_PROP + 1, R), device=o.device) if self.training else None
            res = F.render_forward(self.tensors(), o, d, cam_flat, nears, fars, jitter, training=self.training,
                                   depth_clip_chunk=depth_clip_chunk, return_samples=self.training,
                                   **self._render_kwargs())
        rgb = res["rgb"].view(*shape, 3)
        if self._is_concat():  # "rgb" is the 4-channel RGBT image (the kernel returns channel 3 as `thermal`)
            rgb = torch.cat([rgb, res["thermal"].view(*shape, 1)], dim=-1)
        outputs: Dict[str, Any] = {
            "rgb": rgb,
            "accumulation": res["accumulation"].view(*shape, 1),
            "depth": res["depth"].view(*shape, 1),
            "expected_depth": res["expected_depth"].view(*shape, 1),
        }
        if self.training:
            outputs["weights_list"] = res["weights_list"]
            outputs["ray_samples_list"] = res["sdist_list"]  # spacing bins [R,S+1] per level
        for i in range(cfg.num_proposal_iterations):
            outputs[f"prop_depth_{i}"] = res[f"prop_depth_{i}"].view(*shape, 1)
        if self._has_thermal_head():
            outputs["thermal"] = res["thermal"].view(*shape, 1)
        return outputs

    @torch.no_grad()
    def get_outputs_for_camera_ray_bundle(self, camera_ray_bundle) -> Dict[str, Tensor]:
        """nerfstudio Model.get_outputs_for_camera_ray_bundle (called at renderer.py:185,
        evaluator.py:79).  The reference loops over eval_num_rays_per_chunk slices; here the
        whole image is ONE launch and the only chunk-dependent quantity (the expected-depth
        clip range) is evaluated per chunk inside the kernel, so results are identical."""
        input_device = camera_ray_bundle.directions.device
        image_shape = tuple(camera_ray_bundle.origins.shape[:-1])
        flat = RayBundle(
            origins=camera_ray_bundle.origins.reshape(-1, 3).to(self.device),
            directions=camera_ray_bundle.directions.reshape(-1, 3).to(self.device),
            camera_indices=(camera_ray_bundle.camera_indices.reshape(-1, 1).to(self.device)
                            if camera_ray_bundle.camera_indices is not None else None),
            nears=camera_ray_bundle.nears.reshape(-1, 1).to(self.device) if camera_ray_bundle.nears is not None else None,
            fars=camera_ray_bundle.fars.reshape(-1, 1).to(self.device) if camera_ray_bundle.fars is not None else None,
        )
        if flat.nears is None or flat.fars is None:
            flat.nears, flat.fars = None, None  # constants are folded into the kernel arguments
        was_training = self.training
        outputs = KernelModelMixin.get_outputs(self, flat, depth_clip_chunk=self.config.eval_num_rays_per_chunk)
        assert was_training == self.training
        res = {}
        for k, v in outputs.items():
            if not isinstance(v, Tensor):
                continue  # nerfstudio skips non-tensor outputs (weights_list etc.)
            res[k] = v.view(*image_shape, -1).to(input_device)
        # RenderedImageModality.RGB.value == "img" (rendered_image_modalities.py:5) while get_outputs emits
        # "rgb": the alias lets the unchanged render_video_script.py default modalities work
        res["img"] = res["rgb"]
        return res

    @torch.no_grad()
    def get_outputs_for_camera(self, cameras, camera_idx: int) -> Dict[str, Tensor]:
        """``get_outputs_for_camera_ray_bundle(cameras.generate_rays(camera_indices=camera_idx))``
        (renderer.py:183-187, evaluator.py:69-79) as one launch: the rays of a perspective camera are generated
        inside the kernel, so no [H,W,3] origin/direction tensors are built, stored or re-read.  ``cameras``
        exposes nerfstudio's ``Cameras`` attributes (camera_to_worlds, fx, fy, cx, cy, width, height).
        Eval mode only (training batches are random pixels, not whole frames)."""
        if self.training:
            raise RuntimeError("get_outputs_for_camera is an eval-mode call")

        def scalar(v):
            v = v[camera_idx] if (torch.is_tensor(v) and v.dim() > 0) else v
            return float(v)

        c2w = cameras.camera_to_worlds[camera_idx].detach().to("cpu", torch.float32)
        H, W = int(scalar(cameras.height)), int(scalar(cameras.width))
        cam = F.pack_camera(c2w, scalar(cameras.fx), scalar(cameras.fy), scalar(cameras.cx), scalar(cameras.cy), W, H)
        kw = self._render_kwargs()
        res = F.render_forward(self.tensors(), None, None, camera=cam, training=False,
                               depth_clip_chunk=self.config.eval_num_rays_per_chunk, **kw)
        out = {k: v.view(H, W, -1) for k, v in res.items() if isinstance(v, Tensor)}
        if self._is_concat():
            out["rgb"] = torch.cat([out["rgb"], out["thermal"]], dim=-1)
        if not self._has_thermal_head():
            out.pop("thermal", None)
        out["img"] = out["rgb"]
        return out

    # ------------------------------------------------------------------ losses / metrics
    def _fused_losses(self, outputs, batch) -> Dict[str, Tensor]:
        """tnf_losses over the training outputs, evaluated once per step (metrics + loss dict share it)."""
        cache = outputs.get("_b200_losses")
        if cache is None:
            # the reference keeps the thermal GT on the host and moves it here (thermal_dataset.py:18-20,
            # thermal_nerf_model.py:319); non_blocking keeps that copy from draining the stream when the
            # host tensor is pinned (a blocking .to() waits for the forward kernel before the loss can launch)
            full = batch["image"].to(self.device, non_blocking=True)
            image = full[..., :3].reshape(-1, 3).float()
            if self._is_concat():
                # ConcatNerfModel.get_loss_dict (concat_nerfacto_model.py:197-211): the batch image is RGBT; the
                # renderer's "random" background is blended into the prediction only (rgbt_renderer.py:134-140)
                pred = outputs["rgb"].reshape(-1, 4)
                noise = torch.rand_like(pred)
                cache = F.losses(
                    {"rgb": pred[:, :3], "thermal": pred[:, 3], "accumulation": outputs["accumulation"].reshape(-1),
                     "weights_list": outputs["weights_list"], "sdist_list": outputs["ray_samples_list"]},
                    image, full[..., 3].reshape(-1).float(), interlevel_mult=self.config.interlevel_loss_mult,
                    distortion_mult=self.config.distortion_loss_mult, concat_noise=noise)
                outputs["_b200_losses"] = cache
                return cache
            if self._has_thermal_head():
                thermal = batch["thermal"].to(self.device, non_blocking=True).reshape(-1).float()
                pred_th = outputs["thermal"].reshape(-1)
            else:  # nerfacto field: no thermal term (use_thermal_loss is False below); any [R] tensors do
                thermal = pred_th = torch.zeros(image.shape[0], dtype=torch.float32, device=self.device)
            w = outputs["weights_list"]
            cache = F.losses(
                {"rgb": outputs["rgb"].reshape(-1, 3), "thermal": pred_th,
                 "weights_list": w, "sdist_list": outputs["ray_samples_list"]},
                image, thermal, interlevel_mult=self.config.interlevel_loss_mult,
                distortion_mult=self.config.distortion_loss_mult, use_rgb_loss=self.field.pass_rgb_gradients,
                use_thermal_loss=self._has_thermal_head() and self.field.pass_thermal_gradients)
            outputs["_b200_losses"] = cache
        return cache

    def get_metrics_dict(self, outputs, batch) -> Dict[str, Tensor]:
        """NerfactoModel.get_metrics_dict (inherited by the reference): psnr, and in training the
        distortion metric that get_loss_dict consumes (thermal_nerf_model.py:303-305)."""
        metrics: Dict[str, Tensor] = {}
        if self.training:
            losses = self._fused_losses(outputs, batch)
            metrics["distortion"] = losses["distortion_loss"] / self.config.distortion_loss_mult
        with torch.no_grad():
            if self._is_concat():  # psnr over the four RGBT channels, without the loss's random background
                gt = batch["image"].to(self.device)
                mse = torch.mean((outputs["rgb"].detach() - gt.reshape(outputs["rgb"].shape)) ** 2)
            elif self.training and self.field.pass_rgb_gradients:
                mse = losses["rgb_loss"].detach()  # the fused loss kernel already reduced MSE(rgb, gt)
            else:
                gt_rgb = batch["image"].to(self.device)
                mse = torch.mean((outputs["rgb"].detach() - gt_rgb[..., :3]) ** 2)
            metrics["psnr"] = -10.0 * torch.log10(mse)
        cam_metrics = getattr(self.camera_optimizer, "get_metrics_dict", None)
        if cam_metrics is not None:  # nerfstudio CameraOptimizer: camera_opt_translation / _rotation
            cam_metrics(metrics)
        return metrics

    def get_loss_dict(self, outputs, batch, metrics_dict=None) -> Dict[str, Tensor]:
        """thermal_nerf_model.py:277-326.  background_color="last_sample" makes
        blend_background_for_loss_computation the identity on (pred, gt)."""
        if self.training:
            assert metrics_dict is not None and "distortion" in metrics_dict  # thermal_nerf_model.py:302
            loss_dict = dict(self._fused_losses(outputs, batch))
            cam_loss = getattr(self.camera_optimizer, "get_loss_dict", None)
            if self._is_concat() and cam_loss is not None:  # only ConcatNerfModel adds it (concat_nerfacto_model.py:231)
                cam_loss(loss_dict)
            return loss_dict
        loss_dict: Dict[str, Tensor] = {}
        if self._is_concat():  # concat_nerfacto_model.py:197-211 outside training: the blended colour term only
            image = batch["image"].to(self.device)
            pred = outputs["rgb"] + torch.rand_like(outputs["rgb"]) * (1.0 - outputs["accumulation"])
            return {"rgb_loss": torch.nn.functional.mse_loss(image.reshape(pred.shape), pred)}
        image = batch["image"].to(self.device)[..., :3]
        if self.field.pass_rgb_gradients:
            loss_dict["rgb_loss"] = torch.nn.functional.mse_loss(image, outputs["rgb"])
        if self._has_thermal_head() and self.field.pass_thermal_gradients:
            loss_dict["thermal"] = torch.nn.functional.mse_loss(outputs["thermal"], batch["thermal"].to(self.device))
        return loss_dict


def check_supported_config(cfg) -> None:
    """The options libtnf_b200 compiles; everything else raises instead of silently taking another path."""
    if cfg.predict_normals or getattr(cfg, "use_transient_embedding", False) or cfg.use_gradient_scaling:
        raise ValueError("predict_normals / use_transient_embedding / use_gradient_scaling are not on the "
                         "thermal-nerf hot path and are not built into libtnf_b200")
    if cfg.proposal_initial_sampler != "piecewise" or not cfg.use_single_jitter:
        raise ValueError("libtnf_b200 implements the default piecewise initial sampler with single jitter")
    if cfg.num_proposal_iterations != L.TNF_NUM_PROP or cfg.use_same_proposal_network:
        raise ValueError("libtnf_b200 is built for 2 distinct proposal networks (nerfacto default)")
    if cfg.background_color != "last_sample":
        raise ValueError("libtnf_b200 implements background_color='last_sample' (nerfacto default)")
    if getattr(cfg, "camera_optimizer_mode", "off") not in ("off", "SO3xR3"):
        raise ValueError("libtnf_b200 supports camera_optimizer_mode 'off' or 'SO3xR3'")


class ThermalNerfactoModel(KernelModelMixin, nn.Module):
    """ThermalNerfactoModel (thermo_nerf/nerfacto_config/thermal_nerfacto.py:28-84) on libtnf_b200: the base class of
    the hierarchy, as in the reference (``evaluator.py:76`` asserts ``isinstance(model, ThermalNerfactoModel)`` for
    every model type).  Constructor without the thermal metadata requirement; with a plain
    ``ThermalNerfactoModelConfig`` there is no "thermal" output and only the rgb / interlevel / distortion losses."""

    config: ThermalNerfactoModelConfig

    def __init__(self, config: ThermalNerfactoModelConfig, scene_box, num_train_data: int,
                 metadata: Optional[dict] = None, **kwargs) -> None:
        if config.thermal_head and not isinstance(self, ThermalNerfModel):
            raise ValueError("ThermalNerfactoModel is the model without a thermal head (thermal_head=False)")
        if config.concat_head and not isinstance(self, ConcatNerfModel):
            raise ValueError("the RGBT colour head (concat_head=True) belongs to ConcatNerfModel")
        super().__init__()
        self.config = config
        self.scene_box = scene_box if hasattr(scene_box, "aabb") else _SceneBox(torch.as_tensor(scene_box))
        self.num_train_data = num_train_data
        self.kwargs = kwargs
        self.max_temperature = config.max_temperature
        self.min_temperature = config.min_temperature
        self.device_indicator_param = nn.Parameter(torch.empty(0))
        self.populate_modules()
        self._tensors = None

    # ------------------------------------------------------------------ construction
    def populate_modules(self) -> None:
        cfg = self.config
        check_supported_config(cfg)
        aabb = torch.as_tensor(self.scene_box.aabb, dtype=torch.float32)
        self.field = ThermalNerfactoTField(
            aabb, num_images=self.num_train_data, hidden_dim=cfg.hidden_dim, num_levels=cfg.num_levels,
            max_res=cfg.max_res, base_res=cfg.base_res, features_per_level=cfg.features_per_level,
            log2_hashmap_size=cfg.log2_hashmap_size, hidden_dim_color=cfg.hidden_dim_color,
            hidden_dim_transient=cfg.hidden_dim_transient,
            use_average_appearance_embedding=cfg.use_average_appearance_embedding,
            appearance_embedding_dim=cfg.appearance_embed_dim,
            pass_thermal_gradients=getattr(cfg, "pass_thermal_gradients", False),
            thermal_head=cfg.thermal_head, use_contraction=not cfg.disable_scene_contraction,
            rgb_out_dim=4 if cfg.concat_head else 3)
        self.camera_optimizer = CameraOptimizer(self.num_train_data, cfg.camera_optimizer_mode)
        self.proposal_networks = nn.ModuleList()
        for i in range(cfg.num_proposal_iterations):
            a = dict(cfg.proposal_net_args_list[min(i, len(cfg.proposal_net_args_list) - 1)])
            if a.pop("use_linear", False):
                raise ValueError("use_linear proposal networks are not built into libtnf_b200")
            self.proposal_networks.append(HashMLPDensityField(aabb, use_contraction=not cfg.disable_scene_contraction, **a))
        # density_fns / renderers of the reference's populate_modules (thermal_nerf_model.py:127-208): the fused
        # kernel does not go through them, callers that compose the modules by hand can
        self.density_fns = [net.density_fn for net in self.proposal_networks]
        from .surface import ThermalRenderer

        self.thermal_renderer = ThermalRenderer()
        # ProposalNetworkSampler state (annealing + update schedule)
        self._anneal = 1.0
        self._steps_since_update = 0
        self._step = 0
        self.step = 0

    @property
    def device(self) -> torch.device:
        return self.device_indicator_param.device

    # ------------------------------------------------------------------ nerfstudio Model surface
    def get_param_groups(self) -> Dict[str, List[nn.Parameter]]:
        groups: Dict[str, List[nn.Parameter]] = {
            "proposal_networks": list(self.proposal_networks.parameters()),
            "fields": list(self.field.parameters()),
        }
        self.camera_optimizer.get_param_groups(param_groups=groups)
        return groups

    def update_schedule(self, step: int) -> float:  # thermal_nerf_model.py:152-161
        return float(np.clip(np.interp(step, [0, self.config.proposal_warmup], [0, self.config.proposal_update_every]),
                             1, self.config.proposal_update_every))

    def get_training_callbacks(self, training_callback_attributes=None) -> List[TrainingCallback]:
        callbacks = []
        if self.config.use_proposal_weight_anneal:
            N = self.config.proposal_weights_anneal_max_num_iters

            def set_anneal(step: int) -> None:
                self.step = step
                train_frac = float(np.clip(step / N, 0, 1))
                b = self.config.proposal_weights_anneal_slope
                self._anneal = b * train_frac / ((b - 1) * train_frac + 1)

            callbacks.append(TrainingCallback([BEFORE_TRAIN_ITERATION], 1, set_anneal))

        def step_cb(step: int) -> None:  # ProposalNetworkSampler.step_cb
            self._step = step
            self._steps_since_update += 1

        callbacks.append(TrainingCallback([AFTER_TRAIN_ITERATION], 1, step_cb))
        return callbacks

    # ------------------------------------------------------------------ evaluation (Evaluator, evaluator.py:79-87)
    lpips: Optional[Callable[[Tensor, Tensor], Tensor]] = None  # LPIPS needs pretrained weights: plug a callable in

    @staticmethod
    def psnr(gt: Tensor, pred: Tensor) -> Tensor:
        """torchmetrics PeakSignalNoiseRatio(data_range=1.0), as NerfactoModel.psnr."""
        return -10.0 * torch.log10(torch.mean((gt - pred) ** 2))

    @staticmethod
    def ssim(gt: Tensor, pred: Tensor) -> Tensor:
        """torchmetrics.functional.structural_similarity_index_measure with its defaults (11x11 gaussian window,
        sigma 1.5, k1 0.01, k2 0.03, data_range taken from the data, reflect padding cropped away), [N,C,H,W]."""
        data_range = torch.maximum(pred.max() - pred.min(), gt.max() - gt.min())
        c1, c2 = (0.01 * data_range) ** 2, (0.03 * data_range) ** 2
        k, sigma, pad = 11, 1.5, 5
        x = torch.arange(k, dtype=gt.dtype, device=gt.device) - (k - 1) / 2
        g1 = torch.exp(-(x / sigma) ** 2 / 2)
        g1 = g1 / g1.sum()
        C = gt.shape[1]
        win = (g1[:, None] * g1[None, :]).expand(C, 1, k, k)
        p = torch.nn.functional.pad(pred, (pad, pad, pad, pad), mode="reflect")
        t = torch.nn.functional.pad(gt, (pad, pad, pad, pad), mode="reflect")
        stack = torch.cat([p, t, p * p, t * t, p * t])
        out = torch.nn.functional.conv2d(stack, win, groups=C)
        n = pred.shape[0]
        mu_p, mu_t, e_pp, e_tt, e_pt = (out[i * n:(i + 1) * n] for i in range(5))
        s_pp, s_tt, s_pt = e_pp - mu_p * mu_p, e_tt - mu_t * mu_t, e_pt - mu_p * mu_t
        ssim_map = ((2 * mu_p * mu_t + c1) * (2 * s_pt + c2)) / ((mu_p * mu_p + mu_t * mu_t + c1) * (s_pp + s_tt + c2))
        return ssim_map[..., pad:-pad, pad:-pad].reshape(n, -1).mean(-1).mean()

    def _lpips(self, gt: Tensor, pred: Tensor) -> float:
        return float(self.lpips(gt, pred)) if self.lpips is not None else float("nan")

    def _nerfacto_metrics_and_images(self, outputs, batch) -> Tuple[Dict[str, float], Dict[str, Tensor], Tensor, Tensor]:
        """NerfactoModel.get_image_metrics_and_images: psnr / ssim / lpips of the colour image + the viewer panels.
        The accumulation / depth panels are grey-scale (nerfstudio's default colour maps come from matplotlib)."""
        dev = self.device
        gt_rgb = batch["image"].to(dev)[..., :3]
        rgb = outputs["rgb"]

        def gray(t: Tensor) -> Tensor:  # colormaps.apply_float_colormap(colormap="gray")
            return t.repeat(1, 1, 3) if t.shape[-1] == 1 else t

        acc = outputs["accumulation"]
        depth = outputs["depth"]
        near, far = float(depth.min()), float(depth.max())
        depth_n = torch.clip((depth - near) / (far - near + 1e-10), 0, 1)
        images = {"img": torch.cat([gt_rgb, rgb], dim=1), "accumulation": gray(acc), "depth": gray(depth_n)}
        for i in range(self.config.num_proposal_iterations):
            pd = outputs[f"prop_depth_{i}"]
            images[f"prop_depth_{i}"] = gray(torch.clip((pd - float(pd.min())) / (float(pd.max() - pd.min()) + 1e-10), 0, 1))
        g4, p4 = torch.moveaxis(gt_rgb, -1, 0)[None], torch.moveaxis(rgb, -1, 0)[None]
        metrics: Dict[str, float] = {"psnr": float(self.psnr(g4, p4)), "ssim": float(self.ssim(g4, p4)),
                                     "lpips": self._lpips(g4, p4)}
        return metrics, images, g4, p4

    def get_image_metrics_and_images(self, outputs: Dict[str, Tensor], batch: Dict[str, Tensor],
                                     threshold: Optional[float] = None) -> Tuple[Dict[str, float], Dict[str, Tensor]]:
        """ThermalNerfactoModel.get_image_metrics_and_images (thermal_nerfacto.py:47-84): the "RGB" images are the
        thermal images for the thermal-nerfacto method, so the temperature MAE is taken on rgb.  Evaluation-only glue
        in plain PyTorch on the already rendered [H,W,C] outputs.  LPIPS needs a pretrained network that is not
        shipped here: assign ``model.lpips`` (e.g. torchmetrics' LearnedPerceptualImagePatchSimilarity) or the lpips
        entries are NaN."""
        if self._is_concat():
            return self._concat_metrics_and_images(outputs, batch, threshold)
        metrics, images, g4, p4 = self._nerfacto_metrics_and_images(outputs, batch)
        metrics["mae_foreground"] = float(self.mae_thermal(g4, p4, threshold=threshold))
        metrics["mae"] = float(self.mae_thermal(g4, p4, threshold=None))
        return metrics, images

    def mae_thermal(self, gt: Tensor, pred: Tensor, threshold: Optional[float] = None) -> Tensor:
        """thermal_metrics.py:5-34."""
        if threshold:
            idx = torch.where(gt < threshold) if self.config.cold else torch.where(gt > threshold)
            gt, pred = gt[idx], pred[idx]
        span = self.max_temperature - self.min_temperature
        return torch.mean(torch.abs((gt * span + self.min_temperature) - (pred * span + self.min_temperature)))


class ThermalNerfModel(ThermalNerfactoModel):
    """ThermalNerfModel (thermo_nerf/thermal_nerf/thermal_nerf_model.py:60-393) on libtnf_b200: config :46-56, ctor
    :67-84, populate_modules :86-208, get_outputs :210-275, get_loss_dict :277-326; a subclass of
    ThermalNerfactoModel as in the reference.  Same names, argument meaning, output keys / shapes and error
    behaviour; the arithmetic is one call into libtnf_b200.so."""

    config: ThermalNerfModelConfig

    def __init__(self, config: ThermalNerfModelConfig, metadata: dict, scene_box, num_train_data: int,
                 **kwargs) -> None:
        if config.thermal_head and "thermal" not in metadata.keys():  # thermal_nerf_model.py:75-76
            raise ValueError("Thermal images not found in metadata.")
        super().__init__(config, scene_box, num_train_data, metadata=metadata, **kwargs)

    def get_image_metrics_and_images(self, outputs: Dict[str, Tensor], batch: Dict[str, Tensor],
                                     threshold: Optional[float] = None) -> Tuple[Dict[str, float], Dict[str, Tensor]]:
        """thermal_nerf_model.py:328-398 on top of the base class: psnr_thermal / ssim_thermal / lpips_thermal,
        mae_thermal and mae_thermal_foreground (thermal_metrics.py), plus the side-by-side images the viewer logs."""
        if not self.config.thermal_head:
            return super().get_image_metrics_and_images(outputs, batch, threshold=threshold)
        # the reference reaches the base method through super() *without* its threshold
        # (thermal_nerf_model.py:339), so both colour-image entries are the unthresholded MAE
        metrics, images = super().get_image_metrics_and_images(outputs, batch, threshold=None)
        dev = self.device

        def gray(t: Tensor) -> Tensor:
            return t.repeat(1, 1, 3) if t.shape[-1] == 1 else t

        gt_th = batch["thermal"].to(dev)  # thermal_nerf_model.py:347 (the reference forgets the .to() at :355)
        th = outputs["thermal"]
        images["thermal"] = gray(th)
        images["thermal_combined"] = torch.cat([gray(gt_th), gray(th)], dim=1)
        gt4, th4 = torch.moveaxis(gt_th, -1, 0)[None], torch.moveaxis(th, -1, 0)[None]
        metrics["psnr_thermal"] = float(self.psnr(gt4, th4))
        metrics["ssim_thermal"] = float(self.ssim(gt4, th4))
        metrics["lpips_thermal"] = self._lpips(torch.repeat_interleave(gt4, 3, dim=1), torch.repeat_interleave(th4, 3, dim=1))
        metrics["mae_thermal_foreground"] = float(self.mae_thermal(gt4, th4, threshold=threshold))
        metrics["mae_thermal"] = float(self.mae_thermal(gt4, th4, threshold=None))
        return metrics, images


@dataclass
class ConcatNerfModelConfig(ThermalNerfactoModelConfig):
    """ConcatNerfModelConfig (thermo_nerf/rgb_concat/concat_nerfacto_model.py:52-57): the ``concat_nerf`` model type
    of train_eval_script.py:74-78."""

    _target: Type = field(default_factory=lambda: ConcatNerfModel)
    concat_head: bool = True


class ConcatNerfModel(ThermalNerfactoModel):
    """ConcatNerfModel (rgb_concat/concat_nerfacto_model.py:60-324) on libtnf_b200: the ablation baseline that
    renders temperature as a fourth colour channel.  One RGBT colour head 63-64-64-4 (concat_field.py:65-75), no
    thermal head; ``outputs["rgb"]`` is [*, 4]; RGBTRenderer with its default "random" background, i.e. the plain
    weighted sum (rgbt_renderer.py:63-71); the training loss blends ``rand_like(pred) (1 - accumulation)`` into the
    prediction (:197-211); evaluation metrics are taken on channel 3 (:250-296)."""

    config: ConcatNerfModelConfig

    def populate_modules(self) -> None:
        super().populate_modules()
        from .surface import RGBTRenderer

        self.renderer_rgb = RGBTRenderer()

    def _concat_metrics_and_images(self, outputs, batch, threshold=None):
        dev = self.device
        gt = batch["image"].to(dev)
        pred = outputs["rgb"]

        def gray(t: Tensor) -> Tensor:
            return t.repeat(1, 1, 3) if t.shape[-1] == 1 else t

        depth = outputs["depth"]
        near, far = float(depth.min()), float(depth.max())
        images = {"img": torch.cat([gt, pred], dim=1), "accumulation": gray(outputs["accumulation"]),
                  "depth": gray(torch.clip((depth - near) / (far - near + 1e-10), 0, 1))}
        for i in range(self.config.num_proposal_iterations):
            pd = outputs[f"prop_depth_{i}"]
            images[f"prop_depth_{i}"] = gray(torch.clip((pd - float(pd.min())) / (float(pd.max() - pd.min()) + 1e-10), 0, 1))
        g4 = torch.moveaxis(gt, -1, 0)[None][:, 3, :, :].unsqueeze(0)      # concat_nerfacto_model.py:270-271
        p4 = torch.moveaxis(pred, -1, 0)[None][:, 3, :, :].unsqueeze(0)
        metrics = {"psnr": float(self.psnr(g4, p4)), "ssim": float(self.ssim(g4, p4)),
                   "lpips": self._lpips(torch.repeat_interleave(g4, 3, dim=1), torch.repeat_interleave(p4, 3, dim=1)),
                   "mae_thermal_foreground": float(self.mae_thermal(g4, p4, threshold=threshold)),
                   "mae_thermal": float(self.mae_thermal(g4, p4, threshold=None))}
        return metrics, images
