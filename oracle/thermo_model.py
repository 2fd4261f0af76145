"""Oracle model: ThermoNeRF ``get_outputs`` / losses restated on plain PyTorch.

TEST INFRASTRUCTURE (see oracle/__init__.py).  The module tree mirrors the
nerfstudio/ThermoNeRF attribute names so ``state_dict()`` keys match the ones a
nerfstudio-trained ThermoNeRF checkpoint carries after its ``_model.`` prefix
(names recalled, not verifiable offline).

Follows, line by line:
* thermo_nerf/thermal_nerf/thermal_nerf_model.py:86-208  populate_modules
* thermo_nerf/thermal_nerf/thermal_nerf_model.py:210-275 get_outputs
* thermo_nerf/thermal_nerf/thermal_nerf_model.py:277-326 get_loss_dict
* thermo_nerf/thermal_nerf/thermal_field.py:108-201      field get_outputs / forward
* thermo_nerf/thermal_nerf/thermal_renderer.py:113-149   thermal renderer
"""

from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import torch
from torch import Tensor, nn

from . import nerfstudio_math as M

__all__ = ["OracleConfig", "OracleRays", "OracleThermalNerf", "make_synthetic_rays"]


@dataclass
class OracleConfig:
    """NerfactoModelConfig defaults inherited by ThermalNerfModelConfig
    (thermal_nerf_model.py:46-56)."""

    near_plane: float = 0.05
    far_plane: float = 1000.0
    hidden_dim: int = 64
    hidden_dim_color: int = 64
    hidden_dim_transient: int = 64  # width of mlp_thermal's output (thermal_field.py:94)
    geo_feat_dim: int = 15
    num_levels: int = 16
    base_res: int = 16
    max_res: int = 2048
    log2_hashmap_size: int = 19
    features_per_level: int = 2
    appearance_embed_dim: int = 32
    num_proposal_samples_per_ray: Tuple[int, ...] = (256, 96)
    num_nerf_samples_per_ray: int = 48
    proposal_net_args_list: List[dict] = field(
        default_factory=lambda: [
            {"hidden_dim": 16, "log2_hashmap_size": 17, "num_levels": 5, "max_res": 128},
            {"hidden_dim": 16, "log2_hashmap_size": 17, "num_levels": 5, "max_res": 256},
        ]
    )
    proposal_weights_anneal_slope: float = 10.0
    proposal_weights_anneal_max_num_iters: int = 1000
    interlevel_loss_mult: float = 1.0
    distortion_loss_mult: float = 0.002
    disable_scene_contraction: bool = False
    pass_thermal_gradients: bool = True
    camera_optimizer_mode: str = "SO3xR3"
    # ---- switches for details SURVEY Appendix A marks "recalled" ----
    use_average_appearance_embedding: bool = True
    reset_near_plane_at_eval: bool = True  # NearFarCollider: eval renders from t=0 (A.2)
    sh_on_unit_remapped_dirs: bool = True  # SH evaluated directly on (d+1)/2 (A.5)
    average_init_density: float = 1.0  # thermal_field.py:86 passes 1.0 positionally
    # "thermal": ThermalNerfModel (RGB head + temperature head).  "concat": ConcatNerfModel
    # (rgb_concat/concat_nerfacto_model.py:60-197): one 4-channel RGBT colour head
    # (concat_field.py:65-75), no temperature head, RGBTRenderer() with its default "random" background.
    head: str = "thermal"


@dataclass
class OracleRays:
    """Duck-typed stand-in for nerfstudio ``RayBundle`` (flat [R] batch)."""

    origins: Tensor  # [R,3]
    directions: Tensor  # [R,3]
    camera_indices: Tensor  # [R,1] int64
    pixel_area: Optional[Tensor] = None
    nears: Optional[Tensor] = None
    fars: Optional[Tensor] = None

    def __len__(self) -> int:
        return self.origins.shape[0]


class _HashEncoding(nn.Module):
    def __init__(self, num_levels, min_res, max_res, log2_hashmap_size, features_per_level=2, hash_init_scale=1e-3):
        super().__init__()
        self.num_levels = num_levels
        self.log2_hashmap_size = log2_hashmap_size
        self.features_per_level = features_per_level
        self.register_buffer("scalings", M.hash_scalings(num_levels, min_res, max_res))
        table = torch.rand(size=(2**log2_hashmap_size * num_levels, features_per_level)) * 2 - 1
        self.hash_table = nn.Parameter(table * hash_init_scale)

    def forward(self, x: Tensor) -> Tensor:
        return M.hash_encode(x, self.hash_table, self.scalings, self.log2_hashmap_size)


class _MLP(nn.Module):
    """nerfstudio ``MLP`` torch implementation: Linear(+bias) stack, ReLU between."""

    def __init__(self, in_dim, num_layers, layer_width, out_dim, out_activation=None):
        super().__init__()
        dims = [in_dim] + [layer_width] * (num_layers - 1) + [out_dim]
        self.layers = nn.ModuleList([nn.Linear(dims[i], dims[i + 1]) for i in range(num_layers)])
        self.out_activation = out_activation

    def forward(self, x: Tensor) -> Tensor:
        for i, layer in enumerate(self.layers):
            x = layer(x)
            if i < len(self.layers) - 1:
                x = torch.relu(x)
        if self.out_activation is not None:
            x = self.out_activation(x)
        return x


class _MLPWithHashEncoding(nn.Module):
    def __init__(self, num_levels, min_res, max_res, log2_hashmap_size, num_layers, layer_width, out_dim):
        super().__init__()
        self.encoder = _HashEncoding(num_levels, min_res, max_res, log2_hashmap_size)
        self.mlp = _MLP(num_levels * 2, num_layers, layer_width, out_dim)

    def forward(self, x: Tensor) -> Tensor:
        return self.mlp(self.encoder(x))


class _Embedding(nn.Module):
    def __init__(self, n, d):
        super().__init__()
        self.embedding = nn.Embedding(n, d)

    def forward(self, idx: Tensor) -> Tensor:
        return self.embedding(idx)

    def mean(self, dim=0):
        return self.embedding.weight.mean(dim)


class _ThermalHead(nn.Module):
    """thermal_field_head.py:15-71 with activation=None (thermal_field.py:18-30)."""

    def __init__(self, in_dim):
        super().__init__()
        self.net = nn.Linear(in_dim, 1)

    def forward(self, x):
        return self.net(x)


class _HashMLPDensityField(nn.Module):
    """nerfstudio ``HashMLPDensityField`` (use_linear=False), built at
    thermal_nerf_model.py:127-148."""

    def __init__(self, cfg: OracleConfig, aabb: Tensor, hidden_dim, log2_hashmap_size, num_levels, max_res, base_res=16):
        super().__init__()
        self.cfg = cfg
        self.register_buffer("aabb", aabb)
        self.encoding = _HashEncoding(num_levels, base_res, max_res, log2_hashmap_size)
        network = _MLP(num_levels * 2, 2, hidden_dim, 1)
        self.mlp_base = nn.Sequential(self.encoding, network)

    def density_fn(self, positions: Tensor) -> Tensor:
        p, selector = M.normalise_positions(positions, self.aabb, not self.cfg.disable_scene_contraction)
        h = self.mlp_base(p.view(-1, 3)).view(*positions.shape[:-1], -1).to(p)
        density = 1.0 * M.trunc_exp(h)  # HashMLPDensityField default average_init_density=1.0
        return density * selector[..., None]


class _ThermalField(nn.Module):
    """ThermalNerfactoTField (thermal_field.py:33-201) over NerfactoField."""

    def __init__(self, cfg: OracleConfig, aabb: Tensor, num_images: int):
        super().__init__()
        self.cfg = cfg
        self.register_buffer("aabb", aabb)
        self.mlp_base = _MLPWithHashEncoding(
            cfg.num_levels, cfg.base_res, cfg.max_res, cfg.log2_hashmap_size, 2, cfg.hidden_dim, 1 + cfg.geo_feat_dim
        )
        self.embedding_appearance = _Embedding(num_images, cfg.appearance_embed_dim)
        assert cfg.head in ("thermal", "concat"), cfg.head
        self.mlp_head = _MLP(
            16 + cfg.geo_feat_dim + cfg.appearance_embed_dim, 3, cfg.hidden_dim_color,
            4 if cfg.head == "concat" else 3, out_activation=torch.sigmoid
        )
        if cfg.head == "thermal":
            self.mlp_thermal = _MLP(cfg.geo_feat_dim, 2, 64, cfg.hidden_dim_transient, out_activation=torch.sigmoid)
            self.field_head_thermal = _ThermalHead(cfg.hidden_dim_transient)
        self.pass_thermal_gradients = cfg.pass_thermal_gradients and cfg.head == "thermal"
        self.pass_rgb_gradients = True

    def get_density(self, positions: Tensor):
        p, selector = M.normalise_positions(positions, self.aabb, not self.cfg.disable_scene_contraction)
        h = self.mlp_base(p.view(-1, 3)).view(*positions.shape[:-1], -1)
        dba, geo = torch.split(h, [1, self.cfg.geo_feat_dim], dim=-1)
        density = self.cfg.average_init_density * M.trunc_exp(dba.to(p))
        return density * selector[..., None], geo

    def get_outputs(self, directions: Tensor, camera_indices: Tensor, geo: Tensor, training: bool):
        """directions [R,S,3] (expanded), camera_indices [R,S] -> rgb [R,S,3], thermal [R,S,1]."""
        cfg = self.cfg
        dn = (directions + 1.0) / 2.0  # get_normalized_directions
        with torch.no_grad():
            d = M.sh4(dn.view(-1, 3) if cfg.sh_on_unit_remapped_dirs else directions.reshape(-1, 3))
        shape = directions.shape[:-1]
        if training:
            app = self.embedding_appearance(camera_indices)
        elif cfg.use_average_appearance_embedding:
            app = torch.ones((*shape, cfg.appearance_embed_dim), device=directions.device) * self.embedding_appearance.mean(dim=0)
        else:
            app = torch.zeros((*shape, cfg.appearance_embed_dim), device=directions.device)
        h = torch.cat([d, geo.reshape(-1, cfg.geo_feat_dim), app.reshape(-1, cfg.appearance_embed_dim)], dim=-1)
        rgb = self.mlp_head(h).view(*shape, -1)
        if cfg.head == "concat":
            return rgb, None  # [R,S,4]: temperature is the fourth colour channel
        th_in = geo.reshape(-1, cfg.geo_feat_dim)
        if not self.pass_thermal_gradients:
            th_in = th_in.detach()
        thermal = self.field_head_thermal(self.mlp_thermal(th_in).view(*shape, -1))
        return rgb, thermal


class _CameraOptimizer(nn.Module):
    def __init__(self, num_cameras: int, mode: str):
        super().__init__()
        self.mode = mode
        self.pose_adjustment = nn.Parameter(torch.zeros((num_cameras, 6)))

    def apply_to_raybundle(self, rays: OracleRays) -> None:
        if self.mode == "off":
            return
        assert self.mode == "SO3xR3"
        c = M.exp_map_so3xr3(self.pose_adjustment[rays.camera_indices.squeeze(-1), :])
        rays.origins = rays.origins + c[:, :3, 3]
        rays.directions = torch.bmm(c[:, :3, :3], rays.directions[..., None]).squeeze(-1)


class OracleThermalNerf(nn.Module):
    """ThermalNerfModel restated (no nerfstudio import)."""

    def __init__(self, cfg: OracleConfig, num_train_data: int, aabb: Optional[Tensor] = None, seed: Optional[int] = 0):
        super().__init__()
        if seed is not None:
            torch.manual_seed(seed)
        self.cfg = cfg
        if aabb is None:
            aabb = torch.tensor([[-1.0, -1.0, -1.0], [1.0, 1.0, 1.0]])
        self.field = _ThermalField(cfg, aabb, num_train_data)
        self.camera_optimizer = _CameraOptimizer(num_train_data, cfg.camera_optimizer_mode)
        self.proposal_networks = nn.ModuleList()
        for i in range(len(cfg.num_proposal_samples_per_ray)):
            a = cfg.proposal_net_args_list[min(i, len(cfg.proposal_net_args_list) - 1)]
            self.proposal_networks.append(
                _HashMLPDensityField(cfg, aabb, a["hidden_dim"], a["log2_hashmap_size"], a["num_levels"], a["max_res"])
            )
        self.anneal = 1.0  # ProposalNetworkSampler._anneal of a fresh sampler

    # --- training callbacks (nerfacto get_training_callbacks, inherited) ---
    def set_anneal_for_step(self, step: int) -> None:
        n, s = self.cfg.proposal_weights_anneal_max_num_iters, self.cfg.proposal_weights_anneal_slope
        frac = min(max(step / n, 0.0), 1.0)
        self.anneal = s * frac / ((s - 1) * frac + 1)

    # --- get_outputs (thermal_nerf_model.py:210-275) ---
    def get_outputs(
        self,
        rays: OracleRays,
        training: bool = False,
        jitter: Optional[Tensor] = None,
        prop_grad: bool = True,
    ) -> Dict[str, object]:
        """``jitter`` [n_levels+1, R, 1]: the ``torch.rand((R,1))`` draws of the
        stratified samplers (training only; drawn here when None)."""
        cfg = self.cfg
        R = len(rays)
        if training:
            rays = OracleRays(rays.origins, rays.directions, rays.camera_indices)
            self.camera_optimizer.apply_to_raybundle(rays)
        near = cfg.near_plane if (training or not cfg.reset_near_plane_at_eval) else 0.0
        ones = torch.ones_like(rays.origins[..., 0:1])
        nears = rays.nears if rays.nears is not None else ones * near
        fars = rays.fars if rays.fars is not None else ones * cfg.far_plane
        to_euclid = M.make_spacing_to_euclid(nears, fars)
        n_prop = len(cfg.num_proposal_samples_per_ray)
        if training and jitter is None:
            jitter = torch.rand((n_prop + 1, R, 1), device=rays.origins.device)

        weights_list: List[Tensor] = []
        sdist_list: List[Tensor] = []
        eucl_list: List[Tensor] = []
        weights = None
        sbins = None
        for lvl in range(n_prop + 1):
            is_prop = lvl < n_prop
            S = cfg.num_proposal_samples_per_ray[lvl] if is_prop else cfg.num_nerf_samples_per_ray
            tr = jitter[lvl] if training else None
            if lvl == 0:
                sbins = M.piecewise_initial_bins(R, S, tr, device=rays.origins.device)
            else:
                annealed = torch.pow(weights, self.anneal)
                sbins = M.pdf_resample_bins(annealed[..., 0], sbins, S, tr)
            eucl = to_euclid(sbins)  # [R,S+1]
            starts, ends = eucl[..., :-1, None], eucl[..., 1:, None]
            positions = rays.origins[:, None, :] + rays.directions[:, None, :] * (starts + ends) / 2
            if is_prop:
                if prop_grad:
                    density = self.proposal_networks[lvl].density_fn(positions)
                else:
                    with torch.no_grad():
                        density = self.proposal_networks[lvl].density_fn(positions)
                weights = M.get_weights(ends - starts, density)
                weights_list.append(weights)
                sdist_list.append(sbins)
                eucl_list.append(eucl)

        # final level: the field
        density, geo = self.field.get_density(positions)
        dirs = rays.directions[:, None, :].expand(R, S, 3)
        cams = rays.camera_indices.view(R, 1).expand(R, S)
        rgb_s, thermal_s = self.field.get_outputs(dirs, cams, geo, training)
        weights = M.get_weights(ends - starts, density)
        weights_list.append(weights)
        sdist_list.append(sbins)
        eucl_list.append(eucl)

        concat = cfg.head == "concat"
        out: Dict[str, object] = {
            # concat: RGBTRenderer "random" background = the plain weighted sum (rgbt_renderer.py:63-71)
            "rgb": M.render_rgbt_no_background(rgb_s, weights, training) if concat
            else M.render_rgb_last_sample(rgb_s, weights, training),
            "accumulation": M.render_accumulation(weights),
            "expected_depth": M.render_depth_expected(weights, starts, ends),
        }
        with torch.no_grad():
            out["depth"] = M.render_depth_median(weights, starts, ends)
        for i in range(n_prop):
            e = eucl_list[i]
            out[f"prop_depth_{i}"] = M.render_depth_median(weights_list[i], e[..., :-1, None], e[..., 1:, None])
        if not concat:
            out["thermal"] = M.render_rgb_last_sample(thermal_s, weights, training)
        # always exposed by the oracle (the reference only keeps them in training)
        out["weights_list"] = weights_list
        out["sdist_list"] = sdist_list
        out["euclid_list"] = eucl_list
        out["field_density"] = density
        out["field_rgb"] = rgb_s
        out["field_thermal"] = thermal_s
        return out

    # --- get_loss_dict (thermal_nerf_model.py:277-326) + inherited metrics ---
    def get_loss_dict(self, outputs, gt_rgb: Tensor, gt_thermal: Optional[Tensor] = None, training: bool = True,
                      background_noise: Optional[Tensor] = None) -> Dict[str, Tensor]:
        """concat head (concat_nerfacto_model.py:197-233): ``gt_rgb`` is the 4-channel RGBT image; the renderer's
        "random" background adds ``rand_like(pred) * (1 - accumulation)`` to the *prediction only*
        (rgbt_renderer.py:134-140) - ``background_noise`` [R,4] is that draw (drawn here when None)."""
        cfg = self.cfg
        loss: Dict[str, Tensor] = {}
        if cfg.head == "concat":
            pred = outputs["rgb"]
            if background_noise is None:
                background_noise = torch.rand_like(pred)
            pred = pred + background_noise * (1.0 - outputs["accumulation"])
            loss["rgb_loss"] = torch.nn.functional.mse_loss(gt_rgb, pred)
        elif self.field.pass_rgb_gradients:
            loss["rgb_loss"] = torch.nn.functional.mse_loss(gt_rgb, outputs["rgb"])
        if training:
            loss["interlevel_loss"] = cfg.interlevel_loss_mult * M.interlevel_loss(
                outputs["weights_list"], outputs["sdist_list"]
            )
            loss["distortion_loss"] = cfg.distortion_loss_mult * M.distortion_loss(
                outputs["weights_list"], outputs["sdist_list"]
            )
        if self.field.pass_thermal_gradients:
            loss["thermal"] = torch.nn.functional.mse_loss(outputs["thermal"], gt_thermal)
        return loss


def make_synthetic_rays(
    R: int,
    num_images: int = 100,
    seed: int = 0,
    radius: float = 0.8,
    image_hw: int = 800,
    focal: float = 1111.1,
    contiguous_pixels: bool = False,
) -> OracleRays:
    """ThermoScenes-shaped synthetic rays: pinhole cameras
    on a sphere of ``radius`` looking at the origin; unit directions.

    contiguous_pixels=False -> random (image, y, x) triples (training batches);
    True -> a row-major run of pixels of camera 0 (render chunks).
    """
    g = torch.Generator().manual_seed(seed)
    if contiguous_pixels:
        cam = torch.zeros(R, dtype=torch.int64)
        start = int(torch.randint(0, max(image_hw * image_hw - R, 1), (1,), generator=g))
        pix = (torch.arange(R) + start) % (image_hw * image_hw)
        py, px = pix // image_hw, pix % image_hw
    else:
        cam = torch.randint(0, num_images, (R,), generator=g)
        py = torch.randint(0, image_hw, (R,), generator=g)
        px = torch.randint(0, image_hw, (R,), generator=g)
    # camera centres on a Fibonacci sphere
    k = torch.arange(num_images, dtype=torch.float32) + 0.5
    phi = torch.acos(1 - 2 * k / num_images)
    theta = torch.pi * (1 + 5**0.5) * k
    centres = radius * torch.stack([torch.cos(theta) * torch.sin(phi), torch.sin(theta) * torch.sin(phi), torch.cos(phi)], -1)
    fwd = -centres / centres.norm(dim=-1, keepdim=True)
    up = torch.tensor([0.0, 0.0, 1.0]).expand_as(fwd)
    right = torch.linalg.cross(fwd, up)
    right = right / right.norm(dim=-1, keepdim=True).clamp_min(1e-6)
    true_up = torch.linalg.cross(right, fwd)
    x = (px.float() + 0.5 - image_hw / 2) / focal
    y = -(py.float() + 0.5 - image_hw / 2) / focal
    d = x[:, None] * right[cam] + y[:, None] * true_up[cam] + fwd[cam]
    d = d / d.norm(dim=-1, keepdim=True)
    return OracleRays(origins=centres[cam].contiguous(), directions=d.contiguous(), camera_indices=cam[:, None])
