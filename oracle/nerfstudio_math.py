"""Functional restatement of the nerfstudio-1.1.5 torch kernels on the hot path.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Plain PyTorch, runs on CPU.
Each function names the nerfstudio-1.1.5 module it restates (third-party, not
under /root/reference) and the reference call site that reaches it.
"""

from __future__ import annotations

from typing import Callable, Optional

import numpy as np
import torch
from torch import Tensor

__all__ = [
    "hash_scalings",
    "hash_indices",
    "hash_encode",
    "contract_linf",
    "normalise_positions",
    "sh4",
    "trunc_exp",
    "spacing_fn",
    "spacing_fn_inv",
    "make_spacing_to_euclid",
    "piecewise_initial_bins",
    "pdf_resample_bins",
    "get_weights",
    "render_rgb_last_sample",
    "render_rgbt_no_background",
    "render_accumulation",
    "render_depth_median",
    "render_depth_expected",
    "ray_samples_to_sdist_from_bins",
    "lossfun_outer",
    "interlevel_loss",
    "distortion_loss",
    "exp_map_so3xr3",
    "mae_thermal",
    "HASH_PRIMES",
]

HASH_PRIMES = (1, 2654435761, 805459861)


# --------------------------------------------------------------------------
# A.4 hash encoding (nerfstudio.field_components.encodings.HashEncoding, torch
# implementation).  Reached from thermal_field.py:62-88 (NerfactoField ctor)
# and thermal_nerf_model.py:127-148 (HashMLPDensityField).
# --------------------------------------------------------------------------
def hash_scalings(num_levels: int, min_res: int, max_res: int) -> Tensor:
    """``floor(min_res * growth**levels)`` evaluated exactly like nerfstudio.

    ``growth`` is a numpy float64 scalar and ``levels`` an int64 tensor; torch
    evaluates ``scalar ** int64_tensor`` in the default dtype (float32), so the
    top level of a 16..2048 grid comes out as 2047, not 2048.
    """
    levels = torch.arange(num_levels)
    growth = np.exp((np.log(max_res) - np.log(min_res)) / (num_levels - 1)) if num_levels > 1 else 1
    return torch.floor(min_res * growth**levels).to(torch.float32)


def hash_indices(coords_i32: Tensor, log2_hashmap_size: int, num_levels: int) -> Tensor:
    """``HashEncoding.hash_fn``: int64 products, xor, mod 2^T, + level offset.

    coords_i32: [..., L, 3] int32  ->  [..., L] int64
    """
    primes = torch.tensor(HASH_PRIMES, dtype=torch.int64, device=coords_i32.device)
    v = coords_i32 * primes  # int32 * int64 -> int64
    x = torch.bitwise_xor(v[..., 0], v[..., 1])
    x = torch.bitwise_xor(x, v[..., 2])
    x = x % (2**log2_hashmap_size)
    x = x + (torch.arange(num_levels, device=coords_i32.device) * (2**log2_hashmap_size))
    return x


def hash_encode(x: Tensor, table: Tensor, scalings: Tensor, log2_hashmap_size: int) -> Tensor:
    """``HashEncoding.pytorch_fwd``: x [..., 3] in [0,1] -> [..., L*F] level-major."""
    assert x.shape[-1] == 3
    L = scalings.shape[0]
    x = x[..., None, :]
    scaled = x * scalings.view(-1, 1).to(x.device)  # [..., L, 3]
    scaled_c = torch.ceil(scaled).type(torch.int32)
    scaled_f = torch.floor(scaled).type(torch.int32)
    offset = scaled - scaled_f

    def h(cx: Tensor, cy: Tensor, cz: Tensor) -> Tensor:
        return hash_indices(torch.cat([cx, cy, cz], dim=-1), log2_hashmap_size, L)

    cx, cy, cz = scaled_c[..., 0:1], scaled_c[..., 1:2], scaled_c[..., 2:3]
    fx, fy, fz = scaled_f[..., 0:1], scaled_f[..., 1:2], scaled_f[..., 2:3]
    f_0 = table[h(cx, cy, cz)]
    f_1 = table[h(cx, fy, cz)]
    f_2 = table[h(fx, fy, cz)]
    f_3 = table[h(fx, cy, cz)]
    f_4 = table[h(cx, cy, fz)]
    f_5 = table[h(cx, fy, fz)]
    f_6 = table[h(fx, fy, fz)]
    f_7 = table[h(fx, cy, fz)]

    ox, oy, oz = offset[..., 0:1], offset[..., 1:2], offset[..., 2:3]
    f_03 = f_0 * ox + f_3 * (1 - ox)
    f_12 = f_1 * ox + f_2 * (1 - ox)
    f_56 = f_5 * ox + f_6 * (1 - ox)
    f_47 = f_4 * ox + f_7 * (1 - ox)
    f0312 = f_03 * oy + f_12 * (1 - oy)
    f4756 = f_47 * oy + f_56 * (1 - oy)
    enc = f0312 * oz + f4756 * (1 - oz)  # [..., L, F]
    return torch.flatten(enc, start_dim=-2, end_dim=-1)


# --------------------------------------------------------------------------
# A.5 field helpers
# --------------------------------------------------------------------------
def contract_linf(x: Tensor) -> Tensor:
    """``SceneContraction(order=inf)`` (thermal_nerf_model.py:94)."""
    mag = torch.linalg.norm(x, ord=float("inf"), dim=-1)[..., None]
    return torch.where(mag < 1, x, (2 - (1 / mag)) * (x / mag))


def normalise_positions(positions: Tensor, aabb: Tensor, use_contraction: bool):
    """Shared head of ``NerfactoField.get_density`` / ``HashMLPDensityField.get_density``.

    Returns (positions in [0,1] with out-of-range points zeroed, selector bool).
    """
    if use_contraction:
        p = contract_linf(positions)
        p = (p + 2.0) / 4.0
    else:
        lengths = aabb[1] - aabb[0]
        p = (positions - aabb[0]) / lengths
    selector = ((p > 0.0) & (p < 1.0)).all(dim=-1)
    p = p * selector[..., None]
    return p, selector


_SH_C = (
    0.28209479177387814,
    0.4886025119029199,
    1.0925484305920792,
    0.9461746957575601,
    0.31539156525251999,
    0.5462742152960396,
    0.5900435899266435,
    2.890611442640554,
    0.4570457994644658,
    0.3731763325901154,
    1.445305721320277,
)


def sh4(directions: Tensor) -> Tensor:
    """``components_from_spherical_harmonics(degree=4, directions)`` -> [..., 16].

    Evaluated directly on whatever it is given; the caller
    (thermal_field.py:117-119) passes ``(d+1)/2``.
    """
    x, y, z = directions[..., 0], directions[..., 1], directions[..., 2]
    xx, yy, zz = x**2, y**2, z**2
    c = torch.zeros((*directions.shape[:-1], 16), dtype=directions.dtype, device=directions.device)
    c[..., 0] = _SH_C[0]
    c[..., 1] = _SH_C[1] * y
    c[..., 2] = _SH_C[1] * z
    c[..., 3] = _SH_C[1] * x
    c[..., 4] = _SH_C[2] * x * y
    c[..., 5] = _SH_C[2] * y * z
    c[..., 6] = _SH_C[3] * zz - _SH_C[4]
    c[..., 7] = _SH_C[2] * x * z
    c[..., 8] = _SH_C[5] * (xx - yy)
    c[..., 9] = _SH_C[6] * y * (3 * xx - yy)
    c[..., 10] = _SH_C[7] * x * y * z
    c[..., 11] = _SH_C[8] * y * (5 * zz - 1)
    c[..., 12] = _SH_C[9] * z * (5 * zz - 3)
    c[..., 13] = _SH_C[8] * x * (5 * zz - 1)
    c[..., 14] = _SH_C[10] * z * (xx - yy)
    c[..., 15] = _SH_C[6] * x * (xx - 3 * yy)
    return c


class _TruncExp(torch.autograd.Function):
    """``nerfstudio.field_components.activations.trunc_exp``."""

    @staticmethod
    def forward(ctx, x):  # type: ignore[override]
        x = x.float()
        ctx.save_for_backward(x)
        return torch.exp(x)

    @staticmethod
    def backward(ctx, g):  # type: ignore[override]
        (x,) = ctx.saved_tensors
        return g * torch.exp(x.clamp(-15, 15))


trunc_exp: Callable[[Tensor], Tensor] = _TruncExp.apply


# --------------------------------------------------------------------------
# A.3 sampler (nerfstudio.model_components.ray_samplers), reached from
# thermal_nerf_model.py:172-179 and :222-224.
# --------------------------------------------------------------------------
def spacing_fn(x: Tensor) -> Tensor:
    """UniformLinDispPiecewiseSampler spacing function."""
    return torch.where(x < 1, x / 2, 1 - 1 / (2 * x))


def spacing_fn_inv(x: Tensor) -> Tensor:
    return torch.where(x < 0.5, 2 * x, 1 / (2 - 2 * x))


def make_spacing_to_euclid(nears: Tensor, fars: Tensor) -> Callable[[Tensor], Tensor]:
    s_near, s_far = spacing_fn(nears), spacing_fn(fars)

    def to_euclid(x: Tensor) -> Tensor:
        return spacing_fn_inv(x * s_far + (1 - x) * s_near)

    return to_euclid


def piecewise_initial_bins(num_rays: int, num_samples: int, t_rand: Optional[Tensor], device=None) -> Tensor:
    """``SpacedSampler.generate_ray_samples`` spacing bins [R, S+1].

    ``t_rand`` is the single-jitter draw ``torch.rand((R, 1))`` of the
    stratified training path; ``None`` = eval (no jitter).
    """
    bins = torch.linspace(0.0, 1.0, num_samples + 1, device=device)[None, ...]
    if t_rand is not None:
        bin_centers = (bins[..., 1:] + bins[..., :-1]) / 2.0
        bin_upper = torch.cat([bin_centers, bins[..., -1:]], -1)
        bin_lower = torch.cat([bins[..., :1], bin_centers], -1)
        bins = bin_lower + (bin_upper - bin_lower) * t_rand
    return bins.expand(num_rays, num_samples + 1)


def pdf_resample_bins(
    weights: Tensor,
    existing_bins: Tensor,
    num_samples: int,
    t_rand: Optional[Tensor],
    histogram_padding: float = 0.01,
    eps: float = 1e-5,
) -> Tensor:
    """``PDFSampler.generate_ray_samples`` (include_original=False).

    weights [R, S_prev] (already annealed), existing_bins [R, S_prev+1] in
    spacing units -> new spacing bins [R, num_samples+1] (detached).
    ``t_rand``: ``torch.rand((R,1))`` single-jitter draw, ``None`` in eval.
    """
    num_bins = num_samples + 1
    weights = weights + histogram_padding
    weights_sum = torch.sum(weights, dim=-1, keepdim=True)
    padding = torch.relu(eps - weights_sum)
    weights = weights + padding / weights.shape[-1]
    weights_sum = weights_sum + padding
    pdf = weights / weights_sum
    cdf = torch.min(torch.ones_like(pdf), torch.cumsum(pdf, dim=-1))
    cdf = torch.cat([torch.zeros_like(cdf[..., :1]), cdf], dim=-1)

    u = torch.linspace(0.0, 1.0 - (1.0 / num_bins), steps=num_bins, device=cdf.device)
    if t_rand is not None:
        u = u.expand(size=(*cdf.shape[:-1], num_bins))
        u = u + t_rand / num_bins
    else:
        u = u + 1.0 / (2 * num_bins)
        u = u.expand(size=(*cdf.shape[:-1], num_bins))
    u = u.contiguous()

    inds = torch.searchsorted(cdf, u, side="right")
    below = torch.clamp(inds - 1, 0, existing_bins.shape[-1] - 1)
    above = torch.clamp(inds, 0, existing_bins.shape[-1] - 1)
    cdf_g0 = torch.gather(cdf, -1, below)
    bins_g0 = torch.gather(existing_bins, -1, below)
    cdf_g1 = torch.gather(cdf, -1, above)
    bins_g1 = torch.gather(existing_bins, -1, above)
    t = torch.clip(torch.nan_to_num((u - cdf_g0) / (cdf_g1 - cdf_g0), 0), 0, 1)
    bins = bins_g0 + t * (bins_g1 - bins_g0)
    return bins.detach()


# --------------------------------------------------------------------------
# A.6 compositing (nerfstudio.cameras.rays.RaySamples.get_weights and
# nerfstudio.model_components.renderers), reached from
# thermal_nerf_model.py:233-243,267-273.
# --------------------------------------------------------------------------
def get_weights(deltas: Tensor, densities: Tensor) -> Tensor:
    """deltas, densities [R,S,1] -> weights [R,S,1]."""
    delta_density = deltas * densities
    alphas = 1 - torch.exp(-delta_density)
    transmittance = torch.cumsum(delta_density[..., :-1, :], dim=-2)
    transmittance = torch.cat(
        [torch.zeros((*transmittance.shape[:1], 1, 1), device=densities.device), transmittance], dim=-2
    )
    transmittance = torch.exp(-transmittance)
    weights = alphas * transmittance
    return torch.nan_to_num(weights)


def render_rgb_last_sample(values: Tensor, weights: Tensor, training: bool) -> Tensor:
    """``RGBRenderer(background_color="last_sample")`` and, with C=1,
    ``ThermalRenderer`` (thermal_renderer.py:49,55-56,68-70,79,136-147)."""
    if not training:
        values = torch.nan_to_num(values)
    comp = torch.sum(weights * values, dim=-2)
    acc = torch.sum(weights, dim=-2)
    comp = comp + values[..., -1, :] * (1.0 - acc)
    if not training:
        comp = torch.clamp(comp, min=0.0, max=1.0)
    return comp


def render_rgbt_no_background(values: Tensor, weights: Tensor, training: bool) -> Tensor:
    """RGBTRenderer with background_color="random" (reference rgb_concat/rgbt_renderer.py:63-71 and :163-174):
    the weighted sum is returned as is ("as if the background was black"); eval adds nan_to_num / clamp."""
    if not training:
        values = torch.nan_to_num(values)
    comp = torch.sum(weights * values, dim=-2)
    if not training:
        comp = torch.clamp(comp, min=0.0, max=1.0)
    return comp


def render_accumulation(weights: Tensor) -> Tensor:
    return torch.sum(weights, dim=-2)


def render_depth_median(weights: Tensor, starts: Tensor, ends: Tensor) -> Tensor:
    steps = (starts + ends) / 2
    cumulative_weights = torch.cumsum(weights[..., 0], dim=-1)
    split = torch.ones((*weights.shape[:-2], 1), device=weights.device) * 0.5
    median_index = torch.searchsorted(cumulative_weights, split, side="left")
    median_index = torch.clamp(median_index, 0, steps.shape[-2] - 1)
    return torch.gather(steps[..., 0], dim=-1, index=median_index)


def render_depth_expected(weights: Tensor, starts: Tensor, ends: Tensor) -> Tensor:
    eps = 1e-10
    steps = (starts + ends) / 2
    depth = torch.sum(weights * steps, dim=-2) / (torch.sum(weights, -2) + eps)
    # NOTE: tensor-global clip (chunk dependent!), SURVEY A.6.
    return torch.clip(depth, steps.min(), steps.max())


# --------------------------------------------------------------------------
# A.7 losses (nerfstudio.model_components.losses), reached from
# thermal_nerf_model.py:298-305 and the inherited get_metrics_dict.
# --------------------------------------------------------------------------
_LOSS_EPS = 1.0e-7


def ray_samples_to_sdist_from_bins(spacing_bins: Tensor) -> Tensor:
    """We carry spacing bins [R,S+1] directly; nerfstudio rebuilds them from
    ``spacing_starts`` / ``spacing_ends`` - identical values."""
    return spacing_bins


def _outer(t0_starts, t0_ends, t1_starts, t1_ends, y1):
    cy1 = torch.cat([torch.zeros_like(y1[..., :1]), torch.cumsum(y1, dim=-1)], dim=-1)
    idx_lo = torch.searchsorted(t1_starts.contiguous(), t0_starts.contiguous(), side="right") - 1
    idx_lo = torch.clamp(idx_lo, min=0, max=y1.shape[-1] - 1)
    idx_hi = torch.searchsorted(t1_ends.contiguous(), t0_ends.contiguous(), side="right")
    idx_hi = torch.clamp(idx_hi, min=0, max=y1.shape[-1] - 1)
    cy1_lo = torch.take_along_dim(cy1[..., :-1], idx_lo, dim=-1)
    cy1_hi = torch.take_along_dim(cy1[..., 1:], idx_hi, dim=-1)
    return cy1_hi - cy1_lo


def lossfun_outer(t, w, t_env, w_env):
    w_outer = _outer(t[..., :-1], t[..., 1:], t_env[..., :-1], t_env[..., 1:], w_env)
    return torch.clip(w - w_outer, min=0) ** 2 / (w + _LOSS_EPS)


def interlevel_loss(weights_list, sdist_list) -> Tensor:
    """weights_list[k] [R,S_k,1]; sdist_list[k] [R,S_k+1]."""
    c = sdist_list[-1].detach()
    w = weights_list[-1][..., 0].detach()
    loss = 0.0
    for sdist, weights in zip(sdist_list[:-1], weights_list[:-1]):
        loss = loss + torch.mean(lossfun_outer(c, w, sdist, weights[..., 0]))
    return loss  # type: ignore[return-value]


def distortion_loss(weights_list, sdist_list) -> Tensor:
    t = sdist_list[-1]
    w = weights_list[-1][..., 0]
    ut = (t[..., 1:] + t[..., :-1]) / 2
    dut = torch.abs(ut[..., :, None] - ut[..., None, :])
    loss_inter = torch.sum(w * torch.sum(w[..., None, :] * dut, dim=-1), dim=-1)
    loss_intra = torch.sum(w**2 * (t[..., 1:] - t[..., :-1]), dim=-1) / 3
    return torch.mean(loss_inter + loss_intra)


# --------------------------------------------------------------------------
# a2: camera optimizer (nerfstudio.cameras.camera_optimizers, lie_groups),
# reached from thermal_nerf_model.py:218-219 and evaluator.py:71-73.
# --------------------------------------------------------------------------
def exp_map_so3xr3(tangent: Tensor) -> Tensor:
    """[N,6] (translation | log-rotation) -> [N,3,4]."""
    log_rot = tangent[:, 3:]
    nrms = (log_rot * log_rot).sum(1)
    rot_angles = torch.clamp(nrms, 1e-4).sqrt()
    rot_angles_inv = 1.0 / rot_angles
    fac1 = rot_angles_inv * rot_angles.sin()
    fac2 = rot_angles_inv * rot_angles_inv * (1.0 - rot_angles.cos())
    skews = torch.zeros((log_rot.shape[0], 3, 3), dtype=log_rot.dtype, device=log_rot.device)
    skews[:, 0, 1] = -log_rot[:, 2]
    skews[:, 0, 2] = log_rot[:, 1]
    skews[:, 1, 0] = log_rot[:, 2]
    skews[:, 1, 2] = -log_rot[:, 0]
    skews[:, 2, 0] = -log_rot[:, 1]
    skews[:, 2, 1] = log_rot[:, 0]
    skews_square = torch.bmm(skews, skews)
    ret = torch.zeros(tangent.shape[0], 3, 4, dtype=tangent.dtype, device=tangent.device)
    ret[:, :3, :3] = (
        fac1[:, None, None] * skews
        + fac2[:, None, None] * skews_square
        + torch.eye(3, dtype=log_rot.dtype, device=log_rot.device)[None]
    )
    ret[:, :3, 3] = tangent[:, :3]
    return ret


def mae_thermal(gt, pred, cold_flag, max_temperature, min_temperature, threshold=None) -> Tensor:
    """thermal_metrics.py:5-34 (pure torch in the reference; restated)."""
    if threshold:
        idx = torch.where(gt < threshold) if cold_flag else torch.where(gt > threshold)
        gt, pred = gt[idx], pred[idx]
    span = max_temperature - min_temperature
    return torch.mean(torch.abs((gt * span + min_temperature) - (pred * span + min_temperature)))
