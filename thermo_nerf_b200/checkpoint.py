"""nerfstudio checkpoint envelope <-> ``ThermalNerfModel``.

The reference loads ``<run>/nerfstudio_models/step-*.ckpt`` with ``torch.load`` and hands
``loaded_state["pipeline"]`` to ``pipeline.load_pipeline`` (thermo_nerf/render/renderer.py:93-113); the file is
the trainer's envelope ``{"step", "pipeline", "optimizers", "scalers"}`` whose ``pipeline`` entry is the pipeline's
``state_dict`` - the model's tensors carry the prefix ``_model.`` (``_model.module.`` when saved from DDP), as in the
reference fixture tests/data/vanilla_nerf/.../nerfstudio_models/test_pipeline.ckpt.  The module tree of
``thermo_nerf_b200.model.ThermalNerfModel`` reproduces the nerfstudio ``implementation="torch"`` tree, so a ThermoNeRF
checkpoint trained with the torch implementation loads key for key; tinycudann checkpoints (one opaque ``params`` blob per
network, a different hash function and grid layout) cannot be converted and are rejected with a clear error.
"""

from __future__ import annotations

from pathlib import Path
from typing import Dict, Mapping, Optional, Tuple, Union

import torch
from torch import Tensor

MODEL_PREFIXES = ("_model.module.", "module._model.", "_model.")

# alternative spellings of the same tensors across nerfstudio releases (torch implementation)
KEY_ALIASES = {
    "field.mlp_base.model.0.hash_table": "field.mlp_base.encoder.hash_table",
    "field.mlp_base.model.0.scalings": "field.mlp_base.encoder.scalings",
    "field.mlp_base.model.0.hash_offset": "field.mlp_base.encoder.hash_offset",
    "field.mlp_base.model.1.layers.0.weight": "field.mlp_base.mlp.layers.0.weight",
    "field.mlp_base.model.1.layers.0.bias": "field.mlp_base.mlp.layers.0.bias",
    "field.mlp_base.model.1.layers.1.weight": "field.mlp_base.mlp.layers.1.weight",
    "field.mlp_base.model.1.layers.1.bias": "field.mlp_base.mlp.layers.1.bias",
}


# model tensors a genuine nerfstudio checkpoint carries that are not part of the hot path and that this model does
# not own: the metric networks registered on the nerfstudio Model (the reference fixture
# tests/data/vanilla_nerf/.../test_pipeline.ckpt holds 20 `_model.lpips.net.*` entries), and derived buffers of the
# hash encodings under either spelling of the proposal networks' encoder (`encoding.` and `mlp_base.0.`)
IGNORED_PREFIXES = ("lpips.", "psnr.", "ssim.", "collider.", "renderer_", "normals_shader.")
IGNORED_SUFFIXES = (".hash_offset",)


def _foreign(key: str) -> bool:
    return key.startswith(IGNORED_PREFIXES) or key.endswith(IGNORED_SUFFIXES)


def extract_model_state(checkpoint: Union[str, Path, Mapping]) -> Tuple[Dict[str, Tensor], int]:
    """(model state_dict without the pipeline prefix, step) from a checkpoint path / loaded envelope / bare
    pipeline state_dict."""
    if isinstance(checkpoint, (str, Path)):
        checkpoint = torch.load(checkpoint, map_location="cpu", weights_only=False)
    step = int(checkpoint.get("step", 0)) if "pipeline" in checkpoint else 0
    pipeline_state = checkpoint["pipeline"] if "pipeline" in checkpoint else checkpoint
    out: Dict[str, Tensor] = {}
    for key, value in pipeline_state.items():
        for prefix in MODEL_PREFIXES:
            if key.startswith(prefix):
                out[key[len(prefix):]] = value
                break
    if not out:
        raise KeyError("no '_model.*' entries: not a nerfstudio pipeline checkpoint")
    return out, step


def load_nerfstudio_checkpoint(model: torch.nn.Module, checkpoint: Union[str, Path, Mapping], strict: bool = True) -> int:
    """Load a ThermoNeRF (torch-implementation) nerfstudio checkpoint into ``model``; returns the training step.
    ``strict`` fails on missing or unexpected keys of the modules this model owns (field, proposal networks, camera
    optimiser); tensors of components outside the hot path (``IGNORED_PREFIXES``: LPIPS / PSNR / SSIM metric networks,
    ``*.hash_offset`` derived buffers) are skipped in either mode - every real nerfstudio checkpoint has them."""
    state, step = extract_model_state(checkpoint)
    if any(k.endswith(".params") or ".tcnn_encoding." in k for k in state):
        raise NotImplementedError(
            "this checkpoint was trained with implementation='tcnn' (opaque 'params' blobs, tcnn hash/grid layout); "
            "libtnf_b200 follows the torch implementation's hash encoding and cannot reinterpret those tables")
    state = {KEY_ALIASES.get(k, k): v for k, v in state.items()}
    own = model.state_dict()
    state = {k: v for k, v in state.items() if k in own or not _foreign(k)}
    # nn.Sequential(encoding, mlp) registers the proposal encoder twice: `encoding.*` and `mlp_base.0.*`
    state = {k: v for k, v in state.items()
             if not (k not in own and ".mlp_base.0." in k and k.replace(".mlp_base.0.", ".encoding.") in state)}
    for k, v in state.items():
        if k in own and tuple(own[k].shape) != tuple(v.shape):
            raise ValueError(f"{k}: checkpoint shape {tuple(v.shape)} != model shape {tuple(own[k].shape)} "
                             "(the model must be built with the run's config)")
    missing, unexpected = model.load_state_dict(state, strict=False)
    missing = [k for k in missing if k != "device_indicator_param"]
    if strict and (missing or unexpected):
        raise KeyError(f"checkpoint does not match the model: missing {missing[:8]}, unexpected {list(unexpected)[:8]}")
    if hasattr(model, "_step"):
        model._step = model.step = step
    return step


def save_nerfstudio_checkpoint(model: torch.nn.Module, path: Union[str, Path], step: int,
                               optimizers: Optional[Mapping[str, torch.optim.Optimizer]] = None,
                               scalers: Optional[Mapping] = None) -> None:
    """Write the trainer's envelope (``Trainer.save_checkpoint``): loadable by nerfstudio's ``load_pipeline`` for the
    model part and by :func:`load_nerfstudio_checkpoint`."""
    blob = {
        "step": int(step),
        "pipeline": {"_model." + k: v.detach().cpu() for k, v in model.state_dict().items()},
        "optimizers": {k: o.state_dict() for k, o in (optimizers or {}).items()},
        "scalers": dict(scalers or {}),
    }
    Path(path).parent.mkdir(parents=True, exist_ok=True)
    torch.save(blob, path)
