"""Generates the committed golden vectors under tests/golden/ from the oracle.

PARITY UNPINNED: the reference's tests hold no numerical fixture for this path and
nerfstudio cannot be imported here, so these vectors pin the *oracle* (seeded
weights -> outputs) rather than a run of the reference.  They guard against drift of the
oracle itself across torch versions / hosts and give the GPU tests a fixed target.

    python tests/golden/make_golden.py

Weights are not stored: they regenerate from the seed (torch CPU RNG), and each fixture
carries a checksum of the regenerated weights so a platform whose RNG differs is detected
rather than silently compared.
"""

from __future__ import annotations

import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

from oracle import make_synthetic_rays  # noqa: E402
from tests.helpers import make_trained_like, oracle_config  # noqa: E402
from oracle import OracleThermalNerf  # noqa: E402

OUT = Path(__file__).resolve().parent

CASES = {
    # name: (R, num_samples, log2_field, log2_prop, num_images, trained_like, training)
    "e2e_mini_r32": (32, (16, 8, 16), 10, 10, 4, True, False),
    "e2e_eval_r256": (256, (256, 96, 48), 14, 12, 8, True, False),
    "e2e_eval_init_r128": (128, (256, 96, 48), 14, 12, 8, False, False),
    "e2e_train_r128": (128, (256, 96, 48), 14, 12, 8, True, True),
}


def weights_checksum(model: torch.nn.Module) -> float:
    return float(sum(p.detach().double().abs().sum() for p in model.parameters()))


def build_case(name: str):
    R, ns, lf, lp, nimg, trained, training = CASES[name]
    cfg = oracle_config(lf, lp, ns)
    model = OracleThermalNerf(cfg, nimg, seed=0)
    if trained:
        make_trained_like(model, 0)
    rays = make_synthetic_rays(R, num_images=nimg, seed=11)
    jitter = None
    if training:
        jitter = torch.rand((3, R, 1), generator=torch.Generator().manual_seed(5))
        model.set_anneal_for_step(300)
    return model, rays, jitter, training


def run_case(name: str):
    model, rays, jitter, training = build_case(name)
    with torch.no_grad():
        out = model.get_outputs(rays, training=training, jitter=jitter)
    keep = {k: v.clone() for k, v in out.items() if isinstance(v, torch.Tensor) and not k.startswith("field_")}
    keep["weights_list"] = [w.clone() for w in out["weights_list"]]
    keep["sdist_list"] = [s.clone() for s in out["sdist_list"]]
    return {
        "case": name,
        "params": CASES[name],
        "weights_checksum": weights_checksum(model),
        "origins": rays.origins, "directions": rays.directions, "camera_indices": rays.camera_indices,
        "jitter": jitter,
        "anneal": model.anneal,
        "outputs": keep,
        "torch_version": str(torch.__version__),
    }


def main() -> None:
    for name in CASES:
        blob = run_case(name)
        # fp16 storage would lose the parity margin; keep fp32 but only the small per-ray tensors + samples
        torch.save(blob, OUT / f"{name}.pt")
        print(name, {k: tuple(v.shape) for k, v in blob["outputs"].items() if isinstance(v, torch.Tensor)})


if __name__ == "__main__":
    main()
