"""Golden vectors from the REFERENCE'S OWN compositing code, executed in the build container.

The two renderer modules of the reference are plain torch apart from their imports, so they run here once the
absent third-party packages are replaced by empty stand-ins that the executed lines never touch:

* thermo_nerf/thermal_nerf/thermal_renderer.py  - ThermalRenderer.forward (thermal_renderer.py:113-149):
  imports `nerfstudio.utils.colors` (only used by get_background_color, which "last_sample" never reaches);
* thermo_nerf/rgb_concat/rgbt_renderer.py       - RGBTRenderer.forward and blend_background_for_loss_computation
  (row f4): additionally imports `nerfacc` (only used for packed samples).

No arithmetic is stubbed: every number in the output file was produced by the reference's source lines.

    python tests/golden/make_reference_renderer_golden.py        # needs /root/reference

Writes tests/golden/reference_renderers.pt."""

import importlib.util
import sys
import types
from pathlib import Path

import torch

REF = Path("/root/reference/thermo_nerf")
OUT = Path(__file__).resolve().parent / "reference_renderers.pt"


def _stand_ins() -> None:
    ns, utils, colors = types.ModuleType("nerfstudio"), types.ModuleType("nerfstudio.utils"), types.ModuleType(
        "nerfstudio.utils.colors")
    colors.COLORS_DICT = {}
    ns.utils, utils.colors = utils, colors
    for name, mod in (("nerfstudio", ns), ("nerfstudio.utils", utils), ("nerfstudio.utils.colors", colors),
                      ("nerfacc", types.ModuleType("nerfacc"))):
        sys.modules.setdefault(name, mod)


def _load(path: Path, name: str):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _weights(g, R, S):
    """Valid volume-rendering weights (alpha compositing of random densities), some rays nearly empty / opaque."""
    sigma = torch.rand(R, S, 1, generator=g) * torch.tensor([0.0, 0.05, 1.0, 30.0])[torch.randint(0, 4, (R, 1, 1), generator=g)]
    alpha = 1 - torch.exp(-sigma)
    trans = torch.cumprod(torch.cat([torch.ones(R, 1, 1), 1 - alpha[:, :-1]], 1), 1)
    return alpha * trans


def main() -> None:
    _stand_ins()
    tr = _load(REF / "thermal_nerf" / "thermal_renderer.py", "ref_thermal_renderer")
    rr = _load(REF / "rgb_concat" / "rgbt_renderer.py", "ref_rgbt_renderer")
    g = torch.Generator().manual_seed(77)
    thermal_cases, rgbt_cases, blend_cases = [], [], []
    for R, S in ((5, 3), (64, 48), (33, 12)):
        for training in (True, False):
            w = _weights(g, R, S)
            th = torch.rand(R, S, 1, generator=g) * 1.4 - 0.2      # a linear head is unbounded: values outside [0,1]
            c4 = torch.rand(R, S, 4, generator=g)
            if not training:                                      # eval sanitises non-finite samples
                th[0, 0, 0], th[1, -1, 0] = float("nan"), float("inf")
                c4[0, 1, 3], c4[2, -1, 0] = float("nan"), float("-inf")
            ren = tr.ThermalRenderer()
            ren.train(training)
            thermal_cases.append({"thermal": th.clone(), "weights": w, "training": training,
                                  "out": ren(th.clone(), w).clone()})
            ren4 = rr.RGBTRenderer()
            ren4.train(training)
            rgbt_cases.append({"rgbt": c4.clone(), "weights": w, "training": training, "out": ren4(c4.clone(), w).clone()})
    # the loss-time background blend of the concat model (default "random" background)
    for seed in (0, 1):
        pred, acc, gt = torch.rand(40, 4, generator=g), torch.rand(40, 1, generator=g), torch.rand(40, 4, generator=g)
        torch.manual_seed(seed)
        p2, g2 = rr.RGBTRenderer().blend_background_for_loss_computation(pred_image=pred, pred_accumulation=acc, gt_image=gt)
        blend_cases.append({"pred": pred, "acc": acc, "gt": gt, "seed": seed, "pred_out": p2.clone(), "gt_out": g2.clone()})
    # a non-default background argument is ignored by ThermalRenderer.combine_thermal (thermal_renderer.py:49)
    w, th = _weights(g, 7, 9), torch.rand(7, 9, 1, generator=g)
    ren = tr.ThermalRenderer(background_color="black")
    ren.train(True)
    forced = {"thermal": th, "weights": w, "out": ren(th, w, background_color="white").clone()}
    torch.save({"thermal": thermal_cases, "rgbt": rgbt_cases, "blend": blend_cases, "thermal_forced_background": forced,
                "source": "thermo_nerf/thermal_nerf/thermal_renderer.py and thermo_nerf/rgb_concat/rgbt_renderer.py executed "
                          "from /root/reference with import-only stand-ins for nerfstudio.utils.colors / nerfacc",
                "torch": str(torch.__version__)}, OUT)
    print(f"wrote {OUT}: {len(thermal_cases)} thermal, {len(rgbt_cases)} rgbt, {len(blend_cases)} blend cases")


if __name__ == "__main__":
    sys.exit(main())
