"""Golden vectors from the reference's `concat_nerf` baseline, executed in the build container over the nerfstudio
stand-ins (tests/golden/nerfstudio_standin.py - read its header for what this does and does not pin).

Executed reference code (from /root/reference, unmodified):
  thermo_nerf/rgb_concat/concat_nerfacto_model.py  ConcatNerfModel.__init__, populate_modules (:60-195), get_loss_dict
                                                   (:197-233), get_metrics_dict (:235-249), get_image_metrics_and_images (:251-324)
  thermo_nerf/rgb_concat/concat_field.py           ConcatNerfactoTField (:9-75)
  thermo_nerf/rgb_concat/rgbt_renderer.py          RGBTRenderer (:17-174)
get_outputs / the field's get_outputs are nerfstudio's (inherited): stand-in code, not reference code.

    python tests/golden/make_reference_concat_golden.py        # needs /root/reference

Writes tests/golden/reference_concat_wiring.pt."""

import sys
from pathlib import Path

import torch

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parents[1]))
sys.path.insert(0, str(HERE))
OUT = HERE / "reference_concat_wiring.pt"


def main() -> None:
    import types

    import nerfstudio_standin as S

    S.install()
    sys.modules.setdefault("nerfacc", types.ModuleType("nerfacc"))
    from oracle import OracleConfig, OracleThermalNerf, make_synthetic_rays
    from tests.helpers import make_trained_like

    sys.path.append("/root/reference")
    import make_reference_wiring_golden as Wg
    from thermo_nerf.rgb_concat.concat_nerfacto_model import ConcatNerfModel, ConcatNerfModelConfig

    cfg = ConcatNerfModelConfig(implementation="torch", max_temperature=33.085, min_temperature=13.896, **Wg.MINI)
    ref = ConcatNerfModel(cfg, scene_box=S.SceneBox(torch.tensor([[-1.0, -1.0, -1.0], [1.0, 1.0, 1.0]])),
                          num_train_data=Wg.NUM_IMAGES)
    ocfg = OracleConfig(log2_hashmap_size=Wg.MINI["log2_hashmap_size"],
                        num_proposal_samples_per_ray=Wg.MINI["num_proposal_samples_per_ray"],
                        num_nerf_samples_per_ray=Wg.MINI["num_nerf_samples_per_ray"],
                        proposal_net_args_list=[{k: v for k, v in a.items() if k != "use_linear"}
                                                for a in Wg.MINI["proposal_net_args_list"]], head="concat")
    oracle = OracleThermalNerf(ocfg, Wg.NUM_IMAGES, seed=13)
    make_trained_like(oracle, 13)
    ref_keys = {k: tuple(v.shape) for k, v in ref.state_dict().items()}
    missing, unexpected = ref.load_state_dict(oracle.state_dict(), strict=False)

    R = 40
    rays = make_synthetic_rays(R, num_images=Wg.NUM_IMAGES, seed=8)
    g = torch.Generator().manual_seed(6)
    jitter = torch.rand(3, R, 1, generator=g)
    batch = {"image": torch.rand(R, 4, generator=g)}  # the concat dataset stacks thermal as the 4th channel

    def bundle():
        return S.RayBundle(origins=rays.origins.clone(), directions=rays.directions.clone(),
                           camera_indices=rays.camera_indices.clone())

    blob = {"state_dict": {k: v.clone() for k, v in oracle.state_dict().items()}, "reference_state_dict_keys": ref_keys,
            "load_missing": list(missing), "load_unexpected": list(unexpected), "mini": {
                k: (list(v) if isinstance(v, tuple) else v) for k, v in Wg.MINI.items()}, "num_images": Wg.NUM_IMAGES,
            "origins": rays.origins, "directions": rays.directions, "camera_indices": rays.camera_indices,
            "jitter": jitter, "batch": batch, "background_color": ref.renderer_rgb.background_color,
            "source": "thermo_nerf/rgb_concat/{concat_nerfacto_model,concat_field,rgbt_renderer}.py executed from "
                      "/root/reference over tests/golden/nerfstudio_standin.py", "torch": str(torch.__version__)}
    ref.train()
    ref.proposal_sampler.jitter = jitter
    ref.proposal_sampler.set_anneal(0.61)
    out = ref(bundle())
    metrics = ref.get_metrics_dict(out, batch)
    torch.manual_seed(123)  # the loss draws rand_like(pred) for its "random" background (rgbt_renderer.py:137-139)
    loss = ref.get_loss_dict(out, batch, metrics)
    ref.zero_grad()
    sum(loss.values()).backward()
    blob["train"] = {"anneal": 0.61, "noise_seed": 123, "output_keys": sorted(out.keys()),
                     "outputs": {k: v.detach().clone() for k, v in out.items() if torch.is_tensor(v)},
                     "loss": {k: v.detach().clone() for k, v in loss.items()},
                     "metric_keys": sorted(metrics), "psnr": metrics["psnr"].detach().clone(),
                     "grads": {k: p.grad.detach().clone() for k, p in ref.named_parameters()
                               if p.grad is not None and "hash_table" not in k}}
    ref.eval()
    ref.proposal_sampler.jitter = None
    ref.proposal_sampler.set_anneal(1.0)
    with torch.no_grad():
        out = ref(bundle())
    blob["eval"] = {"output_keys": sorted(out.keys()), "outputs": {k: v.clone() for k, v in out.items() if torch.is_tensor(v)}}
    # evaluation metrics on synthetic frames: the temperature metrics are taken on channel 3 only (:281-282)
    H, W = 16, 14
    outputs = {"rgb": torch.rand(H, W, 4, generator=g), "accumulation": torch.rand(H, W, 1, generator=g),
               "depth": torch.rand(H, W, 1, generator=g) * 3, "prop_depth_0": torch.rand(H, W, 1, generator=g),
               "prop_depth_1": torch.rand(H, W, 1, generator=g)}
    fbatch = {"image": torch.rand(H, W, 4, generator=g)}
    m, images = ref.get_image_metrics_and_images(outputs, fbatch, threshold=0.4)
    blob["image_metrics"] = {"outputs": outputs, "batch": fbatch, "threshold": 0.4, "metrics": m, "image_keys": sorted(images),
                             "image_shapes": {k: list(v.shape) for k, v in images.items()}}
    torch.save(blob, OUT)
    print(f"wrote {OUT} ({OUT.stat().st_size / 1024:.0f} KiB); missing {missing} unexpected {unexpected}; background "
          f"{ref.renderer_rgb.background_color!r}; loss { {k: float(v) for k, v in loss.items()} }; metrics {m}")


if __name__ == "__main__":
    sys.exit(main())
