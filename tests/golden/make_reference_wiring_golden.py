"""Golden vectors from THE REFERENCE'S OWN model code, executed in the build container over stand-ins for the
absent third-party packages (tests/golden/nerfstudio_standin.py - read its header for what this does and does not pin).

Executed reference code (from /root/reference, unmodified):
  thermo_nerf/thermal_nerf/thermal_nerf_model.py   ThermalNerfModel.__init__, populate_modules (:86-208),
                                                   get_outputs (:210-275), get_loss_dict (:277-326),
                                                   get_image_metrics_and_images (:328-400)
  thermo_nerf/thermal_nerf/thermal_field.py        ThermalNerfactoTField.__init__, get_outputs, forward (:33-201)
  thermo_nerf/thermal_nerf/thermal_field_head.py   BaseThermalFieldHead (:15-71)
  thermo_nerf/thermal_nerf/thermal_renderer.py     ThermalRenderer (:16-149)
  thermo_nerf/nerfacto_config/thermal_nerfacto.py  ThermalNerfactoModel.__init__ (:28-44), get_image_metrics_and_images
                                                   (:46-84), config dataclasses

    python tests/golden/make_reference_wiring_golden.py        # needs /root/reference

Writes tests/golden/reference_model_wiring.pt (weights + inputs + the reference's outputs, losses and gradients)."""

import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(Path(__file__).resolve().parent))
OUT = Path(__file__).resolve().parent / "reference_model_wiring.pt"

NUM_IMAGES, R = 6, 48
MINI = dict(log2_hashmap_size=11, num_proposal_samples_per_ray=(24, 16), num_nerf_samples_per_ray=12,
            proposal_net_args_list=[
                {"hidden_dim": 16, "log2_hashmap_size": 9, "num_levels": 5, "max_res": 128, "use_linear": False},
                {"hidden_dim": 16, "log2_hashmap_size": 9, "num_levels": 5, "max_res": 256, "use_linear": False}])


def build_reference_model(pass_thermal_gradients: bool):
    from thermo_nerf.thermal_nerf.thermal_nerf_model import ThermalNerfModel, ThermalNerfModelConfig
    import nerfstudio_standin as S

    cfg = ThermalNerfModelConfig(implementation="torch", max_temperature=33.085, min_temperature=13.896,
                                 pass_thermal_gradients=pass_thermal_gradients, **MINI)
    box = S.SceneBox(torch.tensor([[-1.0, -1.0, -1.0], [1.0, 1.0, 1.0]]))
    return ThermalNerfModel(cfg, metadata={"thermal": []}, scene_box=box, num_train_data=NUM_IMAGES)


def main() -> None:
    import nerfstudio_standin as S

    S.install()
    from oracle import OracleConfig, OracleThermalNerf, make_synthetic_rays
    from tests.helpers import make_trained_like

    sys.path.append("/root/reference")  # after the repo root: only `thermo_nerf` is taken from there

    # constructor contract (thermal_nerf_model.py:77-78)
    from thermo_nerf.thermal_nerf.thermal_nerf_model import ThermalNerfModel, ThermalNerfModelConfig
    try:
        ThermalNerfModel(ThermalNerfModelConfig(implementation="torch", **MINI), metadata={}, scene_box=None, num_train_data=1)
        raised = None
    except ValueError as e:
        raised = str(e)

    blob = {"cases": [], "missing_thermal_metadata_error": raised,
            "source": "thermo_nerf/thermal_nerf/{thermal_nerf_model,thermal_field,thermal_field_head,thermal_renderer}.py and "
                      "nerfacto_config/thermal_nerfacto.py executed from /root/reference over tests/golden/nerfstudio_standin.py",
            "torch": str(torch.__version__), "mini": {k: (list(v) if isinstance(v, tuple) else v) for k, v in MINI.items()},
            "num_images": NUM_IMAGES}
    # the third case gives the temperature head some contrast (a freshly initialised one is almost constant, and in eval
    # mode its slightly negative output clamps to 0); cases 0 and 1 are the ones the GPU test consumes
    for pass_thermal, contrast in ((True, False), (False, False), (True, True)):
        ref = build_reference_model(pass_thermal)
        # trained-like weights come from an oracle instance; loading them by name also checks the module tree
        ocfg = OracleConfig(log2_hashmap_size=MINI["log2_hashmap_size"],
                            num_proposal_samples_per_ray=MINI["num_proposal_samples_per_ray"],
                            num_nerf_samples_per_ray=MINI["num_nerf_samples_per_ray"],
                            proposal_net_args_list=[{k: v for k, v in a.items() if k != "use_linear"}
                                                    for a in MINI["proposal_net_args_list"]],
                            pass_thermal_gradients=pass_thermal)
        oracle = OracleThermalNerf(ocfg, NUM_IMAGES, seed=11)
        make_trained_like(oracle, 11)
        if contrast:
            with torch.no_grad():
                oracle.field.mlp_thermal.layers[0].weight.mul_(6.0)
                oracle.field.mlp_thermal.layers[1].weight.mul_(4.0)
                oracle.field.field_head_thermal.net.weight.mul_(4.0)
                oracle.field.field_head_thermal.net.bias.fill_(0.45)
        ref_keys = {k: tuple(v.shape) for k, v in ref.state_dict().items()}
        missing, unexpected = ref.load_state_dict(oracle.state_dict(), strict=False)
        state = {k: v.clone() for k, v in oracle.state_dict().items()}

        rays = make_synthetic_rays(R, num_images=NUM_IMAGES, seed=5)
        g = torch.Generator().manual_seed(3)
        jitter = torch.rand(3, R, 1, generator=g)
        batch = {"image": torch.rand(R, 3, generator=g), "thermal": torch.rand(R, 1, generator=g)}

        def bundle():
            return S.RayBundle(origins=rays.origins.clone(), directions=rays.directions.clone(),
                               camera_indices=rays.camera_indices.clone())

        case = {"pass_thermal_gradients": pass_thermal, "thermal_contrast": contrast, "state_dict": state, "reference_state_dict_keys": ref_keys,
                "load_missing": list(missing), "load_unexpected": list(unexpected),
                "origins": rays.origins, "directions": rays.directions, "camera_indices": rays.camera_indices,
                "jitter": jitter, "batch": batch}
        # ---- training mode, annealed proposal weights
        ref.train()
        ref.proposal_sampler.jitter = jitter
        ref.proposal_sampler.set_anneal(0.37)
        out = ref(bundle())
        metrics = ref.get_metrics_dict(out, batch)
        loss = ref.get_loss_dict(out, batch, metrics)
        ref.zero_grad()
        sum(loss.values()).backward()
        case["train"] = {
            "anneal": 0.37,
            "output_keys": sorted(out.keys()),
            "outputs": {k: v.detach().clone() for k, v in out.items() if torch.is_tensor(v)},
            "weights_list": [w.detach().clone() for w in out["weights_list"]],
            "spacing_bins": [rs.spacing_bins().detach().clone() for rs in out["ray_samples_list"]],
            "loss": {k: v.detach().clone() for k, v in loss.items()},
            "distortion": metrics["distortion"].detach().clone(),
            "grads": {k: p.grad.detach().clone() for k, p in ref.named_parameters()
                      if p.grad is not None and "hash_table" not in k},
            "grad_norms": {k: p.grad.norm().detach().clone() for k, p in ref.named_parameters() if p.grad is not None},
            "params_without_grad": sorted(k for k, p in ref.named_parameters() if p.grad is None),
        }
        # ---- eval mode
        ref.eval()
        ref.proposal_sampler.jitter = None
        ref.proposal_sampler.set_anneal(1.0)
        with torch.no_grad():
            out = ref(bundle())
            loss = ref.get_loss_dict(out, batch, None)
        case["eval"] = {"output_keys": sorted(out.keys()),
                        "outputs": {k: v.detach().clone() for k, v in out.items() if torch.is_tensor(v)},
                        "loss": {k: v.detach().clone() for k, v in loss.items()}}
        blob["cases"].append(case)

    # ---- evaluation metrics: ThermalNerfModel.get_image_metrics_and_images (thermal_nerf_model.py:328-400) and
    #      ThermalNerfactoModel.get_image_metrics_and_images (thermal_nerfacto.py:46-84) on synthetic rendered frames.
    #      psnr is the real formula; the ssim / lpips callables are shape- and order-sensitive MARKERS (see the stand-in).
    from thermo_nerf.nerfacto_config.thermal_nerfacto import ThermalNerfactoModel

    blob["image_metrics"] = []
    g = torch.Generator().manual_seed(9)
    H, W = 24, 20
    for cold in (False, True):
        for threshold in (None, 0.4):
            ref = build_reference_model(True)
            ref.config.cold = cold
            ref.eval()
            outputs = {"rgb": torch.rand(H, W, 3, generator=g), "accumulation": torch.rand(H, W, 1, generator=g),
                       "depth": torch.rand(H, W, 1, generator=g) * 4, "expected_depth": torch.rand(H, W, 1, generator=g),
                       "prop_depth_0": torch.rand(H, W, 1, generator=g) * 4, "prop_depth_1": torch.rand(H, W, 1, generator=g) * 4,
                       "thermal": torch.rand(H, W, 1, generator=g)}
            batch = {"image": torch.rand(H, W, 3, generator=g), "thermal": torch.rand(H, W, 1, generator=g)}
            metrics, images = ref.get_image_metrics_and_images(outputs, batch, threshold=threshold)
            metrics_nf, images_nf = ThermalNerfactoModel.get_image_metrics_and_images(ref, outputs=outputs, batch=batch,
                                                                                      threshold=threshold)
            blob["image_metrics"].append({
                "cold": cold, "threshold": threshold, "max_temperature": ref.max_temperature,
                "min_temperature": ref.min_temperature, "outputs": outputs, "batch": batch,
                "metrics": metrics, "image_keys": sorted(images), "image_shapes": {k: list(v.shape) for k, v in images.items()},
                "thermal_image": images["thermal"].clone(), "thermal_combined": images["thermal_combined"].clone(),
                "metrics_nerfacto_track": metrics_nf, "image_keys_nerfacto_track": sorted(images_nf)})
    torch.save(blob, OUT)
    print(f"wrote {OUT} ({OUT.stat().st_size / 1024:.0f} KiB)")
    print("  image metrics:", blob["image_metrics"][1]["metrics"], blob["image_metrics"][1]["metrics_nerfacto_track"])
    for c in blob["cases"]:
        print("  pass_thermal_gradients", c["pass_thermal_gradients"], "missing", c["load_missing"], "unexpected",
              c["load_unexpected"], "train keys", c["train"]["output_keys"], "loss", {k: float(v) for k, v in c["train"]["loss"].items()})


if __name__ == "__main__":
    sys.exit(main())
