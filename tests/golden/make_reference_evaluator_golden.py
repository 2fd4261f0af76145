"""Golden files from THE REFERENCE'S OWN Evaluator (thermo_nerf/evaluator/evaluator.py:15-175), executed in the build
container on the reference's ThermalNerfModel over the nerfstudio stand-ins (tests/golden/nerfstudio_standin.py - read its
header; the ssim / lpips callables are markers, psnr and the temperature MAE are the real formulas).

    python tests/golden/make_reference_evaluator_golden.py        # needs /root/reference

Writes tests/golden/reference_evaluator.pt: per-frame model outputs and batches, the aggregated metrics, metrics.json,
the tree of files save_metrics / save_images wrote, and the uint8 evaluation images."""

import sys
import tempfile
from pathlib import Path
from types import SimpleNamespace

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parents[1]))
sys.path.insert(0, str(HERE))
OUT = HERE / "reference_evaluator.pt"
H, W, FOCAL, FRAMES = 12, 10, 14.0, 3


class EvalCameras:
    """One pinhole camera per loader item, with nerfstudio's generate_rays(camera_indices) -> [H,W] bundle."""

    def __init__(self, c2w: torch.Tensor) -> None:
        self.camera_to_worlds = c2w  # [1,3,4]

    def generate_rays(self, camera_indices):
        import nerfstudio_standin as S
        from oracle.camera_post import generate_rays_np

        assert camera_indices.tolist() == [0]
        o, d, _ = generate_rays_np(self.camera_to_worlds[0].numpy(), FOCAL, FOCAL, W / 2, H / 2, H, W)
        return S.RayBundle(origins=torch.from_numpy(o), directions=torch.from_numpy(d),
                           camera_indices=torch.zeros(H, W, 1, dtype=torch.int64))


def main() -> None:
    import nerfstudio_standin as S

    S.install()
    from oracle import OracleConfig, OracleThermalNerf
    from tests.helpers import make_trained_like
    from thermo_nerf_b200 import sphere_cameras

    sys.path.append("/root/reference")
    import make_reference_wiring_golden as Wg
    from thermo_nerf.evaluator.evaluator import Evaluator
    from thermo_nerf.rendered_image_modalities import RenderedImageModality as Mod

    model = Wg.build_reference_model(True)
    model.config.eval_num_rays_per_chunk = 50  # 120 rays per frame: three chunks
    ocfg = OracleConfig(log2_hashmap_size=Wg.MINI["log2_hashmap_size"],
                        num_proposal_samples_per_ray=Wg.MINI["num_proposal_samples_per_ray"],
                        num_nerf_samples_per_ray=Wg.MINI["num_nerf_samples_per_ray"],
                        proposal_net_args_list=[{k: v for k, v in a.items() if k != "use_linear"}
                                                for a in Wg.MINI["proposal_net_args_list"]])
    oracle = OracleThermalNerf(ocfg, Wg.NUM_IMAGES, seed=21)
    make_trained_like(oracle, 21)
    with torch.no_grad():  # a freshly initialised temperature head is almost constant: give the frames some contrast
        oracle.field.mlp_thermal.layers[0].weight.mul_(6.0)
        oracle.field.mlp_thermal.layers[1].weight.mul_(4.0)
        oracle.field.field_head_thermal.net.weight.mul_(4.0)
        oracle.field.field_head_thermal.net.bias.fill_(0.45)
    model.load_state_dict(oracle.state_dict(), strict=False)
    model.eval()

    cams = sphere_cameras(FRAMES, hw=H, focal=FOCAL)  # poses only (the product helper places cameras on a sphere)
    g = torch.Generator().manual_seed(4)
    loader = [(EvalCameras(cams.camera_to_worlds[i:i + 1].clone()),
               {"image": torch.rand(H, W, 3, generator=g), "thermal": torch.rand(H, W, 1, generator=g)})
              for i in range(FRAMES)]
    # record what the model returns per frame (the evaluator's input on the product side of the comparison)
    frames = []
    inner = model.get_outputs_for_camera_ray_bundle

    def recording(bundle):
        out = inner(bundle)
        frames.append({k: v.clone() for k, v in out.items()})
        return out

    model.get_outputs_for_camera_ray_bundle = recording
    pipeline = SimpleNamespace(model=model, datamanager=SimpleNamespace(setup_eval=lambda: None,
                                                                        fixed_indices_eval_dataloader=loader))
    config = SimpleNamespace(experiment_name="double_robot", method_name="thermal-nerf")
    mods = [Mod.RGB, Mod.THERMAL, Mod.THERMAL_COMBINED, Mod.DEPTH, Mod.ACCUMULATION]
    ev = Evaluator(pipeline, config, job_param_identifier="job-3", modalities_to_save=mods, threshold=0.4)
    with tempfile.TemporaryDirectory() as tmp:
        root = Path(tmp)
        ev.save_metrics(root)
        ev.save_images(mods, root)
        files = sorted(str(p.relative_to(root)) for p in root.rglob("*") if p.is_file())
        texts = {f: (root / f).read_text() for f in files if f.endswith((".json", ".txt"))}
    ev_none = Evaluator(pipeline, config)  # defaults: no identifier, RGB only, no threshold
    with tempfile.TemporaryDirectory() as tmp:
        ev_none.save_metrics(Path(tmp))
        files_none = sorted(str(p.relative_to(tmp)) for p in Path(tmp).rglob("*") if p.is_file())
    blob = {"frames": frames[:FRAMES], "batches": [b for _, b in loader], "camera_to_worlds": cams.camera_to_worlds.clone(),
            "hw": [H, W], "focal": FOCAL, "threshold": 0.4, "identifier": "job-3",
            "modalities": [m.name for m in mods], "metrics": ev._metrics, "benchmark_info": ev._benchmark_info,
            "files": files, "texts": texts,
            "images": {m.name: [torch.from_numpy(np.ascontiguousarray(a)) for a in ev._evaluation_images[m]] for m in mods},
            "default_files": files_none, "default_modalities": [m.name for m in ev_none.modalities_to_save],
            "max_temperature": model.max_temperature, "min_temperature": model.min_temperature,
            "source": "thermo_nerf/evaluator/evaluator.py executed from /root/reference on the reference's ThermalNerfModel over "
                      "tests/golden/nerfstudio_standin.py", "torch": str(torch.__version__)}
    torch.save(blob, OUT)
    print(f"wrote {OUT} ({OUT.stat().st_size / 1024:.0f} KiB); files: {files}")
    print({k: v for k, v in ev._metrics.items() if k.endswith("_mean")})


if __name__ == "__main__":
    sys.exit(main())
