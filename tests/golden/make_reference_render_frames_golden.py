"""Golden frames from THE REFERENCE'S OWN Renderer.render loop (thermo_nerf/render/renderer.py:160-201), executed in the
build container on the reference's ThermalNerfModel over the nerfstudio stand-ins (tests/golden/nerfstudio_standin.py).
matplotlib and imageio are absent here: `plt.colormaps["magma"]` is replaced by the oracle's restatement of matplotlib's
Colormap.__call__ over a synthetic 256-entry table (so the colour map's arithmetic is NOT pinned by this file, the
reference's conversion flow around it is), imageio by a recorder of the file names save_images / save_gif ask it to write.

    python tests/golden/make_reference_render_frames_golden.py        # needs /root/reference

Writes tests/golden/reference_render_frames.pt."""

import sys
import types
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parents[1]))
sys.path.insert(0, str(HERE))
OUT = HERE / "reference_render_frames.pt"
H, W, FOCAL, FRAMES = 10, 12, 13.0, 2


def synthetic_lut() -> np.ndarray:
    x = np.linspace(0, 1, 256)
    return np.stack([x**0.5, 0.5 - 0.5 * np.cos(3 * np.pi * x), 1 - x**2], 1)


class PathCameras:
    """What Renderer.render touches of nerfstudio Cameras: to(), size, generate_rays(camera_indices=int) -> [H,W] bundle."""

    def __init__(self, c2w: torch.Tensor) -> None:
        self.camera_to_worlds = c2w

    def to(self, device):
        return self

    @property
    def size(self) -> int:
        return int(self.camera_to_worlds.shape[0])

    def generate_rays(self, camera_indices: int):
        import nerfstudio_standin as S
        from oracle.camera_post import generate_rays_np

        o, d, _ = generate_rays_np(self.camera_to_worlds[camera_indices].numpy(), FOCAL, FOCAL, W / 2, H / 2, H, W)
        return S.RayBundle(origins=torch.from_numpy(o), directions=torch.from_numpy(d),
                           camera_indices=torch.full((H, W, 1), camera_indices, dtype=torch.int64))


def main() -> None:
    import nerfstudio_standin as S

    S.install()
    from oracle import OracleConfig, OracleThermalNerf
    from oracle.camera_post import ListedColormapLike
    from tests.helpers import make_trained_like
    from thermo_nerf_b200 import sphere_cameras

    cmap = ListedColormapLike(synthetic_lut())
    written = []  # what save_images / save_gif hand to imageio: (function, file name, frames, duration)

    def imwrite(path, image):
        written.append(("imwrite", Path(path).name, 1, None))

    def mimsave(path, frames, duration=None):
        written.append(("mimsave", Path(path).name, len(frames), duration))

    for name, attrs in (("imageio", {"imwrite": imwrite, "mimsave": mimsave}), ("matplotlib", {}), ("matplotlib.pyplot", {"colormaps": {"magma": cmap}}),
                        ("matplotlib.colors", {"Colormap": ListedColormapLike}),
                        ("nerfstudio.cameras.camera_paths", {"get_path_from_json": None}),
                        ("nerfstudio.cameras.cameras", {"Cameras": PathCameras}),
                        ("nerfstudio.models.base_model", {"Model": S.NerfactoModel})):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
    sys.path.append("/root/reference")
    import make_reference_wiring_golden as Wg
    from thermo_nerf.render.renderer import Renderer
    from thermo_nerf.rendered_image_modalities import RenderedImageModality as Mod

    model = Wg.build_reference_model(True)
    model.config.eval_num_rays_per_chunk = 64
    ocfg = OracleConfig(log2_hashmap_size=Wg.MINI["log2_hashmap_size"],
                        num_proposal_samples_per_ray=Wg.MINI["num_proposal_samples_per_ray"],
                        num_nerf_samples_per_ray=Wg.MINI["num_nerf_samples_per_ray"],
                        proposal_net_args_list=[{k: v for k, v in a.items() if k != "use_linear"}
                                                for a in Wg.MINI["proposal_net_args_list"]])
    oracle = OracleThermalNerf(ocfg, Wg.NUM_IMAGES, seed=31)
    make_trained_like(oracle, 31)
    with torch.no_grad():  # a freshly initialised temperature head is almost constant: give the frames some contrast
        oracle.field.mlp_thermal.layers[0].weight.mul_(6.0)
        oracle.field.mlp_thermal.layers[1].weight.mul_(4.0)
        oracle.field.field_head_thermal.net.weight.mul_(4.0)
        oracle.field.field_head_thermal.net.bias.fill_(0.45)
    model.load_state_dict(oracle.state_dict(), strict=False)
    model.eval()
    cams = PathCameras(sphere_cameras(FRAMES, hw=H, focal=FOCAL).camera_to_worlds.clone())

    calls, frames = [], {}
    inner = model.get_outputs_for_camera_ray_bundle

    def recording(bundle):
        out = inner(bundle)
        idx = int(bundle.camera_indices.flatten()[0])
        calls.append(idx)
        frames[idx] = {k: v.clone() for k, v in out.items()}
        return out

    model.get_outputs_for_camera_ray_bundle = recording
    r = Renderer(model)
    mods = [Mod.THERMAL, Mod.DEPTH, Mod.ACCUMULATION]
    r.render(mods, cams)  # default thermal_color_map = plt.colormaps["magma"] (bound at import: the stand-in map)
    rendered = {m.name: [torch.from_numpy(np.ascontiguousarray(a)) for a in r._rendered_images[m]] for m in mods}
    r.save_images(mods, Path("/nonexistent"))          # renderer.py:202-213 (the recorder does not touch the disk)
    r.save_gif(mods, 2.5, Path("/nonexistent"))        # renderer.py:215-228
    # the RGB modality is named "img" but the model emits "rgb" (renderer.py:186-187): the stock loop raises
    try:
        r.render([Mod.RGB], cams)
        rgb_error = None
    except Exception as e:  # noqa: BLE001 - the reference raises a bare Exception
        rgb_error = str(e)
    torch.save({"frames": [frames[i] for i in range(FRAMES)], "camera_to_worlds": cams.camera_to_worlds, "hw": [H, W],
                "focal": FOCAL, "lut": torch.from_numpy(synthetic_lut()), "modalities": [m.name for m in mods],
                "rendered": rendered, "model_calls": calls, "written": [list(w) for w in written], "rgb_modality_error": rgb_error,
                "source": "thermo_nerf/render/renderer.py:160-201 executed from /root/reference over tests/golden/nerfstudio_standin.py",
                "torch": str(torch.__version__)}, OUT)
    print(f"wrote {OUT} ({OUT.stat().st_size / 1024:.0f} KiB); model calls {calls}; RGB error: {rgb_error!r}")


if __name__ == "__main__":
    sys.exit(main())
