"""Runs the REFERENCE's own entry points on the B200 plugin classes, in the build container (needs /root/reference).

Executed in a subprocess by tests/test_plugin_cpu.py (the nerfstudio stand-ins are installed into sys.modules, which
must not leak into the rest of the test session).  What runs here, unmodified, from /root/reference:

  thermo_nerf/thermal_nerf/thermal_nerf_model.py   ThermalNerfModel.__init__ / populate_modules (through super())
  thermo_nerf/thermal_nerf/config_thermal_nerf.py  thermal_nerf_config (module-level TrainerConfig)
  thermo_nerf/nerfacto_config/config_nerfacto.py   thermalnerfacto_config
  thermo_nerf/render/renderer.py                   Renderer(model).render / save_images / save_gif
  thermo_nerf/evaluator/evaluator.py               Evaluator(...).save_metrics -> _compute_metrics (isinstance at :76)

over tests/golden/nerfstudio_standin.py (nerfstudio's interfaces) - with ``thermo_nerf_b200.nerfstudio_plugin``'s
``B200ThermalNerfModel`` as the model.  There is no GPU here, so the one call that would launch a kernel,
``functional.render_forward``, is replaced by a test double that answers from the CPU oracle with the model's own
weights; everything around it - class hierarchy, construction through ``config.setup``, state_dict keys, the sampler
state, chunk handling, output keys, the "img" alias, the Renderer / Evaluator flows - is the real code.  The GPU side
of the same mixin is tests/test_plugin_gpu.py.

Prints one JSON object; exit code 0 = every check passed."""

import importlib
import json
import sys
import tempfile
import types
from pathlib import Path
from types import SimpleNamespace

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(HERE / "golden"))

checks = {}


def check(name, cond, detail=""):
    checks[name] = bool(cond)
    if not cond:
        print(f"FAILED: {name} {detail}", file=sys.stderr)


def install_standins():
    import make_reference_config_golden as Cg
    import make_reference_render_frames_golden as Rg
    import nerfstudio_standin as S
    from oracle.camera_post import ListedColormapLike

    S.install()
    written = []
    cmap = ListedColormapLike(Rg.synthetic_lut())
    for name, attrs in (("imageio", {"imwrite": lambda p, im: written.append(("imwrite", Path(p).name)),
                                     "mimsave": lambda p, fr, duration=None: written.append(("mimsave", Path(p).name))}),
                        ("matplotlib", {}), ("matplotlib.pyplot", {"colormaps": {"magma": cmap}}),
                        ("matplotlib.colors", {"Colormap": ListedColormapLike}),
                        ("nerfstudio.cameras.camera_paths", {"get_path_from_json": None}),
                        ("nerfstudio.cameras.cameras", {"Cameras": Rg.PathCameras}),
                        ("nerfstudio.models.base_model", {"Model": S.NerfactoModel})):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        parent, _, leaf = name.rpartition(".")
        if parent:
            setattr(sys.modules[parent], leaf, m)
    # recording stand-ins for the config classes (TrainerConfig, data managers, optimisers ...), with attribute access
    Cg.Recorded.__getattr__ = lambda self, n: self.__dict__["_kwargs"][n] if n in self.__dict__.get("_kwargs", {}) else (
        _ for _ in ()).throw(AttributeError(n))

    def rec_setattr(self, n, v):
        if n in ("_args", "_kwargs"):
            object.__setattr__(self, n, v)
        else:  # dataclass subclasses of a recorded class never run Recorded.__init__
            self.__dict__.setdefault("_kwargs", {})[n] = v

    Cg.Recorded.__setattr__ = rec_setattr
    sys.meta_path.append(Cg._Finder())
    for name, m in list(sys.modules.items()):
        if name.split(".")[0] in Cg._Finder.PREFIXES and type(m) is types.ModuleType:
            m.__class__ = Cg._RecordingModule
    import dataclasses

    @dataclasses.dataclass
    class VanillaPipelineConfig:
        _target: type = None
        datamanager: object = None
        model: object = None

    sys.modules["nerfstudio.pipelines.base_pipeline"].VanillaPipelineConfig = VanillaPipelineConfig
    importlib.import_module("nerfstudio.data.datasets.base_dataset").InputDataset = type(
        "InputDataset", (Cg.Recorded,), {"exclude_batch_keys_from_device": ["image", "mask"]})
    del sys.modules["nerfstudio.engine.trainer"].TrainerConfig
    sys.path.append("/root/reference")
    return S, Rg, written


def oracle_double(model, Wg):
    """functional.render_forward answered by the CPU oracle carrying the model's weights."""
    from oracle import OracleConfig, OracleRays, OracleThermalNerf

    ocfg = OracleConfig(log2_hashmap_size=Wg.MINI["log2_hashmap_size"],
                        num_proposal_samples_per_ray=Wg.MINI["num_proposal_samples_per_ray"],
                        num_nerf_samples_per_ray=Wg.MINI["num_nerf_samples_per_ray"],
                        proposal_net_args_list=[{k: v for k, v in a.items() if k != "use_linear"}
                                                for a in Wg.MINI["proposal_net_args_list"]])
    oracle = OracleThermalNerf(ocfg, Wg.NUM_IMAGES, seed=0)
    calls = []

    def render_forward(tensors, origins, directions, camera_indices=None, nears=None, fars=None, jitter=None, *,
                       training=False, depth_clip_chunk=0, return_samples=False, **kw):
        assert not training and jitter is None and nears is None and fars is None
        sd = {k: v for k, v in model.state_dict().items() if k in oracle.state_dict()}
        oracle.load_state_dict(sd, strict=False)
        oracle.eval()
        R = origins.shape[0]
        calls.append({"rays": R, "chunk": depth_clip_chunk, "near": kw["near_plane"], "anneal": kw["anneal"],
                      "appearance_mode": kw["appearance_mode"]})
        chunk = depth_clip_chunk or R
        parts = []
        with torch.no_grad():
            for s in range(0, R, chunk):
                rays = OracleRays(origins[s:s + chunk], directions[s:s + chunk],
                                  camera_indices[s:s + chunk].reshape(-1) if camera_indices is not None else None)
                parts.append(oracle.get_outputs(rays, training=False))
        return {k: torch.cat([p[k] for p in parts]) for k in parts[0] if torch.is_tensor(parts[0][k])}

    return render_forward, calls


def main():
    S, Rg, written = install_standins()
    import make_reference_wiring_golden as Wg
    from thermo_nerf.nerfacto_config.thermal_nerfacto import ThermalNerfactoModel, ThermalNerfactoModelConfig
    from thermo_nerf.thermal_nerf.thermal_nerf_model import ThermalNerfModel, ThermalNerfModelConfig

    import thermo_nerf_b200.functional as F
    import thermo_nerf_b200.nerfstudio_plugin as P
    from tests.helpers import make_trained_like

    check("plugin_available", P.AVAILABLE, repr(P.IMPORT_ERROR))
    if not P.AVAILABLE:
        return
    # ---- class hierarchy and construction through the config (train_eval_script.py:94, evaluator.py:76)
    cfg = P.B200ThermalNerfModelConfig(max_temperature=33.085, min_temperature=13.896, precision="fp32", **Wg.MINI)
    check("config_is_reference_config", isinstance(cfg, ThermalNerfModelConfig) and isinstance(cfg, ThermalNerfactoModelConfig))
    box = S.SceneBox(torch.tensor([[-1.0, -1.0, -1.0], [1.0, 1.0, 1.0]]))
    model = cfg.setup(metadata={"thermal": []}, scene_box=box, num_train_data=Wg.NUM_IMAGES)
    check("model_is_reference_model", isinstance(model, ThermalNerfModel) and isinstance(model, ThermalNerfactoModel)
          and type(model) is P.B200ThermalNerfModel)
    check("config_forces_torch_layout", model.config.implementation == "torch")
    try:
        cfg.setup(metadata={}, scene_box=box, num_train_data=Wg.NUM_IMAGES)
        check("ctor_requires_thermal_metadata", False)
    except ValueError as e:  # thermal_nerf_model.py:75-76, raised by the reference's own constructor
        check("ctor_requires_thermal_metadata", "Thermal images not found" in str(e))
    stock = Wg.build_reference_model(True)
    check("state_dict_keys_equal_stock", list(model.state_dict()) == list(stock.state_dict()))
    check("param_groups", set(model.get_param_groups()) == {"proposal_networks", "fields", "camera_opt"})
    # ---- the kernel-backed methods are the mixin's, the rest the reference's
    from thermo_nerf_b200.model import KernelModelMixin

    for name in ("forward", "get_outputs", "get_outputs_for_camera_ray_bundle", "get_metrics_dict", "get_loss_dict"):
        check(f"mixin_{name}", getattr(type(model), name) is getattr(KernelModelMixin, name))
    check("reference_image_metrics_kept",
          type(model).get_image_metrics_and_images is ThermalNerfModel.get_image_metrics_and_images)
    # ---- sampler state lives in the reference's ProposalNetworkSampler
    model.proposal_sampler.set_anneal(0.25)
    check("anneal_from_sampler", model._render_kwargs()["anneal"] == 0.25)
    model.proposal_sampler.set_anneal(1.0)
    check("update_schedule_from_sampler", model._update_schedule(2500) == 2.5 and model._update_schedule(0) == 1.0)
    # ---- field surface: kernel-backed methods bound on the reference's own field instance; no CPU fallback
    from thermo_nerf_b200 import surface

    check("field_surface_bound", model.field.get_density.__func__ is surface.field_get_density
          and model.field.forward.__func__ is surface.field_forward
          and model.density_fns[0].__func__ is surface.density_fn)
    try:
        model.eval()
        model.field.density_fn(torch.zeros(2, 3))
        check("field_surface_no_cpu_path", False)
    except RuntimeError as e:
        check("field_surface_no_cpu_path", "no CPU path" in str(e))
    try:
        model.get_outputs(S.RayBundle(torch.zeros(2, 3), torch.ones(2, 3), camera_indices=torch.zeros(2, 1, dtype=torch.int64)))
        check("get_outputs_no_cpu_path", False)
    except RuntimeError as e:
        check("get_outputs_no_cpu_path", "no CPU path" in str(e))

    # ---- the reference's Renderer.render on the plugin model (weights of the committed golden run)
    from oracle import OracleConfig, OracleThermalNerf

    def load_golden_weights(seed):
        ocfg = OracleConfig(log2_hashmap_size=Wg.MINI["log2_hashmap_size"],
                            num_proposal_samples_per_ray=Wg.MINI["num_proposal_samples_per_ray"],
                            num_nerf_samples_per_ray=Wg.MINI["num_nerf_samples_per_ray"],
                            proposal_net_args_list=[{k: v for k, v in a.items() if k != "use_linear"}
                                                    for a in Wg.MINI["proposal_net_args_list"]])
        o = OracleThermalNerf(ocfg, Wg.NUM_IMAGES, seed=seed)
        make_trained_like(o, seed)
        with torch.no_grad():
            o.field.mlp_thermal.layers[0].weight.mul_(6.0)
            o.field.mlp_thermal.layers[1].weight.mul_(4.0)
            o.field.field_head_thermal.net.weight.mul_(4.0)
            o.field.field_head_thermal.net.bias.fill_(0.45)
        model.load_state_dict(o.state_dict(), strict=False)
        model.eval()

    double, calls = oracle_double(model, Wg)
    F.render_forward = double
    import thermo_nerf_b200.model as Mm

    Mm.F.render_forward = double
    from thermo_nerf.evaluator.evaluator import Evaluator
    from thermo_nerf.render.renderer import Renderer
    from thermo_nerf.rendered_image_modalities import RenderedImageModality as Mod
    from thermo_nerf_b200 import sphere_cameras

    gold = torch.load(HERE / "golden" / "reference_render_frames.pt", weights_only=False)
    load_golden_weights(31)
    model.config.eval_num_rays_per_chunk = 64
    cams = Rg.PathCameras(gold["camera_to_worlds"].clone())
    r = Renderer(model)
    mods = [Mod.THERMAL, Mod.DEPTH, Mod.ACCUMULATION]
    r.render(mods, cams)
    same = all(np.array_equal(np.asarray(a), g.numpy()) for m in mods
               for a, g in zip(r._rendered_images[m], gold["rendered"][m.name]))
    check("renderer_frames_equal_reference_run", same)
    check("one_launch_per_frame_with_reference_chunking", all(c["chunk"] == 64 and c["rays"] == gold["hw"][0] * gold["hw"][1]
                                                              and c["near"] == 0.0 for c in calls), repr(calls[:2]))
    r.save_images(mods, Path("/nonexistent"))
    r.save_gif(mods, 2.5, Path("/nonexistent"))
    check("renderer_writes", len(written) > 0)
    r.render([Mod.RGB], cams)  # the stock model raises here ("img" vs "rgb", renderer.py:186-190); the alias fixes it
    check("rgb_modality_renders", len(r._rendered_images[Mod.RGB]) == cams.size and gold["rgb_modality_error"] is not None)

    # ---- the reference's Evaluator on the plugin model (isinstance at evaluator.py:76 included)
    import make_reference_evaluator_golden as Eg

    gold = torch.load(HERE / "golden" / "reference_evaluator.pt", weights_only=False)
    load_golden_weights(21)
    model.config.eval_num_rays_per_chunk = 50
    c2w = gold["camera_to_worlds"]
    loader = [(Eg.EvalCameras(c2w[i:i + 1].clone()), gold["batches"][i]) for i in range(len(gold["batches"]))]
    pipeline = SimpleNamespace(model=model, datamanager=SimpleNamespace(setup_eval=lambda: None,
                                                                        fixed_indices_eval_dataloader=loader))
    config = SimpleNamespace(experiment_name="double_robot", method_name="thermal-nerf")
    mods = [Mod[m] for m in gold["modalities"]]
    ev = Evaluator(pipeline, config, job_param_identifier=gold["identifier"], modalities_to_save=mods,
                   threshold=gold["threshold"])
    with tempfile.TemporaryDirectory() as tmp:
        ev.save_metrics(Path(tmp))
        ev.save_images(mods, Path(tmp))
        files = sorted(str(p.relative_to(tmp)) for p in Path(tmp).rglob("*") if p.is_file())
    worst = 0.0
    for k, v in gold["metrics"].items():
        a = torch.tensor(ev._metrics[k], dtype=torch.float64).reshape(-1)
        b = torch.tensor(v, dtype=torch.float64).reshape(-1)
        worst = max(worst, float((a - b).abs().max()))
    check("evaluator_metrics_equal_reference_run", set(ev._metrics) == set(gold["metrics"]) and worst < 1e-5, f"{worst}")
    check("evaluator_files_equal_reference_run", files == gold["files"])
    same = all(np.array_equal(np.asarray(a), g.numpy()) for m in mods
               for a, g in zip(ev._evaluation_images[m], gold["images"][m.name]))
    check("evaluator_images_equal_reference_run", same)

    # ---- method configs and install()
    tc = P.b200_thermal_nerf_config()
    check("method_config_model_swapped", type(tc.pipeline.model) is P.B200ThermalNerfModelConfig
          and tc.pipeline.model.eval_num_rays_per_chunk == 1 << 16 and tc.method_name == "b200-thermal-nerf")
    check("method_config_keeps_reference_settings", tc.mixed_precision is True and tc.max_num_iterations == 30000
          and tc.pipeline.datamanager.train_num_rays_per_batch == 4096
          and tc.optimizers["fields"]["optimizer"].lr == 1e-2 and tc.optimizers["fields"]["optimizer"].eps == 1e-15)
    nc = P.b200_thermalnerfacto_config()
    check("nerfacto_config_model_swapped", type(nc.pipeline.model) is P.B200ThermalNerfactoModelConfig
          and isinstance(nc.pipeline.model, ThermalNerfactoModelConfig))
    P.install()
    from thermo_nerf.thermal_nerf.config_thermal_nerf import thermal_nerf_config

    check("install_swaps_module_level_config", type(thermal_nerf_config.pipeline.model) is P.B200ThermalNerfModelConfig
          and isinstance(thermal_nerf_config.pipeline.model, ThermalNerfactoModelConfig))  # train_eval_script.py:94
    stock_cfg = ThermalNerfModelConfig(implementation="torch", **Wg.MINI)
    m2 = stock_cfg.setup(metadata={"thermal": []}, scene_box=box, num_train_data=Wg.NUM_IMAGES)
    check("install_upgrades_stock_configs", type(m2) is P.B200ThermalNerfModel)  # config.yml of a stock run
    m2.load_state_dict(stock.state_dict())
    check("stock_checkpoint_loads", True)
    # ---- nerfacto family: stock NerfactoField (no thermal modules) packs with constant-zero thermal tensors
    ncfg = P.B200ThermalNerfactoModelConfig(**Wg.MINI)
    m3 = ncfg.setup(scene_box=box, num_train_data=Wg.NUM_IMAGES)
    t = m3.tensors()
    check("nerfacto_family", isinstance(m3, ThermalNerfactoModel) and not isinstance(m3, ThermalNerfModel)
          and float(t.field_linears["th2"].weight.abs().sum()) == 0.0 and not m3._has_thermal_head()
          and not any("thermal" in k for k in m3.state_dict()))
    # ---- concat_nerf family: the reference's ConcatNerfModel / ConcatNerfactoTField / RGBTRenderer code builds the
    #      modules, the kernels get head_mode = concat and a [4, 64] last colour layer
    from thermo_nerf.rgb_concat.concat_nerfacto_model import ConcatNerfModel

    ccfg = P.B200ConcatNerfModelConfig(**Wg.MINI)
    m4 = ccfg.setup(scene_box=box, num_train_data=Wg.NUM_IMAGES)
    from thermo_nerf_b200 import _lib as L

    check("concat_family", isinstance(m4, ConcatNerfModel) and isinstance(m4, ThermalNerfactoModel) and m4._is_concat()
          and tuple(m4.tensors().field_linears["rgb2"].weight.shape) == (4, 64)
          and m4._render_kwargs()["head_mode"] == L.HEAD_CONCAT and not m4._has_thermal_head()
          and type(m4).get_image_metrics_and_images is ConcatNerfModel.get_image_metrics_and_images)
    from thermo_nerf.rgb_concat.config_concat_nerfacto import concat_nerf_config

    check("install_swaps_concat_config", type(concat_nerf_config.pipeline.model) is P.B200ConcatNerfModelConfig)


if __name__ == "__main__":
    try:
        main()
    finally:
        print(json.dumps(checks, indent=1))
    sys.exit(0 if checks and all(checks.values()) else 1)
