"""The oracle's composition against THE REFERENCE'S OWN model code.

tests/golden/reference_model_wiring.pt was produced by executing, unmodified, the reference's thermal_nerf_model.py
(constructor, populate_modules, get_outputs, get_loss_dict), thermal_field.py, thermal_field_head.py, thermal_renderer.py
and nerfacto_config/thermal_nerfacto.py over stand-ins for nerfstudio / torchmetrics that carry nerfstudio's
interfaces and the oracle's arithmetic (tests/golden/nerfstudio_standin.py, make_reference_wiring_golden.py).
Agreement here pins the reference's wiring - module construction, the colour head's input order, the temperature head
and its detach switch, renderer inputs, output keys, loss terms / multipliers / argument order, state-dict names - not
nerfstudio's arithmetic (that part of the oracle stays "parity unpinned")."""

from pathlib import Path

import pytest
import torch

from oracle import OracleConfig, OracleRays, OracleThermalNerf

GOLD = Path(__file__).parent / "golden" / "reference_model_wiring.pt"


@pytest.fixture(scope="module")
def blob():
    return torch.load(GOLD, weights_only=True)


def build_oracle(blob, case):
    mini = blob["mini"]
    cfg = OracleConfig(log2_hashmap_size=mini["log2_hashmap_size"],
                       num_proposal_samples_per_ray=tuple(mini["num_proposal_samples_per_ray"]),
                       num_nerf_samples_per_ray=mini["num_nerf_samples_per_ray"],
                       proposal_net_args_list=[{k: v for k, v in a.items() if k != "use_linear"}
                                               for a in mini["proposal_net_args_list"]],
                       pass_thermal_gradients=case["pass_thermal_gradients"])
    o = OracleThermalNerf(cfg, blob["num_images"], seed=0)
    o.load_state_dict(case["state_dict"], strict=True)
    return o


def rays_of(case):
    return OracleRays(case["origins"].clone(), case["directions"].clone(), case["camera_indices"].clone())


def test_constructor_contract_and_module_tree(blob):
    assert blob["missing_thermal_metadata_error"] == "Thermal images not found in metadata."  # thermal_nerf_model.py:77-78
    for case in blob["cases"]:
        # the reference's module tree carries exactly the oracle's parameter names and shapes
        assert case["load_unexpected"] == []
        assert set(case["load_missing"]) <= {"device_indicator_param"}  # nerfstudio Model's own bookkeeping tensor
        ref_keys = {k: tuple(v) for k, v in case["reference_state_dict_keys"].items() if k != "device_indicator_param"}
        assert ref_keys == {k: tuple(v.shape) for k, v in case["state_dict"].items()}
        # names of the thermo-nerf-owned modules (thermal_field.py:89-101)
        for k in ("field.mlp_thermal.layers.0.weight", "field.mlp_thermal.layers.1.bias", "field.field_head_thermal.net.weight"):
            assert k in ref_keys, k


def test_product_model_has_the_reference_state_dict_names(blob):
    from thermo_nerf_b200 import ThermalNerfModel, ThermalNerfModelConfig

    mini = blob["mini"]
    cfg = ThermalNerfModelConfig(log2_hashmap_size=mini["log2_hashmap_size"],
                                 num_proposal_samples_per_ray=tuple(mini["num_proposal_samples_per_ray"]),
                                 num_nerf_samples_per_ray=mini["num_nerf_samples_per_ray"],
                                 proposal_net_args_list=mini["proposal_net_args_list"])
    m = ThermalNerfModel(cfg, {"thermal": []}, torch.tensor([[-1.0, -1, -1], [1, 1, 1]]), blob["num_images"])
    ref = {k: tuple(v) for k, v in blob["cases"][0]["reference_state_dict_keys"].items()}
    ours = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    assert ours == ref
    with pytest.raises(ValueError, match="Thermal images not found in metadata"):
        ThermalNerfModel(cfg, {}, torch.tensor([[-1.0, -1, -1], [1, 1, 1]]), 1)


@pytest.mark.parametrize("idx", [0, 1, 2])
def test_training_outputs_losses_and_gradients(blob, idx):
    case = blob["cases"][idx]
    o = build_oracle(blob, case)
    tr = case["train"]
    o.anneal = tr["anneal"]
    out = o.get_outputs(rays_of(case), training=True, jitter=case["jitter"])
    # output keys: the reference's dict + nothing missing (the oracle adds diagnostic entries of its own)
    assert tr["output_keys"] == ["accumulation", "depth", "expected_depth", "prop_depth_0", "prop_depth_1",
                                 "ray_samples_list", "rgb", "thermal", "weights_list"]
    for k, v in tr["outputs"].items():
        assert out[k].shape == v.shape, k
        assert torch.allclose(out[k], v, atol=1e-6, rtol=1e-6), (k, (out[k] - v).abs().max())
    for a, b in zip(out["weights_list"], tr["weights_list"]):
        assert torch.allclose(a, b, atol=1e-7), (a - b).abs().max()
    for a, b in zip(out["sdist_list"], tr["spacing_bins"]):
        assert torch.allclose(a, b, atol=1e-7)
    loss = o.get_loss_dict(out, case["batch"]["image"], case["batch"]["thermal"], training=True)
    assert set(loss) == set(tr["loss"])
    assert ("thermal" in loss) == case["pass_thermal_gradients"]  # thermal_nerf_model.py:321-324
    for k, v in tr["loss"].items():
        assert torch.allclose(loss[k], v, atol=1e-7, rtol=1e-5), (k, loss[k], v)
    assert torch.allclose(loss["distortion_loss"], 0.002 * tr["distortion"], rtol=1e-6)
    o.zero_grad()
    sum(loss.values()).backward()
    grads = {k: p.grad for k, p in o.named_parameters()}
    for k, g in tr["grads"].items():
        assert grads[k] is not None, k
        scale = g.abs().max().clamp_min(1e-12)
        assert ((grads[k] - g).abs().max() / scale) < 1e-4, (k, (grads[k] - g).abs().max(), scale)
    for k, n in tr["grad_norms"].items():
        assert torch.allclose(grads[k].norm(), n, rtol=1e-4, atol=1e-10), k
    assert sorted(k for k, g in grads.items() if g is None) == [k for k in tr["params_without_grad"]
                                                                if k != "device_indicator_param"]
    if not case["pass_thermal_gradients"]:
        # the temperature head is built but receives no gradient at all when its loss is switched off
        assert any("mlp_thermal" in k for k in tr["params_without_grad"])
        assert any("field_head_thermal" in k for k in tr["params_without_grad"])


@pytest.mark.parametrize("idx", [0, 1, 2])
def test_eval_outputs(blob, idx):
    case = blob["cases"][idx]
    o = build_oracle(blob, case)
    ev = case["eval"]
    with torch.no_grad():
        out = o.get_outputs(rays_of(case), training=False)
        loss = o.get_loss_dict(out, case["batch"]["image"], case["batch"]["thermal"], training=False)
    assert ev["output_keys"] == ["accumulation", "depth", "expected_depth", "prop_depth_0", "prop_depth_1", "rgb", "thermal"]
    for k, v in ev["outputs"].items():
        assert torch.allclose(out[k], v, atol=1e-6, rtol=1e-6), (k, (out[k] - v).abs().max())
    if case["thermal_contrast"]:  # a temperature image with real structure, strictly inside the eval clamp
        t = ev["outputs"]["thermal"]
        assert 0.3 < float(t.min()) < float(t.max()) < 0.8 and float(t.max() - t.min()) > 0.1
    assert set(loss) == set(ev["loss"])
    for k, v in ev["loss"].items():
        assert torch.allclose(loss[k], v, atol=1e-7, rtol=1e-5), k


# ---- evaluation metrics (thermal_nerf_model.py:328-400, thermal_nerfacto.py:46-84) ----
def _marker_ssim(a, b):  # the marker callables of tests/golden/nerfstudio_standin.py (NOT ssim / lpips)
    return 1.0 - (a - 0.5 * b).abs().sum() / 1000.0


def _marker_lpips(a, b):
    assert a.shape[1] == 3 and b.shape[1] == 3
    return ((a - 0.25 * b) ** 2).sum() / 1000.0


def _product_model(blob, case, thermal_head=True):
    from thermo_nerf_b200 import ThermalNerfactoModel, ThermalNerfactoModelConfig, ThermalNerfModel, ThermalNerfModelConfig

    mini = blob["mini"]
    kw = dict(log2_hashmap_size=mini["log2_hashmap_size"],
              num_proposal_samples_per_ray=tuple(mini["num_proposal_samples_per_ray"]),
              num_nerf_samples_per_ray=mini["num_nerf_samples_per_ray"], proposal_net_args_list=mini["proposal_net_args_list"],
              max_temperature=case["max_temperature"], min_temperature=case["min_temperature"], cold=case["cold"])
    aabb = torch.tensor([[-1.0, -1, -1], [1, 1, 1]])
    if thermal_head:
        m = ThermalNerfModel(ThermalNerfModelConfig(**kw), {"thermal": []}, aabb, blob["num_images"])
    else:
        m = ThermalNerfactoModel(ThermalNerfactoModelConfig(**kw), aabb, blob["num_images"])
    m.ssim, m.lpips = _marker_ssim, _marker_lpips
    return m


def test_image_metrics_match_the_reference_method(blob):
    assert len(blob["image_metrics"]) == 4
    for case in blob["image_metrics"]:
        m = _product_model(blob, case)
        metrics, images = m.get_image_metrics_and_images(case["outputs"], case["batch"], threshold=case["threshold"])
        ref = case["metrics"]
        assert set(metrics) == set(ref), (sorted(metrics), sorted(ref))
        for k, v in ref.items():
            assert metrics[k] == pytest.approx(v, rel=2e-5, abs=1e-6), (k, case["cold"], case["threshold"])
        # the colour-image MAE entries come through super() without the threshold (thermal_nerf_model.py:339)
        assert ref["mae_foreground"] == ref["mae"]
        assert set(images) == set(case["image_keys"])
        for k, shp in case["image_shapes"].items():
            assert list(images[k].shape) == shp, k
        assert torch.equal(images["thermal"], case["thermal_image"])
        assert torch.equal(images["thermal_combined"], case["thermal_combined"])


def test_image_metrics_of_the_nerfacto_track_variant(blob):
    for case in blob["image_metrics"]:
        m = _product_model(blob, case, thermal_head=False)
        metrics, images = m.get_image_metrics_and_images(case["outputs"], case["batch"], threshold=case["threshold"])
        ref = case["metrics_nerfacto_track"]
        assert set(metrics) == set(ref)
        for k, v in ref.items():
            assert metrics[k] == pytest.approx(v, rel=2e-5, abs=1e-6), (k, case["cold"], case["threshold"])
        assert set(images) == set(case["image_keys_nerfacto_track"])


# ---- Evaluator (thermo_nerf/evaluator/evaluator.py:15-175 executed from the reference) ----
def test_evaluator_mirror_reproduces_the_reference_evaluator(tmp_path):
    """The reference's Evaluator ran on its own model (tests/golden/make_reference_evaluator_golden.py); here the
    product's Evaluator + ThermalNerfModel.get_image_metrics_and_images replay the same per-frame model outputs and must
    produce the same aggregated metrics, metrics.json, file tree and uint8 evaluation images."""
    import json
    from types import SimpleNamespace

    import numpy as np

    from thermo_nerf_b200 import Evaluator, PinholeCameras, RenderedImageModality, ThermalNerfModel, ThermalNerfModelConfig

    gold = torch.load(Path(__file__).parent / "golden" / "reference_evaluator.pt", weights_only=True)
    H, W = gold["hw"]
    args = [{"hidden_dim": 16, "log2_hashmap_size": 8, "num_levels": 5, "max_res": 128, "use_linear": False}] * 2
    cfg = ThermalNerfModelConfig(log2_hashmap_size=8, proposal_net_args_list=args, max_temperature=gold["max_temperature"],
                                 min_temperature=gold["min_temperature"])
    model = ThermalNerfModel(cfg, {"thermal": []}, torch.tensor([[-1.0, -1, -1], [1, 1, 1]]), 4)
    model.eval()
    model.ssim, model.lpips = _marker_ssim, _marker_lpips
    frames = iter(gold["frames"])
    model.get_outputs_for_camera_ray_bundle = lambda bundle: next(frames)
    loader = [(PinholeCameras(gold["camera_to_worlds"][i:i + 1], gold["focal"], gold["focal"], W / 2, H / 2, W, H), b)
              for i, b in enumerate(gold["batches"])]
    pipeline = SimpleNamespace(model=model, datamanager=SimpleNamespace(setup_eval=lambda: None,
                                                                        fixed_indices_eval_dataloader=loader))
    config = SimpleNamespace(experiment_name="double_robot", method_name="thermal-nerf")
    mods = [RenderedImageModality[n] for n in gold["modalities"]]
    ev = Evaluator(pipeline, config, job_param_identifier=gold["identifier"], modalities_to_save=mods,
                   threshold=gold["threshold"])
    # aggregated metrics: same keys (<k>, <k>_mean, <k>_std), same numbers
    assert set(ev.metrics) == set(gold["metrics"])
    for k, v in gold["metrics"].items():
        assert ev.metrics[k] == pytest.approx(v, rel=2e-5, abs=1e-6), k
    info = ev._benchmark_info
    assert {k: info[k] for k in ("experiment_name", "method_name", "job_param_identifier")} == \
        {k: gold["benchmark_info"][k] for k in ("experiment_name", "method_name", "job_param_identifier")}
    # files
    ev.save_metrics(tmp_path)
    ev.save_images(mods, tmp_path)
    files = sorted(str(p.relative_to(tmp_path)) for p in tmp_path.rglob("*") if p.is_file())
    assert files == gold["files"]
    for name, text in gold["texts"].items():
        ours, ref = json.loads((tmp_path / name).read_text()), json.loads(text)
        if name == "metrics.json":
            assert list(ours) == list(ref) and set(ours["results"]) == set(ref["results"])
            for k, v in ref["results"].items():
                assert ours["results"][k] == pytest.approx(v, rel=2e-5, abs=1e-6), k
        else:
            assert ours == pytest.approx(ref, rel=2e-5, abs=1e-6), name
    # uint8 evaluation images, bit for bit
    for m in mods:
        assert len(ev._evaluation_images[m]) == len(gold["images"][m.name])
        for a, b in zip(ev._evaluation_images[m], gold["images"][m.name]):
            assert a.dtype == np.uint8 and np.array_equal(a, b.numpy()), m
    # defaults of the constructor: RGB only, and without an identifier only metrics.json is written
    assert gold["default_modalities"] == ["RGB"] and gold["default_files"] == ["metrics.json"]


# ---- method configuration (config_thermal_nerf.py:17-49 imported from the reference with recording stand-ins) ----
def test_defaults_follow_the_reference_method_config():
    import inspect
    import json

    from thermo_nerf_b200 import ThermalNerfModelConfig
    from thermo_nerf_b200.data import DevicePixelSampler
    from thermo_nerf_b200.engine import TrainEngine, exponential_decay_lr

    ref = json.loads((Path(__file__).parent / "golden" / "reference_method_configs.json").read_text())["thermal_nerf_config"]
    assert ref["method_name"] == "thermal-nerf" and ref["max_num_iterations"] == 30000 and ref["mixed_precision"] is True
    model = ref["pipeline"]["model"]
    cfg = ThermalNerfModelConfig()
    for k in ("camera_optimizer_mode", "cold", "eval_num_rays_per_chunk", "max_temperature", "min_temperature",
              "pass_thermal_gradients", "use_transient_embedding"):
        assert getattr(cfg, k) == model[k], k
    assert model["thermal_loss_weight"] == 1.0  # declared by the reference but never read (thermal_nerf_model.py:53)
    # rays per batch: the benchmark workload and the device pixel sampler's default
    rays = ref["pipeline"]["datamanager"]["train_num_rays_per_batch"]
    assert rays == 4096 == inspect.signature(DevicePixelSampler.next_train).parameters["num_rays"].default
    # optimisers: the same Adam + exponential decay for both parameter groups
    eng = inspect.signature(TrainEngine.__init__).parameters
    for group in ("proposal_networks", "fields"):
        opt, sch = ref["optimizers"][group]["optimizer"], ref["optimizers"][group]["scheduler"]
        assert opt == {"_class": "AdamOptimizerConfig", "lr": 0.01, "eps": 1e-15}
        assert (eng["lr"].default, eng["eps"].default) == (opt["lr"], opt["eps"])
        assert (eng["lr_final"].default, eng["lr_max_steps"].default) == (sch["lr_final"], sch["max_steps"])
    assert set(ref["optimizers"]) == {"proposal_networks", "fields"}  # no camera_opt group in this method config
    assert exponential_decay_lr(0) == pytest.approx(0.01) and exponential_decay_lr(200000) == pytest.approx(1e-4)
    assert exponential_decay_lr(100000) == pytest.approx(1e-3)  # log-linear interpolation
