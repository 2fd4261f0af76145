"""nerfstudio checkpoint envelope compatibility.  The envelope shape asserted here is the
one of the reference fixture tests/data/vanilla_nerf/training-job/vanilla-nerf/date/nerfstudio_models/
test_pipeline.ckpt: keys {'step','pipeline','optimizers','scalers'}, model tensors under '_model.'."""

import pytest
import torch

from oracle import OracleConfig, OracleThermalNerf
from thermo_nerf_b200 import ThermalNerfModel, ThermalNerfModelConfig
from thermo_nerf_b200.checkpoint import extract_model_state, load_nerfstudio_checkpoint, save_nerfstudio_checkpoint


def _small():
    args = [{"hidden_dim": 16, "log2_hashmap_size": 10, "num_levels": 5, "max_res": 128, "use_linear": False},
            {"hidden_dim": 16, "log2_hashmap_size": 10, "num_levels": 5, "max_res": 256, "use_linear": False}]
    cfg = ThermalNerfModelConfig(log2_hashmap_size=10, proposal_net_args_list=args)
    model = ThermalNerfModel(cfg, {"thermal": []}, torch.tensor([[-1.0, -1, -1], [1, 1, 1]]), 6)
    ocfg = OracleConfig(log2_hashmap_size=10, proposal_net_args_list=[{k: v for k, v in a.items() if k != "use_linear"}
                                                                       for a in args])
    return model, OracleThermalNerf(ocfg, 6, seed=3)


@pytest.mark.parametrize("prefix", ["_model.", "_model.module.", "module._model."])
def test_load_reference_style_envelope(prefix):
    model, trained = _small()
    src = trained.state_dict()
    envelope = {"step": 1234, "pipeline": {prefix + k: v.clone() for k, v in src.items()}, "optimizers": {"fields": {}},
                "scalers": {}}
    envelope["pipeline"]["datamanager.train_ray_generator.image_coords"] = torch.zeros(2, 2, 2)  # not the model's
    step = load_nerfstudio_checkpoint(model, envelope)
    assert step == 1234 and model._step == 1234
    got = model.state_dict()
    for k, v in src.items():
        assert torch.equal(got[k], v), k


def test_key_aliases_and_shape_errors(tmp_path):
    model, trained = _small()
    src = {("field.mlp_base.model.0.hash_table" if k == "field.mlp_base.encoder.hash_table" else k): v
           for k, v in trained.state_dict().items()}
    src = {k.replace("field.mlp_base.mlp.layers", "field.mlp_base.model.1.layers"): v for k, v in src.items()}
    load_nerfstudio_checkpoint(model, {"step": 7, "pipeline": {"_model." + k: v for k, v in src.items()}})
    assert torch.equal(model.field.mlp_base.encoder.hash_table, trained.field.mlp_base.encoder.hash_table)
    assert torch.equal(model.field.mlp_base.mlp.layers[1].weight, trained.field.mlp_base.mlp.layers[1].weight)
    bad = {"_model." + k: v for k, v in trained.state_dict().items()}
    bad["_model.field.mlp_head.layers.0.weight"] = torch.zeros(64, 62)
    with pytest.raises(ValueError):
        load_nerfstudio_checkpoint(model, {"step": 0, "pipeline": bad})
    with pytest.raises(NotImplementedError):
        load_nerfstudio_checkpoint(model, {"step": 0, "pipeline": {"_model.field.mlp_base.params": torch.zeros(8)}})
    with pytest.raises(KeyError):
        extract_model_state({"step": 0, "pipeline": {"datamanager.x": torch.zeros(1)}})
    missing = {k: v for k, v in bad.items() if "mlp_thermal" not in k and "mlp_head.layers.0.weight" not in k}
    with pytest.raises(KeyError):
        load_nerfstudio_checkpoint(model, {"step": 0, "pipeline": missing})
    load_nerfstudio_checkpoint(model, {"step": 0, "pipeline": missing}, strict=False)


def test_save_round_trip_with_optimizer_state(tmp_path):
    model, trained = _small()
    model.load_state_dict(trained.state_dict(), strict=False)
    opt = torch.optim.Adam(model.field.parameters(), lr=1e-2, eps=1e-15)
    path = tmp_path / "nerfstudio_models" / "step-000000042.ckpt"
    save_nerfstudio_checkpoint(model, path, 42, optimizers={"fields": opt})
    blob = torch.load(path, weights_only=False)
    assert set(blob) == {"step", "pipeline", "optimizers", "scalers"} and blob["step"] == 42
    assert all(k.startswith("_model.") for k in blob["pipeline"])
    assert set(blob["optimizers"]) == {"fields"}
    other, _ = _small()
    assert load_nerfstudio_checkpoint(other, path) == 42
    for (k, a), (_, b) in zip(sorted(model.state_dict().items()), sorted(other.state_dict().items())):
        assert torch.equal(a, b), k


REFERENCE_FIXTURE = "/root/reference/tests/data/vanilla_nerf/training-job/vanilla-nerf/date/nerfstudio_models/test_pipeline.ckpt"


def test_foreign_keys_of_real_checkpoints_are_skipped_in_strict_mode():
    """A genuine nerfstudio checkpoint carries metric networks (`_model.lpips.net.*`), derived `hash_offset` buffers and
    the proposal encoder under two names (`encoding.*` and `mlp_base.0.*`): strict mode accepts those and still fails
    on a missing or unexpected key of a module the model owns."""
    model, trained = _small()
    pipe = {"_model." + k: v.clone() for k, v in trained.state_dict().items()}
    for i in range(20):  # the key pattern of the reference fixture
        pipe[f"_model.lpips.net.lin{i % 5}.model.1.weight"] = torch.zeros(1, 8, 1, 1)
    pipe["_model.field.mlp_base.encoder.hash_offset"] = torch.arange(16)
    for i in range(2):
        pipe[f"_model.proposal_networks.{i}.encoding.hash_offset"] = torch.arange(5)
        pipe[f"_model.proposal_networks.{i}.mlp_base.0.hash_offset"] = torch.arange(5)
        pipe[f"_model.proposal_networks.{i}.mlp_base.0.hash_table"] = pipe[f"_model.proposal_networks.{i}.encoding.hash_table"]
        pipe[f"_model.proposal_networks.{i}.mlp_base.0.scalings"] = pipe[f"_model.proposal_networks.{i}.encoding.scalings"]
    assert load_nerfstudio_checkpoint(model, {"step": 9, "pipeline": pipe}, strict=True) == 9
    assert torch.equal(model.proposal_networks[1].encoding.hash_table, trained.proposal_networks[1].encoding.hash_table)
    pipe["_model.field.mlp_extra.weight"] = torch.zeros(2)  # an unexpected key of an owned module still fails
    with pytest.raises(KeyError):
        load_nerfstudio_checkpoint(model, {"step": 9, "pipeline": pipe}, strict=True)


@pytest.mark.skipif(not __import__("os").path.exists(REFERENCE_FIXTURE), reason="needs the reference checkout")
def test_envelope_of_the_reference_fixture_checkpoint():
    """The reference's own fixture (a vanilla-NeRF run, so its field keys are not this model's): the envelope, the
    `_model.` prefix and the step are read from the real file, and its LPIPS keys are the ones strict mode skips."""
    from thermo_nerf_b200.checkpoint import _foreign

    state, step = extract_model_state(REFERENCE_FIXTURE)
    assert step == 1 and len(state) == 69
    lp = [k for k in state if k.startswith("lpips.")]
    assert len(lp) == 20 and all(_foreign(k) for k in lp)
    assert any(k.startswith("field_coarse.mlp_base.layers.") for k in state)
    model, _ = _small()
    with pytest.raises(KeyError):  # a vanilla-NeRF checkpoint is not a ThermoNeRF one
        load_nerfstudio_checkpoint(model, REFERENCE_FIXTURE)
