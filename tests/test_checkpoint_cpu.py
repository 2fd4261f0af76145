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
