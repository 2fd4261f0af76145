"""The nerfstudio plugin (thermo_nerf_b200/nerfstudio_plugin.py) under the REFERENCE's own callers.

``tests/plugin_reference_check.py`` runs in a subprocess (it installs nerfstudio stand-ins into sys.modules): it
imports the reference from /root/reference, builds ``B200ThermalNerfModel`` - a subclass of the reference's
``ThermalNerfModel`` - through ``config.setup`` and drives the reference's unmodified ``Renderer.render`` and
``Evaluator`` with it (uint8 frames, metrics and files equal to the committed runs of the stock model), checks the
``isinstance`` assertions of train_eval_script.py:94 / evaluator.py:76, the method configs, ``install()`` and the
nerfacto family.  Skipped where /root/reference does not exist (the GPU box)."""

import json
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]


@pytest.mark.skipif(not Path("/root/reference/thermo_nerf").exists(), reason="needs the reference checkout")
def test_reference_entry_points_run_on_the_plugin_model():
    r = subprocess.run([sys.executable, str(ROOT / "tests" / "plugin_reference_check.py")], capture_output=True,
                       text=True, timeout=600, cwd=str(ROOT))
    start = r.stdout.rfind("\n{") + 1 if "\n{" in r.stdout else r.stdout.find("{")
    checks = json.loads(r.stdout[start:])
    failed = [k for k, v in checks.items() if not v]
    assert r.returncode == 0 and not failed, f"failed checks: {failed}\n{r.stderr[-2000:]}"
    assert len(checks) >= 30


def test_plugin_is_import_guarded():
    """Without nerfstudio / thermo_nerf the module imports, says why it is unavailable, and its entry points raise
    ImportError (no silent fallback to the stand-alone classes)."""
    import thermo_nerf_b200.nerfstudio_plugin as P
    from thermo_nerf_b200 import Renderer

    if P.AVAILABLE:
        pytest.skip("the reference is importable in this interpreter")
    assert isinstance(P.IMPORT_ERROR, ImportError)
    for fn in (P.install, P.b200_thermal_nerf_config, P.b200_thermalnerfacto_config, P.method_specification):
        with pytest.raises(ImportError, match="thermo_nerf"):
            fn()
    with pytest.raises(ImportError):
        Renderer.from_pipeline_path(Path("."), Path("."))


def test_plugin_classes_over_foreign_bases():
    """make_plugin_classes: the mixin's methods win over the base's, the config keeps the base's fields."""
    from dataclasses import dataclass

    from thermo_nerf_b200.model import KernelModelMixin, ThermalNerfModelConfig
    from thermo_nerf_b200.nerfstudio_plugin import make_plugin_classes, upgrade_config

    class Base:  # stands for the reference's ThermalNerfModel
        def get_outputs(self, rb):
            raise AssertionError("the reference's eager path must not run")

        def get_image_metrics_and_images(self, *a, **k):
            return "reference"

    M, C = make_plugin_classes(Base, ThermalNerfModelConfig, name="X")
    assert M.get_outputs is KernelModelMixin.get_outputs and M.get_loss_dict is KernelModelMixin.get_loss_dict
    assert M.get_image_metrics_and_images is Base.get_image_metrics_and_images
    assert issubclass(C, ThermalNerfModelConfig) and C().precision == "tc_fp16" and C()._target is M
    up = upgrade_config(ThermalNerfModelConfig(max_temperature=40.0, cold=True), C)
    assert type(up) is C and up.max_temperature == 40.0 and up.cold is True and up._target is M
