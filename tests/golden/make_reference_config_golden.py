"""The reference's method configurations, read by IMPORTING its config modules in the build container
(thermo_nerf/thermal_nerf/config_thermal_nerf.py:17-49 and thermo_nerf/nerfacto_config/config_nerfacto.py:14-53) with
recording stand-ins for the nerfstudio config classes: every constructor keeps its keyword arguments, so the tree below is
exactly what the reference passes (rays per batch, chunk sizes, optimiser and scheduler settings, precision flag).

    python tests/golden/make_reference_config_golden.py        # needs /root/reference

Writes tests/golden/reference_method_configs.json."""

import dataclasses
import importlib.abc
import importlib.machinery
import json
import sys
import types
from pathlib import Path

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parents[1]))
sys.path.insert(0, str(HERE))
OUT = HERE / "reference_method_configs.json"


class Recorded:
    def __init__(self, *args, **kwargs) -> None:
        self._args, self._kwargs = args, kwargs

    def __class_getitem__(cls, item):
        return cls


class _RecordingModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        cls = type(name, (Recorded,), {})
        setattr(self, name, cls)
        return cls


class _Finder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    PREFIXES = ("nerfstudio", "mlflow", "tyro", "torchmetrics", "nerfacc", "imageio", "matplotlib", "cv2", "flirimageextractor")

    def find_spec(self, fullname, path, target=None):
        if fullname.split(".")[0] in self.PREFIXES and fullname not in sys.modules:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        return _RecordingModule(spec.name)

    def exec_module(self, module):
        module.__path__ = []


def tree(x):
    if dataclasses.is_dataclass(x) and not isinstance(x, type):
        # the reference's own config dataclasses: the fields they declare themselves, plus inherited (stand-in) fields
        # only where the reference changed the value - stand-in defaults are not reference data
        own = set()
        for c in type(x).__mro__:
            if c.__module__.startswith("thermo_nerf"):
                own |= set(vars(c).get("__annotations__", {}))
        d = {"_class": type(x).__name__}
        for f in dataclasses.fields(x):
            if f.name == "_target":
                continue
            default = f.default if f.default is not dataclasses.MISSING else (
                f.default_factory() if f.default_factory is not dataclasses.MISSING else dataclasses.MISSING)
            value = getattr(x, f.name)
            if f.name in own or not (default is not dataclasses.MISSING and value == default):
                d[f.name] = tree(value)
        return d
    if isinstance(x, Recorded):
        d = {"_class": type(x).__name__}
        if x._args:
            d["_args"] = [tree(a) for a in x._args]
        d.update({k: tree(v) for k, v in x._kwargs.items() if k != "_target"})
        return d
    if isinstance(x, dict):
        return {str(k): tree(v) for k, v in x.items()}
    if isinstance(x, (list, tuple)):
        return [tree(v) for v in x]
    if isinstance(x, (int, float, str, bool)) or x is None:
        return x
    return repr(x)


def main() -> None:
    import nerfstudio_standin as S

    S.install()                      # functional stand-ins for what the model modules need ...
    sys.meta_path.append(_Finder())  # ... and recording ones for every other third-party name
    for name, m in list(sys.modules.items()):
        if name.split(".")[0] in _Finder.PREFIXES and type(m) is types.ModuleType:
            m.__class__ = _RecordingModule  # names the functional stand-ins lack are recorded as well

    @dataclasses.dataclass
    class VanillaPipelineConfig:  # the reference subclasses this one as a dataclass (pipeline_tracking.py:10-17)
        _target: type = None
        datamanager: object = None
        model: object = None

    sys.modules["nerfstudio.pipelines.base_pipeline"].VanillaPipelineConfig = VanillaPipelineConfig
    import importlib

    # nerfstudio InputDataset keeps this class attribute, which the reference extends (thermal_dataset.py:18-20)
    importlib.import_module("nerfstudio.data.datasets.base_dataset").InputDataset = type(
        "InputDataset", (Recorded,), {"exclude_batch_keys_from_device": ["image", "mask"]})
    del sys.modules["nerfstudio.engine.trainer"].TrainerConfig  # the annotation-only placeholder: record it instead
    sys.path.append("/root/reference")
    from thermo_nerf.nerfacto_config.config_nerfacto import thermalnerfacto_config
    from thermo_nerf.thermal_nerf.config_thermal_nerf import thermal_nerf_config

    blob = {"thermal_nerf_config": tree(thermal_nerf_config), "thermalnerfacto_config": tree(thermalnerfacto_config),
            "source": "thermo_nerf/thermal_nerf/config_thermal_nerf.py and thermo_nerf/nerfacto_config/config_nerfacto.py "
                      "imported from /root/reference with recording stand-ins for the nerfstudio config classes"}
    OUT.write_text(json.dumps(blob, indent=1, sort_keys=True))
    print(json.dumps(blob["thermal_nerf_config"], indent=1)[:3000])


if __name__ == "__main__":
    sys.exit(main())
