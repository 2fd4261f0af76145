"""The drop-in binding: subclasses of the REFERENCE's own classes, for an installation that has nerfstudio and
thermo-nerf (``pip install thermo-nerf``), so that ``train_eval_script.py`` and ``render_video_script.py`` run
unchanged on the sm_100a kernels.

* ``B200ThermalNerfModel(ThermalNerfModel)``            <- thermo_nerf/thermal_nerf/thermal_nerf_model.py:60-393
  ``B200ThermalNerfModelConfig(ThermalNerfModelConfig)`` <- :46-56 (hence a ``ThermalNerfactoModelConfig``, which
  ``train_eval_script.py:94`` asserts, and ``isinstance(model, ThermalNerfactoModel)`` of ``evaluator.py:76`` holds)
* ``B200ThermalNerfactoModel`` / ``...Config``           <- thermo_nerf/nerfacto_config/thermal_nerfacto.py (the
  ``nerfacto`` / ``thermal-nerfacto`` model types of train_eval_script.py:66-73)
* ``B200ConcatNerfModel`` / ``...Config``                <- thermo_nerf/rgb_concat/concat_nerfacto_model.py (the
  ``concat_nerf`` model type, train_eval_script.py:74-78)
* ``b200_thermal_nerf_config`` / ``b200_thermalnerfacto_config``: the reference's method configs
  (thermal_nerf/config_thermal_nerf.py:17-48, nerfacto_config/config_nerfacto.py:14-53) with the model swapped
* ``install()``: swaps the model configs inside the reference's module-level ``TrainerConfig`` objects and makes
  ``ThermalNerfModelConfig.setup`` build the B200 class, so the unchanged scripts - and ``config.yml`` files of
  runs trained with the stock model - construct this model; ``python -m thermo_nerf_b200.run <script module> ...``
  does that and then runs the script.

``populate_modules`` lets the reference build its modules exactly as it always does (nerfstudio's
``implementation="torch"`` parameter layout: state_dict keys and checkpoints are unchanged), then
``get_outputs`` / ``get_outputs_for_camera_ray_bundle`` / ``get_metrics_dict`` / ``get_loss_dict`` go to libtnf_b200
(``KernelModelMixin``, the same code the stand-alone classes of ``model.py`` use) and the field / proposal-network
instances get the kernel-backed ``get_density`` / ``get_outputs`` / ``forward`` / ``density_fn`` of ``surface.py``.
Everything else - ``get_param_groups``, ``get_training_callbacks`` (annealing, ``step_cb``),
``get_image_metrics_and_images`` (torchmetrics / LPIPS), the camera optimiser - stays the reference's / nerfstudio's.

The classes are built by ``make_plugin_classes(ref_model, ref_config)`` so that the test-suite can build them over
stand-ins of the reference's base classes where nerfstudio is not installable; at import time they are built over the
real ones when ``thermo_nerf`` imports.
"""

from __future__ import annotations

import copy
import types
from dataclasses import dataclass, field, fields, is_dataclass
from typing import Any, Dict, Literal, Optional, Tuple, Type

from .model import KernelModelMixin, check_supported_config


def _bind_field_surface(model) -> None:
    """Kernel-backed Field surface on the nerfstudio module instances the reference built."""
    from . import surface

    f = model.field
    for name, fn in (("get_density", surface.field_get_density), ("get_outputs", surface.field_get_outputs),
                     ("forward", surface.field_forward), ("density_fn", surface.density_fn)):
        object.__setattr__(f, name, types.MethodType(fn, f))
    for net in model.proposal_networks:
        object.__setattr__(net, "density_fn", types.MethodType(surface.density_fn, net))
    # the reference captured the class methods at thermal_nerf_model.py:127-148: rebuild the list
    model.density_fns = [net.density_fn for net in model.proposal_networks]


def make_plugin_classes(ref_model: type, ref_config: type, *, name: str = "B200ThermalNerfModel",
                        field_head_names: Any = None, field_head_names_t: Any = None,
                        thermal_head: bool = True, concat_head: bool = False) -> Tuple[type, type]:
    """(model class, config class) deriving from the given reference classes.  ``thermal_head=False`` for the
    ThermalNerfactoModel family, whose field is a stock NerfactoField (no "thermal" output, no thermal loss);
    ``concat_head=True`` for ConcatNerfModel (4-channel RGBT colour head, rgb_concat/)."""
    has_thermal_head = bool(thermal_head) and not concat_head
    is_concat = bool(concat_head)

    class _Model(KernelModelMixin, ref_model):  # type: ignore[misc, valid-type]
        def populate_modules(self) -> None:
            check_supported_config(self.config)
            self.config.implementation = "torch"  # nerfstudio's torch parameter layout (no tcnn blobs)
            super().populate_modules()
            if field_head_names is not None:
                self.field._field_head_names = field_head_names
            if field_head_names_t is not None:
                self.field._field_head_names_t = field_head_names_t
            _bind_field_surface(self)
            self.field.thermal_head = has_thermal_head
            if not has_thermal_head:
                self.field.pass_thermal_gradients = False
                if not hasattr(self.field, "pass_rgb_gradients"):
                    self.field.pass_rgb_gradients = True
            self._tensors = None

        def _has_thermal_head(self) -> bool:
            return has_thermal_head

        def _is_concat(self) -> bool:
            return is_concat

    _Model.__name__ = _Model.__qualname__ = name
    _Model.__doc__ = f"{ref_model.__name__} on libtnf_b200 (see thermo_nerf_b200.nerfstudio_plugin)."

    @dataclass
    class _Config(ref_config):  # type: ignore[misc, valid-type]
        _target: Type = field(default_factory=lambda: _Model)
        precision: Literal["fp32", "tc_fp16"] = "tc_fp16"
        """fp32: exact-arithmetic kernels; tc_fp16: tensor-core kernels (fp16 forward / bf16 backward operands)."""

    _Config.__name__ = _Config.__qualname__ = name + "Config"
    _Config.__doc__ = f"{ref_config.__name__} whose _target is {name}."
    return _Model, _Config


def upgrade_config(cfg, config_cls):
    """A ``config_cls`` instance carrying every field of the reference config ``cfg`` (its ``_target`` excepted)."""
    if isinstance(cfg, config_cls):
        return cfg
    kw = {f.name: copy.deepcopy(getattr(cfg, f.name)) for f in fields(cfg) if f.name != "_target" and f.init}
    return config_cls(**kw)


AVAILABLE = False
IMPORT_ERROR: Optional[BaseException] = None
try:  # the real reference (which imports nerfstudio at module top)
    from thermo_nerf.nerfacto_config.thermal_nerfacto import (  # type: ignore[import-not-found]
        ThermalNerfactoModel as _RefNerfactoModel,
        ThermalNerfactoModelConfig as _RefNerfactoConfig,
    )
    from thermo_nerf.thermal_nerf.thermal_nerf_model import (  # type: ignore[import-not-found]
        ThermalNerfModel as _RefModel,
        ThermalNerfModelConfig as _RefConfig,
    )

    AVAILABLE = True
except Exception as e:  # nerfstudio / thermo_nerf not installed (this build container, the GPU test box)
    IMPORT_ERROR = e

if AVAILABLE:
    from nerfstudio.field_components.field_heads import FieldHeadNames as _FHN  # type: ignore[import-not-found]
    from thermo_nerf.thermal_nerf.thermal_field_head import FieldHeadNamesT as _FHNT  # type: ignore[import-not-found]

    B200ThermalNerfModel, B200ThermalNerfModelConfig = make_plugin_classes(
        _RefModel, _RefConfig, name="B200ThermalNerfModel", field_head_names=_FHN, field_head_names_t=_FHNT)
    B200ThermalNerfactoModel, B200ThermalNerfactoModelConfig = make_plugin_classes(
        _RefNerfactoModel, _RefNerfactoConfig, name="B200ThermalNerfactoModel", field_head_names=_FHN,
        thermal_head=False)
    _plugin_classes = [B200ThermalNerfModel, B200ThermalNerfModelConfig, B200ThermalNerfactoModel,
                       B200ThermalNerfactoModelConfig]
    try:  # the concat_nerf baseline (imports nerfacc through rgbt_renderer.py)
        from thermo_nerf.rgb_concat.concat_nerfacto_model import (  # type: ignore[import-not-found]
            ConcatNerfModel as _RefConcatModel,
            ConcatNerfModelConfig as _RefConcatConfig,
        )

        B200ConcatNerfModel, B200ConcatNerfModelConfig = make_plugin_classes(
            _RefConcatModel, _RefConcatConfig, name="B200ConcatNerfModel", field_head_names=_FHN, concat_head=True)
        _plugin_classes += [B200ConcatNerfModel, B200ConcatNerfModelConfig]
    except Exception:  # pragma: no cover - nerfacc missing
        _RefConcatModel = _RefConcatConfig = None
    # pickled-YAML config.yml files name the classes by module path: keep them importable from here
    for _c in _plugin_classes:
        _c.__module__ = __name__


def _require() -> None:
    if not AVAILABLE:
        raise ImportError(
            "thermo_nerf_b200.nerfstudio_plugin needs the reference package (thermo_nerf) and nerfstudio: "
            f"{IMPORT_ERROR!r}.  Without them use the stand-alone classes of thermo_nerf_b200.model.")


def b200_thermal_nerf_config():
    """thermal_nerf_config (thermal_nerf/config_thermal_nerf.py:17-48) with the model swapped: same datamanager,
    optimisers (Adam 1e-2 / eps 1e-15, exponential decay to 1e-4 over 200k steps), mixed precision, 4096 rays per
    batch, eval chunk 1 << 16."""
    _require()
    from thermo_nerf.thermal_nerf.config_thermal_nerf import thermal_nerf_config  # type: ignore[import-not-found]

    cfg = copy.deepcopy(thermal_nerf_config)
    cfg.method_name = "b200-thermal-nerf"
    cfg.pipeline.model = upgrade_config(cfg.pipeline.model, B200ThermalNerfModelConfig)
    return cfg


def b200_thermalnerfacto_config():
    """thermalnerfacto_config (nerfacto_config/config_nerfacto.py:14-53: the ``nerfacto`` / ``thermal-nerfacto``
    model types of train_eval_script.py:66-73) with the model swapped."""
    _require()
    from thermo_nerf.nerfacto_config.config_nerfacto import thermalnerfacto_config  # type: ignore[import-not-found]

    cfg = copy.deepcopy(thermalnerfacto_config)
    cfg.method_name = "b200-" + str(cfg.method_name)
    cfg.pipeline.model = upgrade_config(cfg.pipeline.model, B200ThermalNerfactoModelConfig)
    return cfg


_INSTALLED = False


def install() -> None:
    """Route the reference's own entry points to the B200 model, in place:

    * ``thermal_nerf_config.pipeline.model`` (the object ``train_eval_script.py:59`` assigns to ``parameters.model``)
      becomes a ``B200ThermalNerfModelConfig`` with the same field values;
    * ``ThermalNerfModelConfig.setup`` builds ``B200ThermalNerfModel``, so a ``config.yml`` written by a run of the
      stock model (``Renderer.extract_pipeline``, render/renderer.py:70-115, unpickles it and calls
      ``config.pipeline.setup``) loads its checkpoint into the B200 model - state_dict keys are identical."""
    global _INSTALLED
    _require()
    if _INSTALLED:
        return
    from thermo_nerf.thermal_nerf import config_thermal_nerf as ref_cfg_mod  # type: ignore[import-not-found]

    from thermo_nerf.nerfacto_config import config_nerfacto as ref_nerfacto_mod  # type: ignore[import-not-found]

    ref_cfg_mod.thermal_nerf_config.pipeline.model = upgrade_config(ref_cfg_mod.thermal_nerf_config.pipeline.model,
                                                                    B200ThermalNerfModelConfig)
    ref_nerfacto_mod.thermalnerfacto_config.pipeline.model = upgrade_config(
        ref_nerfacto_mod.thermalnerfacto_config.pipeline.model, B200ThermalNerfactoModelConfig)
    if _RefConcatConfig is not None:
        from thermo_nerf.rgb_concat import config_concat_nerfacto as ref_concat_mod  # type: ignore[import-not-found]

        ref_concat_mod.concat_nerf_config.pipeline.model = upgrade_config(
            ref_concat_mod.concat_nerf_config.pipeline.model, B200ConcatNerfModelConfig)
    stock_setup = _RefNerfactoConfig.setup

    def setup(self, **kwargs):
        if type(self) is _RefConfig:
            return upgrade_config(self, B200ThermalNerfModelConfig).setup(**kwargs)
        if type(self) is _RefNerfactoConfig:
            return upgrade_config(self, B200ThermalNerfactoModelConfig).setup(**kwargs)
        if _RefConcatConfig is not None and type(self) is _RefConcatConfig:
            return upgrade_config(self, B200ConcatNerfModelConfig).setup(**kwargs)
        return stock_setup(self, **kwargs)

    _RefNerfactoConfig.setup = setup  # ThermalNerfModelConfig inherits it
    _INSTALLED = True


def method_specification():
    """``nerfstudio.plugins.types.MethodSpecification`` for ``ns-train b200-thermal-nerf`` (entry point
    ``nerfstudio.method_configs``)."""
    _require()
    from nerfstudio.plugins.types import MethodSpecification  # type: ignore[import-not-found]

    return MethodSpecification(config=b200_thermal_nerf_config(),
                               description="ThermoNeRF on hand-written sm_100a kernels (libtnf_b200).")
