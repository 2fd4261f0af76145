"""thermo_nerf_b200: the ThermoNeRF volumetric-render hot path as hand-written sm_100a CUDA.

Layout: ``csrc/`` (kernels + C ABI, built into ``lib/libtnf_b200.so``), ``_lib`` (ctypes
binding of ``include/tnf_b200.h``), ``functional`` (tensor-level entry points),
``modules`` / ``model`` (host-side mirror of the reference's nerfstudio plugin surface).
"""

from . import _lib
from .functional import ModelTensors, adam_step, losses, render, render_forward
from .model import (ConcatNerfModel, ConcatNerfModelConfig, ThermalNerfactoModel, ThermalNerfactoModelConfig,
                    ThermalNerfModel, ThermalNerfModelConfig)
from .optim import FusedAdam
from .modules import (CameraOptimizer, FieldHeadNames, FieldHeadNamesT, HashMLPDensityField, ThermalFieldHead,
                      ThermalNerfactoTField)
from .rays import PinholeCameras, RayBundle, orbit_cameras, sphere_cameras
from .render import RenderedImageModality, Renderer
from .data import DevicePixelSampler
from .evaluator import Evaluator

__all__ = [
    "ModelTensors", "render_forward", "render", "losses", "adam_step", "FusedAdam", "ThermalNerfModel", "ThermalNerfModelConfig", "ThermalNerfactoModel", "ThermalNerfactoModelConfig", "ConcatNerfModel", "ConcatNerfModelConfig", "CameraOptimizer",
    "FieldHeadNames", "FieldHeadNamesT", "HashMLPDensityField", "ThermalFieldHead", "ThermalNerfactoTField",
    "PinholeCameras", "RayBundle", "orbit_cameras", "sphere_cameras", "Renderer", "RenderedImageModality", "DevicePixelSampler", "Evaluator",
]
