"""ctypes binding of ``libtnf_b200.so`` (C ABI declared in ``include/tnf_b200.h``).

The structures below mirror the header field for field.  There is no fallback: if the
shared library is missing or does not load, every compute entry point raises.
"""

from __future__ import annotations

import ctypes as C
import os
from pathlib import Path
from typing import Optional

TNF_ABI_VERSION = 9
TNF_MAX_LEVELS = 16
TNF_MAX_PROP_LEVELS = 8
TNF_MAX_SAMPLES = 256
TNF_MAX_FIELD_SAMPLES = 64
TNF_NUM_PROP = 2

TNF_OK = 0
TNF_ERR_INVALID_ARGUMENT = -1
TNF_ERR_UNSUPPORTED_CONFIG = -2
TNF_ERR_WORKSPACE_TOO_SMALL = -3
TNF_ERR_CUDA = -4

APPEARANCE_ZEROS, APPEARANCE_MEAN, APPEARANCE_LOOKUP = 0, 1, 2
PRECISION_FP32, PRECISION_TC_FP16 = 0, 1
HEAD_THERMAL, HEAD_CONCAT = 0, 1

_fp = C.c_void_p  # device pointers cross the ABI as plain addresses


class TnfHashGrid(C.Structure):
    _fields_ = [
        ("table", _fp),
        ("scalings", C.c_float * TNF_MAX_LEVELS),
        ("num_levels", C.c_int32),
        ("log2_size", C.c_int32),
    ]


class TnfLinear(C.Structure):
    _fields_ = [("weight", _fp), ("bias", _fp)]


class TnfDensityNet(C.Structure):
    _fields_ = [("grid", TnfHashGrid), ("l0", TnfLinear), ("l1", TnfLinear)]


class TnfField(C.Structure):
    _fields_ = [
        ("grid", TnfHashGrid),
        ("base0", TnfLinear),
        ("base1", TnfLinear),
        ("rgb0", TnfLinear),
        ("rgb1", TnfLinear),
        ("rgb2", TnfLinear),
        ("th0", TnfLinear),
        ("th1", TnfLinear),
        ("th2", TnfLinear),
        ("appearance", _fp),
        ("num_images", C.c_int32),
        ("_pad", C.c_int32),
    ]


class TnfModel(C.Structure):
    _fields_ = [
        ("prop", TnfDensityNet * TNF_NUM_PROP),
        ("field", TnfField),
        ("num_samples", C.c_int32 * (TNF_NUM_PROP + 1)),
        ("training", C.c_int32),
        ("near_plane", C.c_float),
        ("far_plane", C.c_float),
        ("anneal", C.c_float),
        ("use_contraction", C.c_int32),
        ("aabb", C.c_float * 6),
        ("appearance_mode", C.c_int32),
        ("precision", C.c_int32),
        ("detach_thermal_geo", C.c_int32),
        ("head_mode", C.c_int32),
    ]


class TnfCamera(C.Structure):
    _fields_ = [
        ("c2w", C.c_float * 12),
        ("fx", C.c_float),
        ("fy", C.c_float),
        ("cx", C.c_float),
        ("cy", C.c_float),
        ("width", C.c_int32),
        ("height", C.c_int32),
    ]


class TnfDataset(C.Structure):
    _fields_ = [
        ("images", _fp),
        ("thermal", _fp),
        ("camera_to_worlds", _fp),
        ("intrinsics", _fp),
        ("num_images", C.c_int32),
        ("height", C.c_int32),
        ("width", C.c_int32),
        ("channels", C.c_int32),
        ("images_uint8", C.c_int32),
        ("thermal_uint8", C.c_int32),
    ]


class TnfRays(C.Structure):
    _fields_ = [
        ("origins", _fp),
        ("directions", _fp),
        ("camera_indices", _fp),
        ("nears", _fp),
        ("fars", _fp),
        ("jitter", _fp),
        ("num_rays", C.c_int64),
        ("from_camera", C.c_int32),
        ("_pad", C.c_int32),
        ("first_pixel", C.c_int64),
        ("camera", TnfCamera),
    ]


class TnfOutputs(C.Structure):
    _fields_ = [
        ("rgb", _fp),
        ("thermal", _fp),
        ("depth", _fp),
        ("expected_depth", _fp),
        ("accumulation", _fp),
        ("prop_depth", _fp * TNF_NUM_PROP),
        ("weights", _fp * (TNF_NUM_PROP + 1)),
        ("sdist", _fp * (TNF_NUM_PROP + 1)),
        ("field_features", _fp),
        ("field_samples", _fp),
    ]


class TnfLinearGrad(C.Structure):
    _fields_ = [("weight", _fp), ("bias", _fp)]


class TnfDensityNetGrad(C.Structure):
    _fields_ = [("table", _fp), ("l0", TnfLinearGrad), ("l1", TnfLinearGrad)]


class TnfFieldGrad(C.Structure):
    _fields_ = [
        ("table", _fp),
        ("base0", TnfLinearGrad),
        ("base1", TnfLinearGrad),
        ("rgb0", TnfLinearGrad),
        ("rgb1", TnfLinearGrad),
        ("rgb2", TnfLinearGrad),
        ("th0", TnfLinearGrad),
        ("th1", TnfLinearGrad),
        ("th2", TnfLinearGrad),
        ("appearance", _fp),
    ]


class TnfModelGrad(C.Structure):
    _fields_ = [("prop", TnfDensityNetGrad * TNF_NUM_PROP), ("field", TnfFieldGrad), ("ray_origins", _fp),
                ("ray_directions", _fp)]


class TnfSaved(C.Structure):
    _fields_ = [
        ("sdist", _fp * (TNF_NUM_PROP + 1)),
        ("weights", _fp * (TNF_NUM_PROP + 1)),
        ("field_features", _fp),
        ("field_samples", _fp),
    ]


class TnfOutputGrads(C.Structure):
    _fields_ = [
        ("rgb", _fp),
        ("thermal", _fp),
        ("accumulation", _fp),
        ("weights", _fp * (TNF_NUM_PROP + 1)),
    ]


class TnfLossArgs(C.Structure):
    _fields_ = [
        ("weights", _fp * (TNF_NUM_PROP + 1)),
        ("sdist", _fp * (TNF_NUM_PROP + 1)),
        ("rgb", _fp),
        ("thermal", _fp),
        ("gt_rgb", _fp),
        ("gt_thermal", _fp),
        ("num_rays", C.c_int64),
        ("num_samples", C.c_int32 * (TNF_NUM_PROP + 1)),
        ("interlevel_mult", C.c_float),
        ("distortion_mult", C.c_float),
        ("use_rgb_loss", C.c_int32),
        ("use_thermal_loss", C.c_int32),
        ("grad_scale", C.c_float),
        ("losses", _fp),
        ("g_rgb", _fp),
        ("g_thermal", _fp),
        ("g_weights", _fp * (TNF_NUM_PROP + 1)),
        ("concat", C.c_int32),
        ("_pad", C.c_int32),
        ("accumulation", _fp),
        ("noise", _fp),
        ("g_accumulation", _fp),
    ]


TNF_ADAM_MAX_TENSORS = 48


class TnfAdamTensor(C.Structure):
    _fields_ = [
        ("param", _fp),
        ("grad", _fp),
        ("exp_avg", _fp),
        ("exp_avg_sq", _fp),
        ("numel", C.c_int64),
        ("lr", C.c_float),
        ("_pad", C.c_int32),
    ]


TNF_MAX_PEERS = 16
TNF_PEER_FLAG_SLOTS = 4
TNF_PEER_FLAG_TIMEOUT = 64
TNF_PEER_FLAG_WORDS = 128
TNF_PEER_PUSH = 0
TNF_PEER_MULTIMEM = 3
TNF_MAX_ADAM_SEGMENTS = 4
TNF_IPC_HANDLE_BYTES = 64


class TnfPeerArena(C.Structure):
    _fields_ = [
        ("grads", _fp * TNF_MAX_PEERS),
        ("params", _fp * TNF_MAX_PEERS),
        ("flags", _fp * TNF_MAX_PEERS),
        ("world_size", C.c_int32),
        ("rank", C.c_int32),
        ("numel", C.c_int64),
    ]


class TnfAdamSegment(C.Structure):
    _fields_ = [
        ("begin", C.c_int64),
        ("end", C.c_int64),
        ("step", C.c_int64),
        ("lr", C.c_float),
        ("active", C.c_int32),
    ]


# every symbol include/tnf_b200.h declares (tests check the library exports all of them)
EXPORTED_SYMBOLS = (
    "tnf_version",
    "tnf_last_error",
    "tnf_forward_workspace_bytes",
    "tnf_render_forward",
    "tnf_render_forward_staged",
    "tnf_render_backward_staged",
    "tnf_generate_rays",
    "tnf_postprocess_frame",
    "tnf_sample_batch",
    "tnf_backward_workspace_bytes",
    "tnf_render_backward",
    "tnf_backward_stage_mask",
    "tnf_field_density",
    "tnf_field_heads",
    "tnf_composite",
    "tnf_losses",
    "tnf_adam_step",
    "tnf_peer_enable_access",
    "tnf_peer_alloc",
    "tnf_peer_free",
    "tnf_peer_open_handle",
    "tnf_peer_close_handle",
    "tnf_peer_barrier",
    "tnf_peer_adam_step",
    "tnf_peer_adam_reduce",
    "tnf_peer_adam_multimem",
    "tnf_peer_adam_range",
    "tnf_peer_gather_params",
)


class TnfError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"libtnf_b200 error {code}: {message}")
        self.code = code


def library_path() -> Path:
    env = os.environ.get("TNF_B200_LIB")
    if env:
        return Path(env)
    return Path(__file__).resolve().parent / "lib" / "libtnf_b200.so"


_LIB: Optional[C.CDLL] = None


def load() -> C.CDLL:
    """Load the shared library (once).  Raises if it is absent - there is no CPU path."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = library_path()
    if not path.exists():
        raise ImportError(
            f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  thermo_nerf_b200 has no CPU or PyTorch fallback."
        )
    lib = C.CDLL(str(path))
    lib.tnf_version.restype = C.c_int
    lib.tnf_version.argtypes = []
    lib.tnf_last_error.restype = C.c_char_p
    lib.tnf_last_error.argtypes = []
    lib.tnf_forward_workspace_bytes.restype = C.c_size_t
    lib.tnf_forward_workspace_bytes.argtypes = [C.c_int64, C.c_int64]
    lib.tnf_render_forward.restype = C.c_int
    lib.tnf_render_forward.argtypes = [
        C.POINTER(TnfModel),
        C.POINTER(TnfRays),
        C.POINTER(TnfOutputs),
        C.c_int64,
        C.c_void_p,
        C.c_size_t,
        C.c_void_p,
    ]
    lib.tnf_render_forward_staged.restype = C.c_int
    lib.tnf_render_forward_staged.argtypes = lib.tnf_render_forward.argtypes + [C.c_void_p]
    lib.tnf_generate_rays.restype = C.c_int
    lib.tnf_generate_rays.argtypes = [C.POINTER(TnfCamera), C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_void_p]
    lib.tnf_postprocess_frame.restype = C.c_int
    lib.tnf_postprocess_frame.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int32, C.c_void_p,
                                          C.c_void_p, C.c_void_p]
    lib.tnf_sample_batch.restype = C.c_int
    lib.tnf_sample_batch.argtypes = [C.POINTER(TnfDataset), C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.tnf_backward_workspace_bytes.restype = C.c_size_t
    lib.tnf_backward_workspace_bytes.argtypes = [C.POINTER(TnfModel), C.c_int64]
    lib.tnf_render_backward.restype = C.c_int
    lib.tnf_render_backward.argtypes = [
        C.POINTER(TnfModel),
        C.POINTER(TnfRays),
        C.POINTER(TnfSaved),
        C.POINTER(TnfOutputGrads),
        C.POINTER(TnfModelGrad),
        C.c_void_p,
        C.c_size_t,
        C.c_void_p,
    ]
    lib.tnf_render_backward_staged.restype = C.c_int
    lib.tnf_render_backward_staged.argtypes = lib.tnf_render_backward.argtypes + [C.c_void_p, C.c_int32]
    lib.tnf_backward_stage_mask.restype = C.c_int
    lib.tnf_backward_stage_mask.argtypes = [C.c_int]
    lib.tnf_field_density.restype = C.c_int
    lib.tnf_field_density.argtypes = [C.POINTER(TnfModel), C.c_int32, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p,
                                      C.c_void_p]
    lib.tnf_field_heads.restype = C.c_int
    lib.tnf_field_heads.argtypes = [C.POINTER(TnfModel), C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p,
                                    C.c_void_p, C.c_void_p]
    lib.tnf_composite.restype = C.c_int
    lib.tnf_composite.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                  C.c_void_p, C.c_void_p]
    lib.tnf_losses.restype = C.c_int
    lib.tnf_losses.argtypes = [C.POINTER(TnfLossArgs), C.c_void_p]
    lib.tnf_adam_step.restype = C.c_int
    lib.tnf_adam_step.argtypes = [
        C.POINTER(TnfAdamTensor),
        C.c_int32,
        C.c_double,
        C.c_double,
        C.c_float,
        C.c_int64,
        C.c_float,
        C.c_void_p,
        C.c_void_p,
        C.c_int32,
        C.c_void_p,
    ]
    lib.tnf_peer_enable_access.restype = C.c_int
    lib.tnf_peer_enable_access.argtypes = [C.c_int32]
    lib.tnf_peer_alloc.restype = C.c_int
    lib.tnf_peer_alloc.argtypes = [C.c_size_t, C.POINTER(C.c_void_p), C.c_char_p]
    lib.tnf_peer_free.restype = C.c_int
    lib.tnf_peer_free.argtypes = [C.c_void_p]
    lib.tnf_peer_open_handle.restype = C.c_int
    lib.tnf_peer_open_handle.argtypes = [C.c_char_p, C.POINTER(C.c_void_p)]
    lib.tnf_peer_close_handle.restype = C.c_int
    lib.tnf_peer_close_handle.argtypes = [C.c_void_p]
    lib.tnf_peer_barrier.restype = C.c_int
    lib.tnf_peer_barrier.argtypes = [C.POINTER(TnfPeerArena), C.c_int32, C.c_uint32, C.c_void_p]
    lib.tnf_peer_adam_step.restype = C.c_int
    lib.tnf_peer_adam_step.argtypes = [C.POINTER(TnfPeerArena), C.c_void_p, C.c_void_p, C.POINTER(TnfAdamSegment),
                                       C.c_int32, C.c_double, C.c_double, C.c_float, C.c_void_p]
    lib.tnf_peer_adam_multimem.restype = C.c_int
    lib.tnf_peer_adam_multimem.argtypes = [C.POINTER(TnfPeerArena), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                           C.POINTER(TnfAdamSegment), C.c_int32, C.c_double, C.c_double, C.c_float,
                                           C.c_void_p]
    lib.tnf_peer_adam_range.restype = C.c_int
    lib.tnf_peer_adam_range.argtypes = [C.POINTER(TnfPeerArena), C.c_int32, C.c_int64, C.c_int64, C.c_int32, C.c_void_p,
                                        C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(TnfAdamSegment), C.c_int32,
                                        C.c_double, C.c_double, C.c_float, C.c_void_p]
    lib.tnf_peer_adam_reduce.restype = C.c_int
    lib.tnf_peer_adam_reduce.argtypes = lib.tnf_peer_adam_step.argtypes
    lib.tnf_peer_gather_params.restype = C.c_int
    lib.tnf_peer_gather_params.argtypes = [C.POINTER(TnfPeerArena), C.POINTER(TnfAdamSegment), C.c_int32, C.c_void_p]
    got = lib.tnf_version()
    if got != TNF_ABI_VERSION:
        raise ImportError(f"{path}: ABI version {got}, binding expects {TNF_ABI_VERSION}")
    _LIB = lib
    return lib


def check(code: int) -> None:
    if code != TNF_OK:
        raise TnfError(code, load().tnf_last_error().decode("utf-8", "replace"))
