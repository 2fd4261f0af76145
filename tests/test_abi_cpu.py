"""The C-ABI library loads without a GPU, exports every symbol include/tnf_b200.h declares,
its ctypes mirror has the same struct layout as the C header, and argument validation
works before any CUDA call.  No compute calls here."""

import ctypes as C
import re
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
HEADER = ROOT / "include" / "tnf_b200.h"


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g

    g.build()
    from thermo_nerf_b200 import _lib

    return _lib.load()


def declared_symbols():
    text = re.sub(r"/\*.*?\*/", "", HEADER.read_text(), flags=re.S)
    return sorted(set(re.findall(r"\b(tnf_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported(lib):
    from thermo_nerf_b200 import _lib

    syms = declared_symbols()
    assert syms, "no declarations found in the header"
    assert sorted(_lib.EXPORTED_SYMBOLS) == syms
    for s in syms:
        assert hasattr(lib, s), s


def test_abi_version_and_error_string(lib):
    from thermo_nerf_b200 import _lib

    assert lib.tnf_version() == _lib.TNF_ABI_VERSION == 9
    assert isinstance(lib.tnf_last_error(), bytes)


def test_ctypes_layout_matches_c_header(tmp_path):
    from thermo_nerf_b200 import _lib

    names = ["TnfHashGrid", "TnfLinear", "TnfDensityNet", "TnfField", "TnfModel", "TnfCamera", "TnfRays", "TnfOutputs",
             "TnfLinearGrad", "TnfDensityNetGrad", "TnfFieldGrad", "TnfModelGrad", "TnfSaved", "TnfOutputGrads",
             "TnfLossArgs", "TnfAdamTensor", "TnfPeerArena", "TnfAdamSegment", "TnfDataset"]
    # every field of every struct, taken from the ctypes mirror (a field the header lacks fails the gcc compile)
    probes = {n: [f[0] for f in getattr(_lib, n)._fields_] for n in names}
    src = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{HEADER}"', "int main(void){"]
    for n in names:
        src.append(f'printf("{n} %zu\\n", sizeof({n}));')
        for f in probes.get(n, []):
            src.append(f'printf("{n}.{f} %zu\\n", offsetof({n}, {f}));')
    src.append("return 0;}")
    c = tmp_path / "layout.c"
    c.write_text("\n".join(src))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-std=c99", "-o", str(exe), str(c)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout
    for line in out.strip().splitlines():
        key, val = line.split()
        if "." in key:
            s, f = key.split(".")
            assert getattr(getattr(_lib, s), f).offset == int(val), key
        else:
            assert C.sizeof(getattr(_lib, key)) == int(val), key


def test_integration_doc_names_every_entry_point():
    """INTEGRATION.md's table maps each exported symbol to the reference interface it replaces."""
    from thermo_nerf_b200 import _lib

    doc = (ROOT / "INTEGRATION.md").read_text()
    doc = doc.replace("tnf_peer_alloc/free/open_handle/close_handle/enable_access",
                      "tnf_peer_alloc tnf_peer_free tnf_peer_open_handle tnf_peer_close_handle tnf_peer_enable_access")
    doc = doc.replace("tnf_*_workspace_bytes", "tnf_forward_workspace_bytes tnf_backward_workspace_bytes")
    missing = [n for n in _lib.EXPORTED_SYMBOLS if n not in doc]
    assert not missing, missing


def test_header_cites_the_reference_for_each_entry_point():
    """Each declaration's preceding comment block cites reference file:line (or says what plumbing it is)."""
    text = HEADER.read_text()
    assert len(re.findall(r"[a-z_]+\.py:\d+", text)) >= 12
    for anchor in ("thermal_nerf_model.py:210", "thermal_nerf_model.py:277", "config_thermal_nerf.py:32",
                   "renderer.py:183", "renderer.py:189", "pipeline_tracking.py"):
        assert anchor in text, anchor


def test_workspace_size(lib):
    assert lib.tnf_forward_workspace_bytes(0, 0) >= 8
    assert lib.tnf_forward_workspace_bytes(640000, 65536) >= 2 * 10 * 4
    assert lib.tnf_forward_workspace_bytes(100, 0) >= 8


def test_argument_validation_happens_before_any_cuda_call(lib):
    from thermo_nerf_b200 import _lib

    m, r, o = _lib.TnfModel(), _lib.TnfRays(), _lib.TnfOutputs()
    rc = lib.tnf_render_forward(C.byref(m), C.byref(r), C.byref(o), 0, None, 0, None)
    assert rc == _lib.TNF_ERR_INVALID_ARGUMENT
    assert b"null" in lib.tnf_last_error()
    rc = lib.tnf_render_forward(None, C.byref(r), C.byref(o), 0, None, 0, None)
    assert rc == _lib.TNF_ERR_INVALID_ARGUMENT
    # unsupported architecture is reported as such
    for k in range(2):
        m.prop[k].grid.table = 16
        m.prop[k].grid.num_levels = 99
    rc = lib.tnf_render_forward(C.byref(m), C.byref(r), C.byref(o), 0, None, 0, None)
    assert rc == _lib.TNF_ERR_UNSUPPORTED_CONFIG
    with pytest.raises(_lib.TnfError):
        _lib.check(rc)


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from thermo_nerf_b200 import _lib

    monkeypatch.setenv("TNF_B200_LIB", str(tmp_path / "nope.so"))
    monkeypatch.setattr(_lib, "_LIB", None)
    with pytest.raises(ImportError):
        _lib.load()


def test_training_entry_points_validate_arguments(lib):
    from thermo_nerf_b200 import _lib

    rc = lib.tnf_losses(None, None)
    assert rc == _lib.TNF_ERR_INVALID_ARGUMENT
    a = _lib.TnfLossArgs()
    rc = lib.tnf_losses(C.byref(a), None)
    assert rc == _lib.TNF_ERR_INVALID_ARGUMENT and b"losses" in lib.tnf_last_error()
    rc = lib.tnf_adam_step(None, 3, 0.9, 0.999, 1e-15, 1, 1.0, None, None, 0, None)
    assert rc == _lib.TNF_ERR_INVALID_ARGUMENT
    rc = lib.tnf_adam_step(None, 0, 0.9, 0.999, 1e-15, 1, 1.0, None, None, 0, None)
    assert rc == _lib.TNF_OK  # nothing to do
    t = (_lib.TnfAdamTensor * 1)()
    t[0].param = t[0].grad = t[0].exp_avg = t[0].exp_avg_sq = 256
    t[0].numel = 16
    rc = lib.tnf_adam_step(t, 1, 0.9, 0.999, 1e-15, 0, 1.0, None, None, 0, None)
    assert rc == _lib.TNF_ERR_INVALID_ARGUMENT and b"step" in lib.tnf_last_error()
    m, r = _lib.TnfModel(), _lib.TnfRays()
    rc = lib.tnf_render_backward(C.byref(m), C.byref(r), None, None, None, None, 0, None)
    assert rc == _lib.TNF_ERR_INVALID_ARGUMENT
    assert lib.tnf_backward_workspace_bytes(None, 10) >= 16


def test_camera_and_postprocess_entry_points_validate_arguments(lib):
    from thermo_nerf_b200 import _lib

    rc = lib.tnf_generate_rays(None, 0, 4, None, None, None, None)
    assert rc == _lib.TNF_ERR_INVALID_ARGUMENT and b"camera" in lib.tnf_last_error()
    cam = _lib.TnfCamera()
    cam.width, cam.height, cam.fx, cam.fy = 4, 3, 10.0, 10.0
    rc = lib.tnf_generate_rays(C.byref(cam), 10, 4, None, None, None, None)  # pixels 10..13 of a 12-pixel image
    assert rc == _lib.TNF_ERR_INVALID_ARGUMENT and b"outside" in lib.tnf_last_error()
    assert lib.tnf_generate_rays(C.byref(cam), 0, 0, None, None, None, None) == _lib.TNF_OK  # nothing to do
    rc = lib.tnf_generate_rays(C.byref(cam), 0, 12, None, None, None, None)
    assert rc == _lib.TNF_ERR_INVALID_ARGUMENT and b"null" in lib.tnf_last_error()
    # an image without its output (and vice versa) is rejected; an empty call is a no-op
    assert lib.tnf_postprocess_frame(256, None, 16, None, 0, None, None, None) == _lib.TNF_ERR_INVALID_ARGUMENT
    assert lib.tnf_postprocess_frame(None, None, 16, None, 0, None, None, None) == _lib.TNF_OK
    assert lib.tnf_postprocess_frame(None, 256, 16, 256, 0, None, 256, None) == _lib.TNF_ERR_INVALID_ARGUMENT  # lut_n
    # camera rays are an eval-mode input of the forward and never of the backward
    m, r = _lib.TnfModel(), _lib.TnfRays()
    r.from_camera = 1
    rc = lib.tnf_render_backward(C.byref(m), C.byref(r), None, None, None, None, 0, None)
    assert rc == _lib.TNF_ERR_INVALID_ARGUMENT


def test_peer_entry_points_validate_arguments(lib):
    from thermo_nerf_b200 import _lib

    assert lib.tnf_peer_barrier(None, 0, 1, None) == _lib.TNF_ERR_INVALID_ARGUMENT
    a = _lib.TnfPeerArena()
    a.world_size, a.rank = 2, 5
    assert lib.tnf_peer_barrier(C.byref(a), 0, 1, None) == _lib.TNF_ERR_INVALID_ARGUMENT  # rank outside the world
    a.rank = 1
    assert lib.tnf_peer_barrier(C.byref(a), 0, 1, None) == _lib.TNF_ERR_INVALID_ARGUMENT  # null flag blocks
    for r in range(2):
        a.flags[r] = a.grads[r] = a.params[r] = 256
    assert lib.tnf_peer_barrier(C.byref(a), 7, 1, None) == _lib.TNF_ERR_INVALID_ARGUMENT  # slot
    a.numel = 20  # not a multiple of 4 * world_size
    seg = (_lib.TnfAdamSegment * 1)()
    assert lib.tnf_peer_adam_step(C.byref(a), 256, 256, seg, 1, 0.9, 0.999, 1e-15, None) == _lib.TNF_ERR_INVALID_ARGUMENT
    a.numel = 16
    seg[0].begin, seg[0].end, seg[0].step, seg[0].lr, seg[0].active = 0, 8, 1, 1e-2, 1
    rc = lib.tnf_peer_adam_step(C.byref(a), 256, 256, seg, 1, 0.9, 0.999, 1e-15, None)
    assert rc == _lib.TNF_ERR_INVALID_ARGUMENT and b"segments cover" in lib.tnf_last_error()
    seg[0].end, seg[0].step = 16, 0
    rc = lib.tnf_peer_adam_step(C.byref(a), 256, 256, seg, 1, 0.9, 0.999, 1e-15, None)
    assert rc == _lib.TNF_ERR_INVALID_ARGUMENT and b"step" in lib.tnf_last_error()


def test_sample_batch_validates_arguments(lib):
    from thermo_nerf_b200 import _lib

    assert lib.tnf_sample_batch(None, None, 4, None, None, None, None, None, None, None) == _lib.TNF_ERR_INVALID_ARGUMENT
    ds = _lib.TnfDataset()
    ds.images = ds.camera_to_worlds = ds.intrinsics = 256
    ds.num_images, ds.height, ds.width, ds.channels = 2, 4, 4, 2
    rc = lib.tnf_sample_batch(C.byref(ds), 256, 4, 256, 256, 256, None, None, None, None)
    assert rc == _lib.TNF_ERR_INVALID_ARGUMENT and b"channels" in lib.tnf_last_error()
    ds.channels = 3
    assert lib.tnf_sample_batch(C.byref(ds), None, 0, None, None, None, None, None, None, None) == _lib.TNF_OK
    assert lib.tnf_sample_batch(C.byref(ds), None, 4, None, None, None, None, None, None, None) == _lib.TNF_ERR_INVALID_ARGUMENT
