"""CPU oracle for the ThermoNeRF volumetric-render hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``thermo_nerf_b200/`` may import this
package: only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs use it, and only as the checker or
as the timed CPU baseline - never as the product path.

PARITY UNPINNED.  The arithmetic of the path lives in the third-party package
``nerfstudio==1.1.5`` (reference ``pyproject.toml:11``, ``uv.lock:2743-2744``),
which is neither vendored under /root/reference nor installable here, and the
reference's own tests never run a forward pass of this path
(``tests/test_renderer.py:31-69``).  This package therefore *restates* the
published nerfstudio-1.1.5 ``implementation="torch"`` algorithm (SURVEY.md
Appendix A) and anchors on the reference's call sites:

* ``thermo_nerf/thermal_nerf/thermal_nerf_model.py:86-275``  (wiring, get_outputs)
* ``thermo_nerf/thermal_nerf/thermal_field.py:33-201``       (field + thermal head)
* ``thermo_nerf/thermal_nerf/thermal_renderer.py:26-149``    (thermal compositing)
* ``thermo_nerf/thermal_nerf/thermal_metrics.py:5-34``       (MAE de-normalisation)

The only arithmetic of the path's neighbourhood that the reference itself can run here,
``thermal_metrics.mae_thermal``, IS pinned: tests/golden/reference_thermal_metrics.pt holds its outputs
(tests/golden/make_reference_golden.py) and ``nerfstudio_math.mae_thermal`` reproduces them bit for bit.

The 8-bit thermal ground-truth convention (x / 255, then (max - min) x + min) is pinned against the reference's
own fixture tests/data/thermal/* (tests/golden/reference_thermal_image_kat.pt, tests/test_camera_post_cpu.py).

Also pinned by executing the reference's own code in the build container (generators under tests/golden/):
ThermalRenderer / RGBTRenderer compositing bit for bit (reference_renderers.pt), and - over stand-ins that carry
nerfstudio's interfaces with THIS package's arithmetic - the wiring of thermal_nerf_model.py, thermal_field.py,
thermal_field_head.py, thermal_nerfacto.py, evaluator.py and renderer.py's render loop (reference_model_wiring.pt,
reference_evaluator.pt, reference_render_frames.pt).  nerfstudio's arithmetic itself stays a restatement.

Every detail flagged "recalled" in SURVEY.md Appendix A is a named switch in
``OracleConfig`` so it can be flipped if real nerfstudio source ever becomes
available.
"""

from .nerfstudio_math import *  # noqa: F401,F403
from .thermo_model import *  # noqa: F401,F403
from .camera_post import *  # noqa: F401,F403
