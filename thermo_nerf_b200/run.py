"""Run one of the reference's scripts, unchanged, on the B200 model:

    python -m thermo_nerf_b200.run thermo_nerf.scripts.train_eval_script --data ... --model_type thermal-nerf
    python -m thermo_nerf_b200.run thermo_nerf.scripts.render_video_script --model_uri ... --camera_path_filename ...

``nerfstudio_plugin.install()`` swaps the model config inside the reference's module-level method configs
(``thermal_nerf_config`` read at train_eval_script.py:59, ``thermalnerfacto_config`` at :66-73) and makes
``ThermalNerfModelConfig.setup`` build ``B200ThermalNerfModel`` (so ``Renderer.from_pipeline_path``,
render_video_script.py:69, loads a stock run's ``config.yml`` + checkpoint into it); the script module then runs
as ``__main__`` with the remaining arguments.
"""

from __future__ import annotations

import runpy
import sys


def main(argv=None) -> None:
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv or argv[0] in ("-h", "--help"):
        print(__doc__)
        raise SystemExit(0 if argv else 2)
    from . import nerfstudio_plugin

    nerfstudio_plugin.install()
    module, rest = argv[0], argv[1:]
    sys.argv = [module] + rest
    runpy.run_module(module, run_name="__main__", alter_sys=True)


if __name__ == "__main__":
    main()
