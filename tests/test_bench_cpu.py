"""bench.py contract checks that need no GPU: the reference arm prints exactly one JSON line with the keys the
driver reads (both modes), and the product arm refuses to run without a CUDA device instead of falling back."""

import json
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent

BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "impl", "cpu_baseline", "e2e"}


def _run(*args, timeout=240):
    return subprocess.run([sys.executable, str(ROOT / "bench.py"), *args], capture_output=True, text=True, timeout=timeout)


@pytest.mark.parametrize("mode,metric,unit", [("train", "train_rays_per_s", "rays/s"), ("render", "render_mpix_per_s", "Mpix/s")])
def test_reference_arm_prints_one_json_line(mode, metric, unit):
    p = _run("--impl", "reference", "--mode", mode, "--steps", "1", "--warmup", "1", "--ref-rays", "64")
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, p.stdout
    d = json.loads(lines[0])
    assert BASE_KEYS <= set(d), sorted(BASE_KEYS - set(d))
    assert d["impl"] == "reference" and d["metric"] == metric and d["unit"] == unit and d["higher_is_better"] is True
    assert d["value"] > 0 and d["steps"] == 1 and d["vs_baseline"] is None and "workload" in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_exit_without_work():
    import os

    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    p = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                        "--warmup", "1", "--ref-rays", "64"], capture_output=True, text=True, timeout=120, env=env)
    assert p.returncode == 0 and p.stdout.strip() == ""


def test_product_arm_needs_cuda():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    p = _run("--steps", "1", "--warmup", "1")
    assert p.returncode != 0 and "no CPU fallback" in (p.stderr + p.stdout)
