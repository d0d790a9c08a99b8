"""bench.py's reference arm (the reference's own all-pairs algorithm on the host cores, the one
place besides tests/ and smoke() that may execute oracle/) runs without a GPU: check the JSON line
the driver parses, and that the product arm refuses to run without the CUDA path."""
import json
import os
import subprocess
import sys

from conftest import ROOT, has_gpu

import pytest


def _run(*args):
    env = dict(os.environ, OMP_NUM_THREADS=os.environ.get("OMP_NUM_THREADS", "4"))
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True,
                          text=True, timeout=600, env=env)


def test_reference_arm_prints_the_contract_line():
    r = _run("--impl", "reference", "--steps", "1", "--warmup", "1")
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["n_gpus"] == 1 and line["steps"] == 1
    assert line["metric"].startswith("SPH particle-updates/s") and line["unit"] == "particle-updates/s"
    assert line["higher_is_better"] is True and line["value"] > 0 and line["gpu_launches"] == 0
    assert line["config"]["particles_total"] == 256 ** 3
    cb, e2e = line["cpu_baseline"], line["e2e"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and "all-pairs" in cb["sample"] and cb["value"] == line["value"]
    assert e2e == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


@pytest.mark.skipif(has_gpu(), reason="checks the no-GPU failure mode")
def test_product_arm_fails_loudly_without_a_gpu():
    r = _run("--steps", "1", "--warmup", "1", "--no-cpu-baseline")
    assert r.returncode != 0
    assert "cuda" in (r.stderr + r.stdout).lower()


def test_egl_reference_runner_builds_and_reports_a_missing_gl_stack(tmp_path):
    """tools/egl_shader_runner.c (the reference's unmodified shaders, headless): plain
    C99 without GL headers; on a machine without libEGL it must say so and exit 3 (what bench.py
    records under cpu_baseline.reference_shaders_egl)."""
    exe = str(tmp_path / "egl_shader_runner")
    subprocess.check_call(["gcc", "-O2", "-std=c99", "-Wall", "-Wextra", "-Werror",
                           os.path.join(ROOT, "tools", "egl_shader_runner.c"), "-o", exe, "-ldl"])
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 2 and "usage" in r.stderr
    r = subprocess.run([exe, str(tmp_path), "1"], capture_output=True, text=True)
    if r.returncode == 3:
        assert "EGL unavailable" in r.stderr
    else:                                   # a GL-capable machine: the empty directory has no shaders
        assert r.returncode == 2 and "cannot read" in r.stderr
