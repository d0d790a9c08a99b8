"""make_golden.py -- generate tests/golden/*.npz by EXECUTING the reference's own shader text.

Run once in the build container (needs /root/reference; the GPU box does not have it):

    python tests/golden/make_golden.py [--nan-onset]

The three compute shaders are read, unmodified, from /root/reference/NPR-SPH/ and executed by
the GLSL-subset SIMT interpreter in glsl_simt.py; the initial block, the uniform defaults and the
dispatch size are parsed out of the reference's Main.cpp text (make_grid / ConstantsUniform /
BoundaryUniform / NUM_WORK_GROUPS), and the passes are dispatched in display()'s order
(Main.cpp:295-303).  Nothing from the reference is copied into the repo -- only the numeric
outputs are stored.  The fixtures pin oracle/sph_oracle.c (tests/test_oracle_golden.py).

Fixtures (float32, little endian):
  default_10k.npz   the reference scene exactly (N = 10,000): for steps 0..2 the full input
                    state (pos, vel) and, after each pass, its outputs.
  small_2000.npz    NUM_PARTICLES re-#defined to 2000 (so dt = 1/2000 as integrate_comp.glsl:33
                    says) with non-default uniforms and a tight box, 4 steps, full states:
                    exercises the wall clamps on all six faces.
  nan_onset.npz     (--nan-onset, ~25 min) the default scene run until the first NaN appears.
"""
from __future__ import annotations

import argparse
import os
import re
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from glsl_simt import Shader  # noqa: E402

REF = "/root/reference/NPR-SPH"
PASSES = ("rho_pres_comp.glsl", "force_comp.glsl", "integrate_comp.glsl")   # Main.cpp:59-61


def read(name):
    with open(os.path.join(REF, name)) as f:
        return f.read()


def parse_host_defaults():
    """Defaults the reference's host code uploads (parsed from Main.cpp, not restated)."""
    src = read("Main.cpp")

    def macro(name):
        return re.search(rf"#define\s+{name}\s+([^\s/]+)", src).group(1)

    def floats(text):
        return [float(x.rstrip("f")) for x in re.findall(r"-?\d+\.?\d*(?:[eE][+-]?\d+)?f?", text)]

    cu = re.search(r"struct ConstantsUniform\s*{(.*?)}\s*ConstantsData", src, re.S).group(1)
    uniforms = {m.group(1): float(m.group(2).rstrip("f"))
                for m in re.finditer(r"float\s+(\w+)\s*=\s*([-\d.eE]+f?)\s*;", cu)}
    bu = re.search(r"struct BoundaryUniform\s*{(.*?)}\s*BoundaryData", src, re.S).group(1)
    for m in re.finditer(r"glm::vec4\s+(\w+)\s*=\s*glm::vec4\((.*?)\)\s*;", bu):
        uniforms[m.group(1)] = tuple(floats(m.group(2)))
    grid = re.search(r"make_grid\(\)\s*{(.*?)return positions", src, re.S).group(1)
    dims = [int(x) for x in re.findall(r"<\s*(\d+)\s*;", grid)]
    return {"uniforms": uniforms, "dims": dims, "radius": np.float32(macro("PARTICLE_RADIUS").rstrip("f")),
            "num_particles": int(macro("NUM_PARTICLES")), "num_groups": int(macro("NUM_WORK_GROUPS"))}


def initial_block(dims, spacing):
    """make_grid() + init_particles(), Main.cpp:488-521: (float)i * PARTICLE_RADIUS, i outermost."""
    nx, ny, nz = dims
    i, j, k = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
    pos = np.zeros((nx * ny * nz, 4), np.float32)
    pos[:, 0] = i.ravel().astype(np.float32) * np.float32(spacing)
    pos[:, 1] = j.ravel().astype(np.float32) * np.float32(spacing)
    pos[:, 2] = k.ravel().astype(np.float32) * np.float32(spacing)
    pos[:, 3] = 1.0
    z = lambda: np.zeros_like(pos)
    return {"pos": pos, "vel": z(), "force": z(), "extras": z()}


def run_case(shaders, groups, state, uniforms, steps, record):
    buffers = {"particles": state}
    for s in range(steps):
        record(s, "in", state)
        for name, sh in zip(("rho", "force", "integrate"), shaders):
            t = time.time()
            sh.dispatch(groups, buffers, uniforms)
            print(f"  step {s} {name:9s} {time.time() - t:6.1f} s", flush=True)
            record(s, name, state)


def gen_default(host):
    shaders = [Shader(read(f)) for f in PASSES]
    state = initial_block(host["dims"], host["radius"])
    out = {}
    sub = np.arange(0, host["num_particles"], 5)
    out["subset"] = sub.astype(np.int32)

    def record(s, what, st):
        if what == "in":
            if s > 0:
                out[f"s{s}_in_pos"] = st["pos"][:, :3].copy()
                out[f"s{s}_in_vel"] = st["vel"][:, :3].copy()
        elif what == "rho":
            out[f"s{s}_rho_p"] = st["extras"][sub, :2].copy()
        elif what == "force":
            out[f"s{s}_force"] = st["force"][sub, :3].copy()
        else:
            out[f"s{s}_out_pos"] = st["pos"][sub, :3].copy()
            out[f"s{s}_out_vel"] = st["vel"][sub, :3].copy()

    run_case(shaders, host["num_groups"], state, host["uniforms"], 3, record)
    out["w_lanes"] = np.stack([state["pos"][:, 3], state["vel"][:, 3], state["force"][:, 3],
                               state["extras"][:, 2], state["extras"][:, 3]])
    np.savez_compressed(os.path.join(HERE, "default_10k.npz"), **out)


def gen_small(host):
    n = 2000
    shaders = [Shader(read(f), {"NUM_PARTICLES": n}) for f in PASSES]
    state = initial_block((10, 20, 10), host["radius"])
    rng = np.random.default_rng(20261017)
    state["pos"][:, :3] += rng.uniform(-1e-3, 1e-3, (n, 3)).astype(np.float32)
    state["vel"][:, :3] = rng.normal(0.0, 2.0, (n, 3)).astype(np.float32)
    uniforms = {"mass": 1.3e-4, "smoothing_coeff": 2.5, "visc": 40.0, "resting_rho": 1000.0,
                "upper": (0.048, 0.0945, 0.047, 1.0), "lower": (-0.002, -0.004, -0.003, 1.0)}
    out = {"uniforms": np.array([uniforms[k] for k in ("mass", "smoothing_coeff", "visc", "resting_rho")], np.float32),
           "upper": np.array(uniforms["upper"], np.float32), "lower": np.array(uniforms["lower"], np.float32),
           "num_particles": np.int32(n)}

    def record(s, what, st):
        if what == "in":
            out[f"s{s}_in_pos"] = st["pos"][:, :3].copy()
            out[f"s{s}_in_vel"] = st["vel"][:, :3].copy()
        elif what == "rho":
            out[f"s{s}_rho_p"] = st["extras"][:, :2].copy()
        elif what == "force":
            out[f"s{s}_force"] = st["force"][:, :3].copy()
        else:
            out[f"s{s}_out_pos"] = st["pos"][:, :3].copy()
            out[f"s{s}_out_vel"] = st["vel"][:, :3].copy()

    run_case(shaders, 2, state, uniforms, 4, record)
    np.savez_compressed(os.path.join(HERE, "small_2000.npz"), **out)


def gen_nan_onset(host, max_steps=70):
    shaders = [Shader(read(f)) for f in PASSES]
    state = initial_block(host["dims"], host["radius"])
    buffers = {"particles": state}
    first, ids = -1, np.zeros(0, np.int32)
    for s in range(max_steps):
        t = time.time()
        for sh in shaders:
            sh.dispatch(host["num_groups"], buffers, host["uniforms"])
        bad = np.isnan(state["pos"][:, :3]).any(axis=1)
        print(f"  step {s}: {bad.sum()} NaN particles ({time.time() - t:.0f} s)", flush=True)
        if bad.any():
            first, ids = s, np.nonzero(bad)[0].astype(np.int32)
            break
    np.savez_compressed(os.path.join(HERE, "nan_onset.npz"), first_nan_step=np.int32(first), nan_ids=ids)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--nan-onset", action="store_true")
    ap.add_argument("--only", default="")
    a = ap.parse_args()
    host = parse_host_defaults()
    print("parsed from Main.cpp:", host)
    if a.nan_onset:
        gen_nan_onset(host)
    else:
        if a.only in ("", "small"):
            gen_small(host)
        if a.only in ("", "default"):
            gen_default(host)
