"""The C ABI driven by a compiled C program (tests/consumer/consumer.c), not by ctypes: create ->
'p' -> step -> download, checksum compared with the same run through the harness."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT

SRC = os.path.join(ROOT, "tests", "consumer", "consumer.c")
LIBDIR = os.path.join(ROOT, "npr-sph_b200", "lib")


def build_consumer(tmp_path):
    exe = str(tmp_path / "consumer")
    subprocess.check_call(["gcc", "-std=c99", "-O1", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
                           SRC, "-o", exe, "-L", LIBDIR, "-lnprsph", f"-Wl,-rpath,{LIBDIR}"])
    return exe


def test_c_consumer_compiles_and_links_against_the_header(tmp_path):
    """CPU side: the header is valid C99 and every symbol the program uses resolves in libnprsph.so."""
    assert os.path.exists(build_consumer(tmp_path))


def fnv1a(b: bytes) -> int:
    h = 1469598103934665603
    for chunk in np.frombuffer(b, np.uint8).tolist():
        h = ((h ^ chunk) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return h


@pytest.mark.gpu
def test_c_consumer_matches_the_ctypes_run(tmp_path, sph):
    exe = build_consumer(tmp_path)
    out = subprocess.run([exe, "5"], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr + out.stdout
    fields = out.stdout.split()
    got = dict(zip(fields[0::2], fields[1::2]))
    assert got["particles"] == "10000" and got["steps"] == "5" and got["positions_match"] == "1"
    sim = sph.Simulation()
    sim.toggle_pause()
    sim.step(5)
    rec = sim.download()
    assert int(got["cell_subdiv"]) == sim.stats().cell_subdiv == 4      # automatic: h = 4 lattice spacings
    assert np.array_equal(sim.download_positions(), rec[:, 0:4])
    assert int(got["checksum"], 16) == fnv1a(rec.tobytes()), "the C program and the harness must see the same bytes"
