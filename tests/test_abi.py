"""The C-ABI library loads and exports every symbol include/nprsph.h declares (no compute calls
without a GPU), and the POD layouts match the reference's GLSL/C++ structs."""
import ctypes as C
import os
import re

import pytest

from conftest import ROOT, has_gpu


def _declared_symbols():
    with open(os.path.join(ROOT, "include", "nprsph.h")) as f:
        text = re.sub(r"/\*.*?\*/", " ", f.read(), flags=re.S)
    return sorted(set(re.findall(r"\b(nprsph_[a-z_0-9]+)\s*\(", text)))


def test_every_declared_symbol_is_exported_and_bound(sph):
    lib = sph.load()
    names = _declared_symbols()
    assert len(names) >= 30
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/nprsph.h but not exported"
    assert sorted(sph.binding.SYMBOLS) == names, "binding.py must mirror the header one to one"
    assert lib.nprsph_abi_version() == 2


def test_pod_layouts_match_the_reference_structs(sph):
    # struct Particle (Main.cpp:93-99): 4 x vec4, stride 64, offsets 0/16/32/48
    assert sph.PARTICLE_DTYPE.itemsize == 64
    assert [sph.PARTICLE_DTYPE.fields[k][1] for k in ("pos", "vel", "force", "extras")] == [0, 16, 32, 48]
    # ConstantsUniform (Main.cpp:110-116): 4 floats = 16 B; BoundaryUniform (:118-122): upper, lower = 32 B
    assert C.sizeof(sph.Constants) == 16 and C.sizeof(sph.Boundary) == 32
    assert sph.Boundary.upper.offset == 0 and sph.Boundary.lower.offset == 16
    cfg = sph.default_config()
    assert cfg.struct_size == C.sizeof(sph.Config)
    # shader constants made run-time keep the reference's values
    assert cfg.particle_radius == pytest.approx(0.005) and cfg.gas_const == 2000.0
    assert list(cfg.gravity) == [0.0, pytest.approx(-9806.65), 0.0]
    assert cfg.damping == pytest.approx(0.3) and cfg.dt == pytest.approx(1e-4)
    assert cfg.pi == pytest.approx(3.141592741)


def test_ctypes_mirrors_have_the_c_compilers_layout(sph, tmp_path):
    """Every struct that crosses the C ABI: size and field offsets as gcc sees include/nprsph.h
    against the ctypes mirrors the tests and bench.py bind through."""
    import subprocess
    from nprsph_b200 import dist as D
    pairs = {"nprsph_constants": sph.Constants, "nprsph_boundary": sph.Boundary, "nprsph_config": sph.Config,
             "nprsph_stats": sph.Stats, "nprsph_slider": sph.Slider, "nprsph_collider": sph.Collider,
             "nprsph_dist_config": D.DistConfig, "nprsph_dist_info": D.DistInfo}
    rename = {("nprsph_slider", "default"): "def"}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "nprsph.h"', 'int main(void) {']
    for cname, ct in pairs.items():
        lines.append(f'  printf("{cname} size %zu\\n", sizeof({cname}));')
        for fname, *_ in ct._fields_:
            cf = rename.get((cname, fname), fname)
            lines.append(f'  printf("{cname} {fname} %zu\\n", offsetof({cname}, {cf}));')
    lines += ['  printf("nprsph_particle size %zu\\n", sizeof(nprsph_particle));', '  return 0;', '}']
    src = tmp_path / "probe.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "probe"
    subprocess.check_call(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    seen = {}
    for line in subprocess.check_output([str(exe)], text=True).splitlines():
        cname, field, value = line.split()
        seen[(cname, field)] = int(value)
    assert seen[("nprsph_particle", "size")] == 64
    for cname, ct in pairs.items():
        assert seen[(cname, "size")] == C.sizeof(ct), cname
        for fname, *_ in ct._fields_:
            assert seen[(cname, fname)] == getattr(ct, fname).offset, (cname, fname)


def test_constants_window_sliders_are_the_references_widgets(sph):
    """nprsph_slider_info is pure host code: label, range and default of the four ImGui sliders
    (Main.cpp:242-245; defaults Main.cpp:110-116, the smoothing default 4 lies outside its 7..10)."""
    lib = sph.load()
    want = [(b"Mass", 0.01, 0.1, 0.02), (b"Smoothing", 7.0, 10.0, 4.0),
            (b"Viscosity", 1000.0, 5000.0, 3000.0), (b"Resting Density", 1000.0, 5000.0, 1000.0)]
    for sid, (label, lo, hi, default) in enumerate(want):
        s = sph.Slider()
        assert lib.nprsph_slider_info(sid, C.byref(s)) == sph.OK
        assert s.label == label
        assert (s.min, s.max, s.default) == pytest.approx((lo, hi, default))
    assert lib.nprsph_slider_info(len(want), C.byref(sph.Slider())) != sph.OK


@pytest.mark.skipif(has_gpu(), reason="checks the no-GPU failure mode")
def test_create_fails_loudly_without_a_gpu(sph):
    with pytest.raises(sph.NprSphError) as ei:
        sph.Simulation()
    assert ei.value.code == sph.ERR_CUDA and "no CPU path" in str(ei.value)


def test_product_never_touches_the_oracle():
    """libnprsph.so and its harness must not link, load or import anything under oracle/."""
    pkg = os.path.join(ROOT, "npr-sph_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".cu", ".cuh", ".cpp", ".h", ".py")) or fn == "Makefile":
                with open(os.path.join(dirpath, fn)) as f:
                    src = f.read()
                for pat in (r'#\s*include\s*[<"][^>"]*oracle', r"\bimport\s+oracle", r"\bfrom\s+oracle",
                            r"libsphoracle", r"liballpairs", r"oracle/_build"):
                    assert not re.search(pat, src), f"{fn} links/loads the oracle ({pat})"
