"""CPU-side checks of the boundary: the C-ABI library loads, exports every
symbol include/sphb200.h declares, derives the constructor constants like the
reference, generates the reference's scenes, and fails loudly without a GPU."""
import os
import re

import numpy as np
import pytest

import smoothed_particle_hydrodynamics_b200 as S
from smoothed_particle_hydrodynamics_b200 import binding
from oracle import scenes
from oracle.port import OracleSPH

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "sphb200.h")).read()
    declared = set(re.findall(r"\b(sphb200_[a-z0-9_]+)\s*\(", header))
    bound = {name for name, _, _ in binding.API}
    assert declared == bound, declared ^ bound
    L = S.lib()
    for name in declared:
        assert hasattr(L, name), name


def test_binding_argument_counts_match_the_header():
    """ctypes does not check arity: a prototype that drifts from include/sphb200.h would
    silently pass garbage.  Count the parameters of every declaration."""
    header = open(os.path.join(ROOT, "include", "sphb200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    protos = dict(re.findall(r"\b(sphb200_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", header))
    for name, args, _ in binding.API:
        params = protos[name].strip()
        n = 0 if params in ("", "void") else params.count(",") + 1
        assert n == len(args), "%s: header has %d parameters, binding passes %d" % (name, n, len(args))


def test_default_params_are_the_reference_constructor_literals():
    p = S.default_params()
    assert (p.particle_count, p.grid_x, p.grid_y, p.grid_z, p.examine_count) == (32768, 32, 32, 32, 32)
    d = S.derive(p)
    o = OracleSPH(init_scene=False)    # oracle_derive restates sph.cpp:47-95
    for a, b in [(d.h2, o.p.h2), (d.h_times2_inv, o.p.h_times2_inv), (d.kernel1, o.p.kernel1),
                 (d.kernel2, o.p.kernel2), (d.kernel3, o.p.kernel3), (d.max_x, o.p.max_x),
                 (d.softening, o.p.softening), (d.cfl_limit2, o.p.cfl_limit2), (d.h_scaled9, o.p.hs9)]:
        assert np.float32(a) == np.float32(b)
    assert list(d.central_pos) == list(o.p.central_pos)
    assert d.total_steps == 1000 and d.grid_cell_count == 32768


def test_sphere_scene_is_the_reference_constructor_scene(golden_default):
    pos, vel = S.scene_sphere(S.default_params())
    assert np.array_equal(pos, golden_default["pos0"])
    assert np.array_equal(vel, golden_default["vel0"])


def test_lattice_scene_matches_numpy_restatement():
    d = scenes.lattice_spacing(0.1, 40)
    a = S.scene_lattice(32, 16, 32, d, origin=(0.2, 3.2, 0.6))
    b = scenes.lattice_scene(32, 16, 32, d, origin=(0.2, 3.2, 0.6))
    assert np.array_equal(a, b)
    # any id range can be generated independently (slab ranks make their own particles)
    c = S.scene_lattice(32, 16, 32, d, origin=(0.2, 3.2, 0.6), first_id=5000, count=777)
    assert np.array_equal(c, b[5000:5777])
    with pytest.raises(S.SphError):
        S.scene_lattice(4, 4, 4, d, first_id=60, count=10)


def test_named_scenes_match_the_scene_rule_tables():
    """sphb200_scene_config / _generate (dam-break, box-drop; configs 2-5) against the independent
    numpy restatement of the scene rule in oracle/scenes.py: sites, grown box, spacing, rest
    density and every position bit for bit."""
    for name in S.SCENES:
        cfg = scenes.CONFIGS[name]
        for nu in (30.0, 40.0, 60.0, 120.0):
            p, lat = S.scene_config(name, nu)
            d = scenes.lattice_spacing(0.1, nu)
            sp = scenes.scene_params(nu=nu)
            origin = [v * 0.2 for v in cfg["origin_vox"]]
            need = [int(np.ceil(o / 0.2 + s * float(d) / 0.2)) + 2 for o, s in zip(origin, cfg["sites"])]
            assert (lat.nx, lat.ny, lat.nz) == cfg["sites"] and np.float32(lat.spacing) == d
            assert (p.grid_x, p.grid_y, p.grid_z) == tuple(max(g, m) for g, m in zip(cfg["grid"], need))
            assert np.float32(p.rho0) == np.float32(sp["rho0"]) and p.particle_count == np.prod(cfg["sites"])
            assert (p.neighbor_mode, p.use_uniform_gravity, p.use_wall_collision) == (S.FULL, 1, 1)
            assert tuple(p.gravity) == (0.0, np.float32(-9.8), 0.0) and p.central_mass == 0.0
            assert np.array_equal(np.float32(origin), np.float32(list(lat.origin)))
    p, lat = S.scene_config("boxdrop_16m")
    a = S.scene_generate(lat, first_id=123456, count=4096)
    b = scenes.lattice_scene(256, 128, 512, lat.spacing, origin=list(lat.origin), first_id=123456, count=4096)
    assert np.array_equal(a, b)
    with pytest.raises(S.SphError):
        S.scene_config(99)


def test_invalid_params_are_rejected():
    for kw in (dict(examine_count=4), dict(h=0.0), dict(grid=(0, 4, 4)), dict(neighbor_mode=7),
               dict(grid=(2048, 2048, 2048))):
        with pytest.raises(S.SphError):
            S.derive(S.default_params(**kw))


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(S.SphError) as e:
        S.SPH()
    assert "no CUDA device" in str(e.value) and e.value.code == -2


def test_reference_gui_call_sites_compile_against_the_facade():
    """SURVEY 8(f3) without Qt: tests/conformance/reference_callsites.cpp repeats every statement
    with which main.cpp:20-69, visualization.cpp:144-157 / 175-193 / 327-335, sphconfig.cpp:56-95
    and widget.cpp:108-125 touch `class SPH`; it must compile against host/sph.h as is."""
    import subprocess
    host = os.path.join(ROOT, "smoothed_particle_hydrodynamics_b200", "host")
    r = subprocess.run(["g++", "-std=c++11", "-fsyntax-only", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                        "-I", host, os.path.join(ROOT, "tests", "conformance", "reference_callsites.cpp")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    # and the facade declares every public name of the reference's class (src/sph.h:20-84)
    decl = open(os.path.join(host, "sph.h")).read()
    for name in ("isStopped", "isPaused", "getParticles", "getParticleCount", "getGridCellCounts", "getParticleBounds",
                 "getInteractionRadius2", "getGrid", "getCellSize", "getGravity", "setGravity", "getStiffness",
                 "setStiffness", "getViscosityScalar", "setViscosityScalar", "getTimeStep", "setTimeStep", "getDamping",
                 "setDamping", "getCflLimit", "setCflLimit", "run", "step", "pauseResume", "stopSimulation",
                 "updateElapsed", "stepFinished"):
        assert re.search(r"\b%s\b" % name, decl), name
