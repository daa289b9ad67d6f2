"""Pins the CPU oracle (oracle/sph_oracle.c): (1) against the committed golden
fixtures, which are outputs of the unmodified reference (tests/golden/
make_golden.py); (2) against the compiled reference itself (oracle/_ref) when
it is present.  CPU only."""
import numpy as np
import pytest

from conftest import sparse_lists
from oracle import scenes
from oracle import refharness
from oracle.port import FULL, SAMPLED, OracleSPH

eq = lambda a, b: np.array_equal(a, b, equal_nan=True)  # noqa: E731


def test_default_scene_init_matches_reference_ctor(golden_default):
    o = OracleSPH()   # oracle_init_sphere: srand(42) rotating sphere (sph.cpp:361-425)
    assert eq(o.pos, golden_default["pos0"])
    assert eq(o.vel, golden_default["vel0"])


def test_default_scene_sampled_steps_bit_exact(golden_default):
    g = golden_default
    o = OracleSPH()
    for s in (1, 2):
        o.step(SAMPLED)
        assert eq(o.voxel_ids, g["voxel_ids_%d" % s])
        assert eq(o.members, g["grid_members_%d" % s])
        assert eq(o.count, g["nbr_count_%d" % s])
        nb, nd = sparse_lists(o.nbr, o.dist, o.count)
        assert eq(nb, g["nbr_idx_%d" % s])          # ordered lists, bit exact
        assert eq(nd, g["nbr_dist_%d" % s])
        assert eq(np.array([o.ekin, o.epot], np.float32), g["energy_%d" % s])
        if s == 1:
            assert eq(o.start, g["grid_start_1"])
            assert eq(o.rho, g["density_1"])
            assert eq(o.acc, g["acc_1"])
            assert eq(o.pos, g["pos_1"])
            assert eq(o.vel, g["vel_1"])
    # sanity numbers the survey probed: 6121 neighbours at step 1
    assert int(g["nbr_count_1"].astype(np.int64).sum()) == 6121


def test_default_scene_energy_log_20_steps(golden_default):
    o = OracleSPH()
    en = []
    for _ in range(20):
        o.step(SAMPLED)
        en.append((o.ekin, o.epot))
    assert eq(np.array(en, np.float32), golden_default["energy_20"])
    # step-0 energies the reference logs: 4.69595e+06, -8.37892e+06
    assert "%.5e" % en[0][0] == "4.69595e+06" and "%.5e" % en[0][1] == "-8.37892e+06"


def _full_oracle(g):
    cfg = scenes.CONFIGS["dambreak_16k"]
    nx, ny, nz = cfg["sites"]
    sp = scenes.scene_params()
    o = OracleSPH(n=nx * ny * nz, grid=cfg["grid"], examine=int(g["examine"]), init_scene=False,
                  rho0=sp["rho0"], stiffness=sp["stiffness"], viscosity=sp["viscosity"], central_mass=0.0,
                  gravity=sp["gravity"], time_step=sp["time_step"], central_pos=g["central_pos"])
    o.set_state(g["pos0"], g["vel0"])
    return o


def test_scene_generator_reproduces_fixture_input(golden_full):
    cfg = scenes.CONFIGS["dambreak_16k"]
    pos = scenes.lattice_scene(*cfg["sites"], scenes.lattice_spacing(0.1, 40))
    assert eq(pos, golden_full["pos0"])


def test_full_mode_matches_reference_physics(golden_full):
    g = golden_full
    o = _full_oracle(g)
    for s in (1, 2):
        o.step(FULL, use_gravity=True, use_walls=True)
        assert eq(o.voxel_ids, g["voxel_ids_%d" % s])
        assert eq(o.fine_keys, g["fine_keys_%d" % s])
        assert eq(o.count, g["nbr_count_%d" % s])
        nb, _ = sparse_lists(o.nbr, None, o.count)
        assert eq(nb, g["nbr_idx_%d" % s])
        assert eq(o.rho, g["density_%d" % s])
        assert eq(o.pos, g["pos_%d" % s])
        assert eq(o.vel, g["vel_%d" % s])
        if s == 1:
            assert eq(o.acc, g["acc_1"])


def test_full_mode_is_all_within_h_brute_force(golden_full):
    """FULL mode == every j != i with d2 < h2 (the 27 fine cells lose nothing)."""
    g = golden_full
    o = _full_oracle(g)
    o.voxelize()
    o.find(FULL)
    rng = np.random.default_rng(0)
    pick = rng.choice(o.n, 300, replace=False)
    p = o.pos
    for i in pick:
        d = p[i] - p
        d2 = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]
        want = np.flatnonzero(d2 < o.p.h2)
        want = want[want != i]
        got = np.sort(o.nbr[i, :o.count[i]])
        assert eq(got, want.astype(np.uint32))


@pytest.mark.skipif(not refharness.available("golden"), reason="oracle/_ref not built (no /root/reference here)")
class TestAgainstCompiledReference:
    def test_sampled_random_state_three_steps(self):
        r = refharness.RefSPH("golden")
        rng = np.random.default_rng(7)
        n = r.n
        # denser than the default scene so the sampler's windows actually fire
        pos = (rng.random((n, 3)) * 1.6 + 2.4).astype(np.float32)
        vel = rng.normal(0, 5, (n, 3)).astype(np.float32)
        mass = (rng.random(n) + 0.5).astype(np.float32)
        r.set_state(pos, vel, mass)
        o = OracleSPH(init_scene=False)
        o.set_state(pos, vel, mass)
        for _ in range(3):
            r.step()
            o.step(SAMPLED)
            ids, coords = r.voxels()
            start, members = r.grid()
            idx, dist = r.neighbors()
            cnt = r.neighbor_counts()
            assert eq(ids, o.voxel_ids) and eq(coords, o.voxel_xyz)
            assert eq(start, o.start) and eq(members, o.members)
            assert eq(cnt, o.count) and cnt.max() > 8
            assert eq(sparse_lists(idx, dist, cnt)[0], sparse_lists(o.nbr, o.dist, o.count)[0])
            assert eq(sparse_lists(idx, dist, cnt)[1], sparse_lists(o.nbr, o.dist, o.count)[1])
            assert eq(r.density(), o.rho)
            assert eq(r.acceleration(), o.acc)
            p, v, _ = r.state()
            assert eq(p, o.pos) and eq(v, o.vel)
            assert eq(np.float32(r.energies()), np.float32((o.ekin, o.epot)))

    def test_out_of_box_particles_are_clamped_like_the_reference(self):
        r = refharness.RefSPH("golden")
        pos, vel, mass = r.state()
        pos = pos.copy()
        pos[:100] = -0.3
        pos[100:200] = 7.0
        pos[200:210, 1] = np.nan
        r.set_state(pos, vel, mass)
        o = OracleSPH(init_scene=False)
        o.set_state(pos, vel, mass)
        r.step()
        o.step(SAMPLED)
        assert eq(r.voxels()[0], o.voxel_ids)
        assert eq(r.neighbor_counts(), o.count)
        assert eq(r.state()[0], o.pos)

    def test_walls_and_gravity_switches(self):
        cfg = scenes.CONFIGS["dambreak_16k"]
        nx, ny, nz = cfg["sites"]
        n = nx * ny * nz
        pos = scenes.lattice_scene(nx, ny, nz, scenes.lattice_spacing(0.1, 40))
        vel = np.random.default_rng(3).normal(0, 40, (n, 3)).astype(np.float32)   # many wall hits
        sp = scenes.scene_params()
        kw = dict(rho0=sp["rho0"], stiffness=sp["stiffness"], viscosity=sp["viscosity"], central_mass=0.0,
                  gravity=sp["gravity"], time_step=sp["time_step"])
        r = refharness.RefSPH("golden")
        r.resize(n, *cfg["grid"], 96)
        r.set_params(**kw)
        r.set_state(pos, vel, np.ones(n, np.float32))
        o = OracleSPH(n=n, grid=cfg["grid"], examine=96, init_scene=False, **kw)
        o.set_params(central_pos=list(r.params().central_pos))
        o.set_state(pos, vel)
        for _ in range(3):
            r.step_phased(True, True, True)
            o.step(FULL, True, True)
            p, v, _ = r.state()
            assert eq(p, o.pos) and eq(v, o.vel)
            assert eq(r.acceleration(), o.acc)
