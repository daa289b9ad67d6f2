"""Generate tests/golden/*.npz by RUNNING THE UNMODIFIED REFERENCE
(oracle/_ref/libsphref_golden.so = /root/reference/src/sph.cpp compiled in
place with -O2 -ffp-contract=off, see oracle/Makefile).  Run in the build
container only:   python tests/golden/make_golden.py

The reference ships no tests and no golden vectors, so these
fixtures -- outputs of the reference's own code -- are the pin for the oracle
restatement and for the CUDA path on boxes where /root/reference is absent.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import scenes  # noqa: E402
from oracle.refharness import RefSPH  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def sparse_lists(idx, dist, cnt):
    m = np.arange(idx.shape[1])[None, :] < cnt[:, None]
    return idx[m].astype(np.uint32), dist[m].astype(np.float32)


def default_scene():
    """SPH::SPH() seeded sphere, SPH::step() verbatim, 2 steps."""
    r = RefSPH("golden")
    pos0, vel0, mass = r.state()
    out = dict(pos0=pos0, vel0=vel0)
    for s in (1, 2):
        r.step()
        ids, coords = r.voxels()
        start, members = r.grid()
        idx, dist = r.neighbors()
        cnt = r.neighbor_counts()
        nb, nd = sparse_lists(idx, dist, cnt)
        pos, vel, _ = r.state()
        ek, ep = r.energies()
        out.update({
            "voxel_ids_%d" % s: ids.astype(np.uint16), "grid_members_%d" % s: members.astype(np.uint16),
            "nbr_count_%d" % s: cnt.astype(np.int8), "nbr_idx_%d" % s: nb.astype(np.uint16),
            "nbr_dist_%d" % s: nd, "energy_%d" % s: np.array([ek, ep], np.float32),
        })
        if s == 1:   # FP fields only for the first step (chaotic afterwards)
            out.update({"grid_start_1": start, "density_1": r.density(), "acc_1": r.acceleration(),
                        "pos_1": pos, "vel_1": vel})
    # energies of a 20-step run (statistical parity of longer runs)
    en = []
    r2 = RefSPH("golden")
    for s in range(20):
        r2.step()
        en.append(r2.energies())
    out["energy_20"] = np.array(en, np.float32)
    np.savez_compressed(os.path.join(OUT, "default_scene.npz"), **out)


def full_scene(name, vel_sigma, steps, seed):
    """FULL neighbour mode on a small dam-break block: harness all-within-h
    search feeding the reference's computeDensity / computeAcceleration /
    integrate (+ dead wall code), uniform gravity on."""
    cfg = scenes.CONFIGS[name]
    nx, ny, nz = cfg["sites"]
    n = nx * ny * nz
    d = scenes.lattice_spacing(0.1, 40)
    pos = scenes.lattice_scene(nx, ny, nz, d)
    vel = (np.random.default_rng(seed).normal(0, vel_sigma, size=(n, 3))).astype(np.float32)
    sp = scenes.scene_params()
    E = 96
    r = RefSPH("golden")
    r.resize(n, *cfg["grid"], E)
    r.set_params(rho0=sp["rho0"], stiffness=sp["stiffness"], viscosity=sp["viscosity"],
                 central_mass=0.0, gravity=sp["gravity"], time_step=sp["time_step"])
    r.set_state(pos, vel, np.ones(n, np.float32))
    out = dict(pos0=pos, vel0=vel, examine=np.int32(E), central_pos=np.array(list(r.params().central_pos), np.float32))
    for s in range(1, steps + 1):
        mx = r.step_phased(True, True, True)
        assert mx <= E
        idx, dist = r.neighbors()
        cnt = r.neighbor_counts()
        nb, nd = sparse_lists(idx, dist, cnt)
        p, v, _ = r.state()
        ids, _ = r.voxels()
        out.update({
            "voxel_ids_%d" % s: ids.astype(np.uint16), "fine_keys_%d" % s: r.fine_keys().astype(np.uint16),
            "nbr_count_%d" % s: cnt.astype(np.int16), "nbr_idx_%d" % s: nb.astype(np.uint16),
            "density_%d" % s: r.density(), "pos_%d" % s: p, "vel_%d" % s: v,
        })
        if s == 1:
            out["acc_1"] = r.acceleration()
    np.savez_compressed(os.path.join(OUT, "full_%s.npz" % name), **out)


if __name__ == "__main__":
    default_scene()
    full_scene("dambreak_16k", vel_sigma=3.0, steps=2, seed=1)
    for f in sorted(os.listdir(OUT)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(OUT, f)) // 1024, "KiB")
