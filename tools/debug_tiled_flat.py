"""Tiled (packed) vs flat density sweep on the dense 128K lattice: where do they differ?"""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import smoothed_particle_hydrodynamics_b200 as S
from oracle import scenes
from test_gpu_parity import _full_params
F = S.Field
cfg = scenes.CONFIGS["dambreak_128k"]
nx, ny, nz = cfg["sites"]
n = nx * ny * nz
nu = float(sys.argv[1]) if len(sys.argv) > 1 else 60
pos = scenes.lattice_scene(nx, ny, nz, scenes.lattice_spacing(0.1, nu))
vel = np.zeros((n, 3), np.float32)
for steps in (1, 2, 3):
    out = []
    for variant in (0, 1):
        sph = S.SPH(_full_params(cfg, n, 160, variant), init_scene=False)
        sph.upload(pos, vel)
        sph.step_n(steps)
        out.append((sph.download(F.NEIGHBOR_COUNT), sph.download(F.DENSITY), sph.download(F.POSITION)))
        sph.close()
    rel = np.abs(out[0][1] - out[1][1]) / np.abs(out[1][1])
    bad = np.nonzero(rel > 1e-5)[0]
    print("steps", steps, "counts equal", np.array_equal(out[0][0], out[1][0]), "max rel rho", rel.max(), "n>1e-5", len(bad),
          "pos maxdiff", np.abs(out[0][2] - out[1][2]).max())
    for i in bad[:10]:
        print("  ", i, out[0][1][i], out[1][1][i], out[0][0][i], pos[i])
