import sys, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import smoothed_particle_hydrodynamics_b200 as S
from oracle import scenes
from oracle.port import FULL, OracleSPH
import test_gpu_parity as T
F = S.Field
cfg = scenes.CONFIGS["dambreak_16k"]; nx, ny, nz = cfg["sites"]; n = nx*ny*nz
pos = scenes.lattice_scene(nx, ny, nz, scenes.lattice_spacing(0.1, 40))
vel = np.random.default_rng(11).normal(0, 3.0, (n, 3)).astype(np.float32)
mass = (np.random.default_rng(5).random(n) * 0.2 + 0.9).astype(np.float32)
p = T._full_params(cfg, n, 96, 0)
sph = S.SPH(p, init_scene=False); o = T._full_oracle(cfg, n, 96, p)
sph.upload(pos, vel, mass); o.set_state(pos, vel, mass)
for s in range(2):
    o.step(FULL, True, True); sph.step_n(1)
    acc = sph.download(F.ACCELERATION).astype(np.float64); ref = o.acc.astype(np.float64)
    norm = np.linalg.norm(ref, axis=1); err = np.abs(acc-ref).max(axis=1)/np.maximum(norm,1e-30)
    i = int(err.argmax())
    print("step", s, "max", err.max(), "n>1e-4", (err>1e-4).sum(), "n>1e-5", (err>1e-5).sum(), "worst", i, acc[i], ref[i], "count", o.count[i], "rho", o.rho[i], "rho0", o.p.rho0)
    # viscosity vs pressure magnitude for the worst particle: recompute in float64 from the oracle list
    sph.upload(o.pos, o.vel, mass)
