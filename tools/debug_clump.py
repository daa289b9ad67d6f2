import numpy as np, sys
sys.path.insert(0, '.')
import smoothed_particle_hydrodynamics_b200 as S
from oracle.port import OracleSPH, FULL
from oracle import scenes
F = S.Field
rng = np.random.default_rng(3)
parts = [rng.random((6000, 3)) * np.array([2.4, 1.6, 1.6]), rng.normal(0, 0.10, (3000, 3)) + np.array([0.83, 0.79, 0.81]),
 rng.normal(0, 0.20, (4000, 3)) + np.array([1.6, 0.6, 1.0]), rng.random((300, 3)) * 0.2 - 0.25, rng.random((300, 3)) * 0.2 + np.array([2.4, 1.6, 1.6])]
pos = np.concatenate(parts).astype(np.float32); pos[100]=pos[101]; pos[200,1]=np.nan
n = len(pos); vel = rng.normal(0,1,(n,3)).astype(np.float32)
sp = scenes.scene_params()
p = S.default_params(particle_count=n, grid=(12,8,8), examine_count=1024, neighbor_mode=S.FULL, use_uniform_gravity=1, use_wall_collision=0,
   rho0=sp["rho0"], stiffness=sp["stiffness"], viscosity=sp["viscosity"], central_mass=0.0, gravity=sp["gravity"], time_step=sp["time_step"], kernel_variant=1)
sph = S.SPH(p, init_scene=False)
o = OracleSPH(n=n, grid=(12,8,8), examine=1024, init_scene=False, rho0=p.rho0, stiffness=p.stiffness, viscosity=p.viscosity, central_mass=0.0, gravity=list(p.gravity), time_step=p.time_step)
sph.upload(pos, vel); o.set_state(pos, vel)
o.step(FULL, True, False); sph.step_n(1)
acc = sph.download(F.ACCELERATION); rho = sph.download(F.DENSITY)
fin = np.isfinite(o.acc).all(1)
err = np.abs(acc.astype(np.float64)-o.acc).max(1)/np.maximum(np.linalg.norm(o.acc.astype(np.float64),axis=1),1e-30)
err[~fin] = 0
for i in np.argsort(err)[-6:]:
    pi = (o.rho[i]-p.rho0)*p.stiffness
    print(i, "err", err[i], "gpu", acc[i], "ref", o.acc[i], "rho", rho[i], o.rho[i], "p_i", pi, "cnt", o.count[i], "s", p.viscosity/pi if pi>0 else p.viscosity)
