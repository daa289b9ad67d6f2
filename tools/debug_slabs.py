import numpy as np, sys
sys.path.insert(0, '.')
import smoothed_particle_hydrodynamics_b200 as S
from oracle import scenes
sys.path.insert(0, 'tests')
from test_gpu_slabs import _params, _gather
F = S.Field
nranks = int(sys.argv[1]) if len(sys.argv) > 1 else 4
cfg = dict(scenes.CONFIGS["dambreak_128k"]); cfg["grid"] = (40, 16, 32)
nx, ny, nz = cfg["sites"]; n = nx*ny*nz
d = scenes.lattice_spacing(0.1, 40)
pos = scenes.lattice_scene(nx, ny, nz, d, origin=(0.0, 0.0, 0.9))
rng = np.random.default_rng(5)
vel = rng.normal(0, 2.0, (n, 3)).astype(np.float32); vel[:, 2] += 25.0
mass = (rng.random(n) * 0.2 + 0.9).astype(np.float32)
ref = S.SPH(_params(cfg, n), init_scene=False); ref.upload(pos, vel, mass)
p = _params(cfg, n)
layers = S.slab_layers(cfg["grid"][2], nranks)
vz = S.voxel_layer(pos[:, 2], ref.derived.h_times2_inv, cfg["grid"][2])
slabs = []
for r, (z0, z1) in enumerate(layers):
    s = S.SlabSPH(p, r, nranks, z0, z1)
    own = np.flatnonzero((vz >= z0) & (vz < z1))
    s.upload_slab(pos[own], vel[own], mass[own], own.astype(np.uint32)); slabs.append(s)
print("layers", layers, "owned", [s.local_count() for s in slabs])
for step in range(1, 21):
    prev = ref.download(F.POSITION)
    ref.step_n(1); S.step_virtual_slabs(slabs, 1)
    a, g = _gather(slabs, F.ACCELERATION); ar = ref.download(F.ACCELERATION)
    nanmis = (np.isnan(a) != np.isnan(ar)).any(1).sum()
    bad = np.flatnonzero((a != ar).any(1) & ~np.isnan(ar).any(1))
    vzp = S.voxel_layer(prev[:, 2], ref.derived.h_times2_inv, cfg["grid"][2])
    rel = np.abs(a[bad]-ar[bad]).max(1)/np.maximum(np.linalg.norm(ar[bad],axis=1),1e-30) if bad.size else np.zeros(0)
    print("step", step, "nanmis", nanmis, "mismatch", bad.size, "layers", np.unique(vzp[bad]), "max rel", rel.max() if bad.size else 0, "counts", [s.local_count() for s in slabs],
          "rho eq", np.array_equal(_gather(slabs, F.DENSITY)[0], ref.download(F.DENSITY)))
    if bad.size:
        i = bad[np.argmax(rel)]
        print("   worst", i, a[i], ar[i], "z", prev[i,2], "vz", vzp[i], "frac in voxel", prev[i,2]/0.2 - vzp[i])
