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
    ps, _ = _gather(slabs, F.POSITION); pr = ref.download(F.POSITION)
    vs, _ = _gather(slabs, F.VELOCITY); vr = ref.download(F.VELOCITY)
    pb = np.flatnonzero(~((ps == pr) | (np.isnan(ps) & np.isnan(pr))).all(1))
    vb = np.flatnonzero(~((vs == vr) | (np.isnan(vs) & np.isnan(vr))).all(1))
    if pb.size or vb.size:
        print("   POS mismatch", pb.size, "VEL mismatch", vb.size)
        for i in pb[:4]:
            print("      ", i, "slab", ps[i], vs[i], "ref", pr[i], vr[i], "prev", prev[i], "acc slab", a[i], "acc ref", ar[i])
    if bad.size:
        i = bad[np.argmax(rel)]
        print("   worst", i, a[i], ar[i], "z", prev[i,2], "vz", vzp[i], "frac in voxel", prev[i,2]/0.2 - vzp[i])
    if bad.size and step >= 10:
        # ghost densities of every slab against the single run (ghost slots keep the pre-step position)
        rho_ref = ref.download(F.DENSITY)
        key = {tuple(p): i for i, p in enumerate(map(tuple, prev))}
        for r, s in enumerate(slabs):
            P = s.download(F.POSITION); R = s.download(F.DENSITY)
            own_pos, own_g = s.download_slab(F.POSITION)
            owned = set(own_g.tolist())
            nbad = nmatch = 0
            ex = []
            for slot in range(P.shape[0]):
                g = key.get(tuple(P[slot]))
                if g is None or g in owned:
                    continue
                z0, z1 = layers[r]
                zz = prev[g][2]
                inner = (z0 * 0.2 - 0.1 <= zz < z0 * 0.2) or (z1 * 0.2 <= zz < z1 * 0.2 + 0.1)
                if not inner:
                    continue
                nmatch += 1
                if R[slot] != rho_ref[g]:
                    nbad += 1
                    if len(ex) < 5:
                        ex.append((g, float(R[slot]), float(rho_ref[g]), prev[g].tolist()))
            # particles in this slab's arrays that are neither matched ghosts nor in its owned z-range
            z0, z1 = layers[r]
            odd = [(int(g), own_pos[i].tolist()) for i, g in enumerate(own_g) if not (z0 * 0.2 <= own_pos[i][2] < z1 * 0.2)]
            print("   slab", r, "owned outside its range (post-step positions):", len(odd), odd[:6])
            print("   slab", r, "ghost slots matched", nmatch, "density differs", nbad, ex)
        break
