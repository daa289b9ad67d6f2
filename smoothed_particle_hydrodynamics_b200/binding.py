"""ctypes binding of include/sphb200.h + a Python mirror of the reference's
``class SPH`` public interface (src/sph.h:20-84) for tests and benchmarks."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

SAMPLED, FULL = 0, 1


class SphError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("sphb200 error %d: %s" % (code, msg))
        self.code = code


class SphParams(C.Structure):
    _fields_ = [
        ("particle_count", C.c_int),
        ("grid_x", C.c_int), ("grid_y", C.c_int), ("grid_z", C.c_int),
        ("examine_count", C.c_int),
        ("neighbor_mode", C.c_int),
        ("use_uniform_gravity", C.c_int),
        ("use_wall_collision", C.c_int),
        ("h", C.c_float),
        ("simulation_scale", C.c_float),
        ("time_step", C.c_float),
        ("rho0", C.c_float),
        ("stiffness", C.c_float),
        ("viscosity", C.c_float),
        ("damping", C.c_float),
        ("cfl_limit", C.c_float),
        ("grav_constant", C.c_float),
        ("central_mass", C.c_float),
        ("central_pos", C.c_float * 3),
        ("softening", C.c_float),
        ("gravity", C.c_float * 3),
        ("kernel_variant", C.c_int),
        ("enable_timers", C.c_int),
        ("reserved", C.c_int * 6),
    ]


class SphDerived(C.Structure):
    _fields_ = [
        ("h2", C.c_float), ("h_times2", C.c_float), ("h_times2_inv", C.c_float),
        ("h_scaled", C.c_float), ("h_scaled2", C.c_float), ("h_scaled6", C.c_float), ("h_scaled9", C.c_float),
        ("kernel1", C.c_float), ("kernel2", C.c_float), ("kernel3", C.c_float),
        ("cell_size", C.c_float), ("max_x", C.c_float), ("max_y", C.c_float), ("max_z", C.c_float),
        ("cfl_limit2", C.c_float),
        ("central_pos", C.c_float * 3), ("softening", C.c_float),
        ("grid_cell_count", C.c_int),
        ("total_steps", C.c_int),
    ]


class SphStepReport(C.Structure):
    _fields_ = [("e_kin", C.c_float), ("e_pot", C.c_float), ("nbr_total", C.c_longlong), ("nbr_max", C.c_int),
                ("nbr_min", C.c_int), ("phase_ms", C.c_float * 6), ("step_index", C.c_longlong)]


class SphSceneLattice(C.Structure):
    _fields_ = [("nx", C.c_int), ("ny", C.c_int), ("nz", C.c_int), ("spacing", C.c_float),
                ("origin", C.c_float * 3), ("seed", C.c_uint32)]


SCENES = {"dambreak_16k": 0, "dambreak_128k": 1, "dambreak_1m": 2, "dambreak_16m": 3, "boxdrop_16m": 4}


class Field:
    POSITION, VELOCITY, MASS, DENSITY, ACCELERATION, NEIGHBOR_COUNT, VOXEL_ID, VOXEL_COORD = range(8)
    GRID_START, GRID_MEMBERS, CELL_COUNT, NEIGHBOR_INDEX, NEIGHBOR_DISTANCE, FINE_KEY = range(8, 14)


# every symbol include/sphb200.h declares: (name, argtypes, restype)
_VP = C.c_void_p
API = [
    ("sphb200_default_params", [C.POINTER(SphParams)], C.c_int),
    ("sphb200_derive", [C.POINTER(SphParams), C.POINTER(SphDerived)], C.c_int),
    ("sphb200_create", [C.POINTER(SphParams), C.c_int, C.POINTER(_VP)], C.c_int),
    ("sphb200_destroy", [_VP], C.c_int),
    ("sphb200_last_error", [_VP], C.c_char_p),
    ("sphb200_get_params", [_VP, C.POINTER(SphParams)], C.c_int),
    ("sphb200_set_params", [_VP, C.POINTER(SphParams)], C.c_int),
    ("sphb200_get_derived", [_VP, C.POINTER(SphDerived)], C.c_int),
    ("sphb200_set_stream", [_VP, _VP], C.c_int),
    ("sphb200_scene_sphere", [C.POINTER(SphParams), _VP, _VP], C.c_int),
    ("sphb200_scene_lattice", [C.c_int, C.c_int, C.c_int, C.c_float, C.POINTER(C.c_float), C.c_uint32,
                               C.c_longlong, C.c_longlong, _VP], C.c_int),
    ("sphb200_scene_config", [C.c_int, C.c_float, C.POINTER(SphParams), C.POINTER(SphSceneLattice)], C.c_int),
    ("sphb200_scene_generate", [C.POINTER(SphSceneLattice), C.c_longlong, C.c_longlong, _VP, _VP], C.c_int),
    ("sphb200_upload_state", [_VP, _VP, _VP, _VP], C.c_int),
    ("sphb200_download", [_VP, C.c_int, _VP, C.c_size_t], C.c_int),
    ("sphb200_step", [_VP, C.c_int], C.c_int),
    ("sphb200_synchronize", [_VP], C.c_int),
    ("sphb200_step_host", [_VP, _VP, _VP, _VP], C.c_int),
    ("sphb200_build_neighbor_lists", [_VP], C.c_int),
    ("sphb200_build_neighbor_lists_visited", [_VP], C.c_int),
    ("sphb200_get_step_report", [_VP, C.POINTER(SphStepReport)], C.c_int),
    ("sphb200_snapshot_request", [_VP, C.c_int], C.c_int),
    ("sphb200_snapshot_read", [_VP, C.c_int, _VP, C.c_size_t, _VP, C.c_size_t, C.POINTER(C.c_longlong)], C.c_int),
    ("sphb200_get_energies", [_VP, C.POINTER(C.c_float), C.POINTER(C.c_float)], C.c_int),
    ("sphb200_get_neighbor_stats", [_VP, C.POINTER(C.c_longlong), C.POINTER(C.c_int), C.POINTER(C.c_int)], C.c_int),
    ("sphb200_get_timings", [_VP, C.POINTER(C.c_float)], C.c_int),
    ("sphb200_get_launch_count", [_VP, C.POINTER(C.c_longlong)], C.c_int),
    ("sphb200_comm_unique_id", [_VP], C.c_int),
    ("sphb200_comm_init", [_VP, C.c_int, C.c_int, _VP, C.c_int, C.c_int], C.c_int),
    ("sphb200_get_local_count", [_VP, C.POINTER(C.c_int), C.POINTER(C.c_int)], C.c_int),
    ("sphb200_upload_slab", [_VP, C.c_int, _VP, _VP, _VP, _VP], C.c_int),
    ("sphb200_download_slab", [_VP, C.c_int, _VP, C.c_size_t, _VP, C.POINTER(C.c_int)], C.c_int),
    ("sphb200_slab_pack", [_VP], C.c_int),
    ("sphb200_slab_transfer", [_VP, C.c_int, _VP], C.c_int),
    ("sphb200_slab_unpack", [_VP], C.c_int),
    ("sphb200_slab_step_local", [_VP], C.c_int),
    ("sphb200_slab_set_halo_capacity", [_VP, C.c_longlong], C.c_int),
    ("sphb200_slab_connect", [_VP, _VP], C.c_int),
    ("sphb200_slab_put_mode", [_VP], C.c_int),
    ("sphb200_slab_status", [_VP], C.c_int),
]


def lib_path():
    # SPHB200_LIB: an alternative build of the same library (tools/build_variant.py, A/B timing only)
    return os.environ.get("SPHB200_LIB") or os.path.join(_HERE, "libsphb200.so")


_lib = None


def lib():
    """Loads libsphb200.so (built by `python -m smoothed_particle_hydrodynamics_b200.build`).
    Fails loudly when it is missing: there is no other implementation to fall back to."""
    global _lib
    if _lib is None:
        path = lib_path()
        if not os.path.exists(path):
            raise SphError(-2, "%s is missing: run `python -m smoothed_particle_hydrodynamics_b200.build` "
                               "(there is no CPU fallback)" % path)
        L = C.CDLL(path)
        for name, args, res in API:
            fn = getattr(L, name)
            fn.argtypes = args
            fn.restype = res
        _lib = L
    return _lib


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def default_params(**kw):
    p = SphParams()
    rc = lib().sphb200_default_params(C.byref(p))
    if rc:
        raise SphError(rc, "default_params")
    _apply(p, kw)
    return p


def _apply(p, kw):
    for k, v in kw.items():
        if k in ("central_pos", "gravity"):
            arr = getattr(p, k)
            for i in range(3):
                arr[i] = float(v[i])
        elif k == "grid":
            p.grid_x, p.grid_y, p.grid_z = [int(g) for g in v]
        else:
            if not hasattr(p, k):
                raise AttributeError("SphParams has no field %r" % k)
            setattr(p, k, v)


def derive(p):
    d = SphDerived()
    rc = lib().sphb200_derive(C.byref(p), C.byref(d))
    if rc:
        raise SphError(rc, lib().sphb200_last_error(None).decode())
    return d


def scene_sphere(p):
    """initParticlePolitionsSphere (sph.cpp:361-425): the constructor's scene."""
    pos = np.empty((p.particle_count, 3), np.float32)
    vel = np.empty((p.particle_count, 3), np.float32)
    rc = lib().sphb200_scene_sphere(C.byref(p), _ptr(pos), _ptr(vel))
    if rc:
        raise SphError(rc, "scene_sphere")
    return pos, vel


def scene_lattice(nx, ny, nz, spacing, origin=(0.0, 0.0, 0.0), seed=42, first_id=0, count=None, out=None):
    if count is None:
        count = nx * ny * nz - first_id
    pos = out if out is not None else np.empty((count, 3), np.float32)
    org = (C.c_float * 3)(*[float(o) for o in origin])
    rc = lib().sphb200_scene_lattice(nx, ny, nz, float(spacing), org, seed, first_id, count, _ptr(pos))
    if rc:
        raise SphError(rc, "scene_lattice: bad arguments")
    return pos


def scene_config(name, nu=40.0, **overrides):
    """(SphParams, SphSceneLattice) of a named throughput scene (sphb200_scene_config)."""
    p, lat = SphParams(), SphSceneLattice()
    rc = lib().sphb200_scene_config(SCENES[name] if isinstance(name, str) else int(name), float(nu),
                                    C.byref(p), C.byref(lat))
    if rc:
        raise SphError(rc, "scene_config: unknown scene or bad nu")
    _apply(p, overrides)
    return p, lat


def scene_generate(lat, first_id=0, count=None, out=None, vel_out=None):
    """Positions (float32 [count,3]) of a configured scene; vel_out, when given, is zeroed."""
    if count is None:
        count = lat.nx * lat.ny * lat.nz - first_id
    pos = out if out is not None else np.empty((count, 3), np.float32)
    rc = lib().sphb200_scene_generate(C.byref(lat), first_id, count, _ptr(pos),
                                      _ptr(vel_out) if vel_out is not None else None)
    if rc:
        raise SphError(rc, "scene_generate: bad arguments")
    return pos


class Particle:
    """Host mirror with the reference's member names (src/particle.h:13-18)."""

    def __init__(self, n):
        self.mMass = np.zeros(n, np.float32)
        self.mDensity = np.zeros(n, np.float32)
        self.mPosition = np.zeros(3 * n, np.float32)
        self.mVelocity = np.zeros(3 * n, np.float32)
        self.mAcceleration = np.zeros(3 * n, np.float32)
        self.mNeighborCount = np.zeros(n, np.int32)


class SPH:
    """Python mirror of the reference's `class SPH` public interface
    (src/sph.h:20-84) over the C ABI.  `SPH()` with no arguments is the
    reference constructor: default parameters + the seeded sphere scene."""

    def __init__(self, params=None, device=-1, init_scene=None, **kw):
        self._lib = lib()
        self._h = _VP()
        p = params if params is not None else default_params()
        _apply(p, kw)
        rc = self._lib.sphb200_create(C.byref(p), device, C.byref(self._h))
        if rc:
            msg = self._lib.sphb200_last_error(self._h if self._h else None).decode()
            if self._h:
                self._lib.sphb200_destroy(self._h)
                self._h = _VP()
            raise SphError(rc, msg)
        self._particles = Particle(p.particle_count)
        if init_scene is None:
            init_scene = params is None and not kw
        if init_scene:
            pos, vel = scene_sphere(p)
            self.upload(pos, vel)

    # ---- plumbing ---------------------------------------------------------
    def _check(self, rc):
        if rc:
            raise SphError(rc, self._lib.sphb200_last_error(self._h).decode())

    def close(self):
        if self._h:
            self._lib.sphb200_destroy(self._h)
            self._h = _VP()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def params(self):
        p = SphParams()
        self._check(self._lib.sphb200_get_params(self._h, C.byref(p)))
        return p

    @property
    def derived(self):
        d = SphDerived()
        self._check(self._lib.sphb200_get_derived(self._h, C.byref(d)))
        return d

    def set_params(self, **kw):
        p = self.params
        _apply(p, kw)
        self._check(self._lib.sphb200_set_params(self._h, C.byref(p)))

    def set_stream(self, cuda_stream):
        self._check(self._lib.sphb200_set_stream(self._h, _VP(cuda_stream) if cuda_stream else None))

    def upload(self, pos, vel, mass=None):
        n = self.params.particle_count
        pos = np.ascontiguousarray(pos, np.float32).reshape(-1)
        vel = np.ascontiguousarray(vel, np.float32).reshape(-1)
        assert pos.size == 3 * n and vel.size == 3 * n, (pos.size, vel.size, n)
        if mass is not None:
            mass = np.ascontiguousarray(mass, np.float32).reshape(-1)
            assert mass.size == n
        self._check(self._lib.sphb200_upload_state(self._h, _ptr(pos), _ptr(vel), _ptr(mass)))

    def upload_ptr(self, pos_ptr, vel_ptr, mass_ptr=None):
        self._check(self._lib.sphb200_upload_state(self._h, _VP(pos_ptr), _VP(vel_ptr),
                                                   _VP(mass_ptr) if mass_ptr else None))

    def step_host_ptr(self, pos_ptr, vel_ptr, mass_ptr=None):
        self._check(self._lib.sphb200_step_host(self._h, _VP(pos_ptr), _VP(vel_ptr),
                                                _VP(mass_ptr) if mass_ptr else None))

    def download(self, field):
        p = self.params
        n, E = p.particle_count, p.examine_count
        cells = p.grid_x * p.grid_y * p.grid_z
        shape, dt = {
            Field.POSITION: ((n, 3), np.float32), Field.VELOCITY: ((n, 3), np.float32),
            Field.MASS: ((n,), np.float32), Field.DENSITY: ((n,), np.float32),
            Field.ACCELERATION: ((n, 3), np.float32), Field.NEIGHBOR_COUNT: ((n,), np.int32),
            Field.VOXEL_ID: ((n,), np.int32), Field.VOXEL_COORD: ((n, 3), np.int32),
            Field.GRID_START: ((cells + 1,), np.int32), Field.GRID_MEMBERS: ((n,), np.uint32),
            Field.CELL_COUNT: ((cells,), np.int32), Field.NEIGHBOR_INDEX: ((n, E), np.uint32),
            Field.NEIGHBOR_DISTANCE: ((n, E), np.float32), Field.FINE_KEY: ((n,), np.int32),
        }[field]
        out = np.empty(shape, dt)
        self._check(self._lib.sphb200_download(self._h, field, _ptr(out), out.nbytes))
        return out

    def step_n(self, n_steps):
        """n device-resident steps, asynchronous (sphb200_step)."""
        self._check(self._lib.sphb200_step(self._h, int(n_steps)))

    def synchronize(self):
        self._check(self._lib.sphb200_synchronize(self._h))

    def build_neighbor_lists(self, visited=False):
        """visited=True: from the hit-mask stream the force sweep of the last step walked."""
        if visited:
            self._check(self._lib.sphb200_build_neighbor_lists_visited(self._h))
        else:
            self._check(self._lib.sphb200_build_neighbor_lists(self._h))

    def step_report(self):
        r = SphStepReport()
        self._check(self._lib.sphb200_get_step_report(self._h, C.byref(r)))
        return r

    def snapshot_request(self, positions=True, cell_counts=False):
        self._check(self._lib.sphb200_snapshot_request(self._h, (1 if positions else 0) | (2 if cell_counts else 0)))

    def snapshot_read(self, have=-1, wait=False, positions=True, cell_counts=False):
        """(step_index, pos[n,3] or None, counts[cells] or None) of the newest completed snapshot;
        step_index == have means there is nothing newer."""
        p = self.params
        n, cells = p.particle_count, p.grid_x * p.grid_y * p.grid_z
        pos = np.empty((n, 3), np.float32) if positions else None
        cnt = np.empty(cells, np.int32) if cell_counts else None
        idx = C.c_longlong(have)
        self._check(self._lib.sphb200_snapshot_read(self._h, 1 if wait else 0, _ptr(pos), pos.nbytes if positions else 0,
                                                    _ptr(cnt), cnt.nbytes if cell_counts else 0, C.byref(idx)))
        return idx.value, pos, cnt

    def energies(self):
        ek, ep = C.c_float(), C.c_float()
        self._check(self._lib.sphb200_get_energies(self._h, C.byref(ek), C.byref(ep)))
        return ek.value, ep.value

    def neighbor_stats(self):
        t, mx, mn = C.c_longlong(), C.c_int(), C.c_int()
        self._check(self._lib.sphb200_get_neighbor_stats(self._h, C.byref(t), C.byref(mx), C.byref(mn)))
        return t.value, mx.value, mn.value

    def timings_ms(self):
        a = (C.c_float * 6)()
        self._check(self._lib.sphb200_get_timings(self._h, a))
        return list(a)

    def launch_count(self):
        v = C.c_longlong()
        self._check(self._lib.sphb200_get_launch_count(self._h, C.byref(v)))
        return v.value

    # ---- the reference's public interface (src/sph.h:22-70) -----------------
    def step(self):
        """SPH::step() (sph.cpp:190-304): one step, then refresh the host mirror
        the GL view reads (visualization.cpp:144-157)."""
        self.step_n(1)
        self._particles.mPosition[:] = self.download(Field.POSITION).reshape(-1)

    def getParticles(self):
        p = self._particles
        p.mPosition[:] = self.download(Field.POSITION).reshape(-1)
        p.mVelocity[:] = self.download(Field.VELOCITY).reshape(-1)
        p.mMass[:] = self.download(Field.MASS)
        p.mDensity[:] = self.download(Field.DENSITY)
        p.mAcceleration[:] = self.download(Field.ACCELERATION).reshape(-1)
        p.mNeighborCount[:] = self.download(Field.NEIGHBOR_COUNT)
        return p

    def getParticleCount(self):
        return self.params.particle_count

    def getGridCellCounts(self):
        p = self.params
        return p.grid_x, p.grid_y, p.grid_z

    def getParticleBounds(self):
        d = self.derived
        return d.max_x, d.max_y, d.max_z

    def getInteractionRadius2(self):
        return self.derived.h_scaled2

    def getCellSize(self):
        return self.derived.cell_size

    def getGrid(self):
        """mGrid as (start, members): cell c holds members[start[c]:start[c+1]]."""
        return self.download(Field.GRID_START), self.download(Field.GRID_MEMBERS)

    def getGravity(self):
        return tuple(self.params.gravity)

    def setGravity(self, g):
        self.set_params(gravity=g)

    def getStiffness(self):
        return self.params.stiffness

    def setStiffness(self, v):
        self.set_params(stiffness=v)

    def getViscosityScalar(self):
        return self.params.viscosity

    def setViscosityScalar(self, v):
        self.set_params(viscosity=v)

    def getTimeStep(self):
        return self.params.time_step

    def setTimeStep(self, v):
        self.set_params(time_step=v)

    def getDamping(self):
        return self.params.damping

    def setDamping(self, v):
        self.set_params(damping=v)

    def getCflLimit(self):
        return self.params.cfl_limit

    def setCflLimit(self, v):
        self.set_params(cfl_limit=v)


def slab_layers(grid_z, nranks, boundaries=None):
    """Voxel-layer ranges [z0, z1) of `nranks` z-slabs tiling [0, grid_z).
    Equal thickness unless explicit interior `boundaries` are given."""
    if boundaries is None:
        boundaries = [(grid_z * r) // nranks for r in range(1, nranks)]
    edges = [0] + list(boundaries) + [grid_z]
    assert len(edges) == nranks + 1 and all(b > a for a, b in zip(edges, edges[1:])), edges
    return [(edges[r], edges[r + 1]) for r in range(nranks)]


def voxel_layer(pos_z, h_times2_inv, grid_z):
    """Global voxel layer of a z coordinate, as voxelizeParticles computes it
    (sph.cpp:452-463): one f32 multiply, floor, clamp.  Host-side slab assignment."""
    v = np.floor(np.asarray(pos_z, np.float32) * np.float32(h_times2_inv))
    v = np.where(np.isfinite(v), v, -2147483648.0)
    return np.clip(v, 0, grid_z - 1).astype(np.int64)


class SlabSPH(SPH):
    """One z-slab of a multi-GPU run (one per rank / GPU).  `params.grid_z` is the
    GLOBAL grid and `params.particle_count` the slot capacity of this slab."""

    def __init__(self, params, rank, nranks, z0, z1, nccl_id=None, device=-1, halo_capacity=None):
        super().__init__(params, device=device, init_scene=False)
        self.rank, self.nranks, self.z0, self.z1 = rank, nranks, z0, z1
        idbuf = None
        if nccl_id is not None:
            idbuf = (C.c_ubyte * 128).from_buffer_copy(bytes(nccl_id))
        self._check(self._lib.sphb200_comm_init(self._h, rank, nranks, idbuf, z0, z1))
        if halo_capacity is not None:
            self._check(self._lib.sphb200_slab_set_halo_capacity(self._h, int(halo_capacity)))

    @staticmethod
    def unique_id():
        buf = (C.c_ubyte * 128)()
        rc = lib().sphb200_comm_unique_id(buf)
        if rc:
            raise SphError(rc, lib().sphb200_last_error(None).decode())
        return bytes(buf)

    def upload_slab(self, pos, vel, mass, gids):
        pos = np.ascontiguousarray(pos, np.float32).reshape(-1)
        vel = np.ascontiguousarray(vel, np.float32).reshape(-1)
        gids = np.ascontiguousarray(gids, np.uint32).reshape(-1)
        n = gids.size
        assert pos.size == 3 * n and vel.size == 3 * n
        if mass is not None:
            mass = np.ascontiguousarray(mass, np.float32).reshape(-1)
        self._check(self._lib.sphb200_upload_slab(self._h, n, _ptr(pos), _ptr(vel), _ptr(mass), _ptr(gids)))

    def download_slab(self, field):
        cap = self.params.particle_count
        comps, dt = {Field.POSITION: (3, np.float32), Field.VELOCITY: (3, np.float32), Field.MASS: (1, np.float32),
                     Field.DENSITY: (1, np.float32), Field.ACCELERATION: (3, np.float32),
                     Field.NEIGHBOR_COUNT: (1, np.int32)}[field]
        out = np.empty((cap, comps), dt)
        gids = np.empty(cap, np.uint32)
        cnt = C.c_int()
        self._check(self._lib.sphb200_download_slab(self._h, field, _ptr(out), out.nbytes, _ptr(gids), C.byref(cnt)))
        out = out[:cnt.value]
        return (out if comps > 1 else out[:, 0]), gids[:cnt.value]

    def local_count(self):
        o, g = C.c_int(), C.c_int()
        self._check(self._lib.sphb200_get_local_count(self._h, C.byref(o), C.byref(g)))
        return o.value, g.value

    def pack(self):
        self._check(self._lib.sphb200_slab_pack(self._h))

    def transfer_to(self, direction, other):
        self._check(self._lib.sphb200_slab_transfer(self._h, direction, other._h))

    def unpack(self):
        self._check(self._lib.sphb200_slab_unpack(self._h))

    def step_local(self):
        self._check(self._lib.sphb200_slab_step_local(self._h))

    def status(self):
        self._check(self._lib.sphb200_slab_status(self._h))

    def connect_up(self, upper):
        """Put mode between this virtual rank and the one above it (before the first exchange)."""
        self._check(self._lib.sphb200_slab_connect(self._h, upper._h))

    def put_mode(self):
        return bool(self._lib.sphb200_slab_put_mode(self._h))


def step_virtual_slabs(slabs, n_steps=1):
    """Steps a list of virtual-rank slabs (one process, any devices) in lockstep:
    pack everywhere, move each message to its neighbour, unpack, local step."""
    for _ in range(n_steps):
        for s in slabs:
            s.pack()
        for r, s in enumerate(slabs):
            if r > 0:
                s.transfer_to(0, slabs[r - 1])
            if r + 1 < len(slabs):
                s.transfer_to(1, slabs[r + 1])
        for s in slabs:
            s.unpack()
            s.step_local()
