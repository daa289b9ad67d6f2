"""ctypes binding of oracle/_ref/libsphref_{golden,timing}.so.

TEST INFRASTRUCTURE ONLY -- only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this.  The shared objects are
the UNMODIFIED reference (/root/reference/src/sph.cpp) compiled in place behind
oracle/ref_harness.cpp; see oracle/Makefile.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


class RefParams(C.Structure):
    _fields_ = [
        ("particle_count", C.c_int),
        ("grid_x", C.c_int), ("grid_y", C.c_int), ("grid_z", C.c_int),
        ("examine_count", C.c_int),
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
        ("kernel1", C.c_float), ("kernel2", C.c_float), ("kernel3", C.c_float),
        ("h2", C.c_float), ("h_times2", C.c_float), ("h_times2_inv", C.c_float),
        ("max_x", C.c_float), ("max_y", C.c_float), ("max_z", C.c_float),
    ]


def lib_path(kind="golden"):
    return os.path.join(_HERE, "_ref", "libsphref_%s.so" % kind)


def available(kind="golden"):
    return os.path.exists(lib_path(kind))


_libs = {}


def _load(kind):
    if kind in _libs:
        return _libs[kind]
    lib = C.CDLL(lib_path(kind))
    vp = C.c_void_p
    lib.ref_create.restype = vp
    for name, args, res in [
        ("ref_destroy", [vp], None),
        ("ref_resize", [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int], None),
        ("ref_get_params", [vp, C.POINTER(RefParams)], None),
        ("ref_set_params", [vp, C.POINTER(RefParams)], None),
        ("ref_set_state", [vp, vp, vp, vp], None),
        ("ref_get_state", [vp, vp, vp, vp], None),
        ("ref_get_density", [vp, vp], None),
        ("ref_set_density", [vp, vp], None),
        ("ref_get_acceleration", [vp, vp], None),
        ("ref_get_neighbor_counts", [vp, vp], None),
        ("ref_get_neighbors", [vp, vp, vp], None),
        ("ref_set_neighbors", [vp, vp, vp, vp], None),
        ("ref_get_voxels", [vp, vp, vp], None),
        ("ref_get_grid", [vp, vp, vp], None),
        ("ref_get_fine_keys", [vp, vp], None),
        ("ref_get_energies", [vp, C.POINTER(C.c_float), C.POINTER(C.c_float)], None),
        ("ref_get_neighbor_stats", [vp, C.POINTER(C.c_longlong), C.POINTER(C.c_int), C.POINTER(C.c_int)], None),
        ("ref_get_timers_ms", [vp, vp], None),
        ("ref_get_phase_ns", [vp, vp], None),
        ("ref_step", [vp], None),
        ("ref_run", [vp], None),
        ("ref_voxelize", [vp], None),
        ("ref_find_sampled", [vp], None),
        ("ref_find_full", [vp], C.c_int),
        ("ref_density", [vp], None),
        ("ref_accel", [vp, C.c_int], None),
        ("ref_integrate", [vp, C.c_int, C.c_int], None),
        ("ref_step_phased", [vp, C.c_int, C.c_int, C.c_int], C.c_int),
    ]:
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = res
    _libs[kind] = lib
    return lib


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class RefSPH:
    """The reference `SPH` object (seeded sphere scene by default), probed."""

    def __init__(self, kind="golden"):
        self.lib = _load(kind)
        self.h = C.c_void_p(self.lib.ref_create())

    # -- configuration -----------------------------------------------------
    def params(self):
        p = RefParams()
        self.lib.ref_get_params(self.h, C.byref(p))
        return p

    def set_params(self, **kw):
        p = self.params()
        for k, v in kw.items():
            if k in ("central_pos", "gravity"):
                arr = getattr(p, k)
                for i in range(3):
                    arr[i] = float(v[i])
            else:
                if not hasattr(p, k):
                    raise AttributeError(k)
                setattr(p, k, v)
        self.lib.ref_set_params(self.h, C.byref(p))

    def resize(self, n, gx, gy, gz, examine=32):
        self.lib.ref_resize(self.h, int(n), int(gx), int(gy), int(gz), int(examine))

    @property
    def n(self):
        return self.params().particle_count

    @property
    def examine(self):
        return self.params().examine_count

    # -- state ---------------------------------------------------------------
    def set_state(self, pos=None, vel=None, mass=None):
        def prep(a, k):
            if a is None:
                return None
            a = np.ascontiguousarray(a, dtype=np.float32).reshape(-1)
            assert a.size == self.n * k, (a.size, self.n, k)
            return a
        pos, vel, mass = prep(pos, 3), prep(vel, 3), prep(mass, 1)
        self.lib.ref_set_state(self.h, _ptr(pos), _ptr(vel), _ptr(mass))

    def state(self):
        n = self.n
        pos = np.empty((n, 3), np.float32)
        vel = np.empty((n, 3), np.float32)
        mass = np.empty(n, np.float32)
        self.lib.ref_get_state(self.h, _ptr(pos), _ptr(vel), _ptr(mass))
        return pos, vel, mass

    def density(self):
        a = np.empty(self.n, np.float32)
        self.lib.ref_get_density(self.h, _ptr(a))
        return a

    def set_density(self, rho):
        rho = np.ascontiguousarray(rho, dtype=np.float32)
        assert rho.size == self.n
        self.lib.ref_set_density(self.h, _ptr(rho))

    def acceleration(self):
        a = np.empty((self.n, 3), np.float32)
        self.lib.ref_get_acceleration(self.h, _ptr(a))
        return a

    def neighbor_counts(self):
        a = np.empty(self.n, np.int32)
        self.lib.ref_get_neighbor_counts(self.h, _ptr(a))
        return a

    def neighbors(self):
        n, e = self.n, self.examine
        idx = np.empty((n, e), np.uint32)
        dist = np.empty((n, e), np.float32)
        self.lib.ref_get_neighbors(self.h, _ptr(idx), _ptr(dist))
        return idx, dist

    def set_neighbors(self, idx, dist, counts):
        n, e = self.n, self.examine
        idx = np.ascontiguousarray(idx, dtype=np.uint32)
        dist = np.ascontiguousarray(dist, dtype=np.float32)
        counts = np.ascontiguousarray(counts, dtype=np.int32)
        assert idx.shape == (n, e) and dist.shape == (n, e) and counts.shape == (n,)
        self.lib.ref_set_neighbors(self.h, _ptr(idx), _ptr(dist), _ptr(counts))

    def voxels(self):
        n = self.n
        ids = np.empty(n, np.int32)
        coords = np.empty((n, 3), np.int32)
        self.lib.ref_get_voxels(self.h, _ptr(ids), _ptr(coords))
        return ids, coords

    def grid(self):
        p = self.params()
        cells = p.grid_x * p.grid_y * p.grid_z
        start = np.empty(cells + 1, np.int32)
        members = np.empty(p.particle_count, np.uint32)
        self.lib.ref_get_grid(self.h, _ptr(start), _ptr(members))
        return start, members

    def fine_keys(self):
        a = np.empty(self.n, np.int32)
        self.lib.ref_get_fine_keys(self.h, _ptr(a))
        return a

    def energies(self):
        ek, ep = C.c_float(), C.c_float()
        self.lib.ref_get_energies(self.h, C.byref(ek), C.byref(ep))
        return ek.value, ep.value

    def neighbor_stats(self):
        t, mx, mn = C.c_longlong(), C.c_int(), C.c_int()
        self.lib.ref_get_neighbor_stats(self.h, C.byref(t), C.byref(mx), C.byref(mn))
        return t.value, mx.value, mn.value

    def phase_ns(self):
        a = np.empty(6, np.float64)
        self.lib.ref_get_phase_ns(self.h, _ptr(a))
        return a

    # -- stepping ----------------------------------------------------------
    def step(self):
        """SPH::step() verbatim (sph.cpp:190-304)."""
        self.lib.ref_step(self.h)

    def voxelize(self):
        self.lib.ref_voxelize(self.h)

    def find_sampled(self):
        self.lib.ref_find_sampled(self.h)

    def find_full(self):
        return self.lib.ref_find_full(self.h)

    def compute_density(self):
        self.lib.ref_density(self.h)

    def compute_acceleration(self, use_gravity=False):
        self.lib.ref_accel(self.h, int(use_gravity))

    def integrate(self, use_gravity=False, use_walls=False):
        self.lib.ref_integrate(self.h, int(use_gravity), int(use_walls))

    def step_phased(self, full_mode=False, use_gravity=False, use_walls=False):
        """Phase loops in step()'s order with ns timers; returns max neighbour count."""
        return self.lib.ref_step_phased(self.h, int(full_mode), int(use_gravity), int(use_walls))

    def close(self):
        if self.h:
            self.lib.ref_destroy(self.h)
            self.h = None
