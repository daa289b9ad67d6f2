"""ctypes binding of oracle/_build/libsph_oracle.so (oracle/sph_oracle.c).

TEST INFRASTRUCTURE ONLY -- the CPU restatement of the reference's step used as
the parity checker.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline leg may import this; the product never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(_HERE, "_build", "libsph_oracle.so")


class OracleParams(C.Structure):
    _fields_ = [
        ("particle_count", C.c_int),
        ("grid_x", C.c_int), ("grid_y", C.c_int), ("grid_z", C.c_int),
        ("examine_count", C.c_int),
        ("h", C.c_float),
        ("simulation_scale", C.c_float),
        ("time_step", C.c_float),
        ("rho0", C.c_float), ("stiffness", C.c_float), ("viscosity", C.c_float),
        ("damping", C.c_float), ("cfl_limit", C.c_float),
        ("grav_constant", C.c_float), ("central_mass", C.c_float),
        ("central_pos", C.c_float * 3),
        ("softening", C.c_float),
        ("gravity", C.c_float * 3),
        ("h2", C.c_float), ("h_times2", C.c_float), ("h_times2_inv", C.c_float),
        ("hs", C.c_float), ("hs2", C.c_float), ("hs6", C.c_float), ("hs9", C.c_float),
        ("kernel1", C.c_float), ("kernel2", C.c_float), ("kernel3", C.c_float),
        ("max_x", C.c_float), ("max_y", C.c_float), ("max_z", C.c_float),
        ("cfl_limit2", C.c_float),
    ]


def build(force=False):
    src = os.path.join(_HERE, "sph_oracle.c")
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "port"], stdout=subprocess.DEVNULL)
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB)
        for f in ("oracle_default_params", "oracle_derive", "oracle_init_sphere", "oracle_voxelize",
                  "oracle_build_lists", "oracle_order_cells_by_x", "oracle_find_sampled", "oracle_fine_keys",
                  "oracle_find_full",
                  "oracle_density", "oracle_acceleration", "oracle_integrate"):
            getattr(_lib, f).restype = None
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


SAMPLED, FULL = 0, 1


class OracleSPH:
    """CPU restatement of SPH (sph.cpp) with the harness switches (FULL
    neighbour mode, uniform gravity, wall collision)."""

    def __init__(self, n=None, grid=None, examine=32, init_scene=True, **kw):
        self.lib = lib()
        self.p = OracleParams()
        self.lib.oracle_default_params(C.byref(self.p))
        if n is not None:
            self.p.particle_count = int(n)
        if grid is not None:
            self.p.grid_x, self.p.grid_y, self.p.grid_z = [int(g) for g in grid]
        self.p.examine_count = int(examine)
        self.set_params(**kw)
        n = self.p.particle_count
        self.pos = np.zeros((n, 3), np.float32)
        self.vel = np.zeros((n, 3), np.float32)
        self.mass = np.ones(n, np.float32)
        self.rho = np.zeros(n, np.float32)
        self.acc = np.zeros((n, 3), np.float32)
        self.count = np.zeros(n, np.int32)
        self.ekin = self.epot = 0.0
        if init_scene:
            self.lib.oracle_init_sphere(C.byref(self.p), _p(self.pos), _p(self.vel))

    def set_params(self, **kw):
        for k, v in kw.items():
            if k in ("central_pos", "gravity"):
                arr = getattr(self.p, k)
                for i in range(3):
                    arr[i] = float(v[i])
            else:
                if not hasattr(self.p, k):
                    raise AttributeError(k)
                setattr(self.p, k, v)
        self.lib.oracle_derive(C.byref(self.p))

    @property
    def n(self):
        return self.p.particle_count

    def set_state(self, pos, vel, mass=None):
        self.pos[...] = np.asarray(pos, np.float32).reshape(-1, 3)
        self.vel[...] = np.asarray(vel, np.float32).reshape(-1, 3)
        if mass is not None:
            self.mass[...] = np.asarray(mass, np.float32)

    # ---- phases ---------------------------------------------------------
    def voxelize(self):
        n, p = self.n, self.p
        cells = p.grid_x * p.grid_y * p.grid_z
        self.voxel_ids = np.empty(n, np.int32)
        self.voxel_xyz = np.empty((n, 3), np.int32)
        self.lib.oracle_voxelize(C.byref(p), _p(self.pos), _p(self.voxel_ids), _p(self.voxel_xyz))
        self.start = np.empty(cells + 1, np.int32)
        self.members = np.empty(n, np.uint32)
        self.lib.oracle_build_lists(n, cells, _p(self.voxel_ids), _p(self.start), _p(self.members))

    def find(self, mode):
        n, p = self.n, self.p
        E = p.examine_count
        self.nbr = np.zeros((n, E), np.uint32)
        self.dist = np.zeros((n, E), np.float32)
        if mode == SAMPLED:
            self.lib.oracle_find_sampled(C.byref(p), _p(self.pos), _p(self.voxel_xyz), _p(self.start),
                                         _p(self.members), _p(self.nbr), _p(self.dist), _p(self.count))
        else:
            fcells = 8 * p.grid_x * p.grid_y * p.grid_z
            self.fine_xyz = np.empty((n, 3), np.int32)
            self.fine_keys = np.empty(n, np.int32)
            self.lib.oracle_fine_keys(C.byref(p), _p(self.pos), _p(self.voxel_xyz), _p(self.fine_xyz),
                                      _p(self.fine_keys))
            self.fstart = np.empty(fcells + 1, np.int32)
            self.fmembers = np.empty(n, np.uint32)
            self.lib.oracle_build_lists(n, fcells, _p(self.fine_keys), _p(self.fstart), _p(self.fmembers))
            self.lib.oracle_order_cells_by_x(fcells, _p(self.fstart), _p(self.fmembers), _p(self.pos))
            self.lib.oracle_find_full(C.byref(p), _p(self.pos), _p(self.fine_xyz), _p(self.fstart),
                                      _p(self.fmembers), _p(self.nbr), _p(self.dist), _p(self.count))
            if int(self.count.max(initial=0)) > E:
                raise RuntimeError("examine_count %d too small for FULL mode (max %d)" % (E, self.count.max()))

    def compute_density(self):
        self.lib.oracle_density(C.byref(self.p), _p(self.mass), _p(self.nbr), _p(self.dist), _p(self.count),
                                _p(self.rho))

    def compute_acceleration(self, use_gravity=False):
        self.lib.oracle_acceleration(C.byref(self.p), _p(self.pos), _p(self.vel), _p(self.mass), _p(self.rho),
                                     _p(self.nbr), _p(self.dist), _p(self.count), int(use_gravity), _p(self.acc))

    def integrate(self, use_gravity=False, use_walls=False):
        ek, ep = C.c_float(), C.c_float()
        self.lib.oracle_integrate(C.byref(self.p), _p(self.pos), _p(self.vel), _p(self.acc), _p(self.mass),
                                  int(use_gravity), int(use_walls), C.byref(ek), C.byref(ep))
        self.ekin, self.epot = ek.value, ep.value

    def step_density_only(self, mode=SAMPLED):
        """The first three phases of step(): binning, neighbour search, density (large-N checks)."""
        self.voxelize()
        self.find(mode)
        self.compute_density()

    def step(self, mode=SAMPLED, use_gravity=False, use_walls=False):
        """SPH::step() order (sph.cpp:190-304)."""
        self.voxelize()
        self.find(mode)
        self.compute_density()
        self.compute_acceleration(use_gravity)
        self.integrate(use_gravity, use_walls)
