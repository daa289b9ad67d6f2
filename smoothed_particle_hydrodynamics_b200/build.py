"""Builds smoothed_particle_hydrodynamics_b200/libsphb200.so in tree with nvcc for sm_100a.

    python -m smoothed_particle_hydrodynamics_b200.build [--force] [--verbose]

nvcc cross-compiles without a GPU.  The .so is git-ignored but travels to the GPU
box with the repo snapshot.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libsphb200.so")
OBJ = os.path.join(HERE, "build")

CU_SOURCES = ["sph_capi.cu", "sph_grid.cu", "sph_sampled.cu", "sph_full.cu", "sph_comm.cu"]
CPP_SOURCES = ["sph_scene.cpp"]
HEADERS = ["sph_internal.h", "sph_math.cuh", os.path.join(ROOT, "include", "sphb200.h")]

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-ffp-contract=off", "-I", os.path.join(ROOT, "include"), "-I", CSRC,
]


def _newer(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    hdrs = [h if os.path.isabs(h) else os.path.join(CSRC, h) for h in HEADERS]
    objs = []
    procs = []
    for src in CU_SOURCES + CPP_SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ, os.path.splitext(src)[0] + ".o")
        objs.append(o)
        if force or _newer(o, [s, os.path.abspath(__file__)] + hdrs):
            cmd = [NVCC] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o]
            procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            failed = True
            sys.stderr.write("nvcc failed on %s:\n%s\n" % (src, out))
        elif verbose and out:
            sys.stdout.write("== %s\n%s\n" % (src, out))
    if failed:
        raise RuntimeError("build of libsphb200.so failed")
    if force or procs or _newer(LIB, objs):
        cmd = [NVCC, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a",
                                                    "-Xcompiler", "-fPIC", "-cudart", "static", "-ldl"]
        subprocess.check_call(cmd)
    return LIB


HOST = os.path.join(HERE, "host")
HEADLESS = os.path.join(HERE, "sph_headless")
FACADE_CHECK = os.path.join(HERE, "facade_check")


def build_host(force=False):
    """The C++ facade (host/sph.cpp, reference `class SPH` interface) + the headless driver."""
    hdrs = [os.path.join(HOST, f) for f in ("sph.h", "particle.h", "vec3.h")] + [LIB]
    # sph_headless = the reference's `./sph r`; facade_check = GPU checks of the readback path (tests/test_gpu_headless.py)
    for exe, main in ((HEADLESS, "headless_main.cpp"), (FACADE_CHECK, "facade_check.cpp")):
        srcs = [os.path.join(HOST, f) for f in ("sph.cpp", main)]
        if force or _newer(exe, srcs + hdrs):
            subprocess.check_call(["g++", "-std=c++11", "-O2", "-pthread", "-I", os.path.join(ROOT, "include"), "-I", HOST] +
                                  srcs + ["-L", HERE, "-lsphb200", "-Wl,-rpath,$ORIGIN", "-o", exe])
    return HEADLESS


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
    print(build_host(force="--force" in sys.argv))
