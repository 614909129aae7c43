"""Build cice_b200/_build/libevp_b200.so with nvcc for sm_100a (in-tree; travels with gpurun).

    python -m cice_b200.build [--force]

The kernel translation unit is compiled twice: namespace `exact` with -fmad=false (bit-identical
to the CPU oracle) and namespace `fast` with nvcc's default FMA contraction.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "_build")
LIB = os.path.join(OUT, "libevp_b200.so")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ARCH + ["-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "-ccbin", "/usr/bin/g++",
                 "-I", os.path.join(ROOT, "include"), "-I", SRC]

UNITS = [
    ("evp_kernels_exact.o", "evp_kernels.cu", ["-DEVP_NS=exact", "-fmad=false"]),
    ("evp_kernels_fast.o", "evp_kernels.cu", ["-DEVP_NS=fast"]),
    ("evp_persist_exact.o", "evp_persist.cu", ["-DEVP_NS=exact", "-fmad=false"]),
    ("evp_persist_fast.o", "evp_persist.cu", ["-DEVP_NS=fast"]),
    ("evp_tstream_exact.o", "evp_tstream.cu", ["-DEVP_NS=exact", "-fmad=false"]),
    ("evp_tstream_fast.o", "evp_tstream.cu", ["-DEVP_NS=fast"]),
    ("evp_cgrid_exact.o", "evp_cgrid.cu", ["-DEVP_NS=exact", "-fmad=false"]),
    ("evp_cgrid_fast.o", "evp_cgrid.cu", ["-DEVP_NS=fast"]),
    ("evp_halo.o", "evp_halo.cu", []),
    ("evp_abi.o", "evp_abi.cu", []),
]


def _sources():
    return [os.path.join(SRC, f) for f in os.listdir(SRC)] + [os.path.join(ROOT, "include", "evp_b200.h")]


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(s) > t for s in _sources() + [os.path.abspath(__file__)])


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    os.makedirs(OUT, exist_ok=True)
    units = [u for u in UNITS if os.path.exists(os.path.join(SRC, u[1]))]

    def cc(u):
        obj, src, extra = u
        cmd = [NVCC] + COMMON + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(SRC, src), "-o", os.path.join(OUT, obj)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode:
            raise RuntimeError("nvcc failed: " + " ".join(cmd) + "\n" + r.stdout + r.stderr)
        return obj, r.stderr

    with ThreadPoolExecutor(max_workers=6) as ex:
        logs = list(ex.map(cc, units))
    if verbose:
        for obj, log in logs:
            print("==", obj)
            print(log)
    cmd = [NVCC] + ARCH + ["-shared", "-ccbin", "/usr/bin/g++", "-o", LIB] + [os.path.join(OUT, u[0]) for u in units] + ["-lnccl"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode:
        raise RuntimeError("link failed: " + " ".join(cmd) + "\n" + r.stdout + r.stderr)
    return LIB


HOST_LIB = os.path.join(OUT, "libdyn_evp_b200.so")
HOST_CALLER = os.path.join(OUT, "host_caller")


def build_host(force=False):
    """the C++ host mirror of the reference seam (host/dyn_evp_b200.cpp) and the test caller that
    stands in for the Fortran driver (tests/host_caller.cpp); plain g++, linked against libevp_b200.so."""
    build(force=False)
    srcs = [os.path.join(HERE, "host", "dyn_evp_b200.cpp"), os.path.join(HERE, "host", "dyn_evp_b200.hpp"),
            os.path.join(ROOT, "tests", "host_caller.cpp"), LIB]
    if not force and os.path.exists(HOST_CALLER) and all(os.path.getmtime(s) <= os.path.getmtime(HOST_CALLER) for s in srcs):
        return HOST_CALLER
    inc = ["-I", os.path.join(ROOT, "include"), "-I", os.path.join(HERE, "host")]
    cmds = [
        ["/usr/bin/g++", "-O2", "-std=c++17", "-fPIC", "-shared"] + inc + [srcs[0], "-o", HOST_LIB, "-L", OUT, "-levp_b200",
                                                                        "-Wl,-rpath,$ORIGIN"],
        ["/usr/bin/g++", "-O2", "-std=c++17"] + inc + [srcs[2], "-o", HOST_CALLER, "-L", OUT, "-ldyn_evp_b200", "-levp_b200",
                                                       "-Wl,-rpath,$ORIGIN", "-Wl,-rpath-link," + OUT],
    ]
    for cmd in cmds:
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode:
            raise RuntimeError("host build failed: " + " ".join(cmd) + "\n" + r.stdout + r.stderr)
    return HOST_CALLER


if __name__ == "__main__":
    if "--host" in sys.argv:
        print(build_host(force="--force" in sys.argv))
        sys.exit(0)
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
