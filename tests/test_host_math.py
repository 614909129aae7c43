"""The DEVICE arithmetic compiled for the HOST: cice_b200/csrc/evp_math.cuh (stress_point, stepu_point -- the source the CUDA
kernels are built from) is compiled with g++ -ffp-contract=off behind a few intrinsic stand-ins (tests/host_math.cpp) and run for
one subcycle on one block; the result must equal the oracle bit for bit.  A slip in the device math is then caught on a machine
without a GPU.  (The MUFU-seeded div_fast / sqrt_fast paths are device-only and are covered by the -m gpu tests.)"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from cice_b200 import abi, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class KParams(C.Structure):   # cice_b200/csrc/evp_internal.h
    _fields_ = [(n, C.c_double) for n in ("arlx1i", "denom1", "revp", "brlx", "e_factor", "epp2i", "capping", "Ktens", "u0", "cosw",
                                          "sinw", "rhow", "deltaminEVP")] + [("visc_method", C.c_int)]


@pytest.fixture(scope="module")
def host_math(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("hm") / "libhost_math.so")
    cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
    if not os.path.exists(os.path.join(cuda_inc, "cuda_runtime.h")):
        pytest.skip("CUDA headers not found")
    cmd = ["/usr/bin/g++", "-O2", "-ffp-contract=off", "-std=c++17", "-fPIC", "-shared", "-I", cuda_inc, "-I", os.path.join(ROOT, "cice_b200", "csrc"),
           "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "host_math.cpp"), "-o", out]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return C.CDLL(out)


@pytest.mark.parametrize("kw", [dict(seed=121), dict(seed=122, revised_evp=True), dict(), dict(seed=123, kmt="continents")],
                         ids=["S2", "S2-revised", "S1", "S2-continents"])
def test_device_math_on_the_host_equals_the_oracle(oracle_mod, host_math, kw):
    c = synth.make_case("tiny", ndte=1, **kw)
    if kw.get("seed") == 121:
        c.params.update(capping=0.0, Ktens=0.2, cosw=0.9, sinw=0.4358898943540674)   # general branches too
    ref = c.copy_fields()
    oracle_mod.evp_run_bgrid(c.grid, c.params, ref)
    g, f = c.grid, c.copy_fields()
    nxb, nyb = g["nx_block"], g["ny_block"]
    n = nxb * nyb
    k = KParams(**{nm: float(c.params.get(nm, 0.0)) for nm, _ in KParams._fields_[:-1]})
    sig = np.ascontiguousarray(np.stack([f[nm][0] for nm in abi.STRESS]))
    geo = np.ascontiguousarray(np.stack([np.asarray(g[nm][0]) for nm in abi.GRID_STATIC]))
    assert abi.GRID_STATIC == ("dxT", "dyT", "dxhy", "dyhx", "cxp", "cyp", "cxm", "cym", "DminTarea", "uarear")
    inp = np.ascontiguousarray(np.stack([f[nm][0] for nm in ("cdn_ocnU", "aiU", "uocnU", "vocnU", "waterxU", "wateryU", "forcexU", "forceyU",
                                                             "umassdti", "fmU", "TbU")]))
    diag = np.ascontiguousarray(np.stack([f[nm][0] for nm in ("strintxU", "strintyU", "taubxU", "taubyU")]))
    u, v = f["uvel"][0].copy(), f["vvel"][0].copy()
    strength = np.ascontiguousarray(f["strength"][0])
    scratch = np.zeros((8, n))
    pd = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    pi = lambda a: a.ctypes.data_as(C.POINTER(C.c_int32))
    mT, mU = np.ascontiguousarray(f["iceTmask"][0]), np.ascontiguousarray(f["iceUmask"][0])
    rc = host_math.host_math_one_subcycle(nxb, nyb, int(g["ilo"][0]), int(g["ihi"][0]), int(g["jlo"][0]), int(g["jhi"][0]), C.byref(k),
                                          pi(mT), pi(mU), pd(sig), pd(u), pd(v), pd(geo), pd(strength), pd(inp), pd(diag), pd(scratch))
    assert rc == 0
    inner = (slice(int(g["jlo"][0]) - 1, int(g["jhi"][0])), slice(int(g["ilo"][0]) - 1, int(g["ihi"][0])))
    for q, nm in enumerate(abi.STRESS):     # stresses: every listed T cell (interior + N/E ghost); compare the whole block
        assert np.array_equal(sig[q].view(np.int64), ref[nm][0].view(np.int64)), nm
    for q, nm in enumerate(("strintxU", "strintyU", "taubxU", "taubyU")):
        assert np.array_equal(diag[q].view(np.int64), ref[nm][0].view(np.int64)), nm
    assert np.array_equal(u[inner].view(np.int64), ref["uvel"][0][inner].view(np.int64))    # ghosts: the halo is not part of this check
    assert np.array_equal(v[inner].view(np.int64), ref["vvel"][0][inner].view(np.int64))
