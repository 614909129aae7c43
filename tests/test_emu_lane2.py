"""The two-lanes-per-cell B-grid kernel (cice_b200/csrc/evp_lane2.cuh, EVP_B200_FUSED_VARIANT=40..53) run thread by thread ON THE
HOST: tests/emu_lane2.cpp compiles the kernel text unchanged with g++ -ffp-contract=off, makes every CUDA thread of a CTA a host
thread (pthread barriers for __syncthreads / the named barrier, a mailbox for the lane shuffle) and runs several subcycles of the
ping-pong on one block.  The result must equal the oracle bit for bit -- stresses, velocities including the on-rank cyclic ghost
copies, and the last subcycle's diagnostics -- for every patch shape and both lane mappings.  This is the CPU-side check of the
kernel's index logic; the GPU-side check is tests/test_gpu_parity.py::test_two_lane_variants."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from cice_b200 import abi, synth



class KParams(C.Structure):   # cice_b200/csrc/evp_internal.h
    _fields_ = [(n, C.c_double) for n in ("arlx1i", "denom1", "revp", "brlx", "e_factor", "epp2i", "capping", "Ktens", "u0", "cosw",
                                          "sinw", "rhow", "deltaminEVP")] + [("visc_method", C.c_int)]


ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHAPES = {0: "32x8 shuffle", 1: "16x8 shuffle", 2: "32x4 shuffle", 3: "16x16 shuffle", 4: "32x8 warp pairs", 5: "32x4 warp pairs"}


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("emu") / "libemu_lane2.so")
    cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
    if not os.path.exists(os.path.join(cuda_inc, "cuda_runtime.h")):
        pytest.skip("CUDA headers not found")
    cmd = ["/usr/bin/g++", "-O2", "-ffp-contract=off", "-std=c++17", "-fPIC", "-shared", "-pthread", "-I", cuda_inc,
           "-I", os.path.join(ROOT, "cice_b200", "csrc"), "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "emu_lane2.cpp"), "-o", out]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return C.CDLL(out)


def run_emulated(emu, c, shape):
    g, f = c.grid, c.copy_fields()
    assert g["nblocks"] == 1
    nxb, nyb = g["nx_block"], g["ny_block"]
    k = KParams(**{nm: float(c.params.get(nm, 0.0)) for nm, _ in KParams._fields_[:-1]})
    sig = np.ascontiguousarray(np.stack([f[nm][0] for nm in abi.STRESS]))
    geo = np.ascontiguousarray(np.stack([np.asarray(g[nm][0]) for nm in abi.GRID_STATIC]))
    inp = np.ascontiguousarray(np.stack([f[nm][0] for nm in ("cdn_ocnU", "aiU", "uocnU", "vocnU", "waterxU", "wateryU", "forcexU", "forceyU",
                                                             "umassdti", "fmU", "TbU")]))
    diag = np.zeros((4, nyb, nxb))
    u, v = f["uvel"][0].copy(), f["vvel"][0].copy()
    strength = np.ascontiguousarray(f["strength"][0])
    pd = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    pi = lambda a: a.ctypes.data_as(C.POINTER(C.c_int32))
    mT, mU = np.ascontiguousarray(f["iceTmask"][0]), np.ascontiguousarray(f["iceUmask"][0])
    cyc = abi.BNDY_NAMES["cyclic"]
    rc = emu.emu_lane2_run(shape, nxb, nyb, int(g["ew_boundary_type"] == cyc), int(g["ns_boundary_type"] == cyc), C.byref(k),
                           int(c.params["ndte"]), pi(mT), pi(mU), pd(sig), pd(u), pd(v), pd(geo), pd(strength), pd(inp), pd(diag))
    assert rc == 0
    out = {nm: sig[q] for q, nm in enumerate(abi.STRESS)}
    out.update(uvel=u, vvel=v, strintxU=diag[0], strintyU=diag[1], taubxU=diag[2], taubyU=diag[3])
    return out


CASES = {
    "tiny-4sub": dict(config="tiny", ndte=4, seed=131),
    "tiny-5sub-revised": dict(config="tiny", ndte=5, seed=132, revised_evp=True),
    "wide-3sub": dict(config="tiny", nx=70, ny=17, ndte=3, seed=133, kmt="continents"),
    "doubly-cyclic-3sub": dict(config="tiny", nx=33, ny=22, ndte=3, seed=134, ns="cyclic"),
    "S1-2sub": dict(config="tiny", ndte=2),
    # the N/E ghost T cells fall on the overlap row/column of the last patch (the `own` exception): 30 = 2*15, 21 = 3*7 = 7*3
    "aligned-16x8-32x4": dict(config="tiny", nx=30, ny=21, ndte=2, seed=135),
    "aligned-32x8-16x16": dict(config="tiny", nx=62, ny=30, ndte=2, seed=136, ns="cyclic"),   # 62 = 2*31, 30 = 2*15
}


@pytest.mark.parametrize("shape", sorted(SHAPES), ids=[SHAPES[s].replace(" ", "-") for s in sorted(SHAPES)])
@pytest.mark.parametrize("case", sorted(CASES))
def test_two_lane_kernel_on_the_host_equals_the_oracle(oracle_mod, emu, case, shape):
    c = synth.make_case(**CASES[case])
    if case == "tiny-4sub":
        c.params.update(capping=0.0, Ktens=0.2, cosw=0.9, sinw=0.4358898943540674)   # general branches too
    ref = c.copy_fields()
    oracle_mod.evp_run_bgrid(c.grid, c.params, ref)
    got = run_emulated(emu, c, shape)
    for nm in abi.STRESS + ("uvel", "vvel", "strintxU", "strintyU", "taubxU", "taubyU"):
        assert np.array_equal(got[nm].view(np.int64), ref[nm][0].view(np.int64)), (nm, int((got[nm] != ref[nm][0]).sum()))
