"""KERNEL_PERSISTENT (cice_b200/csrc/evp_persist.cu) run thread by thread ON THE HOST, every CTA at once.

tests/emu_persist.cpp includes the kernel translation unit unchanged (EVP_HOST_EMU) on top of tests/cuda_emu.h's concurrent mode
(one host thread per CUDA thread of EVERY CTA; per-CTA barriers and dynamic shared memory; GCC atomics for the tiles' progress
counters) and takes tiling, shared-memory layout and (slot, thread) tables from the product's planner (evp_persist_plan.h) for a
pretended SM count and CTA size.  The whole loop on one block must equal the oracle bit for bit: stresses, velocities including the
on-rank cyclic ghost copies, the last subcycle's diagnostics.  What the -m gpu tests add: the hardware's division / square-root
seeds, real memory ordering between SMs, timing."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from cice_b200 import abi, synth
from tests.test_emu_bgrid import KParams, ROOT


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("emu") / "libemu_persist.so")
    cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
    if not os.path.exists(os.path.join(cuda_inc, "cuda_runtime.h")):
        pytest.skip("CUDA headers not found")
    cmd = ["/usr/bin/g++", "-O2", "-ffp-contract=off", "-std=c++17", "-fPIC", "-shared", "-pthread", "-I", cuda_inc,
           "-I", os.path.join(ROOT, "cice_b200", "csrc"), "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "tests"),
           os.path.join(ROOT, "tests", "emu_persist.cpp"), "-o", out]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return C.CDLL(out)


def run_emulated(emu, c, nthreads, num_sms, force_k63):
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
    info = np.zeros(8, dtype=np.int32)
    rc = emu.emu_persist_run(nthreads, num_sms, force_k63, nxb, nyb, int(g["ew_boundary_type"] == cyc), int(g["ns_boundary_type"] == cyc),
                             C.byref(k), int(c.params["ndte"]), pi(mT), pi(mU), pd(sig), pd(u), pd(v), pd(geo), pd(strength), pd(inp),
                             pd(diag), pi(info))
    assert rc == 0, rc
    out = {nm: sig[q] for q, nm in enumerate(abi.STRESS)}
    out.update(uvel=u, vvel=v, strintxU=diag[0], strintyU=diag[1], taubxU=diag[2], taubyU=diag[3])
    return out, info


# (case, threads per CTA, pretended SMs, force the 6 T + 3 U instantiation)
CASES = {
    "tiny-5sub-6tiles": (dict(config="tiny", ndte=5, seed=231), 64, 6, 0),
    "tiny-4sub-revised-k63": (dict(config="tiny", ndte=4, seed=232, revised_evp=True), 64, 6, 1),
    "narrow-last-column-and-row": (dict(config="tiny", nx=23, ny=19, ndte=4, seed=233, kmt="continents"), 64, 6, 0),
    "doubly-cyclic-4tiles": (dict(config="tiny", nx=22, ny=20, ndte=5, seed=234, ns="cyclic", kmt="none"), 64, 6, 1),
    "one-tile-column-cyclic": (dict(config="tiny", nx=9, ny=30, ndte=4, seed=235, kmt="none"), 64, 4, 0),
    "128-threads-3x3": (dict(config="tiny", nx=36, ny=33, ndte=3, seed=236), 128, 9, 1),
    "general-branches": (dict(config="tiny", ndte=4, seed=131), 64, 8, 0),
}


@pytest.mark.parametrize("case", sorted(CASES))
def test_persistent_kernel_text_on_the_host_equals_the_oracle(oracle_mod, emu, case):
    kw, nthreads, sms, k63 = CASES[case]
    c = synth.make_case(**kw)
    if case == "general-branches":
        c.params.update(capping=0.0, Ktens=0.2, cosw=0.9, sinw=0.4358898943540674)
    ref = c.copy_fields()
    oracle_mod.evp_run_bgrid(c.grid, c.params, ref)
    got, info = run_emulated(emu, c, nthreads, sms, k63)
    assert info[0] * info[1] <= sms and info[0] * info[1] > 1, info
    for nm in abi.STRESS + ("uvel", "vvel", "strintxU", "strintyU", "taubxU", "taubyU"):
        assert np.array_equal(got[nm].view(np.int64), ref[nm][0].view(np.int64)), (nm, int((got[nm] != ref[nm][0]).sum()), info.tolist())
