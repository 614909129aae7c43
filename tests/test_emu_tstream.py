"""The tile-streaming kernel (cice_b200/csrc/evp_tstream.cu, KERNEL_TSTREAM) run thread by thread ON THE HOST.

tests/emu_tstream.cpp includes the kernel translation unit unchanged with EVP_HOST_EMU defined: TMA box loads become synchronous
copies with the hardware's zero fill outside the tensor, mbarriers counters of completed phases (evp_tma.cuh), the CTAs of the
persistent grid concurrent groups of host threads.  Several subcycles on one block must equal the oracle bit for bit for both
instantiations (12 rows x 1 CTA per SM, 6 rows x 2), for the planner's own cut and for cuts that put segment joints, short last
segments, one-block items and more CTAs than items into play.  The hardware's TMA unit, the real mbarrier phases and the timing are
what the -m gpu twin (tests/test_gpu_parity.py::test_tstream_bitwise) adds."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from cice_b200 import abi, synth
from tests.test_emu_bgrid import KParams

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FIELDS = abi.STRESS + ("uvel", "vvel", "strintxU", "strintyU", "taubxU", "taubyU")


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("emu") / "libemu_tstream.so")
    cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
    if not os.path.exists(os.path.join(cuda_inc, "cuda_runtime.h")):
        pytest.skip("CUDA headers not found")
    cmd = ["/usr/bin/g++", "-O2", "-ffp-contract=off", "-std=c++17", "-fPIC", "-shared", "-pthread", "-I", cuda_inc,
           "-I", os.path.join(ROOT, "cice_b200", "csrc"), "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "tests"),
           os.path.join(ROOT, "tests", "emu_tstream.cpp"), "-o", out]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return C.CDLL(out)


def run_emulated(emu, c, rows, nb=0, nctas=0):
    g, f = c.grid, c.copy_fields()
    assert g["nblocks"] == 1
    nxb, nyb = g["nx_block"], g["ny_block"]
    k = KParams(**{nm: float(c.params.get(nm, 0.0)) for nm, _ in KParams._fields_[:-1]})
    sig = np.ascontiguousarray(np.stack([f[nm][0] for nm in abi.STRESS]))
    geo = np.ascontiguousarray(np.stack([np.asarray(g[nm][0]) for nm in abi.GRID_STATIC]))
    inp = np.ascontiguousarray(np.stack([f[nm][0] for nm in ("cdn_ocnU", "aiU", "uocnU", "vocnU", "waterxU", "wateryU", "forcexU", "forceyU",
                                                             "umassdti", "fmU", "TbU")]))
    HTN = np.ascontiguousarray(synth.scatter(c.X["HTN"], c.blocks)[0])
    HTE = np.ascontiguousarray(synth.scatter(c.X["HTE"], c.blocks)[0])
    diag = np.zeros((4, nyb, nxb))
    u, v = f["uvel"][0].copy(), f["vvel"][0].copy()
    strength = np.ascontiguousarray(f["strength"][0])
    pd = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    pi = lambda a: a.ctypes.data_as(C.POINTER(C.c_int32))
    mT, mU = np.ascontiguousarray(f["iceTmask"][0]), np.ascontiguousarray(f["iceUmask"][0])
    cyc = abi.BNDY_NAMES["cyclic"]
    plan = np.zeros(5, dtype=np.int32)
    emu.emu_tstream_run.argtypes = [C.c_int] * 8 + [C.POINTER(KParams), C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_int32)] + \
        [C.POINTER(C.c_double)] * 6 + [C.c_double] + [C.POINTER(C.c_double)] * 3 + [C.POINTER(C.c_int32)]
    rc = emu.emu_tstream_run(rows, nb, nctas, nxb, nyb, int(g["ew_boundary_type"] == cyc), int(g["ns_boundary_type"] == cyc), 0, C.byref(k),
                             int(c.params["ndte"]), pi(mT), pi(mU), pd(sig), pd(u), pd(v), pd(geo), pd(HTN), pd(HTE), 1e-11, pd(strength),
                             pd(inp), pd(diag), pi(plan))
    assert rc == 0, rc
    out = {nm: sig[q] for q, nm in enumerate(abi.STRESS)}
    out.update(uvel=u, vvel=v, strintxU=diag[0], strintyU=diag[1], taubxU=diag[2], taubyU=diag[3])
    return out, plan


CASES = {
    "tiny-4sub": dict(config="tiny", ndte=4, seed=131),
    "tiny-5sub-revised": dict(config="tiny", ndte=5, seed=132, revised_evp=True),
    "wide-3sub": dict(config="tiny", nx=70, ny=17, ndte=3, seed=133, kmt="continents"),
    "doubly-cyclic-3sub": dict(config="tiny", nx=33, ny=22, ndte=3, seed=134, ns="cyclic"),
    "S1-2sub": dict(config="tiny", ndte=2),
    "tall-3sub": dict(config="tiny", nx=40, ny=75, ndte=3, seed=137, kmt="continents"),
    # sub-domain edges on strip / segment joints: 60 = 2*30 U columns; 23 = 2*12 - 1 and 47 = 2*(2*12 - 1) + 1 U rows
    "aligned-60x23": dict(config="tiny", nx=60, ny=23, ndte=2, seed=135),
    "aligned-30x47-cyclic": dict(config="tiny", nx=30, ny=47, ndte=2, seed=136, ns="cyclic"),
    "wide-5-strips": dict(config="tiny", nx=131, ny=22, ndte=2, seed=138),
}
# (rows, blocks per segment or 0 = planner's, CTAs or 0 = planner's)
CUTS = {"12rows-planned": (12, 0, 0), "12rows-nb2": (12, 2, 3), "12rows-nb2-1cta": (12, 2, 1), "6rows-planned": (6, 0, 0), "6rows-nb2": (6, 2, 5),
        "6rows-nb3-many-ctas": (6, 3, 64)}


def reference(oracle_mod, c):
    ref = c.copy_fields()
    oracle_mod.evp_run_bgrid(c.grid, c.params, ref)
    return ref


@pytest.mark.parametrize("cut", list(CUTS))
@pytest.mark.parametrize("case", sorted(CASES))
def test_tstream_text_on_the_host_equals_the_oracle(oracle_mod, emu, case, cut):
    c = synth.make_case(**CASES[case])
    if case == "tiny-4sub":
        c.params.update(capping=0.0, Ktens=0.2, cosw=0.9, sinw=0.4358898943540674)   # general branches too
    ref = reference(oracle_mod, c)
    got, plan = run_emulated(emu, c, *CUTS[cut])
    for nm in FIELDS:
        bad = np.argwhere(got[nm].view(np.int64) != ref[nm][0].view(np.int64))
        assert len(bad) == 0, (nm, len(bad), bad[:6].tolist(), plan.tolist())


def test_tstream_fallbacks_on_the_host(oracle_mod, emu):
    """zero and denormal-range operands send the interleaved division / square root to the built-in operators"""
    c = synth.make_case("tiny", seed=21, ndte=3)
    for n in ("uvel", "vvel", "uocnU", "vocnU", "forcexU", "forceyU", "waterxU", "wateryU"):
        c.fields[n][...] = 0.0
    c.fields["strength"][...] *= 1e-300
    ref = reference(oracle_mod, c)
    got, _ = run_emulated(emu, c, 12)
    for nm in FIELDS:
        assert np.array_equal(got[nm].view(np.int64), ref[nm][0].view(np.int64)), nm


def test_planner_covers_the_subdomain_evenly(emu):
    """every U row / column belongs to exactly one (strip, segment); the CTAs' block counts differ by little at the sizes of configs[4]"""
    out = np.zeros(5, dtype=np.int32)
    for nx, ny, sms, rows in ((3600, 2400, 148, 12), (3600, 2400, 148, 6), (900, 1200, 148, 12), (320, 384, 148, 12), (100, 116, 148, 6), (5, 3, 148, 12)):
        emu.emu_tstream_cut(nx, ny, sms, rows, out.ctypes.data_as(C.POINTER(C.c_int32)))
        nstrips, nseg, nb, nitems, ctas = out.tolist()
        H = rows * nb
        assert nstrips * 30 >= nx > (nstrips - 1) * 30
        assert nseg * (H - 1) >= ny > (nseg - 1) * (H - 1)
        assert nitems == nstrips * nseg and 1 <= ctas <= min(nitems, sms * (1 if rows == 12 else 2))
        blocks = np.zeros(ctas, dtype=np.int64)
        for it in range(nitems):
            j0 = 1 + (it // nstrips) * (H - 1)
            blocks[it % ctas] += min(nb, (ny + 2 - j0 + rows - 1) // rows)
        if nx >= 900:
            assert blocks.max() <= 1.06 * blocks.mean(), (nx, ny, rows, nb, blocks.max(), blocks.mean())


def test_random_cuts_on_the_host(oracle_mod, emu):
    """random sub-domain sizes, blocks per segment and grid sizes (seeded): whatever the cut, the same bits"""
    rng = np.random.default_rng(20261017)
    for _ in range(10):
        nx, ny = int(rng.integers(20, 95)), int(rng.integers(20, 60))
        rows = int(rng.choice([12, 6]))
        nb = int(rng.integers(0, 5))                   # 0: the planner's own choice; 1: every block is a segment of its own
        nctas = int(rng.integers(0, 9))
        c = synth.make_case("tiny", nx=nx, ny=ny, ndte=2, seed=int(rng.integers(1, 1000)), ns=str(rng.choice(["closed", "cyclic"])),
                            kmt=str(rng.choice(["none", "continents"])))
        ref = reference(oracle_mod, c)
        got, plan = run_emulated(emu, c, rows, nb, nctas)
        for nm in FIELDS:
            assert np.array_equal(got[nm].view(np.int64), ref[nm][0].view(np.int64)), (nm, nx, ny, rows, nb, nctas, plan.tolist())
