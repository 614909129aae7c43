"""The B-grid CUDA kernels run thread by thread ON THE HOST.

tests/emu_bgrid.cpp includes the kernel translation unit cice_b200/csrc/evp_kernels.cu unchanged (EVP_HOST_EMU: launchers compiled
out, the inline-PTX helpers of evp_ptx.cuh replaced by their plain C++ meaning) on top of tests/cuda_emu.h, which makes every CUDA
thread of a CTA a host thread (pthread barriers for __syncthreads, counters for named barriers, mailboxes for warp shuffles, GCC
atomics).  Several subcycles of the ping-pong on one block must equal the oracle bit for bit -- stresses, velocities including the
on-rank cyclic ghost copies, the last subcycle's diagnostics -- for the split kernels, every selectable form of the fused kernel
(the default one included) and the in-kernel-halo instantiation without peers.  This is the CPU-side check of the kernels' index logic, ownership rules and hand-overs; the hardware's own
division / square-root seeds, timing and inter-CTA memory ordering are what the -m gpu tests add."""
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
# (kind, sub) of emu_bgrid_run
KERNELS = {
    "split": (0, 0),
    "fused-spec-loads": (1, 1), "fused-cp-async": (1, 2), "fused-stream": (1, 3),
    "fused-resident-default": (1, 4), "fused-spec-loads-IL": (1, 5), "fused-stream-IL": (1, 7),
    "p2p-no-peers-edge-first": (5, 0), "p2p-no-peers-no-counter": (5, 1),
}


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("emu") / "libemu_bgrid.so")
    cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
    if not os.path.exists(os.path.join(cuda_inc, "cuda_runtime.h")):
        pytest.skip("CUDA headers not found")
    cmd = ["/usr/bin/g++", "-O2", "-ffp-contract=off", "-std=c++17", "-fPIC", "-shared", "-pthread", "-I", cuda_inc,
           "-I", os.path.join(ROOT, "cice_b200", "csrc"), "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "emu_bgrid.cpp"), "-o", out]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return C.CDLL(out)


def run_emulated(emu, c, kind, sub):
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
    rc = emu.emu_bgrid_run(kind, sub, nxb, nyb, int(g["ew_boundary_type"] == cyc), int(g["ns_boundary_type"] == cyc), C.byref(k),
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


@pytest.mark.parametrize("kernel", list(KERNELS))
@pytest.mark.parametrize("case", sorted(CASES))
def test_kernel_text_on_the_host_equals_the_oracle(oracle_mod, emu, case, kernel):
    c = synth.make_case(**CASES[case])
    if case == "tiny-4sub":
        c.params.update(capping=0.0, Ktens=0.2, cosw=0.9, sinw=0.4358898943540674)   # general branches too
    ref = c.copy_fields()
    oracle_mod.evp_run_bgrid(c.grid, c.params, ref)
    got = run_emulated(emu, c, *KERNELS[kernel])
    for nm in abi.STRESS + ("uvel", "vvel", "strintxU", "strintyU", "taubxU", "taubyU"):
        assert np.array_equal(got[nm].view(np.int64), ref[nm][0].view(np.int64)), (nm, int((got[nm] != ref[nm][0]).sum()))


def test_fast_path_fallbacks_on_the_host(oracle_mod, emu):
    """zero and denormal-range operands: the hand-scheduled division / square root must report out-of-range and the built-in
    operators take over (the wiring of the fallbacks; the hardware seeds themselves are covered by the -m gpu twin of this test)."""
    c = synth.make_case("tiny", seed=21, ndte=3)
    for n in ("uvel", "vvel", "uocnU", "vocnU", "forcexU", "forceyU", "waterxU", "wateryU"):
        c.fields[n][...] = 0.0
    c.fields["strength"][...] *= 1e-300
    ref = c.copy_fields()
    oracle_mod.evp_run_bgrid(c.grid, c.params, ref)
    for kernel in ("fused-resident-default", "fused-stream-IL"):
        got = run_emulated(emu, c, *KERNELS[kernel])
        for nm in abi.STRESS + ("uvel", "vvel", "strintxU", "strintyU", "taubxU", "taubyU"):
            assert np.array_equal(got[nm].view(np.int64), ref[nm][0].view(np.int64)), (kernel, nm)


def test_post_loop_kernels_on_the_host(oracle_mod, emu):
    """deform_kernel and finish_kernel (`deformations`, `dyn_finish`) against the oracle, which is itself pinned to
    the reference source for both (tests/test_oracle.py)."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import ref_translit as rt
    pd = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    pi = lambda a: a.ctypes.data_as(C.POINTER(C.c_int32))
    emu.emu_deform_run.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_int32)] + [C.POINTER(C.c_double)] * 7 + [C.c_double]
    emu.emu_finish_run.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_int32)] + [C.POINTER(C.c_double)] * 5 + [C.c_double] * 3
    # deformations
    c, d = rt.deform_inputs(synth)
    g = c.grid
    nxb, nyb = g["nx_block"], g["ny_block"]
    ref = oracle_mod.deformations(g, c.fields["iceTmask"], c.fields["uvel"], c.fields["vvel"], {k: v.copy() for k, v in d.items()},
                                  c.params["e_factor"])
    geo = np.ascontiguousarray(np.stack([np.asarray(g[nm][0]) for nm in abi.GRID_STATIC]))
    out = np.ascontiguousarray(np.stack([d[k][0] for k in rt.DFIELDS]))
    A = lambda a: np.ascontiguousarray(a[0], dtype=np.float64)
    mT = np.ascontiguousarray(c.fields["iceTmask"][0], dtype=np.int32)
    u, v, dxU, dyU, tar = A(c.fields["uvel"]), A(c.fields["vvel"]), A(d["dxU"]), A(d["dyU"]), A(d["tarear"])
    assert emu.emu_deform_run(nxb, nyb, pi(mT), pd(u), pd(v), pd(geo), pd(dxU), pd(dyU), pd(tar), pd(out), float(c.params["e_factor"])) == 0
    for q, k in enumerate(rt.DFIELDS):
        assert np.array_equal(out[q].view(np.int64), ref[k][0].view(np.int64)), k
    # dyn_finish
    c, f, d = rt.finish_inputs(synth, oracle_mod)
    g = c.grid
    nxb, nyb = g["nx_block"], g["ny_block"]
    ref = oracle_mod.dyn_finish(g, f, {k: v.copy() for k, v in d.items()}, c.params["rhow"], c.params["cosw"], c.params["sinw"])
    inp = np.ascontiguousarray(np.stack([f[nm][0] for nm in ("cdn_ocnU", "aiU", "uocnU", "vocnU", "waterxU", "wateryU", "forcexU", "forceyU",
                                                             "umassdti", "fmU", "TbU")]))
    mU = np.ascontiguousarray(f["iceUmask"][0], dtype=np.int32)
    u, v, sx, sy = A(f["uvel"]), A(f["vvel"]), A(d["strocnxU"]), A(d["strocnyU"])
    assert emu.emu_finish_run(nxb, nyb, pi(mU), pd(u), pd(v), pd(inp), pd(sx), pd(sy), float(c.params["rhow"]), float(c.params["cosw"]),
                              float(c.params["sinw"])) == 0
    assert np.array_equal(sx.view(np.int64), ref["strocnxU"][0].view(np.int64))
    assert np.array_equal(sy.view(np.int64), ref["strocnyU"][0].view(np.int64))


@pytest.mark.parametrize("case", ["tiny-4sub", "tiny-5sub-revised", "wide-3sub", "doubly-cyclic-3sub", "aligned-32x8-16x16"])
def test_derived_geometry_kernels_on_the_host(oracle_mod, emu, case):
    """derived-geometry forms: dxhy, dyhx, cxp, cyp, cxm, cym, DminTarea derived in the kernel from HTN, HTE, dxT, dyT with the reference's
    expressions (ice_dyn_shared.F90:384-441).  The device-side check must find no differing cell on the synthetic grids (ghost cells
    outside a closed edge excluded), must notice a perturbed metric array, and the kernels must equal the oracle bit for bit."""
    c = synth.make_case(**CASES[case])
    g = c.grid
    nxb, nyb = g["nx_block"], g["ny_block"]
    geo = np.ascontiguousarray(np.stack([np.asarray(g[nm][0]) for nm in abi.GRID_STATIC]))
    HTN = np.ascontiguousarray(synth.scatter(c.X["HTN"], c.blocks)[0])
    HTE = np.ascontiguousarray(synth.scatter(c.X["HTE"], c.blocks)[0])
    pd = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    emu.emu_set_metric.argtypes = [C.c_int, C.c_int] + [C.POINTER(C.c_double)] * 3 + [C.c_double, C.c_int, C.c_int]
    cyc = abi.BNDY_NAMES["cyclic"]
    skip_e, skip_n = int(g["ew_boundary_type"] != cyc), int(g["ns_boundary_type"] != cyc)
    dmin = 1e-11                                      # deltaminEVP of synth.global_state
    wrong = HTN.copy()
    wrong[nyb // 2, nxb // 2] *= 1.0 + 2.0 ** -40
    assert emu.emu_set_metric(nxb, nyb, pd(geo), pd(wrong), pd(HTE), dmin, skip_e, skip_n) > 0
    assert emu.emu_set_metric(nxb, nyb, pd(geo), pd(HTN), pd(HTE), dmin, skip_e, skip_n) == 0
    ref = c.copy_fields()
    oracle_mod.evp_run_bgrid(c.grid, c.params, ref)
    for sub in (35, 36):
        got = run_emulated(emu, c, 1, sub)
        for nm in abi.STRESS + ("uvel", "vvel", "strintxU", "strintyU", "taubxU", "taubyU"):
            assert np.array_equal(got[nm].view(np.int64), ref[nm][0].view(np.int64)), (sub, nm)


@pytest.mark.parametrize("order", [1, 2], ids=["backwards", "alternating"])
@pytest.mark.parametrize("kernel", ["fused-resident-default", "fused-stream", "p2p-no-peers-edge-first"])
def test_cta_order_does_not_matter(oracle_mod, emu, kernel, order):
    """the emulation runs the CTAs of a launch one after the other, so a hazard BETWEEN CTAs (one reading what another writes in the
    same launch) would show as a dependence on their order: run them backwards and interleaved from both ends."""
    c = synth.make_case(**CASES["wide-3sub"])
    ref = c.copy_fields()
    oracle_mod.evp_run_bgrid(c.grid, c.params, ref)
    emu.emu_set_cta_order(order)
    try:
        got = run_emulated(emu, c, *KERNELS[kernel])
    finally:
        emu.emu_set_cta_order(0)
    for nm in abi.STRESS + ("uvel", "vvel", "strintxU", "strintyU", "taubxU", "taubyU"):
        assert np.array_equal(got[nm].view(np.int64), ref[nm][0].view(np.int64)), nm


@pytest.mark.parametrize("nx,ny", [(24, 20), (62, 21)])
def test_local_halo_kernel_on_the_host_equals_the_oracle_halo(oracle_mod, emu, evp_lib, nx, ny):
    """p2p_fold_kernel: the (uvel,vvel) halo update as one kernel when every source is on the rank (tripole fold on one
    GPU).  The kernel text runs on the host-exported plan (evp_b200_halo_plan, the enumeration the GPU exchange is built from) and
    must reproduce the oracle's halo update -- ghost row, symmetrised top row with its signed zeros, pole points."""
    from cice_b200 import dyn_evp
    c = synth.make_case("tiny", nx=nx, ny=ny, seed=141, ew="cyclic", ns="tripole", kmt="none")
    g = c.grid
    tu, tv = c.fields["uvel"].copy(), c.fields["vvel"].copy()
    for a in (tu, tv):      # ghost cells must come from the update, not from the input
        a[0, 0, :] = a[0, -1, :] = np.nan
        a[0, :, 0] = a[0, :, -1] = np.nan
    oracle_mod.halo_update(g, [tu, tv], field_loc=1, field_type=1)
    ld = dyn_evp.dom_pitch(nx)
    plan = np.asarray(dyn_evp.halo_plan([[1, 1, nx, ny]], 0, nx, ny, g["ew_boundary_type"], g["ns_boundary_type"]), dtype=np.int32)
    assert len(plan) > 0 and (plan[:, 1] == 0).all()
    dom = {}
    for name in ("uvel", "vvel"):
        a = np.full((ny + 2, ld), np.nan)
        a[1:ny + 1, 1:nx + 1] = c.fields[name][0][1:-1, 1:-1]
        a[1:ny + 1, 0], a[1:ny + 1, nx + 1] = a[1:ny + 1, nx].copy(), a[1:ny + 1, 1].copy()    # the compute kernels' E-W wrap stores
        dom[name] = np.ascontiguousarray(a.reshape(-1))
    dst, c1 = np.ascontiguousarray(plan[:, 0]), np.ascontiguousarray(plan[:, 2])
    c2 = np.ascontiguousarray(np.where(plan[:, 3] < 0, plan[:, 2], plan[:, 4]).astype(np.int32))
    code = np.ascontiguousarray(plan[:, 5].astype(np.int8))
    pd, pi = (lambda a: a.ctypes.data_as(C.POINTER(C.c_double))), (lambda a: a.ctypes.data_as(C.POINTER(C.c_int32)))
    assert emu.emu_halo_local(pd(dom["uvel"]), pd(dom["vvel"]), pi(dst), pi(c1), pi(c2), code.ctypes.data_as(C.POINTER(C.c_int8)), len(plan)) == 0
    for name, t in (("uvel", tu), ("vvel", tv)):
        got = dom[name].reshape(ny + 2, ld)[:, :nx + 2]
        want = t[0]
        ok = ~np.isnan(want)
        assert ok[1:-1, :].all() and ok[-1, 1:-1].all()           # interior rows incl. E-W ghosts, and the tripole ghost row
        assert np.array_equal(got[ok].view(np.int64), want[ok].view(np.int64)), name
