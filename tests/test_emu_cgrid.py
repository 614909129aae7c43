"""The C-grid and CD-grid CUDA kernels run thread by thread ON THE HOST (tests/emu_cgrid.cpp: cice_b200/csrc/evp_cgrid.cu included
unchanged on top of tests/cuda_emu.h).  A few subcycles on one block must equal the oracle bit for bit on every inout / out field,
ghost cells included: the five-kernel form, the default fused three-kernel form (stress12U ping-pong, aliased ring points) in four tile
shapes, and the four CD-grid kernels.  CPU-side check of the kernels' index logic; see tests/test_emu_bgrid.py."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from cice_b200 import abi, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("emuc") / "libemu_cgrid.so")
    cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
    if not os.path.exists(os.path.join(cuda_inc, "cuda_runtime.h")):
        pytest.skip("CUDA headers not found")
    cmd = ["/usr/bin/g++", "-O2", "-ffp-contract=off", "-std=c++17", "-fPIC", "-shared", "-pthread", "-I", cuda_inc,
           "-I", os.path.join(ROOT, "cice_b200", "csrc"), "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "tests"),
           os.path.join(ROOT, "tests", "emu_cgrid.cpp"), "-o", out]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return C.CDLL(out)


def _npl(g):
    return g["nx_block"] * g["ny_block"] * g["max_blocks"]


def block_equal(got, ref, names, skip=()):
    """whole block, ghost cells included: the wrap stores and the zero-fills are part of what is checked"""
    bad = []
    for n in names:
        if n in skip:
            continue
        a, b = got[n][0], ref[n][0]
        if not np.array_equal(a.view(np.int64), b.view(np.int64)):
            bad.append(f"{n}: {np.count_nonzero(a != b)} cells differ ({np.count_nonzero(a[1:-1, 1:-1] != b[1:-1, 1:-1])} interior)")
    assert not bad, "\n".join(bad)


CCASES = {
    "tiny-s1": dict(config="tiny", ndte=3),
    "tiny-s2-odd": dict(config="tiny", seed=3, ndte=5),
    "tiny-revised": dict(config="tiny", seed=4, revised_evp=True, ndte=4),
    "tiny-avgstrength": dict(config="tiny", seed=5, visc_method=abi.VISC_AVG_STRENGTH, ndte=3),
    "tiny-closed": dict(config="tiny", seed=6, ew="closed", ns="closed", ndte=3),
    "tiny-cyclic2": dict(config="tiny", seed=7, ew="cyclic", ns="cyclic", kmt="none", ndte=3),
}
FORMS = {0: "five-kernels", 1: "fused-32x8", 2: "fused-32x16", 3: "fused-32x12", 4: "fused-32x4", 5: "fused-32x8-interleaved-momentum"}


@pytest.mark.parametrize("form", sorted(FORMS), ids=[FORMS[k] for k in sorted(FORMS)])
@pytest.mark.parametrize("case", sorted(CCASES))
def test_cgrid_kernel_text_on_the_host_equals_the_oracle(oracle_mod, emu, case, form):
    c = synth.make_ccase(**CCASES[case])
    ref = c.copy_fields()
    oracle_mod.evp_run_cgrid(c.grid, c.cgrid, c.params, ref)
    f = c.copy_fields()
    g, kg = abi.make_grid(c.grid)
    cg, kcg = abi.make_cgrid(c.cgrid, _npl(c.grid))
    p = abi.make_params(c.params)
    s, ks = abi.make_cfields(f, _npl(c.grid))
    assert emu.emu_cgrid_run(form, C.byref(g), C.byref(cg), C.byref(p), C.byref(s)) == 0
    skip = ("etax2U",) if c.params["visc_method"] == abi.VISC_AVG_STRENGTH else ("strengthU",)
    block_equal(f, ref, abi.CFIELDS_INOUT + abi.CFIELDS_OUT, skip)


CDCASES = {
    "tiny-s1": dict(config="tiny", ndte=3),
    "tiny-s2": dict(config="tiny", seed=3, ndte=4),
    "tiny-revised": dict(config="tiny", seed=4, revised_evp=True, ndte=3),
    "tiny-cyclic2": dict(config="tiny", seed=7, ew="cyclic", ns="cyclic", kmt="none", ndte=3),
}


@pytest.mark.parametrize("case", sorted(CDCASES))
def test_cdgrid_kernel_text_on_the_host_equals_the_oracle(oracle_mod, emu, case):
    c = synth.make_cdcase(**CDCASES[case])
    ref = c.copy_fields()
    oracle_mod.evp_run_cdgrid(c.grid, c.cgrid, c.params, ref)
    f = c.copy_fields()
    g, kg = abi.make_grid(c.grid)
    cg, kcg = abi.make_cgrid(c.cgrid, _npl(c.grid))
    p = abi.make_params(c.params)
    s, ks = abi.make_cdfields(f, _npl(c.grid))
    assert emu.emu_cdgrid_run(C.byref(g), C.byref(cg), C.byref(p), C.byref(s)) == 0
    skip = ("zetax2U", "etax2U") if c.params["visc_method"] == abi.VISC_AVG_STRENGTH else ("strengthU",)
    block_equal(f, ref, abi.CDFIELDS_INOUT + abi.CDFIELDS_OUT, skip)


@pytest.mark.parametrize("order", [1, 2], ids=["backwards", "alternating"])
@pytest.mark.parametrize("form", [0, 1], ids=["five-kernels", "fused-32x8"])
def test_cgrid_cta_order_does_not_matter(oracle_mod, emu, form, order):
    """as tests/test_emu_bgrid.py::test_cta_order_does_not_matter: the fused C-grid kernels recompute ring points of neighbouring
    tiles and ping-pong stress12U exactly so that no CTA depends on another one of the same launch."""
    c = synth.make_ccase(**CCASES["tiny-cyclic2"])
    ref = c.copy_fields()
    oracle_mod.evp_run_cgrid(c.grid, c.cgrid, c.params, ref)
    f = c.copy_fields()
    g, kg = abi.make_grid(c.grid)
    cg, kcg = abi.make_cgrid(c.cgrid, _npl(c.grid))
    p = abi.make_params(c.params)
    s, ks = abi.make_cfields(f, _npl(c.grid))
    emu.emu_set_cta_order(order)
    try:
        assert emu.emu_cgrid_run(form, C.byref(g), C.byref(cg), C.byref(p), C.byref(s)) == 0
    finally:
        emu.emu_set_cta_order(0)
    block_equal(f, ref, abi.CFIELDS_INOUT + abi.CFIELDS_OUT, ("strengthU",))


def test_cgrid_interleaved_momentum_fallbacks_on_the_host(oracle_mod, emu):
    """ice and ocean at rest: the square-root operand is zero and the numerators are tiny, so the fast paths of momentum_il_at must
    report out-of-range and hand over to the built-in operators (same bits as the plain form)."""
    c = synth.make_ccase("tiny", seed=9, ndte=3)
    for n in ("uvelE", "vvelE", "uvelN", "vvelN", "uvel", "vvel", "uocnE", "vocnE", "uocnN", "vocnN", "forcexE", "forceyN", "waterxE", "wateryN"):
        c.fields[n][...] = 0.0
    c.fields["strength"][...] *= 1e-300
    ref = c.copy_fields()
    oracle_mod.evp_run_cgrid(c.grid, c.cgrid, c.params, ref)
    f = c.copy_fields()
    g, kg = abi.make_grid(c.grid)
    cg, kcg = abi.make_cgrid(c.cgrid, _npl(c.grid))
    p = abi.make_params(c.params)
    s, ks = abi.make_cfields(f, _npl(c.grid))
    assert emu.emu_cgrid_run(5, C.byref(g), C.byref(cg), C.byref(p), C.byref(s)) == 0
    block_equal(f, ref, abi.CFIELDS_INOUT + abi.CFIELDS_OUT, ("strengthU",))
