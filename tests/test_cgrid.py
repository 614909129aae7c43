"""C grid (configs[2]: gx1 C-grid EVP, evp_algorithm=standard_2d, ndte=600, 1 GPU): oracle properties on the
CPU, bit-exact parity of the CUDA path through the C ABI on the GPU."""
import os

import numpy as np
import pytest

from cice_b200 import abi, synth

CF = abi.CFIELDS_INOUT + tuple(n for n in abi.CFIELDS_OUT if n != "strengthU")


def run_oracle_c(oracle, c, nthreads=0):
    f = c.copy_fields()
    oracle.evp_run_cgrid(c.grid, c.cgrid, c.params, f, nthreads=nthreads)
    return f


def run_gpu_c(evp, c, **over):
    f = c.copy_fields()
    evp.dyn_evp_b200_init(c.grid)
    try:
        evp.dyn_evp_b200_init_cgrid(c.cgrid)
        evp.dyn_evp_b200_run_cgrid(dict(c.params, **over), f)
    finally:
        evp.dyn_evp_b200_finalize()
    return f


@pytest.mark.parametrize("bs", [(12, 10), (7, 9), (24, 5)])
@pytest.mark.parametrize("visc", [abi.VISC_AVG_ZETA, abi.VISC_AVG_STRENGTH], ids=["avg_zeta", "avg_strength"])
def test_cgrid_oracle_decomposition_invariance(oracle_mod, bs, visc):
    """the reference's decomp property (gridsys_suite runs every box case on the C grid too)"""
    ref = synth.make_ccase("tiny", seed=5, visc_method=visc)
    G = {n: synth.gather(v, ref.blocks) for n, v in run_oracle_c(oracle_mod, ref, 1).items() if n in abi.CFIELDS_INOUT}
    c = synth.make_ccase("tiny", seed=5, visc_method=visc, block_size=bs)
    f = run_oracle_c(oracle_mod, c)
    for n in abi.CFIELDS_INOUT:
        assert np.array_equal(synth.gather(f[n], c.blocks), G[n]), n


def test_cgrid_oracle_zero_forcing_stays_at_rest(oracle_mod):
    c = synth.make_ccase("tiny")
    for n in ("uvelE", "vvelE", "uvelN", "vvelN", "uvel", "vvel", "uocnE", "vocnE", "uocnN", "vocnN", "waterxE", "wateryN", "forcexE", "forceyN"):
        c.fields[n][:] = 0.0
    f = run_oracle_c(oracle_mod, c, 1)
    for n in ("uvelE", "vvelN", "uvel", "vvel", "stressmT", "stress12T", "stress12U"):
        assert np.abs(f[n]).max() == 0.0, n


def test_cgrid_oracle_interpolants_are_consistent(oracle_mod):
    """after the loop uvel/vvel are exactly the masked area averages of uvelE/vvelN (ice_dyn_evp.F90:1081-1090)"""
    c = synth.make_ccase("gx3", ndte=15)
    f = run_oracle_c(oracle_mod, c, 1)
    ea, uvm = c.cgrid["earea"][0], c.cgrid["uvm"][0]
    uE = f["uvelE"][0]
    num = uE[1:-1, 1:-1] * ea[1:-1, 1:-1] + uE[2:, 1:-1] * ea[2:, 1:-1]
    want = num / (ea[1:-1, 1:-1] + ea[2:, 1:-1]) * uvm[1:-1, 1:-1]
    assert np.array_equal(f["uvel"][0][1:-1, 1:-1], want)


@pytest.mark.gpu
@pytest.mark.parametrize("cfg,kw", [
    ("tiny", dict()), ("tiny", dict(seed=3)), ("tiny", dict(seed=4, revised_evp=True)),
    ("tiny", dict(seed=5, visc_method=abi.VISC_AVG_STRENGTH)),
    ("tiny", dict(seed=6, ew="closed", ns="closed")), ("tiny", dict(seed=7, ew="cyclic", ns="cyclic", kmt="none")),
    ("gx3", dict(ndte=40)), ("gx3", dict(seed=20260103, ndte=9)),
], ids=["tiny-s1", "tiny-s2", "tiny-revised", "tiny-avgstrength", "tiny-closed", "tiny-cyclic2", "gx3-s1", "gx3-s2"])
def test_cgrid_exact_bitwise(oracle_mod, evp_lib, cfg, kw):
    c = synth.make_ccase(cfg, **kw)
    ref = run_oracle_c(oracle_mod, c)
    got = run_gpu_c(evp_lib, c, mode=abi.MODE_EXACT)
    skip = "etax2U" if c.params["visc_method"] == abi.VISC_AVG_STRENGTH else "strengthU"
    for n in abi.CFIELDS_INOUT + abi.CFIELDS_OUT:
        if n == skip:
            continue
        assert np.array_equal(got[n].view(np.int64), ref[n].view(np.int64)), \
            f"{n}: {np.count_nonzero(got[n] != ref[n])} cells differ, max {np.nanmax(np.abs(got[n] - ref[n])):.3e}"


@pytest.mark.gpu
@pytest.mark.parametrize("bs", [(12, 10), (7, 9)])
def test_cgrid_exact_bitwise_multi_block(oracle_mod, evp_lib, bs):
    c = synth.make_ccase("tiny", seed=8, block_size=bs)
    ref = run_oracle_c(oracle_mod, c)
    got = run_gpu_c(evp_lib, c, mode=abi.MODE_EXACT)
    for n in abi.CFIELDS_INOUT + abi.CFIELDS_OUT:
        if n == "strengthU":
            continue
        assert np.array_equal(got[n].view(np.int64), ref[n].view(np.int64)), n


@pytest.mark.gpu
def test_cgrid_gx1_ndte600(oracle_mod, evp_lib):
    """configs[2]: gx1 C grid, ndte = 600, one GPU."""
    c = synth.make_ccase("gx1", ndte=600)
    ref = run_oracle_c(oracle_mod, c)
    got = run_gpu_c(evp_lib, c, mode=abi.MODE_EXACT)
    for n in abi.CFIELDS_INOUT:
        assert np.array_equal(got[n].view(np.int64), ref[n].view(np.int64)), n


@pytest.mark.gpu
def test_cgrid_refused_where_unsupported(evp_lib):
    c = synth.make_ccase("tiny")
    with pytest.raises(evp_lib.EvpB200Error, match="evp_b200_init first"):
        evp_lib.dyn_evp_b200_init_cgrid(c.cgrid)


@pytest.mark.gpu
@pytest.mark.parametrize("cfg,kw", [("tiny", dict(seed=12)), ("tiny", dict(seed=13, ew="closed", ns="closed")),
                                    ("tiny", dict(seed=14, ew="cyclic", ns="cyclic", kmt="none")),
                                    ("tiny", dict(seed=15, visc_method=abi.VISC_AVG_STRENGTH, block_size=(12, 10))),
                                    ("gx3", dict(ndte=31))],
                         ids=["tiny", "tiny-closed", "tiny-cyclic2", "tiny-avgstrength-4blocks", "gx3-odd"])
def test_cgrid_five_kernel_form(oracle_mod, evp_lib, cfg, kw):
    """params.kernel = SPLIT: the first correct form, five kernels per subcycle (the default is three: kA, kB, k5)."""
    c = synth.make_ccase(cfg, **kw)
    ref = run_oracle_c(oracle_mod, c)
    got = run_gpu_c(evp_lib, c, mode=abi.MODE_EXACT, kernel=abi.KERNEL_SPLIT)
    skip = "etax2U" if c.params["visc_method"] == abi.VISC_AVG_STRENGTH else "strengthU"
    for n in abi.CFIELDS_INOUT + abi.CFIELDS_OUT:
        if n != skip:
            assert np.array_equal(got[n].view(np.int64), ref[n].view(np.int64)), n


@pytest.mark.gpu
def test_cgrid_fused_form_more_cases(oracle_mod, evp_lib):
    """the default three-kernel form (kB with the interleaved square roots / divisions of momentum_il_at) on odd loops, doubly cyclic
    wrap, avg_strength on several blocks, gx3."""
    for cfg, kw in (("tiny", dict(seed=12, ndte=7)), ("tiny", dict(seed=14, ew="cyclic", ns="cyclic", kmt="none")),
                    ("tiny", dict(seed=15, visc_method=abi.VISC_AVG_STRENGTH, block_size=(12, 10))), ("gx3", dict(ndte=31))):
        c = synth.make_ccase(cfg, **kw)
        ref = run_oracle_c(oracle_mod, c)
        got = run_gpu_c(evp_lib, c, mode=abi.MODE_EXACT)
        skip = "etax2U" if c.params["visc_method"] == abi.VISC_AVG_STRENGTH else "strengthU"
        for n in abi.CFIELDS_INOUT + abi.CFIELDS_OUT:
            if n != skip:
                assert np.array_equal(got[n].view(np.int64), ref[n].view(np.int64)), (cfg, kw, n)


@pytest.mark.gpu
def test_cgrid_fused_odd_loop_and_repeat(oracle_mod, evp_lib):
    """the fused form ping-pongs stress12U: an odd ndte ends on the second copy, and a second call must start from it."""
    c = synth.make_ccase("tiny", seed=16, ndte=7)
    ref = c.copy_fields()
    got = c.copy_fields()
    evp_lib.dyn_evp_b200_init(c.grid)
    evp_lib.dyn_evp_b200_init_cgrid(c.cgrid)
    try:
        for _ in range(2):
            oracle_mod.evp_run_cgrid(c.grid, c.cgrid, c.params, ref)
            evp_lib.dyn_evp_b200_run_cgrid(dict(c.params, mode=abi.MODE_EXACT), got)
            for n in abi.CFIELDS_INOUT:
                assert np.array_equal(got[n].view(np.int64), ref[n].view(np.int64)), n
    finally:
        evp_lib.dyn_evp_b200_finalize()


# ---- C grid against vectors generated from the reference's own source text (tests/golden/ref_translit.py) ----
def check_cgrid_against_ref_source_vectors(run):
    import hashlib
    import json
    import os
    import sys
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    sys.path.insert(0, here)
    import ref_translit as rt
    meta = json.load(open(os.path.join(here, "ref_source_vectors.json")))
    full = np.load(os.path.join(here, "ref_source_vectors.npz"))
    assert meta["ccases"] == [dict(kw) for kw in rt.CCASES], "tests/golden/ref_source_vectors.json is stale: regenerate"
    for n, kw in enumerate(rt.CCASES):
        c = synth.make_ccase(**kw)
        f = run(c)
        for k in rt.CFIELDS:
            key = f"ccase{n}_{k}"
            if key in full.files:
                assert np.array_equal(f[k].view(np.int64), full[key].view(np.int64)), \
                    f"{key}: {np.count_nonzero(f[k] != full[key])} cells differ from the reference-source vector"
            h = hashlib.sha256(np.ascontiguousarray(f[k], dtype=np.float64).tobytes()).hexdigest()
            assert h == meta["sha256"][key], f"{key}: differs from the reference-source vector (sha256)"


def test_cgrid_oracle_matches_vectors_from_reference_source(oracle_mod):
    """the C-grid oracle against the output of the reference's own Fortran text, transliterated statement by statement and
    executed (strain_rates_U, strain_rates_Tdt, stressC_T, stressC_U, div_stress_Ex/Ny, stepu_C, stepv_C,
    grid_average_X2Y_1/X2YS/X2YA, visc_replpress): bit for bit, both visc_method settings, classic and revised EVP."""
    check_cgrid_against_ref_source_vectors(lambda c: run_oracle_c(oracle_mod, c))


@pytest.mark.gpu
@pytest.mark.parametrize("kernel", [abi.KERNEL_AUTO, abi.KERNEL_SPLIT], ids=["three-kernel", "five-kernel"])
def test_cgrid_gpu_matches_vectors_from_reference_source(evp_lib, kernel):
    check_cgrid_against_ref_source_vectors(lambda c: run_gpu_c(evp_lib, c, mode=abi.MODE_EXACT, kernel=kernel))


# ---- CD grid ----
def run_oracle_cd(oracle, c, nthreads=0):
    f = c.copy_fields()
    oracle.evp_run_cdgrid(c.grid, c.cgrid, c.params, f, nthreads=nthreads)
    return f


def check_cdgrid_against_ref_source_vectors(run):
    import hashlib
    import json
    import os
    import sys
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    sys.path.insert(0, here)
    import ref_translit as rt
    meta = json.load(open(os.path.join(here, "ref_source_vectors.json")))
    full = np.load(os.path.join(here, "ref_source_vectors.npz"))
    assert meta["cdcases"] == [dict(kw) for kw in rt.CDCASES], "tests/golden/ref_source_vectors.json is stale: regenerate"
    for n, kw in enumerate(rt.CDCASES):
        c = synth.make_cdcase(**kw)
        f = run(c)
        for k in rt.CDFIELDS:
            key = f"cdcase{n}_{k}"
            if key in full.files:
                assert np.array_equal(f[k].view(np.int64), full[key].view(np.int64)), \
                    f"{key}: {np.count_nonzero(f[k] != full[key])} cells differ from the reference-source vector"
            h = hashlib.sha256(np.ascontiguousarray(f[k], dtype=np.float64).tobytes()).hexdigest()
            assert h == meta["sha256"][key], f"{key}: differs from the reference-source vector (sha256)"


def test_cdgrid_oracle_matches_vectors_from_reference_source(oracle_mod):
    """grid_ice = 'CD' (ice_dyn_evp.F90:1123-1275): the oracle against the transliterated reference source (stressCD_T, stressCD_U,
    strain_rates_Tdtsd, strain_rates_U, div_stress_Ex/Ey/Nx/Ny, stepuv_CD, grid averages), bit for bit."""
    check_cdgrid_against_ref_source_vectors(lambda c: run_oracle_cd(oracle_mod, c))


@pytest.mark.parametrize("bs", [(12, 10), (7, 9)])
def test_cdgrid_oracle_decomposition_invariance(oracle_mod, bs):
    c1 = synth.make_cdcase("tiny", seed=96, ndte=5)
    cb = synth.make_cdcase("tiny", seed=96, ndte=5, block_size=bs)
    f1, fb = run_oracle_cd(oracle_mod, c1), run_oracle_cd(oracle_mod, cb)
    for n in abi.CDFIELDS_INOUT:
        assert np.array_equal(synth.gather(fb[n], cb.blocks), synth.gather(f1[n], c1.blocks)), n


def run_gpu_cd(evp, c, **over):
    f = c.copy_fields()
    evp.dyn_evp_b200_init(c.grid)
    try:
        evp.dyn_evp_b200_init_cgrid(c.cgrid)
        evp.dyn_evp_b200_run_cdgrid(dict(c.params, **over), f)
    finally:
        evp.dyn_evp_b200_finalize()
    return f


def _cd_compare(got, ref, params):
    skip = ("zetax2U", "etax2U") if params["visc_method"] == abi.VISC_AVG_STRENGTH else ("strengthU",)
    bad = []
    for n in abi.CDFIELDS_INOUT + abi.CDFIELDS_OUT:
        if n in skip:
            continue
        if not np.array_equal(got[n].view(np.int64), ref[n].view(np.int64)):
            bad.append(f"{n}: {np.count_nonzero(got[n] != ref[n])} cells differ, max {np.nanmax(np.abs(got[n] - ref[n])):.3e}")
    assert not bad, "CD grid not bit-identical:\n  " + "\n  ".join(bad)


@pytest.mark.gpu
@pytest.mark.parametrize("cfg,kw", [
    ("tiny", dict(seed=101)), ("tiny", dict(seed=102, revised_evp=True)), ("tiny", dict(seed=103, visc_method=abi.VISC_AVG_STRENGTH)),
    ("tiny", dict(seed=104, ew="closed", ns="closed")), ("tiny", dict(seed=105, ew="cyclic", ns="cyclic", kmt="none")),
    ("tiny", dict(seed=106, block_size=(12, 10))), ("tiny", dict(seed=107, block_size=(7, 9), visc_method=abi.VISC_AVG_STRENGTH)),
    ("tiny", dict(ndte=9)), ("gx3", dict(seed=108, ndte=12)),
], ids=["tiny", "tiny-revised", "tiny-avgstrength", "tiny-closed", "tiny-cyclic2", "tiny-4blocks", "tiny-12blocks-avgstrength", "tiny-S1",
        "gx3"])
def test_cdgrid_exact_bitwise(oracle_mod, evp_lib, cfg, kw):
    """grid_ice = 'CD' (SURVEY 8a row a13): the CUDA path (four kernels per subcycle) against the oracle, every inout and work
    array, every cell of every block."""
    c = synth.make_cdcase(cfg, **kw)
    ref = run_oracle_cd(oracle_mod, c)
    got = run_gpu_cd(evp_lib, c, mode=abi.MODE_EXACT)
    _cd_compare(got, ref, c.params)


@pytest.mark.gpu
def test_cdgrid_gpu_matches_vectors_from_reference_source(evp_lib):
    check_cdgrid_against_ref_source_vectors(lambda c: run_gpu_cd(evp_lib, c, mode=abi.MODE_EXACT))


@pytest.mark.gpu
def test_cdgrid_repeat_and_fast_mode(oracle_mod, evp_lib):
    """two consecutive calls carry the state; FMA mode stays within 1e-10 on a short loop."""
    c = synth.make_cdcase("tiny", seed=109, ndte=5)
    ref, got = c.copy_fields(), c.copy_fields()
    evp_lib.dyn_evp_b200_init(c.grid)
    evp_lib.dyn_evp_b200_init_cgrid(c.cgrid)
    try:
        for _ in range(2):
            oracle_mod.evp_run_cdgrid(c.grid, c.cgrid, c.params, ref)
            evp_lib.dyn_evp_b200_run_cdgrid(dict(c.params, mode=abi.MODE_EXACT), got)
            _cd_compare(got, ref, c.params)
        fast = c.copy_fields()
        evp_lib.dyn_evp_b200_run_cdgrid(dict(c.params, mode=abi.MODE_FAST), fast)
        one = c.copy_fields()
        oracle_mod.evp_run_cdgrid(c.grid, c.cgrid, c.params, one)
        for n in abi.CDFIELDS_INOUT:
            den = max(np.abs(one[n]).max(), 1e-300)
            assert np.abs(fast[n] - one[n]).max() / den <= 1e-10, n
    finally:
        evp_lib.dyn_evp_b200_finalize()
