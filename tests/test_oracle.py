"""CPU tests of the oracle (the checker must be trusted before it checks anything).

The reference stores no golden vectors for this path and cannot be compiled here.  The oracle is pinned
(1) to vectors generated from the reference's own Fortran text by tests/golden/ref_translit.py (statement-by-statement
transliteration, executed; bit for bit), and (2) by the properties the reference asserts: bit-for-bit results under any
block decomposition (decomp_suite / perf_suite BFB columns) and with eliminated land blocks, 2-D == 1-D formulation
(core1d.F90:196), the halochk closed-form halo values, plus an independent numpy restatement of one subcycle.
"""
import hashlib
import json
import os

import numpy as np
import pytest

from cice_b200 import abi, synth
from tests.util import run_oracle, assert_bitwise

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "oracle_checksums.json")


def gathered(case, f):
    return {n: synth.gather(f[n], case.blocks) for n in abi.FIELDS_INOUT}


@pytest.mark.parametrize("bs", [(12, 10), (8, 7), (24, 5), (5, 20), (7, 9)])
def test_decomposition_invariance_tiny(oracle_mod, bs):
    ref = synth.make_case("tiny", seed=11)
    G = gathered(ref, run_oracle(oracle_mod, ref, 1))
    c = synth.make_case("tiny", seed=11, block_size=bs)
    g2 = gathered(c, run_oracle(oracle_mod, c))
    for n in abi.FIELDS_INOUT:
        assert np.array_equal(G[n], g2[n]), n


@pytest.mark.parametrize("bs", [(25, 29), (32, 40)])
def test_decomposition_invariance_gx3(oracle_mod, bs):
    ref = synth.make_case("gx3", ndte=20)
    G = gathered(ref, run_oracle(oracle_mod, ref, 1))
    c = synth.make_case("gx3", ndte=20, block_size=bs)
    g2 = gathered(c, run_oracle(oracle_mod, c))
    for n in abi.FIELDS_INOUT:
        assert np.array_equal(G[n], g2[n]), n


def test_thread_count_invariance(oracle_mod):
    c = synth.make_case("gx3", ndte=10, block_size=(25, 29), seed=3)
    assert_bitwise(run_oracle(oracle_mod, c, 1), run_oracle(oracle_mod, c, 4))


@pytest.mark.parametrize("cfg,kw", [("tiny", dict(seed=2)), ("gx3", dict(ndte=25)), ("gx3", dict(ndte=8, seed=9, revised_evp=True))])
def test_2d_equals_1d(oracle_mod, cfg, kw):
    """standard_2d vs shared_mem_1d: same arithmetic, gather-indexed, geometry recomputed from HTE/HTN."""
    c = synth.make_case(cfg, **kw)
    f2 = run_oracle(oracle_mod, c, 1)
    f1 = c.copy_fields()
    oracle_mod.evp_run_bgrid_1d(c.grid, c.X["HTE"], c.X["HTN"], 1e-11, c.params, f1, 1)
    for n in abi.FIELDS_INOUT:
        a, b = synth.gather(f2[n], c.blocks), synth.gather(f1[n], c.blocks)
        assert np.array_equal(a, b), n


# ---- independent numpy restatement of ONE subcycle (array form, written from the equations) -------
def numpy_subcycle(c, f):
    g, p = c.grid, c.params
    A = lambda a: a[0]  # single block
    u, v = A(f["uvel"]).copy(), A(f["vvel"]).copy()
    sh = lambda a, di, dj: np.roll(np.roll(a, -dj, 0), -di, 1)  # value at (i+di, j+dj)
    ucc, uee, use_, une = u, sh(u, -1, 0), sh(u, 0, -1), sh(u, -1, -1)
    vcc, vee, vse, vne = v, sh(v, -1, 0), sh(v, 0, -1), sh(v, -1, -1)
    dxT, dyT, cxp, cyp, cxm, cym = (A(g[k]) for k in ("dxT", "dyT", "cxp", "cyp", "cxm", "cym"))
    dxhy, dyhx, dmin = A(g["dxhy"]), A(g["dyhx"]), A(g["DminTarea"])
    div = [cyp * ucc - dyT * uee + cxp * vcc - dxT * vse, cym * uee + dyT * ucc + cxp * vee - dxT * vne,
           cym * une + dyT * use_ + cxm * vne + dxT * vee, cyp * use_ - dyT * une + cxm * vse + dxT * vcc]
    ten = [-cym * ucc - dyT * uee + cxm * vcc + dxT * vse, -cyp * uee + dyT * ucc + cxm * vee + dxT * vne,
           -cyp * une + dyT * use_ + cxp * vne - dxT * vee, -cym * use_ - dyT * une + cxp * vse - dxT * vcc]
    shr = [-cym * vcc - dyT * vee - cxm * ucc - dxT * use_, -cyp * vee + dyT * vcc - cxm * uee - dxT * une,
           -cyp * vne + dyT * vse - cxp * une + dxT * uee, -cym * vse - dyT * vne - cxp * use_ + dxT * ucc]
    mT = A(f["iceTmask"]).astype(bool).copy()
    ny_b, nx_b = mT.shape
    mT[0, :] = False
    mT[:, 0] = False
    P, M, S = [], [], []
    relax = 1.0 - p["arlx1i"] * p["revp"]
    for k in range(4):
        Delta = np.sqrt(div[k] * div[k] + p["e_factor"] * (ten[k] * ten[k] + shr[k] * shr[k]))
        with np.errstate(divide="ignore", invalid="ignore"):
            tmp = p["capping"] * (A(f["strength"]) / np.maximum(Delta, dmin)) + (1 - p["capping"]) * (A(f["strength"]) / (Delta + dmin))
        z = (1 + p["Ktens"]) * tmp
        rp = (1 - p["Ktens"]) * tmp * Delta
        e = p["epp2i"] * z
        sp = (A(f[f"stressp_{k+1}"]) * relax + p["arlx1i"] * (z * div[k] - rp)) * p["denom1"]
        sm = (A(f[f"stressm_{k+1}"]) * relax + p["arlx1i"] * e * ten[k]) * p["denom1"]
        s12 = (A(f[f"stress12_{k+1}"]) * relax + p["arlx1i"] * 0.5 * e * shr[k]) * p["denom1"]
        P.append(np.where(mT, sp, A(f[f"stressp_{k+1}"])))
        M.append(np.where(mT, sm, A(f[f"stressm_{k+1}"])))
        S.append(np.where(mT, s12, A(f[f"stress12_{k+1}"])))
    p111 = 1.0 / 9.0; p055 = p111 * 0.5; p027 = p055 * 0.5; p166 = 1.0 / 6.0; p222 = 2.0 / 9.0; p333 = 1.0 / 3.0
    ssn = lambda X: X[0] + X[1]; sss = lambda X: X[2] + X[3]; sse = lambda X: X[0] + X[3]; ssw = lambda X: X[1] + X[2]
    s1 = lambda X, w: (X[0] + X[2]) * w; s2 = lambda X, w: (X[1] + X[3]) * w
    cs = lambda X, a, b: [a * X[0] + s2(X, b) + (b * 0.5 if a == p111 else p055) * X[2],
                          a * X[1] + s1(X, b) + (b * 0.5 if a == p111 else p055) * X[3],
                          a * X[2] + s2(X, b) + (b * 0.5 if a == p111 else p055) * X[0],
                          a * X[3] + s1(X, b) + (b * 0.5 if a == p111 else p055) * X[1]]
    cp, cm, c12 = cs(P, p111, p055), cs(M, p111, p055), cs(S, p222, p111)  # [ne, nw, sw, se]
    str12ew = 0.5 * dxT * (p333 * sse(S) + p166 * ssw(S)); str12we = 0.5 * dxT * (p333 * ssw(S) + p166 * sse(S))
    str12ns = 0.5 * dyT * (p333 * ssn(S) + p166 * sss(S)); str12sn = 0.5 * dyT * (p333 * sss(S) + p166 * ssn(S))
    st = [None] * 8
    a = 0.25 * dyT * (p333 * ssn(P) + p166 * sss(P)); b = 0.25 * dyT * (p333 * ssn(M) + p166 * sss(M))
    st[0] = -a - b - str12ew + dxhy * (-cp[0] + cm[0]) + dyhx * c12[0]
    st[1] = a + b - str12we + dxhy * (-cp[1] + cm[1]) + dyhx * c12[1]
    a = 0.25 * dyT * (p333 * sss(P) + p166 * ssn(P)); b = 0.25 * dyT * (p333 * sss(M) + p166 * ssn(M))
    st[2] = -a - b + str12ew + dxhy * (-cp[3] + cm[3]) + dyhx * c12[3]
    st[3] = a + b + str12we + dxhy * (-cp[2] + cm[2]) + dyhx * c12[2]
    a = 0.25 * dxT * (p333 * sse(P) + p166 * ssw(P)); b = 0.25 * dxT * (p333 * sse(M) + p166 * ssw(M))
    st[4] = -a + b - str12ns - dyhx * (cp[0] + cm[0]) + dxhy * c12[0]
    st[5] = a - b - str12sn - dyhx * (cp[3] + cm[3]) + dxhy * c12[3]
    a = 0.25 * dxT * (p333 * ssw(P) + p166 * sse(P)); b = 0.25 * dxT * (p333 * ssw(M) + p166 * sse(M))
    st[6] = -a + b + str12ns - dyhx * (cp[1] + cm[1]) + dxhy * c12[1]
    st[7] = a - b + str12sn - dyhx * (cp[2] + cm[2]) + dxhy * c12[2]
    st = [np.where(mT, x, 0.0) for x in st]
    # momentum
    mU = A(f["iceUmask"]).astype(bool)
    F = lambda k: A(f[k])
    vrel = F("aiU") * p["rhow"] * F("cdn_ocnU") * np.sqrt((F("uocnU") - u) * (F("uocnU") - u) + (F("vocnU") - v) * (F("vocnU") - v))
    taux, tauy = vrel * F("waterxU"), vrel * F("wateryU")
    Cb = F("TbU") / (np.sqrt(u * u + v * v) + p["u0"])
    cca = (p["brlx"] + p["revp"]) * F("umassdti") + vrel * p["cosw"] + Cb
    ccb = F("fmU") + np.copysign(1.0, F("fmU")) * vrel * p["sinw"]
    ab2 = cca * cca + ccb * ccb
    sx = A(g["uarear"]) * (st[0] + sh(st[1], 1, 0) + sh(st[2], 0, 1) + sh(st[3], 1, 1))
    sy = A(g["uarear"]) * (st[4] + sh(st[5], 0, 1) + sh(st[6], 1, 0) + sh(st[7], 1, 1))
    cc1 = sx + F("forcexU") + taux + F("umassdti") * (p["brlx"] * u + p["revp"] * u)
    cc2 = sy + F("forceyU") + tauy + F("umassdti") * (p["brlx"] * v + p["revp"] * v)
    with np.errstate(divide="ignore", invalid="ignore"):
        un = np.where(mU, (cca * cc1 + ccb * cc2) / ab2, u)
        vn = np.where(mU, (cca * cc2 - ccb * cc1) / ab2, v)
    out = {f"stressp_{k+1}": P[k] for k in range(4)}
    out.update({f"stressm_{k+1}": M[k] for k in range(4)})
    out.update({f"stress12_{k+1}": S[k] for k in range(4)})
    out.update(uvel=un, vvel=vn, strintxU=np.where(mU, sx, 0), strintyU=np.where(mU, sy, 0),
               taubxU=np.where(mU, -un * Cb, 0), taubyU=np.where(mU, -vn * Cb, 0))
    return out


@pytest.mark.parametrize("cfg,seed", [("tiny", 1), ("gx3", 20260101)])
def test_one_subcycle_against_numpy(oracle_mod, cfg, seed):
    """set S2 (random velocities / stresses / TbU): one subcycle of the C oracle equals an array-form
    numpy evaluation of the same equations bit for bit on every interior cell."""
    c = synth.make_case(cfg, seed=seed, ndte=1)
    fo = run_oracle(oracle_mod, c, 1)
    fn = numpy_subcycle(c, c.fields)
    b = c.blocks
    sl = (slice(b.jlo[0] - 1, b.jhi[0]), slice(b.ilo[0] - 1, b.ihi[0]))
    for n in abi.FIELDS_INOUT:
        assert np.array_equal(fo[n][0][sl], fn[n][sl]), n


def test_zero_forcing_stays_at_rest(oracle_mod):
    c = synth.make_case("tiny")
    for n in ("uvel", "vvel", "uocnU", "vocnU", "waterxU", "wateryU", "forcexU", "forceyU"):
        c.fields[n][:] = 0.0
    f = run_oracle(oracle_mod, c, 1)
    for n in ("uvel", "vvel") + abi.STRESS[4:]:
        assert np.abs(f[n]).max() == 0.0, n


def test_golden_checksums(oracle_mod):
    """regression pin of the oracle itself (generated by tests/golden/make_golden.py from this oracle;
    NOT reference output -- none can be produced in this image)."""
    with open(GOLDEN) as fh:
        gold = json.load(fh)
    for name, spec in gold.items():
        c = synth.make_case(**spec["case"])
        f = run_oracle(oracle_mod, c, 1)
        for n, want in spec["sums"].items():
            got = hashlib.sha256(np.ascontiguousarray(synth.gather(f[n], c.blocks)).tobytes()).hexdigest()
            assert got == want, (name, n)


@pytest.mark.parametrize("cfg,bs", [("gx3", (10, 10)), ("tiny", (4, 4))])
def test_land_block_elimination_is_bit_for_bit(oracle_mod, cfg, bs):
    """ice_domain.F90 drops blocks without ocean before distributing them; the answer on the remaining blocks must not
    change (the reference's decomp_suite compares such runs with cmp).  Ghost cells facing a dropped block are filled
    with zeros by the halo update (ice_boundary.F90:1398-1408)."""
    c = synth.make_case(cfg, block_size=bs, seed=52, ndte=8, kmt="continents")
    nb = c.blocks.nblocks_tot
    land = [n for n in range(nb) if not c.fields["iceTmask"][n][1:-1, 1:-1].any() and not c.fields["iceUmask"][n].any()]
    assert len(land) >= 5
    owner = np.zeros(nb, int)
    owner[land] = -1
    g, f, ids = c.rank_view(owner, 0)
    oracle_mod.evp_run_bgrid(g, c.params, f)
    full = c.copy_fields()
    oracle_mod.evp_run_bgrid(c.grid, c.params, full)
    for n in abi.FIELDS_INOUT:
        assert np.array_equal(f[n].view(np.int64), full[n][ids].view(np.int64)), n


# ---- the oracle against vectors generated from the reference's own source text (tests/golden/ref_translit.py) ----
def _ref_source_vectors():
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    meta = json.load(open(os.path.join(here, "ref_source_vectors.json")))
    full = np.load(os.path.join(here, "ref_source_vectors.npz"))
    return meta, full


def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a, dtype=np.float64).tobytes()).hexdigest()


def check_against_ref_source_vectors(run, cases=None):
    """run(case) -> fields dict after the loop; compared bit for bit with what the transliterated reference source gave."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import ref_translit as rt
    meta, full = _ref_source_vectors()
    assert meta["cases"] == [dict(kw) for kw in rt.CASES], "tests/golden/ref_source_vectors.json is stale: regenerate"
    for n, kw in enumerate(rt.CASES):
        if cases is not None and n not in cases:
            continue
        c = rt.make(synth, kw)
        f = run(c)
        for k in rt.FIELDS:
            got = f[k][0]
            key = f"case{n}_{k}"
            if key in full.files:
                assert np.array_equal(got.view(np.int64), full[key].view(np.int64)), \
                    f"{key}: {np.count_nonzero(got != full[key])} cells differ from the reference-source vector"
            assert _sha(got) == meta["sha256"][key], f"{key}: differs from the reference-source vector (sha256)"


def test_oracle_matches_vectors_from_reference_source(oracle_mod):
    """the reference cannot be built here and stores no vectors for this path, so the vectors were produced by
    executing a statement-by-statement transliteration of the reference's Fortran (stress, stepu, strain_rates, visc_replpress,
    ice_constants) -- see tests/golden/ref_translit.py.  The C oracle must reproduce them bit for bit: classic and revised EVP,
    both visc_replpress branches, turning angle, grounded-ice term, start from rest, gx3."""
    def run(c):
        f = c.copy_fields()
        oracle_mod.evp_run_bgrid(c.grid, c.params, f)
        return f
    check_against_ref_source_vectors(run)


@pytest.mark.skipif(not os.path.isdir("/root/reference/cicecore"), reason="the reference source tree is not on this machine")
def test_reference_source_vectors_regenerate():
    """where /root/reference exists: re-derive two of the cases from the Fortran text and compare with the committed vectors."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import ref_translit as rt
    meta, _ = _ref_source_vectors()
    vec = rt.generate(only=(1, 3))
    assert len(vec) == 2 * len(rt.FIELDS)
    for k, v in vec.items():
        assert _sha(v) == meta["sha256"][k], k


def test_deformations_oracle_matches_vectors_from_reference_source(oracle_mod):
    """`deformations` (ice_dyn_shared.F90:1756-1860): the oracle against the transliterated reference text."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import ref_translit as rt
    meta, full = _ref_source_vectors()
    c, d = rt.deform_inputs(synth)
    got = oracle_mod.deformations(c.grid, c.fields["iceTmask"], c.fields["uvel"], c.fields["vvel"], d, c.params["e_factor"])
    for k in rt.DFIELDS:
        assert np.array_equal(got[k].view(np.int64), full[f"dcase0_{k}"].view(np.int64)), k
        assert _sha(got[k]) == meta["sha256"][f"dcase0_{k}"], k


def test_dyn_finish_oracle_matches_vectors_from_reference_source(oracle_mod):
    """`dyn_finish` (ice_dyn_shared.F90:1291-1365): the oracle against the transliterated reference text, on the
    velocities a three-subcycle loop leaves, with a turning angle (the sinw * sign(fm) terms)."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import ref_translit as rt
    meta, full = _ref_source_vectors()
    c, f, d = rt.finish_inputs(synth, oracle_mod)
    got = oracle_mod.dyn_finish(c.grid, f, d, c.params["rhow"], c.params["cosw"], c.params["sinw"])
    for k in rt.FFIELDS:
        assert np.array_equal(got[k].view(np.int64), full[f"fcase0_{k}"].view(np.int64)), k
        assert _sha(got[k]) == meta["sha256"][f"fcase0_{k}"], k
        assert (got[k] != -7.0).sum() == (f["iceUmask"] != 0).sum()      # exactly the U list is written


@pytest.mark.parametrize("cfg", ["tiny", "gx3", "gx1"])
def test_synthetic_inputs_follow_reference_dyn_prep2(cfg):
    """the loop's time-varying inputs (umassdti, fmU, waterx/y, forcex/y, the U ice mask, new-ice velocities) that
    cice_b200/synth.py produces for tests and bench are, bit for bit, what the reference's own dyn_prep2
    (ice_dyn_shared.F90:593-839, transliterated and executed by tests/golden/ref_translit.py) makes of the same box2001 state."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import ref_translit as rt
    meta, _ = _ref_source_vectors()
    c = synth.make_case(cfg)
    g = c.grid
    inner = (slice(int(g["jlo"][0]) - 1, int(g["jhi"][0])), slice(int(g["ilo"][0]) - 1, int(g["ihi"][0])))
    for k in rt.PFIELDS:
        mine = np.asarray(c.X[k], dtype=np.float64)[inner]
        assert _sha(mine) == meta["sha256"][f"prep2_{cfg}_{k}"], k


@pytest.mark.parametrize("cfg", ["tiny", "gx3", "gx1"])
def test_synthetic_inputs_follow_reference_prep1_and_averages(cfg):
    """the stage before dyn_prep2: dyn_prep1 (tmass, iceTmask; ice_dyn_shared.F90:496-592) and evp()'s T->U averages
    (grid_average_X2YS 'NE' of aice, tmass, uocn, vocn; grid_average_X2YF 'NE' of the wind stress; ice_dyn_evp.F90:433-456) as
    cice_b200/synth.py computes them == the transliterated reference routines on the same box2001 state, bit for bit."""
    meta, _ = _ref_source_vectors()
    c = synth.make_case(cfg)
    g, X = c.grid, c.X
    inner = (slice(int(g["jlo"][0]) - 1, int(g["jhi"][0])), slice(int(g["ilo"][0]) - 1, int(g["ihi"][0])))
    mine = {"tmass": X["tmass"], "iceTmask_interior": X["iceTmask"].astype(np.float64)[inner], "aiU": X["aiU"][inner],
            "umass_i": X["umass_i"], "uocnU": X["uocnU"][inner], "vocnU": X["vocnU"][inner], "strairxU_i": X["strairxU_i"],
            "strairyU_i": X["strairyU_i"]}
    for k, v in mine.items():
        assert _sha(v) == meta["sha256"][f"prep1_{cfg}_{k}"], k


@pytest.mark.parametrize("cfg", ["tiny", "gx3", "gx1"])
def test_synthetic_geometry_follows_reference_init_dyn_shared(cfg):
    """dxhy, dyhx, cyp, cxp, cym, cxm as cice_b200/synth.py derives them from HTE/HTN == the tail of the reference's
    init_dyn_shared (ice_dyn_shared.F90:401-441), transliterated and executed on the same HTE/HTN, bit for bit."""
    meta, _ = _ref_source_vectors()
    c = synth.make_case(cfg)
    g, X = c.grid, c.X
    i0, i1, j0, j1 = int(g["ilo"][0]) - 1, int(g["ihi"][0]), int(g["jlo"][0]) - 1, int(g["jhi"][0])
    for k in ("dxhy", "dyhx"):
        assert _sha(X[k][j0:j1, i0:i1]) == meta["sha256"][f"geom_{cfg}_{k}"], k
    for k in ("cyp", "cxp", "cym", "cxm"):
        assert _sha(X[k][j0:j1 + 1, i0:i1 + 1]) == meta["sha256"][f"geom_{cfg}_{k}"], k


@pytest.mark.skipif(not os.path.isdir("/root/reference/cicecore"), reason="the reference source tree is not on this machine")
@pytest.mark.parametrize("ndte,revised", [(120, False), (240, False), (600, False), (240, True)])
def test_evp_constants_follow_reference_set_evp_parameters(ndte, revised):
    """arlx1i, denom1, revp, brlx, epp2i, e_factor as synth.evp_params computes them == set_evp_parameters
    (ice_dyn_shared.F90:453-486), transliterated and executed."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import ref_translit as rt
    ref = rt.generate_params(ndte, revised)
    mine = synth.evp_params(ndte, revised_evp=revised)
    for k, v in ref.items():
        assert mine[k] == v, k


def _halo_update_stress_literal(grid, a1, a2):
    """ice_HaloUpdate_stress(array1, array2, halo, field_loc_center, field_type_scalar) on one task, loop for loop
    (/root/reference/cicecore/cicedyn/infrastructure/comm/mpi/ice_boundary.F90): the tripole buffer is filled from array2 through the
    "srcBlock > 0, dstBlock < 0" local copies (:7641-7652; addresses built at :8117-8133: the top tripoleRows = nghost+1 physical rows,
    buffer column = global i), and copied out into array1 through the "srcBlock < 0" entries (:7760-7794; addresses :8136-8157), with
    ioffset = joffset = 0 and isign = +1 for a centre-located scalar on a u-fold (:7729-7732, :7702)."""
    nghost, nxg = 1, grid["nx_global"]
    rows = nghost + 1                                   # halo%tripoleRows
    buf = np.zeros((rows + 1, nxg + 1))                 # bufTripole(1:nxGlobal, 1:tripoleRows), 1-based; = fill (0)
    top = [b for b in range(grid["nblocks"]) if grid["j_glob"][b][grid["jhi"][b] - 1] == grid["ny_global"]]
    for b in top:                                       # copy into the buffer
        ib, ie, je = grid["ilo"][b], grid["ihi"][b], grid["jhi"][b]
        for j in range(1, rows + 1):
            for i in range(1, ie - ib + 2):
                buf[j, grid["i_glob"][b][ib + i - 1 - 1]] = a2[b, je - rows + j - 1, ib + i - 1 - 1]
    for b in top:                                       # copy out of the buffer
        ie, je = grid["ihi"][b], grid["jhi"][b]
        for j in range(1, rows + 1):
            for i in range(1, ie + nghost + 1):
                ig = grid["i_glob"][b][i - 1]
                if ig < 1:
                    continue                            # padded column
                isrc, jsrc = nxg - ig + 1, nghost + 3 - j
                jdst = -1 if j > nghost + 1 else je + j - 1
                if isrc < 1:
                    isrc += nxg
                if isrc > nxg:
                    isrc -= nxg
                if 0 < jsrc <= rows and jdst > 0:
                    a1[b, jdst - 1, i - 1] = buf[jsrc, isrc]


@pytest.mark.parametrize("bs", [None, (12, 10), (8, 7)], ids=["1block", "4blocks", "padded-blocks"])
def test_stress_symmetrise_equals_the_literal_halo_update_stress(oracle_mod, bs):
    """the symmetrisation of the stress tensor across the tripole fold after the loop (ice_dyn_evp.F90:1321-1388): the oracle's closed
    form against the reference's twelve ice_HaloUpdate_stress calls restated loop for loop; only north ghost rows of top blocks change."""
    c = synth.make_case("tiny", seed=51, ns="tripole", ew="cyclic", kmt="none", ndte=1, block_size=bs)
    rng = np.random.default_rng(5)
    f = {n: rng.normal(size=c.fields[n].shape) for n in abi.STRESS}
    want = {n: a.copy() for n, a in f.items()}
    for k in ("p", "m", "12"):
        for a, b in ((1, 3), (3, 1), (2, 4), (4, 2)):   # the order of ice_dyn_evp.F90:1347-1386
            _halo_update_stress_literal(c.grid, want[f"stress{k}_{a}"], want[f"stress{k}_{b}"])
    got = {n: a.copy() for n, a in f.items()}
    oracle_mod.stress_symmetrise(c.grid, got)
    changed = 0
    for n in abi.STRESS:
        assert np.array_equal(got[n].view(np.int64), want[n].view(np.int64)), n
        changed += int((got[n] != f[n]).sum())
    assert changed > 0
    # a grid without a fold is left alone
    c2 = synth.make_case("tiny", seed=52, ndte=1)
    g2 = {n: c2.fields[n].copy() for n in abi.STRESS}
    oracle_mod.stress_symmetrise(c2.grid, g2)
    assert all(np.array_equal(g2[n], c2.fields[n]) for n in abi.STRESS)


def _halo_update_literal(grid, a, field_loc, field_type):
    """ice_HaloUpdate2DR8 on one task for a centre- or NE-corner-located field on a u-fold tripole (or any other) grid, with the
    tripole part restated loop for loop from /root/reference/cicecore/cicedyn/infrastructure/comm/mpi/ice_boundary.F90:
      * ordinary ghost cells: the local copies built by ice_HaloMsgCreate (east/west/north/south/corner cases, :7975-8100, 8160-8400)
        give a ghost cell the interior value of the block that owns its global index (cyclic wrap is in i_glob/j_glob,
        ice_blocks.F90:222-276); cells outside a closed/open edge are not touched (no fill requested, :1173-1181);
      * tripole buffer: the top tripoleRows = nghost+1 physical rows, buffer column = global i (:1372-1395, addresses :8117-8133);
      * NE corner on a u-fold: ioffset = joffset = 1 and the top buffer row is symmetrised pairwise, i = 1 .. nxGlobal/2 - 1 with
        iDst = nxGlobal - i (:1689-1703); centre: no offsets, no averaging (:1685-1688);
      * copy-out over every column of every top block, rows j = 1 .. tripoleRows (:1724-1756, addresses :8136-8157):
        iSrc = nxGlobal - i_glob(i) + 1 - ioffset (wrapped), jSrc = nghost + 3 - j - joffset, jDst = jhi + j - 1,
        array(i, jDst) = isign * buf(iSrc, jSrc) when 0 < jSrc <= tripoleRows."""
    nghost, nxg, nyg = 1, grid["nx_global"], grid["ny_global"]
    nb = grid["nblocks"]
    ig_, jg_ = grid["i_glob"], grid["j_glob"]
    tripole = grid["ns_boundary_type"] == abi.BNDY_NAMES["tripole"]
    isign = -1.0 if field_type == 1 else 1.0
    # global interior values as the blocks hold them (sources are physical cells only: the order of the copies does not matter)
    G = np.zeros((nyg + 1, nxg + 1))
    for b in range(nb):
        for j in range(grid["jlo"][b], grid["jhi"][b] + 1):
            for i in range(grid["ilo"][b], grid["ihi"][b] + 1):
                G[jg_[b][j - 1], ig_[b][i - 1]] = a[b, j - 1, i - 1]
    for b in range(nb):
        ilo, ihi, jlo, jhi = grid["ilo"][b], grid["ihi"][b], grid["jlo"][b], grid["jhi"][b]
        for j in range(jlo - nghost, jhi + nghost + 1):
            for i in range(ilo - nghost, ihi + nghost + 1):
                if ilo <= i <= ihi and jlo <= j <= jhi:
                    continue
                gi, gj = ig_[b][i - 1], jg_[b][j - 1]
                if 1 <= gi <= nxg and 1 <= gj <= nyg:
                    a[b, j - 1, i - 1] = G[gj, gi]
    if not tripole:
        return
    rows = nghost + 1
    buf = np.zeros((rows + 1, nxg + 1))
    for j in range(1, rows + 1):
        buf[j, 1:] = G[nyg - rows + j, 1:]
    if field_loc == 1:                       # field_loc_NEcorner, u-fold
        ioffset = joffset = 1
        for i in range(1, nxg // 2):         # do i = 1, nxGlobal/2 - 1
            idst = nxg - i
            x1, x2 = buf[rows, i], buf[rows, idst]
            xavg = 0.5 * (x1 + isign * x2)
            buf[rows, i] = xavg
            buf[rows, idst] = isign * xavg
    else:                                    # field_loc_center
        ioffset = joffset = 0
    for b in range(nb):
        if jg_[b][grid["jhi"][b] - 1] != nyg:
            continue
        ie, je = grid["ihi"][b], grid["jhi"][b]
        for j in range(1, rows + 1):
            for i in range(1, ie + nghost + 1):
                gi = ig_[b][i - 1]
                if gi < 1:
                    continue
                isrc, jsrc = nxg - gi + 1 - ioffset, nghost + 3 - j - joffset
                jdst = je + j - 1
                if isrc < 1:
                    isrc += nxg
                if isrc > nxg:
                    isrc -= nxg
                if 0 < jsrc <= rows and jdst > 0:
                    a[b, jdst - 1, i - 1] = isign * buf[jsrc, isrc]


@pytest.mark.parametrize("bs", [None, (12, 10), (8, 7)], ids=["1block", "4blocks", "padded-blocks"])
@pytest.mark.parametrize("ns", ["tripole", "closed", "cyclic"])
def test_halo_update_equals_the_literal_reference_loops(oracle_mod, bs, ns):
    """the oracle's halo update (closed form, oracle/evp_oracle.c: orc_halo_update) against the reference's own sequence restated loop
    for loop -- tripole buffer, pairwise symmetrisation of the top row, copy-out with the location offsets -- for the velocity pair
    (NE corner, vector: sign flip across the fold) and for a centre-located scalar, bit for bit including the signs of zeros."""
    c = synth.make_case("tiny", seed=71, ns=ns, ew="cyclic", kmt="none", ndte=1, block_size=bs)
    rng = np.random.default_rng(9)
    for loc, typ in ((1, 1), (0, 0)):
        a = rng.normal(size=c.fields["uvel"].shape)
        a[rng.random(a.shape) < 0.2] = 0.0                     # zeros: -0.0 / +0.0 must come out as in the reference
        want, got = a.copy(), a.copy()
        _halo_update_literal(c.grid, want, loc, typ)
        oracle_mod.halo_update(c.grid, [got], field_loc=loc, field_type=typ)
        bad = np.argwhere(got.view(np.int64) != want.view(np.int64))
        assert len(bad) == 0, (loc, typ, len(bad), bad[:6].tolist(), [(got[tuple(q)], want[tuple(q)]) for q in bad[:6]])
