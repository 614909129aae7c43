"""Deterministic synthetic inputs for the EVP subcycling path (numpy only; no oracle, no GPU).

The reference's grid / initial-condition / forcing files live outside its repository
(Fact 3), so every config of BASELINE.json is driven by the analytic `box2001`
set-up the reference ships, evaluated on a stretched rectangular grid:

  grid lengths   ice_grid.F90:3063-3280 (primary_grid_lengths_HTN/HTE), :681-715 (areas)
  land masks     ice_grid.F90:2676-2759 (rectgrid kmt types), :2935-3053 (boxislands),
                 :3333-3437 (makemask: uvm, tmask, umask)
  ice state      ice_init.F90:3467-3469, 3659-3698 (box2001: aice linear in x, hi = 2 m)
  forcing        ice_forcing.F90:5157-5172 (box2001 wind), :5242-5245 (box2001 current)
  dyn geometry   ice_dyn_shared.F90:384-441 (DminTarea, dxhy, dyhx, cxp, cyp, cxm, cym)
  preparation    ice_dyn_shared.F90:496-581 (dyn_prep1), :593-839 (dyn_prep2),
                 ice_grid.F90:4159-4211 ('S' T->U average), :4616-4660 ('F' T->U average),
                 icepack_mechred.F90:1720 (Hibler strength), ice_dyn_shared.F90:453-486 (constants)

Everything is computed on the ghost-extended GLOBAL arrays (ny_global+2, nx_global+2) and then
scattered into the reference's block layout (max_blocks, ny_block, nx_block) -- the memory layout
of Fortran `a(nx_block,ny_block,max_blocks)` -- so the same case can be cut into any block
decomposition with bit-identical cell values (the property the reference's decomp_suite checks).

This module is the stand-in for the Fortran caller's state, not part of the timed path.
"""
import numpy as np

from . import abi
from .decomp import create_blocks, Blocks

# icepack_parameters.F90 defaults
RHOI, RHOS, RHOW, RHOA = 917.0, 330.0, 1026.0, 1.3
PSTAR, CSTAR = 2.75e4, 20.0
CDN_OCN = 0.00536
FCOR_CONST = 1.46e-4  # ice_dyn_shared.F90:339
SECDAY = 86400.0

LOC_CENTER, LOC_NE, LOC_N, LOC_E = 0, 1, 2, 3

CONFIGS = {
    # name: nx_global, ny_global, dx0, ndte, ew, ns, kmt
    "tiny": dict(nx=24, ny=20, dx0=3.0e4, ndte=16, ew="cyclic", ns="closed", kmt="boxislands"),
    "gx3": dict(nx=100, ny=116, dx0=3.0e4, ndte=120, ew="cyclic", ns="closed", kmt="boxislands"),
    "gx1": dict(nx=320, ny=384, dx0=1.0e4, ndte=240, ew="cyclic", ns="closed", kmt="boxislands"),
    "tx1": dict(nx=360, ny=240, dx0=1.0e4, ndte=240, ew="cyclic", ns="tripole", kmt="boxislands"),
    "p1deg": dict(nx=3600, ny=2400, dx0=1.0e3, ndte=240, ew="cyclic", ns="closed", kmt="boxislands"),
}


# ------------------------------------------------------------------------------------------
# ghost-extended global arrays
# ------------------------------------------------------------------------------------------
def extend(a, ew, ns, loc=LOC_CENTER, vector=False, fill=0.0, extrap=False):
    """(ny,nx) global field -> (ny+2,nx+2) with the ghost ring a halo update / scatter would give.

    cyclic: wrapped; open/closed: `fill` (or linear extrapolation when extrap, as
    ice_HaloExtrapolate does for grid lengths, ice_boundary.F90:9056-9166); tripole north:
    folded by field location with the index map halochk states (halochk.F90:688-777), no row-je
    averaging (that belongs to the dynamic halo update, not to initial data)."""
    ny, nx = a.shape
    e = np.full((ny + 2, nx + 2), fill, dtype=a.dtype)
    e[1:-1, 1:-1] = a
    if ew == abi.BNDY_CYCLIC:
        e[1:-1, 0] = a[:, -1]
        e[1:-1, -1] = a[:, 0]
    elif extrap:
        e[1:-1, 0] = 2 * a[:, 0] - a[:, 1]
        e[1:-1, -1] = 2 * a[:, -1] - a[:, -2]
    if ns == abi.BNDY_CYCLIC:
        e[0, :] = e[-2, :]
        e[-1, :] = e[1, :]
    else:
        if extrap:
            e[0, :] = 2 * e[1, :] - e[2, :]
        if ns == abi.BNDY_TRIPOLE:
            sgn = -1.0 if vector else 1.0
            ig = np.arange(0, nx + 2)
            ig[0], ig[-1] = nx, 1  # wrapped global i of the ghost columns (tripole implies cyclic ew)
            ioff = -1 if loc in (LOC_NE, LOC_E) else 0
            joff = -1 if loc in (LOC_NE, LOC_N) else 0
            it = (nx - ig + 1 + ioff + nx - 1) % nx + 1
            jt = ny + joff
            e[-1, :] = sgn * e[jt, it]
        elif extrap:
            e[-1, :] = 2 * e[-2, :] - e[-3, :]
    return e


def scatter(e, blocks: Blocks, max_blocks=None):
    """ghost-extended global (ny+2,nx+2) -> block array (max_blocks, ny_block, nx_block)."""
    nb = blocks.nblocks_tot
    mb = max_blocks or nb
    out = np.zeros((mb, blocks.ny_block, blocks.nx_block), dtype=e.dtype)
    for n in range(nb):
        jj = np.abs(blocks.j_glob[n])
        ii = np.abs(blocks.i_glob[n])
        out[n] = e[np.ix_(jj, ii)]
    return out


def gather(field, blocks: Blocks):
    """block array -> (ny_global, nx_global) from block interiors."""
    g = np.zeros((blocks.ny_global, blocks.nx_global), dtype=field.dtype)
    for n in range(blocks.nblocks_tot):
        ilo, ihi, jlo, jhi = blocks.ilo[n], blocks.ihi[n], blocks.jlo[n], blocks.jhi[n]
        gi = blocks.i_glob[n, ilo - 1:ihi]
        gj = blocks.j_glob[n, jlo - 1:jhi]
        g[np.ix_(gj - 1, gi - 1)] = field[n, jlo - 1:jhi, ilo - 1:ihi]
    return g


def zero_ghosts(field, blocks: Blocks):
    """keep only block interiors (what dyn_prep2 leaves: fields are zeroed over the whole block and
    written at interior ice points only, ice_dyn_shared.F90:712-716, 795-837)."""
    out = np.zeros_like(field)
    for n in range(blocks.nblocks_tot):
        ilo, ihi, jlo, jhi = blocks.ilo[n], blocks.ihi[n], blocks.jlo[n], blocks.jhi[n]
        out[n, jlo - 1:jhi, ilo - 1:ihi] = field[n, jlo - 1:jhi, ilo - 1:ihi]
    return out


# ------------------------------------------------------------------------------------------
# land masks
# ------------------------------------------------------------------------------------------
def kmt_boxislands(nx, ny):
    """ice_grid.F90:2935-3053 (grid_boxislands_kmt); 1-based Fortran loops kept as written."""
    nxb, nyb = int(nx / 20.0), int(ny / 20.0)
    if nxb < 1 or nyb < 1:
        raise ValueError("boxislands requires a larger grid")
    w = np.ones((ny + 1, nx + 1))  # 1-based [j,i]

    def land(i0, i1, j0, j1):
        i0, j0 = max(i0, 1), max(j0, 1)
        i1, j1 = min(i1, nx), min(j1, ny)
        if i1 >= i0 and j1 >= j0:
            w[j0:j1 + 1, i0:i1 + 1] = 0.0

    k = 0
    for j in range(ny, ny - 3 * nyb - 1, -1):  # northeast triangle
        k += 1
        land(nx - 3 * nxb + k, nx, j, j)
    land(1, 1, ny - 3 * nyb, ny)  # northwest docks
    land(1, 2 * nxb, ny - 3 * nyb, ny - nyb - 2)
    land(1, 2 * nxb, ny - nyb, ny - nyb + 1)
    land(1, 1, 2 * nyb, 3 * nyb)  # southwest docks
    land(2, nxb, 1, 2 * nyb)
    land(2 * nxb - 1, 2 * nxb, 1, 2 * nyb)
    land(2 * nxb + 2, 4 * nxb, 1, 2 * nyb)
    land(14 * nxb, 14 * nxb + 1, 14 * nyb, 14 * nyb + 1)  # tiny island
    k = 0
    for i in range(2 * nxb, 4 * nxb + 1):  # X islands: left triangle
        k += 1
        land(i, i, 10 * nyb + k, 14 * nyb - k)
    k = 0
    for j in range(14 * nyb, 12 * nyb - 1, -1):  # upper triangle
        k += 1
        land(2 * nxb + 2 + k, 6 * nxb - 2 - k, j, j)
    k = 0
    for j in range(10 * nyb, 14 * nyb + 1):  # diagonal
        k += 1
        land(2 * nxb + 4 + k, 2 * nxb + 6 + k, j, j)
    k = 0
    for j in range(12 * nyb, 10 * nyb - 1, -1):  # lower right triangle
        k += 1
        land(5 * nxb + k, 8 * nxb, j, j)
    land(10 * nxb, 16 * nxb, 4 * nyb, 5 * nyb)  # bar islands
    land(10 * nxb, 16 * nxb, 6 * nyb + 2, 8 * nyb)
    land(10 * nxb, 16 * nxb, 8 * nyb + 2, 8 * nyb + 3)
    return w[1:, 1:].copy()


def make_kmt(kind, nx, ny, ew, ns):
    """ice_grid.F90:2676-2759."""
    if kind == "boxislands":
        hm = kmt_boxislands(nx, ny)
    elif kind == "none":
        hm = np.ones((ny, nx))
    elif kind == "channel":
        hm = np.zeros((ny, nx))
        hm[2:ny - 2, :] = 1.0
    elif kind == "continents":
        # not a reference option: boxislands plus two solid land masses (a block in the interior, a cap along the southern
        # edge) so that small blocks come out all-land and land-block elimination (ice_domain.F90) can be exercised
        hm = kmt_boxislands(nx, ny) if min(nx, ny) >= 20 else np.ones((ny, nx))
        hm[int(0.35 * ny):int(0.75 * ny), int(0.40 * nx):int(0.80 * nx)] = 0.0
        hm[:int(0.22 * ny), :] = 0.0
    else:
        raise ValueError(kind)
    if ew == abi.BNDY_CLOSED:
        hm[:, :2] = 0.0
        hm[:, -2:] = 0.0
    if ns == abi.BNDY_CLOSED:
        hm[:2, :] = 0.0
        hm[-2:, :] = 0.0
    return hm


# ------------------------------------------------------------------------------------------
# EVP constants: ice_dyn_shared.F90:453-486 (set_evp_parameters), defaults ice_init.F90:418-456
# ------------------------------------------------------------------------------------------
def evp_params(ndte, revised_evp=False, elasticDamp=0.36, e_yieldcurve=2.0, e_plasticpot=2.0,
               Ktens=0.0, capping=1.0, arlx=300.0, brlx=300.0, mode=abi.MODE_EXACT, kernel=abi.KERNEL_AUTO):
    p = dict(ndte=int(ndte), mode=mode, kernel=kernel)
    p["epp2i"] = 1.0 / e_plasticpot ** 2
    p["e_factor"] = e_yieldcurve ** 2 / e_plasticpot ** 4
    if revised_evp:
        p["revp"], p["denom1"], p["arlx1i"], p["brlx"] = 1.0, 1.0, 1.0 / arlx, brlx
    else:
        arlx = 2.0 * elasticDamp * float(ndte)
        p["revp"] = 0.0
        p["arlx1i"] = 1.0 / arlx
        p["brlx"] = float(ndte)
        p["denom1"] = 1.0 / (1.0 + p["arlx1i"])
    p.update(capping=capping, Ktens=Ktens, u0=5.0e-5, cosw=1.0, sinw=0.0, rhow=RHOW, visc_method=0, deltaminEVP=1e-11)
    return p


# ------------------------------------------------------------------------------------------
# the case builder
# ------------------------------------------------------------------------------------------
def _sl(e, di, dj):
    """e shifted so that result[j,i] = e[j+dj, i+di] on the interior (ny,nx) window."""
    ny, nx = e.shape[0] - 2, e.shape[1] - 2
    return e[1 + dj:1 + dj + ny, 1 + di:1 + di + nx]


def global_state(nx, ny, dx0, ew, ns, kmt="boxislands", dt=3600.0, deltaminEVP=1e-11,
                 dyn_area_min=1e-11, dyn_mass_min=1e-10, timesecs=0.0):
    """All loop inputs on ghost-extended global arrays. Returns dict of (ny+2,nx+2) arrays."""
    ig = np.arange(1, nx + 1, dtype=np.float64)[None, :]
    jg = np.arange(1, ny + 1, dtype=np.float64)[:, None]
    pi = np.pi
    # stretched grid: both metric terms dxhy, dyhx are non-zero
    HTN = dx0 * (1.0 + 0.1 * np.sin(2 * pi * ig / nx) * np.cos(pi * jg / ny))
    HTE = dx0 * (1.0 + 0.1 * np.cos(2 * pi * ig / nx) * np.sin(pi * jg / ny))

    # ice_grid.F90:3086-3131
    dxU = 0.5 * (HTN + np.roll(HTN, -1, axis=1))
    dxT = np.empty_like(HTN)
    dxT[1:, :] = 0.5 * (HTN[1:, :] + HTN[:-1, :])
    dxT[0, :] = 2.0 * HTN[1, :] - HTN[2, :]
    # ice_grid.F90:3197-3241
    dyU = np.empty_like(HTE)
    dyU[:-1, :] = 0.5 * (HTE[:-1, :] + HTE[1:, :])
    dyU[-1, :] = 2.0 * HTE[-2, :] - HTE[-3, :]
    dyT = 0.5 * (HTE + np.roll(HTE, 1, axis=1))

    X = {}
    X["HTN"] = extend(HTN, ew, ns, LOC_N, extrap=True)
    X["HTE"] = extend(HTE, ew, ns, LOC_E, extrap=True)
    X["dxT"] = extend(dxT, ew, ns, LOC_CENTER, extrap=True)
    X["dyT"] = extend(dyT, ew, ns, LOC_CENTER, extrap=True)
    X["dxU"] = extend(dxU, ew, ns, LOC_NE, extrap=True)
    X["dyU"] = extend(dyU, ew, ns, LOC_NE, extrap=True)
    # ice_grid.F90:681-715
    X["tarea"] = X["dxT"] * X["dyT"]
    uarea = X["dxU"] * X["dyU"]
    X["uarea"] = uarea
    X["uarear"] = np.where(uarea > 0.0, 1.0 / np.where(uarea > 0.0, uarea, 1.0), 0.0)

    # masks: ice_grid.F90:3333-3437
    hm = extend(make_kmt(kmt, nx, ny, ew, ns), ew, ns, LOC_CENTER)
    uvm_i = np.minimum(np.minimum(_sl(hm, 0, 0), _sl(hm, 1, 0)), np.minimum(_sl(hm, 0, 1), _sl(hm, 1, 1)))
    uvm = extend(uvm_i, ew, ns, LOC_NE)
    X["hm"], X["uvm"] = hm, uvm
    tmask, umask = hm > 0.5, uvm > 0.5

    # dyn geometry: ice_dyn_shared.F90:384-441
    X["DminTarea"] = deltaminEVP * X["tarea"]
    dxhy_i = 0.5 * (_sl(X["HTE"], 0, 0) - _sl(X["HTE"], -1, 0))
    dyhx_i = 0.5 * (_sl(X["HTN"], 0, 0) - _sl(X["HTN"], 0, -1))
    X["dxhy"] = extend(dxhy_i, ew, ns, LOC_CENTER, vector=True, fill=1.0)
    X["dyhx"] = extend(dyhx_i, ew, ns, LOC_CENTER, vector=True, fill=1.0)
    for nm in ("cxp", "cyp", "cxm", "cym"):
        X[nm] = np.zeros((ny + 2, nx + 2))
    He, Hn = X["HTE"], X["HTN"]
    X["cyp"][1:, 1:] = 1.5 * He[1:, 1:] - 0.5 * He[1:, :-1]
    X["cxp"][1:, 1:] = 1.5 * Hn[1:, 1:] - 0.5 * Hn[:-1, 1:]
    X["cym"][1:, 1:] = -(1.5 * He[1:, :-1] - 0.5 * He[1:, 1:])
    X["cxm"][1:, 1:] = -(1.5 * Hn[:-1, 1:] - 0.5 * Hn[1:, 1:])

    # ice state: ice_init.F90:3659-3698 (box2001 distribution, hbar = 2)
    aice_i = np.where(tmask[1:-1, 1:-1], (ig - 0.5) / nx + 0.0 * jg, 0.0)
    aice = extend(aice_i, ew, ns, LOC_CENTER)
    vice = 2.0 * aice
    vsno = np.zeros_like(aice)

    # forcing: ice_forcing.F90:5157-5172, 5242-5245
    period = 4.0 * SECDAY
    amp = np.sin(2 * pi * timesecs / period) - 3.0
    uatm = 5.0 + amp * np.sin(2 * pi * ig / nx) * np.sin(pi * jg / ny)
    vatm = 5.0 + amp * np.sin(pi * ig / nx) * np.sin(2 * pi * jg / ny)
    wind = np.sqrt(uatm ** 2 + vatm ** 2)
    tau = RHOA * 0.0012 * wind
    strax = extend(aice_i * tau * uatm, ew, ns, LOC_CENTER, vector=True)
    stray = extend(aice_i * tau * vatm, ew, ns, LOC_CENTER, vector=True)
    uocn = extend(0.2 * jg / ny - 0.1 + 0.0 * ig, ew, ns, LOC_CENTER, vector=True)
    vocn = extend(-0.2 * ig / nx + 0.1 + 0.0 * jg, ew, ns, LOC_CENTER, vector=True)

    # dyn_prep1: ice_dyn_shared.F90:536-577
    tmass = np.where(tmask, RHOI * vice + RHOS * vsno, 0.0)
    tmphm = tmask & (aice > dyn_area_min) & (tmass > dyn_mass_min)
    t3 = np.zeros((ny, nx), dtype=bool)
    for dj in (-1, 0, 1):
        for di in (-1, 0, 1):
            t3 |= _sl(tmphm, di, dj)
    iceT_i = t3 & tmask[1:-1, 1:-1]
    iceT = extend(iceT_i.astype(np.float64), ew, ns, LOC_CENTER) > 0.5
    X["iceTmask"] = iceT

    # T -> U, state masked: ice_grid.F90:4180-4205
    def t2u_S(w):
        mw = hm * X["tarea"]
        wt = _sl(mw, 0, 0) + _sl(mw, 1, 0) + _sl(mw, 0, 1) + _sl(mw, 1, 1)
        wm = w * mw  # mask*work*wght evaluated as (mask*work)*wght in the reference; mask is 0/1 so identical
        num = _sl(wm, 0, 0) + _sl(wm, 1, 0) + _sl(wm, 0, 1) + _sl(wm, 1, 1)
        return np.where(wt != 0.0, num / np.where(wt != 0.0, wt, 1.0), 0.0)

    # T -> U, flux unmasked: ice_grid.F90:4640-4655
    def t2u_F(w):
        wa = w * X["tarea"]
        return 0.25 * (_sl(wa, 0, 0) + _sl(wa, 1, 0) + _sl(wa, 0, 1) + _sl(wa, 1, 1)) / _sl(X["uarea"], 0, 0)

    umass = t2u_S(tmass)
    aiU = t2u_S(aice)
    cdn_ocnU = t2u_S(np.full_like(aice, CDN_OCN))
    uocnU, vocnU = t2u_S(uocn), t2u_S(vocn)
    strairxU, strairyU = t2u_F(strax), t2u_F(stray)

    # dyn_prep2: ice_dyn_shared.F90:759-837
    iceU = umask[1:-1, 1:-1] & (aiU > dyn_area_min) & (umass > dyn_mass_min)
    cosw, sinw = 1.0, 0.0
    fm = np.where(iceU, FCOR_CONST * umass, 0.0)
    sgn = np.copysign(1.0, fm)
    I = lambda a: np.where(iceU, a, 0.0)
    umassdti = I(umass / dt)
    waterx = I(uocnU * cosw - vocnU * sinw * sgn)
    watery = I(vocnU * cosw + uocnU * sinw * sgn)
    strtltx, strtlty = -fm * vocnU, fm * uocnU  # ssh_stress = 'geostrophic'
    forcex, forcey = I(strairxU + strtltx), I(strairyU + strtlty)
    uvel_i, vvel_i = I(uocnU), I(vocnU)  # new ice points start from the ocean current, :772-775

    def ext_u(a, vector=False):
        return extend(a, ew, ns, LOC_NE, vector=vector)

    X["iceUmask"] = ext_u(iceU.astype(np.float64)) > 0.5
    X["umassdti"], X["fmU"] = ext_u(umassdti), ext_u(fm)
    X["waterxU"], X["wateryU"] = ext_u(waterx), ext_u(watery)
    X["forcexU"], X["forceyU"] = ext_u(forcex), ext_u(forcey)
    X["aiU"], X["cdn_ocnU"] = ext_u(aiU), ext_u(cdn_ocnU)
    X["uocnU"], X["vocnU"] = ext_u(uocnU), ext_u(vocnU)
    X["uvel"], X["vvel"] = ext_u(uvel_i, True), ext_u(vvel_i, True)
    X["strax"], X["stray"], X["uocn"], X["vocn"], X["umass_i"] = strax, stray, uocn, vocn, umass
    X["strairxU_i"], X["strairyU_i"] = strairxU, strairyU   # interior arrays: dyn_prep2 inputs (tests/golden/ref_translit.py)
    X["TbU"] = np.zeros((ny + 2, nx + 2))

    # Hibler strength at ice T cells, then halo: ice_dyn_evp.F90:540-550, 727-728
    X["strength"] = np.where(iceT, PSTAR * vice * np.exp(-CSTAR * (1.0 - aice)), 0.0)
    X["aice"], X["vice"], X["tmass"] = aice, vice, tmass
    return X


def cgrid_state(X, nx, ny, dx0, ew, ns, dt=3600.0, dyn_area_min=1e-11, dyn_mass_min=1e-10, rheo_area_min=1e-11):
    """Extra geometry and per-step inputs of grid_ice='C' on the ghost-extended global arrays, added to X.

    grid lengths at E/N points   ice_grid.F90:3100-3131 (dxE), :3243-3266 (dyN), dxN = HTN, dyE = HTE
    areas, masks                 ice_grid.F90:681-715, :3370-3430 (epm, npm, uvmCD)
    BC ratios                    ice_dyn_evp.F90:226-240
    T -> E/N averages            ice_dyn_evp.F90:441-456 with grid_average_X2YS 'E','N' (ice_grid.F90:4289-4340)
    dyn_prep2 at E, N, U         ice_dyn_evp.F90:564-680 (umaskCD for the U mask, rheofact)
    initial uvelE.. = ocean current at new ice points (ice_dyn_shared.F90:772-775), then the re-interpolations
    of ice_dyn_evp.F90:700-724."""
    ig = np.arange(1, nx + 1, dtype=np.float64)[None, :]
    jg = np.arange(1, ny + 1, dtype=np.float64)[:, None]
    pi = np.pi
    HTN = dx0 * (1.0 + 0.1 * np.sin(2 * pi * ig / nx) * np.cos(pi * jg / ny))
    HTE = dx0 * (1.0 + 0.1 * np.cos(2 * pi * ig / nx) * np.sin(pi * jg / ny))
    Hn1 = np.roll(HTN, -1, axis=1)  # (ip1, j)
    dxE = np.empty_like(HTN)
    dxE[1:, :] = 0.25 * (HTN[1:, :] + Hn1[1:, :] + HTN[:-1, :] + Hn1[:-1, :])
    dxE[0, :] = 0.5 * (2.0 * HTN[1, :] - HTN[2, :] + 2.0 * Hn1[1, :] - Hn1[2, :])
    Hem = np.roll(HTE, 1, axis=1)  # (im1, j)
    dyN = np.empty_like(HTE)
    dyN[:-1, :] = 0.25 * (HTE[:-1, :] + Hem[:-1, :] + HTE[1:, :] + Hem[1:, :])
    dyN[-1, :] = 0.5 * (2.0 * HTE[-2, :] - HTE[-3, :] + 2.0 * Hem[-2, :] - Hem[-3, :])
    X["dxN"], X["dyE"] = X["HTN"], X["HTE"]
    X["dxE"] = extend(dxE, ew, ns, LOC_E, extrap=True)
    X["dyN"] = extend(dyN, ew, ns, LOC_N, extrap=True)
    X["earea"] = X["dxE"] * X["dyE"]
    X["narea"] = X["dxN"] * X["dyN"]
    for a, r in (("earea", "earear"), ("narea", "narear")):
        X[r] = np.where(X[a] > 0.0, 1.0 / np.where(X[a] > 0.0, X[a], 1.0), 0.0)
    hm = X["hm"]
    npm = extend(np.minimum(_sl(hm, 0, 0), _sl(hm, 0, 1)), ew, ns, LOC_N)
    epm = extend(np.minimum(_sl(hm, 0, 0), _sl(hm, 1, 0)), ew, ns, LOC_E)
    uvmCD = extend(_sl(hm, 0, 0) + _sl(hm, 1, 0) + _sl(hm, 0, 1) + _sl(hm, 1, 1), ew, ns, LOC_NE)
    X["npm"], X["epm"] = npm, epm
    for nm in ("ratiodxN", "ratiodxNr", "ratiodyE", "ratiodyEr"):
        X[nm] = np.zeros((ny + 2, nx + 2))
    X["ratiodxN"][1:-1, 1:-1] = -_sl(X["dxN"], 1, 0) / _sl(X["dxN"], 0, 0)
    X["ratiodyE"][1:-1, 1:-1] = -_sl(X["dyE"], 0, 1) / _sl(X["dyE"], 0, 0)
    X["ratiodxNr"][1:-1, 1:-1] = 1.0 / X["ratiodxN"][1:-1, 1:-1]
    X["ratiodyEr"][1:-1, 1:-1] = 1.0 / X["ratiodyE"][1:-1, 1:-1]

    mw = hm * X["tarea"]

    def t2x_S(w, di, dj):  # two-point masked average T -> E (di=1) or N (dj=1)
        wt = _sl(mw, 0, 0) + _sl(mw, di, dj)
        wm = w * mw
        num = _sl(wm, 0, 0) + _sl(wm, di, dj)
        return np.where(wt != 0.0, num / np.where(wt != 0.0, wt, 1.0), 0.0)

    def t2x_F(w, di, dj, area):
        wa = w * X["tarea"]
        return 0.5 * (_sl(wa, 0, 0) + _sl(wa, di, dj)) / _sl(area, 0, 0)

    aice, tmass = X["aice"], X["tmass"]
    pts = {"E": (1, 0, epm, X["earea"], LOC_E), "N": (0, 1, npm, X["narea"], LOC_N)}
    strax, stray = X["strax"], X["stray"]
    uocn, vocn = X["uocn"], X["vocn"]
    cosw, sinw = 1.0, 0.0
    for P, (di, dj, pm, area, loc) in pts.items():
        mass = t2x_S(tmass, di, dj)
        ai = t2x_S(aice, di, dj)
        cdn = t2x_S(np.full_like(aice, CDN_OCN), di, dj)
        uo, vo = t2x_S(uocn, di, dj), t2x_S(vocn, di, dj)
        sx, sy = t2x_F(strax, di, dj, area), t2x_F(stray, di, dj, area)
        ice = (pm[1:-1, 1:-1] > 0.5) & (ai > dyn_area_min) & (mass > dyn_mass_min)
        fm = np.where(ice, FCOR_CONST * mass, 0.0)
        sgn = np.copysign(1.0, fm)
        I = lambda a: np.where(ice, a, 0.0)
        ext = lambda a, vector=False: extend(a, ew, ns, loc, vector=vector)
        X["ice%smask" % P] = ext(ice.astype(np.float64)) > 0.5
        X["%smassdti" % P.lower()] = ext(I(mass / dt))
        X["fm" + P] = ext(fm)
        X["ai" + P], X["cdn_ocn" + P] = ext(ai), ext(cdn)
        X["uocn" + P], X["vocn" + P] = ext(uo), ext(vo)
        X["waterx" + P] = ext(I(uo * cosw - vo * sinw * sgn))
        X["watery" + P] = ext(I(vo * cosw + uo * sinw * sgn))
        X["forcex" + P] = ext(I(sx + -fm * vo))
        X["forcey" + P] = ext(I(sy + fm * uo))
        X["rheofact" + P] = ext(np.where(ice, np.where(ai > rheo_area_min, 1.0, 0.0), 0.0))
        X["Tb" + P] = np.zeros((ny + 2, nx + 2))
        X["uvel" + P], X["vvel" + P] = ext(I(uo), True), ext(I(vo), True)
    # the U mask of the C grid uses umaskCD (ice_dyn_evp.F90:572)
    aiU, umass = X["aiU"][1:-1, 1:-1], None
    iceU = (uvmCD[1:-1, 1:-1] > 1.5) & (X["aiU"][1:-1, 1:-1] > dyn_area_min) & (X["umass_i"] > dyn_mass_min)
    X["iceUmaskC"] = extend(iceU.astype(np.float64), ew, ns, LOC_NE) > 0.5
    # velocities as ice_dyn_evp.F90:700-724 leaves them: uvelN, vvelE and uvel, vvel re-interpolated from uvelE, vvelN
    ea, na = X["earea"], X["narea"]

    def avgA(w, wg, offs):
        wt = sum(_sl(wg, di, dj) for di, dj in offs)
        num = sum(_sl(w * wg, di, dj) for di, dj in offs)
        return np.where(wt != 0.0, num / np.where(wt != 0.0, wt, 1.0), 0.0)

    uE, vN = X["uvelE"], X["vvelN"]
    X["uvelN"] = extend(avgA(uE, ea, [(-1, 0), (0, 0), (-1, 1), (0, 1)]) * npm[1:-1, 1:-1], ew, ns, LOC_N, vector=True)
    X["vvelE"] = extend(avgA(vN, na, [(0, -1), (1, -1), (0, 0), (1, 0)]) * epm[1:-1, 1:-1], ew, ns, LOC_E, vector=True)
    X["uvelC"] = extend(avgA(uE, ea, [(0, 0), (0, 1)]) * X["uvm"][1:-1, 1:-1], ew, ns, LOC_NE, vector=True)
    X["vvelC"] = extend(avgA(vN, na, [(0, 0), (1, 0)]) * X["uvm"][1:-1, 1:-1], ew, ns, LOC_NE, vector=True)
    return X


class CCase:
    """One synthetic C-grid EVP step in the reference's block layout."""

    def __init__(self, blocks, grid, cgrid, params, fields, X):
        self.blocks, self.grid, self.cgrid, self.params, self.fields, self.X = blocks, grid, cgrid, params, fields, X

    def copy_fields(self):
        return {k: v.copy() for k, v in self.fields.items()}


def make_ccase(config="gx3", block_size=None, seed=None, ndte=None, mode=abi.MODE_EXACT, revised_evp=False,
               visc_method=abi.VISC_AVG_ZETA, kmt=None, ew=None, ns=None, nx=None, ny=None, **kw):
    """grid_ice='C' counterpart of make_case (configs[2]: gx1 C grid, ndte=600)."""
    c = dict(CONFIGS[config])
    if nx:
        c["nx"] = nx
    if ny:
        c["ny"] = ny
    ew_i = abi.BNDY_NAMES[ew or c["ew"]]
    ns_i = abi.BNDY_NAMES[ns or c["ns"]]
    if ns_i == abi.BNDY_TRIPOLE:
        raise ValueError("C-grid cases are not generated for tripole grids")
    nxg, nyg = c["nx"], c["ny"]
    bsx, bsy = block_size or (nxg, nyg)
    blocks = create_blocks(nxg, nyg, bsx, bsy, ew_i, ns_i)
    X = global_state(nxg, nyg, c["dx0"], ew_i, ns_i, kmt or c["kmt"], **kw)
    X = cgrid_state(X, nxg, nyg, c["dx0"], ew_i, ns_i)
    shape = (nyg + 2, nxg + 2)
    rng = np.random.Generator(np.random.PCG64(seed)) if seed is not None else None
    iceT = X["iceTmask"]
    for n in ("stresspT", "stressmT", "stress12T", "stress12U"):
        if rng is None:
            X[n] = np.zeros(shape)
        else:
            msk = X["iceUmaskC"] if n.endswith("U") else iceT
            loc = LOC_NE if n.endswith("U") else LOC_CENTER
            X[n] = np.where(msk, extend(rng.normal(0.0, 1.0, (nyg, nxg)), ew_i, ns_i, loc) * 0.2 * X["strength"].max() * 0.05, 0.0)
    if rng is not None:
        for P, loc in (("E", LOC_E), ("N", LOC_N)):
            ice = X["ice%smask" % P][1:-1, 1:-1]
            X["uvel" + P] = extend(np.where(ice, rng.uniform(-0.3, 0.3, (nyg, nxg)), 0.0), ew_i, ns_i, loc, vector=True)
            X["vvel" + P] = extend(np.where(ice, rng.uniform(-0.3, 0.3, (nyg, nxg)), 0.0), ew_i, ns_i, loc, vector=True)
            X["Tb" + P] = np.where(X["ice%smask" % P], extend(rng.uniform(0.0, 5.0, (nyg, nxg)), ew_i, ns_i, loc), 0.0)
        iu = X["iceUmaskC"][1:-1, 1:-1]
        X["uvelC"] = extend(np.where(iu, rng.uniform(-0.3, 0.3, (nyg, nxg)), 0.0), ew_i, ns_i, LOC_NE, vector=True)
        X["vvelC"] = extend(np.where(iu, rng.uniform(-0.3, 0.3, (nyg, nxg)), 0.0), ew_i, ns_i, LOC_NE, vector=True)

    mb = blocks.nblocks_tot
    grid = dict(nx_block=blocks.nx_block, ny_block=blocks.ny_block, nblocks=blocks.nblocks_tot, max_blocks=mb,
                nghost=1, nx_global=nxg, ny_global=nyg, ew_boundary_type=ew_i, ns_boundary_type=ns_i,
                ilo=blocks.ilo, ihi=blocks.ihi, jlo=blocks.jlo, jhi=blocks.jhi, i_glob=blocks.i_glob, j_glob=blocks.j_glob)
    for n in abi.GRID_STATIC:
        grid[n] = scatter(X[n], blocks, mb)
    cgrid = {n: scatter(X[n], blocks, mb) for n in abi.CGRID_STATIC}
    src = {"uvel": "uvelC", "vvel": "vvelC", "emassdti": "emassdti", "nmassdti": "nmassdti"}
    fields = {}
    for n in abi.CFIELDS_ORDER:
        if n in abi.CFIELDS_OUT:
            fields[n] = np.zeros((mb, blocks.ny_block, blocks.nx_block))
        else:
            fields[n] = scatter(X[src.get(n, n)], blocks, mb)
    fields["iceTmask"] = scatter(X["iceTmask"].astype(np.int32), blocks, mb)
    fields["iceUmask"] = zero_ghosts(scatter(X["iceUmaskC"].astype(np.int32), blocks, mb), blocks)
    fields["iceEmask"] = zero_ghosts(scatter(X["iceEmask"].astype(np.int32), blocks, mb), blocks)
    fields["iceNmask"] = zero_ghosts(scatter(X["iceNmask"].astype(np.int32), blocks, mb), blocks)
    params = evp_params(ndte or c["ndte"], revised_evp=revised_evp, mode=mode)
    params["visc_method"] = visc_method
    params["deltaminEVP"] = 1e-11
    return CCase(blocks, grid, cgrid, params, fields, X)


def make_cdcase(config="gx3", block_size=None, seed=None, ndte=None, mode=abi.MODE_EXACT, revised_evp=False,
                visc_method=abi.VISC_AVG_ZETA, kmt=None, ew=None, ns=None, **kw):
    """grid_ice='CD' counterpart of make_ccase (ice_dyn_evp.F90:1123-1293): same geometry and E/N-point coefficients (dyn_prep2
    at E and N yields both components of water*, force*), all four of uvelE, vvelE, uvelN, vvelN prognostic, the stress
    tensor carried at T and at U."""
    cc = make_ccase(config, block_size=block_size, seed=seed, ndte=ndte, mode=mode, revised_evp=revised_evp, visc_method=visc_method,
                    kmt=kmt, ew=ew, ns=ns, **kw)
    X, blocks = cc.X, cc.blocks
    mb = blocks.nblocks_tot
    nyg, nxg = X["aice"].shape[0] - 2, X["aice"].shape[1] - 2
    ew_i, ns_i = cc.grid["ew_boundary_type"], cc.grid["ns_boundary_type"]
    rng = np.random.Generator(np.random.PCG64(seed + 1000)) if seed is not None else None
    for n in ("stresspU", "stressmU"):
        X[n] = (np.zeros((nyg + 2, nxg + 2)) if rng is None else
                np.where(X["iceUmaskC"], extend(rng.normal(0.0, 1.0, (nyg, nxg)), ew_i, ns_i, LOC_NE) * 0.01 * X["strength"].max(), 0.0))
    src = {"uvel": "uvelC", "vvel": "vvelC"}
    fields = {}
    for n in abi.CDFIELDS_ORDER:
        if n in abi.CDFIELDS_OUT:
            fields[n] = np.zeros((mb, blocks.ny_block, blocks.nx_block))
        elif n in cc.fields:
            fields[n] = cc.fields[n].copy()
        else:
            fields[n] = scatter(X[src.get(n, n)], blocks, mb)
    for n in abi.CFIELDS_MASK:
        fields[n] = cc.fields[n].copy()
    return CCase(blocks, cc.grid, cc.cgrid, cc.params, fields, X)


U_PREP = ("umassdti", "fmU", "waterxU", "wateryU", "forcexU", "forceyU", "TbU")  # interior-only after dyn_prep2


def step_inputs(c, max_blocks=None):
    """(static, prep) of cice_b200.dyn_evp.dyn_evp_b200_prep_init / dyn_evp_b200_step_resident for a Case: the T-point inputs of the
    step and the static arrays, in the reference's block layout, from the same global state the U-point inputs of c.fields were
    derived from -- the device-side preparation must reproduce those bit for bit."""
    X, B = c.X, c.blocks
    sc = lambda a: np.ascontiguousarray(scatter(np.asarray(a, dtype=np.float64), B, max_blocks))
    static = dict(hm=sc(X["hm"]), tarea=sc(X["tarea"]), uarea=sc(X["uarea"]), fcor=np.full_like(sc(X["hm"]), FCOR_CONST),
                  umask=np.ascontiguousarray(scatter((X["uvm"] > 0.5).astype(np.int32), B, max_blocks)))
    prep = dict(tmass=sc(X["tmass"]), aice_init=sc(X["aice"]), cdn_ocn=np.full_like(sc(X["hm"]), CDN_OCN), uocn=sc(X["uocn"]),
                vocn=sc(X["vocn"]), strairxT=sc(X["strax"]), strairyT=sc(X["stray"]), strength=c.fields["strength"].copy(),
                iceTmask=c.fields["iceTmask"].copy(), dt=3600.0, dyn_area_min=1e-11, dyn_mass_min=1e-10)
    return static, prep


class Case:
    """One synthetic EVP step in the reference's block layout, ready for the C ABI."""

    def __init__(self, blocks: Blocks, grid, params, fields, X=None):
        self.blocks, self.grid, self.params, self.fields, self.X = blocks, grid, params, fields, X

    def copy_fields(self):
        return {k: v.copy() for k, v in self.fields.items()}

    def rank_view(self, owner, rank):
        """grid/fields restricted to the blocks `rank` owns (what that MPI rank would hold)."""
        ids = np.nonzero(owner == rank)[0]
        b = self.blocks
        g = dict(self.grid)
        g["nblocks"] = g["max_blocks"] = len(ids)
        for n in ("ilo", "ihi", "jlo", "jhi"):
            g[n] = np.ascontiguousarray(self.grid[n][ids])
        g["i_glob"] = np.ascontiguousarray(b.i_glob[ids])
        g["j_glob"] = np.ascontiguousarray(b.j_glob[ids])
        for n in abi.GRID_STATIC:
            g[n] = np.ascontiguousarray(self.grid[n][ids])
        f = {k: np.ascontiguousarray(v[ids]) for k, v in self.fields.items()}
        return g, f, ids


def make_case(config="gx3", block_size=None, seed=None, ndte=None, mode=abi.MODE_EXACT, kernel=abi.KERNEL_AUTO,
              revised_evp=False, kmt=None, ew=None, ns=None, nx=None, ny=None, max_blocks=None, **kw):
    """Build a Case.  seed=None gives set S1 (deterministic box2001 start); an integer seed gives set S2
    (random velocities and stresses on top) for kernel-equivalence tests."""
    c = dict(CONFIGS[config])
    if nx:
        c["nx"] = nx
    if ny:
        c["ny"] = ny
    ew_i = abi.BNDY_NAMES[ew or c["ew"]]
    ns_i = abi.BNDY_NAMES[ns or c["ns"]]
    nxg, nyg = c["nx"], c["ny"]
    bsx, bsy = block_size or (nxg, nyg)
    blocks = create_blocks(nxg, nyg, bsx, bsy, ew_i, ns_i)
    X = global_state(nxg, nyg, c["dx0"], ew_i, ns_i, kmt or c["kmt"], **kw)

    if seed is not None:
        rng = np.random.Generator(np.random.PCG64(seed))
        iceU = X["iceUmask"][1:-1, 1:-1]
        u = np.where(iceU, rng.uniform(-0.3, 0.3, (nyg, nxg)), 0.0)
        v = np.where(iceU, rng.uniform(-0.3, 0.3, (nyg, nxg)), 0.0)
        X["uvel"] = extend(u, ew_i, ns_i, LOC_NE, vector=True)
        X["vvel"] = extend(v, ew_i, ns_i, LOC_NE, vector=True)
        iceT = X["iceTmask"]
        for n in abi.STRESS:
            s = rng.normal(0.0, 1.0, (nyg, nxg))
            X[n] = np.where(iceT, extend(s, ew_i, ns_i, LOC_CENTER) * 0.2 * X["strength"], 0.0)
        X["TbU"] = np.where(X["iceUmask"], extend(rng.uniform(0.0, 5.0, (nyg, nxg)), ew_i, ns_i, LOC_NE), 0.0)
    else:
        for n in abi.STRESS:
            X[n] = np.zeros((nyg + 2, nxg + 2))

    mb = max_blocks or blocks.nblocks_tot
    grid = dict(nx_block=blocks.nx_block, ny_block=blocks.ny_block, nblocks=blocks.nblocks_tot, max_blocks=mb,
                nghost=1, nx_global=nxg, ny_global=nyg, ew_boundary_type=ew_i, ns_boundary_type=ns_i,
                ilo=blocks.ilo, ihi=blocks.ihi, jlo=blocks.jlo, jhi=blocks.jhi,
                i_glob=blocks.i_glob, j_glob=blocks.j_glob)
    for n in abi.GRID_STATIC:
        grid[n] = scatter(X[n], blocks, mb)
    fields = {}
    for n in abi.FIELDS_ORDER:
        if n in ("strintxU", "strintyU", "taubxU", "taubyU"):
            fields[n] = np.zeros((mb, blocks.ny_block, blocks.nx_block))
        else:
            a = scatter(X[n], blocks, mb)
            fields[n] = zero_ghosts(a, blocks) if n in U_PREP else a
    for n in abi.FIELDS_MASK:
        fields[n] = scatter(X[n].astype(np.int32), blocks, mb)
    # iceUmask is defined on block interiors only (dyn_prep2 loop bounds, ice_dyn_shared.F90:759-760)
    fields["iceUmask"] = zero_ghosts(fields["iceUmask"], blocks)
    params = evp_params(ndte or c["ndte"], revised_evp=revised_evp, mode=mode, kernel=kernel)
    return Case(blocks, grid, params, fields, X)
