"""ctypes mirror of include/evp_b200.h (structs and enums only; no library is loaded here).

The product binding (cice_b200.dyn_evp) and the test-side checker binding both
describe their buffers with these structs, so a test feeds both sides from one set of arrays.
"""
import ctypes as C

import numpy as np

ABI_VERSION = 3

BNDY_OPEN, BNDY_CLOSED, BNDY_CYCLIC, BNDY_TRIPOLE = 0, 1, 2, 3
BNDY_NAMES = {"open": BNDY_OPEN, "closed": BNDY_CLOSED, "cyclic": BNDY_CYCLIC, "tripole": BNDY_TRIPOLE}

MODE_EXACT, MODE_FAST = 0, 1
KERNEL_AUTO, KERNEL_SPLIT, KERNEL_FUSED, KERNEL_PERSISTENT, KERNEL_FUSED_STREAM, KERNEL_FUSED_RESIDENT, KERNEL_TSTREAM = 0, 1, 2, 3, 4, 5, 6
KERNEL_NAMES = {"auto": 0, "split": 1, "fused": 2, "persistent": 3, "stream": 4, "resident": 5, "tstream": 6}

UNIQUE_ID_BYTES = 128
KEEP_STRESS, FETCH_STRESS = 1, 2  # evp_b200_run_bgrid_resident flags
STEP_INIT_STATE, STEP_FETCH_DIAG, STEP_FETCH_STATE = 1, 2, 4  # evp_b200_step_resident flags
SSH_GEOSTROPHIC, SSH_COUPLED = 0, 1

_pd = C.POINTER(C.c_double)
_pi = C.POINTER(C.c_int32)

GRID_STATIC = ("dxT", "dyT", "dxhy", "dyhx", "cxp", "cyp", "cxm", "cym", "DminTarea", "uarear")


class Grid(C.Structure):
    _fields_ = (
        [(n, C.c_int32) for n in ("abi_version", "nx_block", "ny_block", "nblocks", "max_blocks", "nghost",
                                  "nx_global", "ny_global", "ew_boundary_type", "ns_boundary_type")]
        + [(n, _pi) for n in ("ilo", "ihi", "jlo", "jhi", "i_glob", "j_glob")]
        + [(n, _pd) for n in GRID_STATIC]
    )


class Params(C.Structure):
    _fields_ = (
        [(n, C.c_int32) for n in ("ndte", "mode", "kernel", "visc_method")]
        + [(n, C.c_double) for n in ("arlx1i", "denom1", "revp", "brlx", "e_factor", "epp2i", "capping",
                                     "Ktens", "u0", "cosw", "sinw", "rhow", "deltaminEVP")]
    )


STRESS = tuple(f"stress{k}_{n}" for k in ("p", "m", "12") for n in (1, 2, 3, 4))
FIELDS_IN = ("strength", "cdn_ocnU", "aiU", "uocnU", "vocnU", "waterxU", "wateryU", "forcexU", "forceyU",
             "umassdti", "fmU", "TbU")
FIELDS_INOUT = STRESS + ("strintxU", "strintyU", "taubxU", "taubyU", "uvel", "vvel")
FIELDS_MASK = ("iceTmask", "iceUmask")
# struct order == argument order of dyn_evp1d_run (ice_dyn_evp1d.F90:119-153)
FIELDS_ORDER = (STRESS + ("strength", "cdn_ocnU", "aiU", "uocnU", "vocnU", "waterxU", "wateryU", "forcexU",
                          "forceyU", "umassdti", "fmU", "strintxU", "strintyU", "TbU", "taubxU", "taubyU",
                          "uvel", "vvel"))


class Fields(C.Structure):
    _fields_ = [(n, _pd) for n in FIELDS_ORDER] + [(n, _pi) for n in FIELDS_MASK]


VISC_AVG_ZETA, VISC_AVG_STRENGTH = 0, 1

CGRID_STATIC = ("dxN", "dyE", "dxE", "dyN", "dxU", "dyU", "tarea", "uarea", "earea", "narea", "earear", "narear",
                "ratiodxN", "ratiodxNr", "ratiodyE", "ratiodyEr", "hm", "uvm", "epm", "npm")


class CGrid(C.Structure):
    _fields_ = [(n, _pd) for n in CGRID_STATIC]


CFIELDS_INOUT = ("uvelE", "vvelE", "uvelN", "vvelN", "uvel", "vvel", "stresspT", "stressmT", "stress12T", "stress12U")
CFIELDS_OUT = ("zetax2T", "etax2T", "etax2U", "strengthU", "divergU", "tensionU", "shearU", "deltaU",
               "strintxE", "strintyN", "taubxE", "taubyN")
CFIELDS_IN = ("strength", "cdn_ocnE", "cdn_ocnN", "aiE", "aiN", "uocnE", "vocnE", "uocnN", "vocnN", "waterxE", "wateryN",
              "forcexE", "forceyN", "emassdti", "nmassdti", "fmE", "fmN", "TbE", "TbN", "rheofactE", "rheofactN")
CFIELDS_ORDER = CFIELDS_INOUT + CFIELDS_OUT + CFIELDS_IN
CFIELDS_MASK = ("iceTmask", "iceUmask", "iceEmask", "iceNmask")


class CFields(C.Structure):
    _fields_ = [(n, _pd) for n in CFIELDS_ORDER] + [(n, _pi) for n in CFIELDS_MASK]


DEFORM_IN = ("dxU", "dyU", "tarear")
# grid_ice = 'CD' (ice_dyn_evp.F90:1123-1293): both velocity components at E and at N, the full stress tensor at T and at U
CDFIELDS_INOUT = ("uvelE", "vvelE", "uvelN", "vvelN", "uvel", "vvel", "stresspT", "stressmT", "stress12T", "stresspU", "stressmU",
                  "stress12U")
CDFIELDS_OUT = ("zetax2T", "etax2T", "zetax2U", "etax2U", "strengthU", "divergU", "tensionU", "shearU", "deltaU",
                "strintxE", "strintyE", "strintxN", "strintyN", "taubxE", "taubyE", "taubxN", "taubyN")
CDFIELDS_IN = ("strength", "cdn_ocnE", "cdn_ocnN", "aiE", "aiN", "uocnE", "vocnE", "uocnN", "vocnN", "waterxE", "wateryE", "waterxN",
               "wateryN", "forcexE", "forceyE", "forcexN", "forceyN", "emassdti", "nmassdti", "fmE", "fmN", "TbE", "TbN", "rheofactE",
               "rheofactN")
CDFIELDS_ORDER = CDFIELDS_INOUT + CDFIELDS_OUT + CDFIELDS_IN


class CDFields(C.Structure):
    _fields_ = [(n, _pd) for n in CDFIELDS_ORDER] + [(n, _pi) for n in CFIELDS_MASK]


def make_cdfields(f, npl_total):
    s, keep = CDFields(), {}
    for n in CDFIELDS_ORDER:
        a = f[n]
        assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"] and a.size == npl_total, n
        keep[n] = a
        setattr(s, n, _ptr(a, C.c_double))
    for n in CFIELDS_MASK:
        a = f[n]
        assert a.dtype == np.int32 and a.flags["C_CONTIGUOUS"] and a.size == npl_total, n
        keep[n] = a
        setattr(s, n, _ptr(a, C.c_int32))
    return s, keep


DEFORM_OUT = ("divu", "shear", "vort", "rdg_conv", "rdg_shear")


class Deform(C.Structure):
    _fields_ = [(n, _pd) for n in DEFORM_IN + DEFORM_OUT] + [("e_factor", C.c_double)]


def make_deform(d, npl_total, e_factor):
    s, keep = Deform(), {}
    for n in DEFORM_IN:
        keep[n] = as_f64(d[n])
        setattr(s, n, _ptr(keep[n], C.c_double))
    for n in DEFORM_OUT:
        a = d[n]
        assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"] and a.size == npl_total, n
        keep[n] = a
        setattr(s, n, _ptr(a, C.c_double))
    s.e_factor = float(e_factor)
    return s, keep


class Finish(C.Structure):   # evp_b200_finish_t
    _fields_ = [("strocnxU", _pd), ("strocnyU", _pd), ("rhow", C.c_double), ("cosw", C.c_double), ("sinw", C.c_double)]


def make_finish(d, npl_total, rhow, cosw, sinw):
    s, keep = Finish(), {}
    for n in ("strocnxU", "strocnyU"):
        a = d[n]
        assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"] and a.size == npl_total, n
        keep[n] = a
        setattr(s, n, _ptr(a, C.c_double))
    s.rhow, s.cosw, s.sinw = float(rhow), float(cosw), float(sinw)
    return s, keep


def _ptr(a, ctype):
    return a.ctypes.data_as(C.POINTER(ctype))


def as_f64(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a


def make_grid(g):
    """Build a Grid struct from a dict of numpy arrays / ints. Returns (struct, keepalive)."""
    keep = {}
    s = Grid()
    s.abi_version = ABI_VERSION
    for n in ("nx_block", "ny_block", "nblocks", "max_blocks", "nghost", "nx_global", "ny_global",
              "ew_boundary_type", "ns_boundary_type"):
        setattr(s, n, int(g[n]))
    for n in ("ilo", "ihi", "jlo", "jhi", "i_glob", "j_glob"):
        keep[n] = np.ascontiguousarray(g[n], dtype=np.int32)
        setattr(s, n, _ptr(keep[n], C.c_int32))
    for n in GRID_STATIC:
        keep[n] = as_f64(g[n])
        assert keep[n].size == s.nx_block * s.ny_block * s.max_blocks, n
        setattr(s, n, _ptr(keep[n], C.c_double))
    return s, keep


def make_cgrid(cg, npl_total):
    s, keep = CGrid(), {}
    for n in CGRID_STATIC:
        keep[n] = as_f64(cg[n])
        assert keep[n].size == npl_total, n
        setattr(s, n, _ptr(keep[n], C.c_double))
    return s, keep


def make_cfields(f, npl_total):
    s, keep = CFields(), {}
    for n in CFIELDS_ORDER:
        a = f[n]
        assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"] and a.size == npl_total, n
        keep[n] = a
        setattr(s, n, _ptr(a, C.c_double))
    for n in CFIELDS_MASK:
        a = f[n]
        assert a.dtype == np.int32 and a.flags["C_CONTIGUOUS"] and a.size == npl_total, n
        keep[n] = a
        setattr(s, n, _ptr(a, C.c_int32))
    return s, keep


def make_params(p):
    s = Params()
    for n, _ in Params._fields_:
        if n in p:
            setattr(s, n, p[n])
    return s


def make_fields(f, npl_total):
    """Fields struct over the arrays in dict f (must already be C-contiguous with the right dtype:
    they are written in place by run calls). Returns (struct, keepalive)."""
    s = Fields()
    keep = {}
    for n in FIELDS_ORDER:
        a = f[n]
        assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"] and a.size == npl_total, n
        keep[n] = a
        setattr(s, n, _ptr(a, C.c_double))
    for n in FIELDS_MASK:
        a = f[n]
        assert a.dtype == np.int32 and a.flags["C_CONTIGUOUS"] and a.size == npl_total, n
        keep[n] = a
        setattr(s, n, _ptr(a, C.c_int32))
    return s, keep


# ---- evp_b200_prep_init / evp_b200_step_resident -----------------------------------------------------
PREP_STATIC = ("hm", "tarea", "uarea", "fcor")
PREP_T = ("tmass", "aice_init", "cdn_ocn", "uocn", "vocn", "ss_tltx", "ss_tlty", "strairxT", "strairyT", "strength")
PREP_OPTIONAL = ("ss_tltx", "ss_tlty", "TbU")


class PrepStatic(C.Structure):
    _fields_ = [(n, _pd) for n in PREP_STATIC] + [("umask", _pi)]


class Prep(C.Structure):
    _fields_ = ([(n, _pd) for n in PREP_T] + [("iceTmask", _pi), ("TbU", _pd)]
                + [(n, C.c_double) for n in ("dt", "dyn_area_min", "dyn_mass_min", "gravit")] + [("ssh_stress", C.c_int32)])


def make_prep_static(d, npl_total):
    s, keep = PrepStatic(), {}
    for n in PREP_STATIC:
        a = d[n]
        assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"] and a.size == npl_total, n
        keep[n] = a
        setattr(s, n, _ptr(a, C.c_double))
    a = d["umask"]
    assert a.dtype == np.int32 and a.flags["C_CONTIGUOUS"] and a.size == npl_total
    keep["umask"] = a
    s.umask = _ptr(a, C.c_int32)
    return s, keep


def make_prep(d, npl_total):
    s, keep = Prep(), {}
    for n in PREP_T + ("TbU",):
        a = d.get(n)
        if a is None:
            assert n in PREP_OPTIONAL, n
            continue
        assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"] and a.size == npl_total, n
        keep[n] = a
        setattr(s, n, _ptr(a, C.c_double))
    a = d["iceTmask"]
    assert a.dtype == np.int32 and a.flags["C_CONTIGUOUS"] and a.size == npl_total
    keep["iceTmask"] = a
    s.iceTmask = _ptr(a, C.c_int32)
    s.dt, s.dyn_area_min, s.dyn_mass_min = float(d["dt"]), float(d["dyn_area_min"]), float(d["dyn_mass_min"])
    s.gravit, s.ssh_stress = float(d.get("gravit", 9.80616)), int(d.get("ssh_stress", SSH_GEOSTROPHIC))
    return s, keep


def make_fields_partial(f, npl_total):
    """Fields with only the arrays present in `f` set (evp_b200_step_resident reads what its flags ask for)."""
    s, keep = Fields(), {}
    for n in FIELDS_ORDER:
        if n in f:
            a = f[n]
            assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"] and a.size == npl_total, n
            keep[n] = a
            setattr(s, n, _ptr(a, C.c_double))
    for n in FIELDS_MASK:
        if n in f:
            a = f[n]
            assert a.dtype == np.int32 and a.flags["C_CONTIGUOUS"] and a.size == npl_total, n
            keep[n] = a
            setattr(s, n, _ptr(a, C.c_int32))
    return s, keep
