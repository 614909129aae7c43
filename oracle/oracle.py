"""ctypes binding of the C oracle (oracle/evp_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this module;
the product package (cice_b200) never does.  Pinned to vectors generated from the reference's source text
(tests/golden/ref_translit.py), not to a compiled reference binary -- see the header of evp_oracle.h.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from cice_b200 import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
_BUILD = os.path.join(_HERE, "_build")
_libs = {}


def build(force=False):
    """compile the oracle with the committed Makefile (gcc only; no reference build system)."""
    need = force or not all(os.path.exists(os.path.join(_BUILD, f"liboracle_{v}.so")) for v in ("exact", "fast"))
    if not need:
        src_m = max(os.path.getmtime(os.path.join(_HERE, f)) for f in ("evp_oracle.c", "evp_oracle_cgrid.c", "evp_oracle.h"))
        src_m = max(src_m, os.path.getmtime(os.path.join(os.path.dirname(_HERE), "include", "evp_b200.h")))
        need = any(os.path.getmtime(os.path.join(_BUILD, f"liboracle_{v}.so")) < src_m for v in ("exact", "fast"))
    if need:
        subprocess.run(["make", "-C", _HERE, "-B"], check=True, capture_output=True)


def _cpu_has_avx2_fma():
    try:
        with open("/proc/cpuinfo") as fh:
            flags = fh.read()
        return " avx2" in flags and " fma" in flags
    except OSError:
        return False


def lib(variant="exact"):
    """variant 'exact' (-O2 -ffp-contract=off: defines the answer) or 'fast' (timed CPU baseline)."""
    if variant == "fast" and not _cpu_has_avx2_fma():
        variant = "exact"
    if variant not in _libs:
        path = os.path.join(_BUILD, f"liboracle_{variant}.so")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        L.orc_evp_run_bgrid.argtypes = [C.POINTER(abi.Grid), C.POINTER(abi.Params), C.POINTER(abi.Fields), C.c_int]
        L.orc_evp_run_bgrid.restype = C.c_int
        L.orc_evp_run_bgrid_1d.argtypes = [C.POINTER(abi.Grid), C.POINTER(C.c_double), C.POINTER(C.c_double),
                                           C.c_double, C.POINTER(abi.Params), C.POINTER(abi.Fields), C.c_int]
        L.orc_evp_run_bgrid_1d.restype = C.c_int
        L.orc_evp_run_cgrid.argtypes = [C.POINTER(abi.Grid), C.POINTER(abi.CGrid), C.POINTER(abi.Params), C.POINTER(abi.CFields), C.c_int]
        L.orc_evp_run_cdgrid.argtypes = [C.POINTER(abi.Grid), C.POINTER(abi.CGrid), C.POINTER(abi.Params), C.POINTER(abi.CDFields), C.c_int]
        L.orc_evp_run_cgrid.restype = C.c_int
        L.orc_deformations.argtypes = [C.POINTER(abi.Grid), C.POINTER(C.c_int32), C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(abi.Deform)]
        L.orc_deformations.restype = C.c_int
        L.orc_dyn_finish.argtypes = [C.POINTER(abi.Grid), C.POINTER(C.c_int32)] + [C.POINTER(C.c_double)] * 7 + [C.POINTER(abi.Finish)]
        L.orc_dyn_finish.restype = C.c_int
        L.orc_halo_update.argtypes = [C.POINTER(abi.Grid), C.POINTER(C.POINTER(C.c_double)), C.c_int, C.c_int, C.c_int]
        L.orc_halo_update.restype = C.c_int
        L.orc_stress_symmetrise.argtypes = [C.POINTER(abi.Grid), C.POINTER(C.POINTER(C.c_double))]
        L.orc_stress_symmetrise.restype = C.c_int
        L.orc_last_error.restype = C.c_char_p
        L.orc_num_threads.restype = C.c_int
        _libs[variant] = L
    return _libs[variant]


def num_threads():
    return lib().orc_num_threads()


def _npl(grid):
    return int(grid["nx_block"]) * int(grid["ny_block"]) * int(grid["max_blocks"])


def evp_run_bgrid(grid, params, fields, nthreads=0, variant="exact"):
    """run the 2-D blocked oracle IN PLACE on the arrays of `fields`."""
    L = lib(variant)
    g, kg = abi.make_grid(grid)
    p = abi.make_params(params)
    f, kf = abi.make_fields(fields, _npl(grid))
    rc = L.orc_evp_run_bgrid(C.byref(g), C.byref(p), C.byref(f), int(nthreads))
    if rc:
        raise RuntimeError("oracle: " + L.orc_last_error().decode())
    return fields


def evp_run_bgrid_1d(grid, HTE, HTN, deltaminEVP, params, fields, nthreads=0, variant="exact"):
    L = lib(variant)
    g, kg = abi.make_grid(grid)
    p = abi.make_params(params)
    f, kf = abi.make_fields(fields, _npl(grid))
    hte = abi.as_f64(HTE)
    htn = abi.as_f64(HTN)
    rc = L.orc_evp_run_bgrid_1d(C.byref(g), hte.ctypes.data_as(C.POINTER(C.c_double)),
                                htn.ctypes.data_as(C.POINTER(C.c_double)), float(deltaminEVP),
                                C.byref(p), C.byref(f), int(nthreads))
    if rc:
        raise RuntimeError("oracle1d: " + L.orc_last_error().decode())
    return fields


def evp_run_cgrid(grid, cgrid, params, cfields, nthreads=0, variant="exact"):
    """run the C-grid oracle IN PLACE on the arrays of `cfields`."""
    L = lib(variant)
    g, kg = abi.make_grid(grid)
    cg, kcg = abi.make_cgrid(cgrid, _npl(grid))
    p = abi.make_params(params)
    f, kf = abi.make_cfields(cfields, _npl(grid))
    rc = L.orc_evp_run_cgrid(C.byref(g), C.byref(cg), C.byref(p), C.byref(f), int(nthreads))
    if rc:
        raise RuntimeError("oracle cgrid failed")
    return cfields


def evp_run_cdgrid(grid, cgrid, params, cdfields, nthreads=0, variant="exact"):
    """grid_ice = 'CD' (ice_dyn_evp.F90:1123-1275), in place."""
    L = lib(variant)
    g, kg = abi.make_grid(grid)
    cg, kc = abi.make_cgrid(cgrid, _npl(grid))
    p = abi.make_params(params)
    f, kf = abi.make_cdfields(cdfields, _npl(grid))
    rc = L.orc_evp_run_cdgrid(C.byref(g), C.byref(cg), C.byref(p), C.byref(f), nthreads)
    if rc:
        raise RuntimeError("oracle cd grid failed")
    return cdfields


def deformations(grid, iceTmask, uvel, vvel, d, e_factor, variant="exact"):
    L = lib(variant)
    g, kg = abi.make_grid(grid)
    s, keep = abi.make_deform(d, _npl(grid), e_factor)
    pd, pi = C.POINTER(C.c_double), C.POINTER(C.c_int32)
    rc = L.orc_deformations(C.byref(g), iceTmask.ctypes.data_as(pi), uvel.ctypes.data_as(pd), vvel.ctypes.data_as(pd), C.byref(s))
    if rc:
        raise RuntimeError("oracle deformations: " + L.orc_last_error().decode())
    return d


def dyn_finish(grid, fields, d, rhow, cosw, sinw, variant="exact"):
    """dyn_finish (ice_dyn_shared.F90:1291-1365) from the loop's final velocities and U-point inputs in `fields`;
    d["strocnxU"], d["strocnyU"] are inout."""
    L = lib(variant)
    g, kg = abi.make_grid(grid)
    s, keep = abi.make_finish(d, _npl(grid), rhow, cosw, sinw)
    pd, pi = C.POINTER(C.c_double), C.POINTER(C.c_int32)
    arr = [np.ascontiguousarray(fields[n], dtype=np.float64) for n in ("uvel", "vvel", "cdn_ocnU", "uocnU", "vocnU", "aiU", "fmU")]
    mask = np.ascontiguousarray(fields["iceUmask"], dtype=np.int32)
    rc = L.orc_dyn_finish(C.byref(g), mask.ctypes.data_as(pi), *[a.ctypes.data_as(pd) for a in arr], C.byref(s))
    if rc:
        raise RuntimeError("oracle dyn_finish: " + L.orc_last_error().decode())
    return d


def stress_symmetrise(grid, fields, variant="exact"):
    """tripole grids: force symmetry of the 12 stress arrays across the fold (ice_dyn_evp.F90:1321-1388), in place"""
    L = lib(variant)
    g, kg = abi.make_grid(grid)
    ptrs = (C.POINTER(C.c_double) * 12)(*[fields[n].ctypes.data_as(C.POINTER(C.c_double)) for n in abi.STRESS])
    if L.orc_stress_symmetrise(C.byref(g), ptrs):
        raise RuntimeError("oracle stress symmetrise: " + L.orc_last_error().decode())


def halo_update(grid, arrays, field_loc=1, field_type=1, variant="exact"):
    L = lib(variant)
    g, kg = abi.make_grid(grid)
    n = len(arrays)
    ptrs = (C.POINTER(C.c_double) * n)(*[a.ctypes.data_as(C.POINTER(C.c_double)) for a in arrays])
    rc = L.orc_halo_update(C.byref(g), ptrs, n, field_loc, field_type)
    if rc:
        raise RuntimeError("oracle halo: " + L.orc_last_error().decode())
