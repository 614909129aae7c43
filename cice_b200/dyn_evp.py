"""Host-side mirror of the reference seam for the EVP path, over the C ABI (include/evp_b200.h).

The reference dispatches its own 1-D solver with three calls (ice_dyn_evp1d.F90:25):

    dyn_evp1d_init()            ice_dyn_evp.F90:153-155
    dyn_evp1d_run(31 arrays)    ice_dyn_evp.F90:848-856
    dyn_evp1d_finalize()

and this module offers the same three with the same argument meaning, so the parity tests read
like a caller of the reference would:

    dyn_evp_b200_init(grid)               static block table + geometry, once
    dyn_evp_b200_run(params, fields)      one dynamics step: the whole ndte subcycle loop, in place
    dyn_evp_b200_finalize()

`fields` is a dict of numpy arrays in the reference's memory layout: Fortran
`a(nx_block,ny_block,max_blocks)` == C-ordered (max_blocks, ny_block, nx_block).  Errors raise
EvpB200Error (the Fortran shim calls abort_ice instead).  There is no CPU fallback.
"""
import ctypes as C

import numpy as np

from . import abi
from ._lib import load, check, EvpB200Error  # noqa: F401

_state = {"grid": None, "keep": None, "npl": 0}


def set_device(ordinal):
    check(load().evp_b200_set_device(int(ordinal)), "evp_b200_set_device")


def allow_partial_domain(yes=True):
    """a single rank whose blocks span less than the domain because land blocks were eliminated (see include/evp_b200.h)."""
    check(load().evp_b200_allow_partial_domain(1 if yes else 0), "evp_b200_allow_partial_domain")


def get_unique_id():
    buf = C.create_string_buffer(abi.UNIQUE_ID_BYTES)
    check(load().evp_b200_get_unique_id(buf), "evp_b200_get_unique_id")
    return buf.raw


def comm_init(rank, nranks, unique_id):
    buf = C.create_string_buffer(bytes(unique_id), abi.UNIQUE_ID_BYTES)
    check(load().evp_b200_comm_init(int(rank), int(nranks), buf), "evp_b200_comm_init")


def dyn_evp_b200_init(grid):
    """grid: dict as built by cice_b200.synth (block table + static geometry of THIS rank)."""
    g, keep = abi.make_grid(grid)
    check(load().evp_b200_init(C.byref(g)), "evp_b200_init")
    _state.update(grid=g, keep=keep, npl=int(grid["nx_block"]) * int(grid["ny_block"]) * int(grid["max_blocks"]))


def _fields(fields):
    if _state["grid"] is None:
        raise EvpB200Error("dyn_evp_b200_init has not been called")
    return abi.make_fields(fields, _state["npl"])


def dyn_evp_b200_run(params, fields):
    """the hot path: host arrays in, host arrays out (inout arrays are updated in place)."""
    p = abi.make_params(params)
    f, keep = _fields(fields)
    check(load().evp_b200_run_bgrid(C.byref(p), C.byref(f)), "evp_b200_run_bgrid")
    return fields


def dyn_evp_b200_run_resident(params, fields, keep_stress=True, fetch_stress=False):
    """one dynamics step with the stresses resident on the device: the host's 12 stress arrays are
    neither read nor written (unless fetch_stress), dyn_prep2's zeroing off the ice is applied on the device."""
    p = abi.make_params(params)
    f, keep = _fields(fields)
    flags = (abi.KEEP_STRESS if keep_stress else 0) | (abi.FETCH_STRESS if fetch_stress else 0)
    check(load().evp_b200_run_bgrid_resident(C.byref(p), C.byref(f), flags), "evp_b200_run_bgrid_resident")
    return fields


def dyn_evp_b200_prep_init(static):
    """static inputs of the device-side step preparation: hm, tarea, uarea, fcor, umask (after dyn_evp_b200_init)."""
    st, keep = abi.make_prep_static(static, _state["npl"])
    check(load().evp_b200_prep_init(C.byref(st)), "evp_b200_prep_init")
    _state["pkeep"] = keep


def dyn_evp_b200_step_resident(params, prep, fields, init_state=False, fetch_diag=False, fetch_state=False):
    """one dynamics step with the preparation on the device (T -> U averages, dyn_prep2) and the whole dynamics state resident:
    `prep` holds the T-point inputs of the step, `fields` receives uvel, vvel (and what the flags ask for)."""
    p = abi.make_params(params)
    pr, k1 = abi.make_prep(prep, _state["npl"])
    f, k2 = abi.make_fields_partial(fields, _state["npl"])
    flags = (abi.STEP_INIT_STATE if init_state else 0) | (abi.STEP_FETCH_DIAG if fetch_diag else 0) | (abi.STEP_FETCH_STATE if fetch_state else 0)
    check(load().evp_b200_step_resident(C.byref(p), C.byref(pr), C.byref(f), flags), "evp_b200_step_resident")
    return fields


def bind(entry, params, fields, prep=None, **flags):
    """A zero-argument callable for one of the per-step entry points with its C structs built ONCE: what a compiled caller does
    (the Fortran shim fills its structs once per run).  entry: "run" (evp_b200_run_bgrid), "run_resident"
    (evp_b200_run_bgrid_resident, keep_stress / fetch_stress) or "step_resident" (evp_b200_step_resident, init_state / fetch_diag /
    fetch_state).  The arrays must stay alive and in place; timing loops use this so that they time the C ABI, not ctypes marshalling."""
    L = load()
    p = abi.make_params(params)
    if entry == "run":
        f, keep = _fields(fields)
        fn, args, what = L.evp_b200_run_bgrid, (C.byref(p), C.byref(f)), "evp_b200_run_bgrid"
    elif entry == "run_resident":
        f, keep = _fields(fields)
        fl = (abi.KEEP_STRESS if flags.get("keep_stress", True) else 0) | (abi.FETCH_STRESS if flags.get("fetch_stress") else 0)
        fn, args, what = L.evp_b200_run_bgrid_resident, (C.byref(p), C.byref(f), fl), "evp_b200_run_bgrid_resident"
    elif entry == "step_resident":
        pr, k1 = abi.make_prep(prep, _state["npl"])
        f, keep = abi.make_fields_partial(fields, _state["npl"])
        keep = (keep, k1, pr)
        fl = ((abi.STEP_INIT_STATE if flags.get("init_state") else 0) | (abi.STEP_FETCH_DIAG if flags.get("fetch_diag") else 0)
              | (abi.STEP_FETCH_STATE if flags.get("fetch_state") else 0))
        fn, args, what = L.evp_b200_step_resident, (C.byref(p), C.byref(pr), C.byref(f), fl), "evp_b200_step_resident"
    else:
        raise ValueError(entry)

    def call(_keep=(p, f, keep)):
        check(fn(*args), what)
    return call


def stress_symmetrise():
    """tripole grids: force the symmetry of the device-resident stresses across the fold (ice_dyn_evp.F90:1321-1388)"""
    check(load().evp_b200_stress_symmetrise(), "evp_b200_stress_symmetrise")


def download_stress(fields):
    """fetch the device-resident stresses into the host arrays (restart / history)."""
    f, keep = _fields(fields)
    check(load().evp_b200_download_stress(C.byref(f)), "evp_b200_download_stress")
    return fields


def dyn_evp_b200_init_cgrid(cgrid):
    """extra static geometry of grid_ice='C' (after dyn_evp_b200_init)."""
    if _state["grid"] is None:
        raise EvpB200Error("evp_b200_init_cgrid: call evp_b200_init first")
    cg, keep = abi.make_cgrid(cgrid, _state["npl"])
    check(load().evp_b200_init_cgrid(C.byref(cg)), "evp_b200_init_cgrid")
    _state["ckeep"] = keep


def dyn_evp_b200_run_cgrid(params, cfields):
    """the C-grid subcycle loop (ice_dyn_evp.F90:936-1101): host arrays in, host arrays out."""
    p = abi.make_params(params)
    f, keep = abi.make_cfields(cfields, _state["npl"])
    check(load().evp_b200_run_cgrid(C.byref(p), C.byref(f)), "evp_b200_run_cgrid")
    return cfields


def dyn_evp_b200_run_cdgrid(params, cdfields):
    """the CD-grid subcycle loop (ice_dyn_evp.F90:1123-1275): host arrays in, host arrays out (after dyn_evp_b200_init_cgrid)."""
    p = abi.make_params(params)
    f, keep = abi.make_cdfields(cdfields, _state["npl"])
    check(load().evp_b200_run_cdgrid(C.byref(p), C.byref(f)), "evp_b200_run_cdgrid")
    return cdfields


def deformations(d, e_factor):
    """`deformations` (ice_dyn_shared.F90:1756-1860) from the velocities the last loop left on the device;
    d: dict with dxU, dyU, tarear (in) and divu, shear, vort, rdg_conv, rdg_shear (inout, in place)."""
    s, keep = abi.make_deform(d, _state["npl"], e_factor)
    check(load().evp_b200_deformations(C.byref(s)), "evp_b200_deformations")
    return d


def dyn_finish(d, rhow, cosw=1.0, sinw=0.0):
    """`dyn_finish` (ice_dyn_shared.F90:1291-1365) from the velocities and U-point inputs the last loop left on the device;
    d["strocnxU"], d["strocnyU"] are inout block arrays (points off the U list keep their values)."""
    s, keep = abi.make_finish(d, _state["npl"], rhow, cosw, sinw)
    check(load().evp_b200_dyn_finish(C.byref(s)), "evp_b200_dyn_finish")
    return d


def pin_host(a):
    """page-lock a numpy array that is passed to the library every step (evp_b200_pin_host)."""
    check(load().evp_b200_pin_host(a.ctypes.data, a.nbytes), "evp_b200_pin_host")


def unpin_host(a):
    check(load().evp_b200_unpin_host(a.ctypes.data), "evp_b200_unpin_host")


def set_metric(HTN, HTE, deltaminEVP):
    """hand HTN, HTE (block arrays) to the library; returns the number of T cells on which the reference's expressions do NOT
    reproduce dxhy, dyhx, cxp, cyp, cxm, cym, DminTarea bit for bit (0 = the derived-geometry kernels, variants 59/63, may be used)."""
    a = np.ascontiguousarray(HTN, dtype=np.float64)
    b = np.ascontiguousarray(HTE, dtype=np.float64)
    assert a.size == _state["npl"] and b.size == _state["npl"]
    bad = C.c_int32(-1)
    pd = C.POINTER(C.c_double)
    check(load().evp_b200_set_metric(a.ctypes.data_as(pd), b.ctypes.data_as(pd), float(deltaminEVP), C.byref(bad)), "evp_b200_set_metric")
    return int(bad.value)


def upload(fields):
    f, keep = _fields(fields)
    check(load().evp_b200_upload(C.byref(f)), "evp_b200_upload")


def subcycle(params):
    p = abi.make_params(params)
    check(load().evp_b200_subcycle(C.byref(p)), "evp_b200_subcycle")


def download(fields):
    f, keep = _fields(fields)
    check(load().evp_b200_download(C.byref(f)), "evp_b200_download")
    return fields


def last_loop_ms():
    v = C.c_double()
    check(load().evp_b200_last_loop_ms(C.byref(v)), "evp_b200_last_loop_ms")
    return v.value


def last_launches():
    v = C.c_int64()
    check(load().evp_b200_last_launches(C.byref(v)), "evp_b200_last_launches")
    return v.value


def stream_handle():
    v = C.c_void_p()
    check(load().evp_b200_stream(C.byref(v)), "evp_b200_stream")
    return v.value


def halo_plan(rects, rank, nx_global, ny_global, ew, ns):
    """host-only: the (uvel,vvel) halo plan of `rank` as an (n,6) int array
    {dst, src1_rank, src1, src2_rank, src2, op}; rects is (nranks,4) {gi0, gj0, nx, ny}."""
    L = load()
    r = np.ascontiguousarray(rects, dtype=np.int32)
    n = C.c_int32()
    pi = C.POINTER(C.c_int32)
    check(L.evp_b200_halo_plan(len(r), r.ctypes.data_as(pi), rank, nx_global, ny_global, ew, ns, C.byref(n), None, 0), "halo_plan")
    out = np.zeros((max(n.value, 1), 6), np.int32)
    check(L.evp_b200_halo_plan(len(r), r.ctypes.data_as(pi), rank, nx_global, ny_global, ew, ns, C.byref(n),
                               out.ctypes.data_as(pi), n.value), "halo_plan")
    return out[:n.value]


def p2p_plan(rects, rank, nx_global, ny_global, ew, ns):
    """host-only: the halo as the in-kernel NVLink form serves it.  Returns (push, fold): push (n,4) {src, dst rank, dst cell, negate},
    fold (m,4) {dst, src1, src2, op}; cells index the owning rank's sub-domain array, staging rows included (dom_cells)."""
    L = load()
    r = np.ascontiguousarray(rects, dtype=np.int32)
    npush, nfold = C.c_int32(), C.c_int32()
    pi = C.POINTER(C.c_int32)
    args = (len(r), r.ctypes.data_as(pi), rank, nx_global, ny_global, ew, ns)
    check(L.evp_b200_p2p_plan(*args, C.byref(npush), None, C.byref(nfold), None, 0), "p2p_plan")
    cap = max(npush.value, nfold.value, 1)
    push, fold = np.zeros((cap, 4), np.int32), np.zeros((cap, 4), np.int32)
    check(L.evp_b200_p2p_plan(*args, C.byref(npush), push.ctypes.data_as(pi), C.byref(nfold), fold.ctypes.data_as(pi), cap), "p2p_plan")
    return push[:npush.value], fold[:nfold.value]


def stress_fold_plan(rects, rank, nx_global, ny_global, ns):
    """host-only: the stress symmetrisation across a tripole fold between ranks.  Returns (seg, cell): seg (n,3) {rank, gi0, nx} of the
    other top-row ranks, cell (m,3) {ghost column, source rank or -1, source column} for the rank's north ghost row."""
    L = load()
    r = np.ascontiguousarray(rects, dtype=np.int32)
    nseg, ncell = C.c_int32(), C.c_int32()
    pi = C.POINTER(C.c_int32)
    args = (len(r), r.ctypes.data_as(pi), rank, nx_global, ny_global, ns)
    check(L.evp_b200_stress_fold_plan(*args, C.byref(nseg), None, C.byref(ncell), None, 0), "stress_fold_plan")
    cap = max(nseg.value, ncell.value, 1)
    seg, cell = np.zeros((cap, 3), np.int32), np.zeros((cap, 3), np.int32)
    check(L.evp_b200_stress_fold_plan(*args, C.byref(nseg), seg.ctypes.data_as(pi), C.byref(ncell), cell.ctypes.data_as(pi), cap), "stress_fold_plan")
    return seg[:nseg.value], cell[:ncell.value]


def dom_pitch(nx):
    return load().evp_b200_dom_pitch(int(nx))


def dom_cells(nx, ny):
    return int(load().evp_b200_dom_cells(int(nx), int(ny)))


def describe():
    return load().evp_b200_describe().decode()


def dyn_evp_b200_finalize():
    check(load().evp_b200_finalize(), "evp_b200_finalize")
    _state.update(grid=None, keep=None, npl=0)
