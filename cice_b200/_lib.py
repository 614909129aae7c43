"""Loader of the CUDA library behind the C ABI.  Fails loudly: there is no CPU fallback."""
import ctypes as C
import os

from . import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_build", "libevp_b200.so")
_lib = None

# every symbol include/evp_b200.h declares
SYMBOLS = (
    "evp_b200_get_unique_id", "evp_b200_comm_init", "evp_b200_set_device", "evp_b200_init", "evp_b200_finalize",
    "evp_b200_last_error", "evp_b200_run_bgrid", "evp_b200_upload", "evp_b200_subcycle", "evp_b200_download",
    "evp_b200_last_loop_ms", "evp_b200_last_launches", "evp_b200_stream", "evp_b200_describe",
    "evp_b200_halo_plan", "evp_b200_dom_pitch", "evp_b200_dom_cells", "evp_b200_p2p_plan", "evp_b200_stress_fold_plan", "evp_b200_init_cgrid", "evp_b200_run_cgrid", "evp_b200_deformations", "evp_b200_dyn_finish", "evp_b200_set_metric", "evp_b200_pin_host", "evp_b200_unpin_host",
    "evp_b200_run_bgrid_resident", "evp_b200_download_stress", "evp_b200_stress_symmetrise", "evp_b200_allow_partial_domain", "evp_b200_run_cdgrid",
    "evp_b200_prep_init", "evp_b200_step_resident",
)


class EvpB200Error(RuntimeError):
    """non-zero return from the C ABI; the reference-side shim maps this to abort_ice()."""


def _preload_nccl():
    """libevp_b200.so needs libnccl.so.2.  Inside a Python process that also uses torch, the NCCL that
    torch bundles must be the one bound to that soname (torch's libtorch_cuda needs symbols newer than
    the system NCCL), so load it first; a Fortran host simply links the system NCCL."""
    import sys
    for d in sys.path:
        cand = os.path.join(d, "nvidia", "nccl", "lib", "libnccl.so.2")
        if os.path.exists(cand):
            try:
                C.CDLL(cand, mode=C.RTLD_GLOBAL)
                return cand
            except OSError:
                pass
    return None


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise EvpB200Error(f"{LIB_PATH} is missing: build it with `python -m cice_b200.build` "
                           "(the EVP path has no CPU fallback)")
    _preload_nccl()
    L = C.CDLL(LIB_PATH)
    pg, pp, pf = C.POINTER(abi.Grid), C.POINTER(abi.Params), C.POINTER(abi.Fields)
    L.evp_b200_get_unique_id.argtypes = [C.c_void_p]
    L.evp_b200_comm_init.argtypes = [C.c_int32, C.c_int32, C.c_void_p]
    L.evp_b200_set_device.argtypes = [C.c_int32]
    L.evp_b200_allow_partial_domain.argtypes = [C.c_int32]
    L.evp_b200_init.argtypes = [pg]
    L.evp_b200_finalize.argtypes = []
    L.evp_b200_run_bgrid.argtypes = [pp, pf]
    L.evp_b200_init_cgrid.argtypes = [C.POINTER(abi.CGrid)]
    L.evp_b200_run_cgrid.argtypes = [pp, C.POINTER(abi.CFields)]
    L.evp_b200_run_cdgrid.argtypes = [pp, C.POINTER(abi.CDFields)]
    L.evp_b200_deformations.argtypes = [C.POINTER(abi.Deform)]
    L.evp_b200_dyn_finish.argtypes = [C.POINTER(abi.Finish)]
    L.evp_b200_pin_host.argtypes = [C.c_void_p, C.c_size_t]
    L.evp_b200_unpin_host.argtypes = [C.c_void_p]
    L.evp_b200_set_metric.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_double, C.POINTER(C.c_int32)]
    L.evp_b200_upload.argtypes = [pf]
    L.evp_b200_run_bgrid_resident.argtypes = [pp, pf, C.c_int32]
    L.evp_b200_download_stress.argtypes = [pf]
    L.evp_b200_subcycle.argtypes = [pp]
    L.evp_b200_download.argtypes = [pf]
    L.evp_b200_last_loop_ms.argtypes = [C.POINTER(C.c_double)]
    L.evp_b200_last_launches.argtypes = [C.POINTER(C.c_int64)]
    L.evp_b200_stream.argtypes = [C.POINTER(C.c_void_p)]
    pi32 = C.POINTER(C.c_int32)
    L.evp_b200_halo_plan.argtypes = [C.c_int32, pi32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, pi32, pi32, C.c_int32]
    L.evp_b200_dom_pitch.argtypes = [C.c_int32]
    L.evp_b200_dom_cells.argtypes = [C.c_int32, C.c_int32]
    L.evp_b200_stress_fold_plan.argtypes = [C.c_int32, pi32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, pi32, pi32, pi32, pi32, C.c_int32]
    L.evp_b200_p2p_plan.argtypes = [C.c_int32, pi32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, pi32, pi32, pi32, pi32, C.c_int32]
    for n in SYMBOLS:
        getattr(L, n).restype = C.c_int
    L.evp_b200_dom_cells.restype = C.c_int64
    L.evp_b200_last_error.restype = C.c_char_p
    L.evp_b200_describe.restype = C.c_char_p
    _lib = L
    return L


def check(rc, what=""):
    if rc != 0:
        raise EvpB200Error(f"{what}: {load().evp_b200_last_error().decode(errors='replace')}")
