import numpy as np

from cice_b200 import abi, synth


def run_oracle(oracle, case, nthreads=0, variant="exact"):
    f = case.copy_fields()
    oracle.evp_run_bgrid(case.grid, case.params, f, nthreads=nthreads, variant=variant)
    return f


def run_gpu(dyn_evp, case, **param_over):
    f = case.copy_fields()
    p = dict(case.params)
    p.update(param_over)
    dyn_evp.dyn_evp_b200_init(case.grid)
    try:
        dyn_evp.dyn_evp_b200_run(p, f)
    finally:
        dyn_evp.dyn_evp_b200_finalize()
    return f


def assert_bitwise(a, b, names=abi.FIELDS_INOUT):
    bad = []
    for n in names:
        x, y = a[n], b[n]
        if not np.array_equal(x.view(np.int64), y.view(np.int64)):
            # +0 / -0 are distinct bit patterns but equal values: report them separately
            d = np.abs(x - y)
            bad.append(f"{n}: {np.count_nonzero(x.view(np.int64) != y.view(np.int64))} cells differ, max|d|={np.nanmax(d):.3e}")
    assert not bad, "not bit-identical:\n  " + "\n  ".join(bad)


def rel_err(a, b, n):
    """norm-relative error of BASELINE.md section 3: maxabs(gpu-ref)/max(maxabs(ref), tiny)."""
    ref = np.abs(b[n]).max()
    return np.abs(a[n] - b[n]).max() / max(ref, 1e-300)


def assert_close(a, b, tol=1e-10, names=abi.FIELDS_INOUT):
    bad = []
    for n in names:
        if np.abs(b[n]).max() == 0.0:
            if np.abs(a[n]).max() != 0.0:
                bad.append(f"{n}: reference is identically zero, got {np.abs(a[n]).max():.3e}")
            continue
        e = rel_err(a, b, n)
        if not e <= tol:
            bad.append(f"{n}: rel err {e:.3e} > {tol:g}")
    assert not bad, "outside tolerance:\n  " + "\n  ".join(bad)
