"""The C++ host mirror of the reference seam (cice_b200/host/dyn_evp_b200.hpp), driven by a compiled
caller that holds the arrays exactly as the Fortran driver would (tests/host_caller.cpp)."""
import os
import subprocess

import numpy as np
import pytest

from cice_b200 import abi, synth
from tests.util import run_oracle, assert_bitwise

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def write_case(path, c, params):
    g = c.grid
    with open(path, "wb") as fh:
        np.array([g["nx_block"], g["ny_block"], g["nblocks"], g["max_blocks"], g["nx_global"], g["ny_global"],
                  g["ew_boundary_type"], g["ns_boundary_type"], params["ndte"], params["mode"], params["kernel"]], np.int32).tofile(fh)
        np.array([params[k] for k in ("arlx1i", "denom1", "revp", "brlx", "e_factor", "epp2i", "capping", "Ktens",
                                      "u0", "cosw", "sinw", "rhow")], np.float64).tofile(fh)
        for k in ("ilo", "ihi", "jlo", "jhi", "i_glob", "j_glob"):
            np.ascontiguousarray(g[k], np.int32).tofile(fh)
        for k in abi.GRID_STATIC:
            np.ascontiguousarray(g[k], np.float64).tofile(fh)
        for k in abi.FIELDS_ORDER:
            np.ascontiguousarray(c.fields[k], np.float64).tofile(fh)
        for k in abi.FIELDS_MASK:
            np.ascontiguousarray(c.fields[k], np.int32).tofile(fh)


def test_host_layer_builds_and_links():
    from cice_b200 import build
    exe = build.build_host()
    assert os.path.exists(exe) and os.path.exists(build.HOST_LIB)
    syms = subprocess.run(["nm", "-DC", build.HOST_LIB], capture_output=True, text=True).stdout
    for name in ("cice_b200::dyn_evp_b200_init", "cice_b200::dyn_evp_b200_run", "cice_b200::dyn_evp_b200_finalize"):
        assert name in syms, name


def test_host_caller_aborts_without_gpu(tmp_path):
    """error convention: no exit() inside the library; the caller turns the non-zero return into the
    reference's abort (exit status 128 like MPI_ABORT(comm,128), comm/mpi/ice_exit.F90)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from cice_b200 import build
    exe = build.build_host()
    c = synth.make_case("tiny")
    write_case(tmp_path / "case.bin", c, c.params)
    r = subprocess.run([exe, str(tmp_path / "case.bin"), str(tmp_path / "out.bin")], capture_output=True, text=True)
    assert r.returncode == 128 and "abort_ice: (dyn_evp_b200_init) ERROR" in r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("kw", [dict(config="gx3", ndte=15, seed=3), dict(config="gx3", ndte=10, block_size=(25, 29), max_blocks=20)],
                         ids=["1block", "16blocks-maxblocks20"])
def test_host_caller_matches_oracle(oracle_mod, tmp_path, kw):
    from cice_b200 import build
    exe = build.build_host()
    c = synth.make_case(**kw)
    ref = run_oracle(oracle_mod, c)
    write_case(tmp_path / "case.bin", c, dict(c.params, mode=abi.MODE_EXACT))
    r = subprocess.run([exe, str(tmp_path / "case.bin"), str(tmp_path / "out.bin")], capture_output=True, text=True)
    assert r.returncode == 0 and "host_caller ok" in r.stdout, r.stdout + r.stderr
    out = np.fromfile(tmp_path / "out.bin", np.float64).reshape(30, *c.fields["uvel"].shape)
    got = {n: out[k] for k, n in enumerate(abi.FIELDS_ORDER)}
    assert_bitwise(got, ref)
