"""CPU-side checks of the boundary: the C-ABI library builds for sm_100a, loads, exports every symbol
include/evp_b200.h declares, and fails loudly (no CPU fallback) when there is no GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from cice_b200 import abi, decomp, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "evp_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(evp_b200_[a-z_0-9]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol(evp_lib):
    from cice_b200 import _lib
    L = _lib.load()
    syms = header_symbols()
    assert len(syms) >= 14
    for s in syms:
        assert hasattr(L, s), f"{s} declared in include/evp_b200.h but not exported"
    assert sorted(_lib.SYMBOLS) == syms


def test_struct_layouts_match_header():
    """ctypes mirrors must have the C layout: 10 int32 + 6 + 10 pointers; 4 int32 + 12 doubles; 30 + 2 pointers."""
    assert C.sizeof(abi.Grid) == 10 * 4 + 16 * 8
    assert C.sizeof(abi.Params) == 4 * 4 + 13 * 8
    assert C.sizeof(abi.Fields) == 32 * 8
    assert C.sizeof(abi.CGrid) == 20 * 8
    assert C.sizeof(abi.CFields) == 47 * 8
    assert C.sizeof(abi.CDFields) == 58 * 8
    assert C.sizeof(abi.Deform) == 9 * 8
    assert C.sizeof(abi.Finish) == 5 * 8
    txt = open(os.path.join(ROOT, "include", "evp_b200.h")).read()
    body = txt[txt.index("typedef struct {", txt.index("Time-varying fields of one call")):txt.index("} evp_b200_fields_t;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = re.findall(r"\*\s*([A-Za-z_0-9]+)", body)
    assert tuple(names) == abi.FIELDS_ORDER + abi.FIELDS_MASK
    for start, end, want in (("Extra static geometry of grid_ice", "} evp_b200_cgrid_t;", abi.CGRID_STATIC),
                             ("Time-varying fields of one C-grid call", "} evp_b200_cfields_t;", abi.CFIELDS_ORDER + abi.CFIELDS_MASK),
                             ("CD grid (SURVEY 8a row a13", "} evp_b200_cdfields_t;", abi.CDFIELDS_ORDER + abi.CFIELDS_MASK)):
        body = txt[txt.index("typedef struct {", txt.index(start)):txt.index(end)]
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        assert tuple(re.findall(r"\*\s*([A-Za-z_0-9]+)", body)) == want, start


def test_no_cpu_fallback_without_gpu(evp_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    c = synth.make_case("tiny")
    with pytest.raises(evp_lib.EvpB200Error):
        evp_lib.dyn_evp_b200_init(c.grid)
    with pytest.raises(evp_lib.EvpB200Error):
        evp_lib.dyn_evp_b200_run(c.params, c.copy_fields())


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "cice_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                for pat in (r"^\s*(from|import)\s+oracle", r"liboracle", r"oracle[/\\]", r"orc_[a-z_]+\s*\("):
                    assert not re.search(pat, txt, flags=re.M), (dp, f, pat)


# ---- block decomposition data contract (ice_blocks.F90:121-330) -------------------------------------
def test_create_blocks_gx3_25x29():
    b = decomp.create_blocks(100, 116, 25, 29, "cyclic", "closed")
    assert (b.nblocks_x, b.nblocks_y, b.nx_block, b.ny_block) == (4, 4, 27, 31)
    assert list(b.i_glob[0][:3]) == [100, 1, 2] and b.i_glob[3][-1] == 1  # cyclic wrap
    assert b.j_glob[0][0] == 0 and b.j_glob[12][-1] == 117  # closed: counted past the edge
    assert (b.ihi == 26).all() and (b.jhi == 30).all()


def test_create_blocks_padding():
    b = decomp.create_blocks(24, 20, 7, 9, "cyclic", "closed")
    assert (b.nblocks_x, b.nblocks_y) == (4, 3)
    last = b.nblocks_x - 1
    assert b.ihi[last] == 4  # 24 = 3*7 + 3 -> three interior columns in the last block
    assert b.i_glob[last][b.ihi[last] - 1] == 24 and b.i_glob[last][b.ihi[last] + 1] == 0  # padding is 0
    top = (b.nblocks_y - 1) * b.nblocks_x
    assert b.j_glob[top][b.jhi[top] - 1] == 20


def test_create_blocks_tripole_ghost_rows_negative():
    b = decomp.create_blocks(40, 30, 20, 15, "cyclic", "tripole")
    assert b.j_glob[-1][-1] == -31


@pytest.mark.parametrize("n,exp", [(1, (1, 1)), (2, (2, 1)), (4, (2, 2)), (8, (4, 2))])
def test_proc_grid(n, exp):
    assert decomp.proc_grid(n, 8, 8) == exp


def test_scatter_gather_roundtrip():
    c = synth.make_case("tiny", block_size=(7, 9), seed=1)
    g = synth.gather(c.fields["uvel"], c.blocks)
    assert np.array_equal(g, c.X["uvel"][1:-1, 1:-1])


def test_rank_view_partitions_blocks():
    c = synth.make_case("gx3", block_size=(25, 29))
    owner, (px, py) = decomp.cartesian_owner(c.blocks, 4)
    assert (px, py) == (2, 2)
    seen = np.zeros(c.blocks.nblocks_tot, int)
    for r in range(4):
        g, f, ids = c.rank_view(owner, r)
        seen[ids] += 1
        assert g["nblocks"] == 4 and f["uvel"].shape[0] == 4
    assert (seen == 1).all()


def test_fortran_shim_types_match_the_header():
    """fortran/ice_dyn_evp_b200.F90 cannot be compiled here (no Fortran compiler), so at least its bind(C) derived types must
    list the members of the C structs in include/evp_b200.h in the same order (a swapped pair would be silent corruption)."""
    txt = open(os.path.join(ROOT, "fortran", "ice_dyn_evp_b200.F90")).read()

    def members(tname):
        m = re.search(rf"type,\s*bind\(C\)\s*::\s*{tname}\b(.*?)end type {tname}", txt, flags=re.S | re.I)
        assert m, tname
        names = []
        for line in m.group(1).splitlines():
            line = line.split("!")[0]
            if "::" in line:
                names += [n.strip() for n in line.split("::", 1)[1].split(",") if n.strip()]
        return tuple(names)

    assert members("evp_b200_fields_t") == abi.FIELDS_ORDER + abi.FIELDS_MASK
    assert members("evp_b200_cgrid_t") == abi.CGRID_STATIC
    assert members("evp_b200_cfields_t") == abi.CFIELDS_ORDER + abi.CFIELDS_MASK
    assert members("evp_b200_grid_t") == tuple(n for n, _ in abi.Grid._fields_)
    assert members("evp_b200_params_t") == tuple(n for n, _ in abi.Params._fields_)
    assert members("evp_b200_deform_t") == tuple(n for n, _ in abi.Deform._fields_)
    assert members("evp_b200_finish_t") == tuple(n for n, _ in abi.Finish._fields_)
    # every C entry point the shim binds exists in the header
    hdr = open(os.path.join(ROOT, "include", "evp_b200.h")).read()
    for name in re.findall(r"bind\(C,\s*name='(\w+)'\)", txt):
        assert re.search(rf"\b{name}\s*\(", hdr), name
