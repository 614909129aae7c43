"""Block decomposition: the data contract of ice_blocks.F90 / ice_domain.F90 that the path honours.

`create_blocks` restates cicecore/cicedyn/infrastructure/ice_blocks.F90:121-330 (block table,
ghost-inclusive global index vectors, padded last blocks); `cartesian_owner` restates the
'cartesian' distribution of cicecore/shared/ice_distribution.F90 (blocks dealt to ranks as a
2-D processor grid) in the only form the GPU path needs: every rank owns a rectangle of blocks.

Arrays use the memory layout of the Fortran originals: a field `a(nx_block,ny_block,max_blocks)`
is a C-ordered numpy array of shape (max_blocks, ny_block, nx_block).
"""
from dataclasses import dataclass

import numpy as np

from . import abi

NGHOST = 1


@dataclass
class Blocks:
    nx_global: int
    ny_global: int
    block_size_x: int
    block_size_y: int
    ew: int
    ns: int
    nx_block: int
    ny_block: int
    nblocks_x: int
    nblocks_y: int
    # per global block n (0-based, n = jblock*nblocks_x + iblock), Fortran 1-based local indices
    ilo: np.ndarray
    ihi: np.ndarray
    jlo: np.ndarray
    jhi: np.ndarray
    iblock: np.ndarray
    jblock: np.ndarray
    i_glob: np.ndarray  # (nblocks_tot, nx_block)
    j_glob: np.ndarray  # (nblocks_tot, ny_block)

    @property
    def nblocks_tot(self):
        return self.nblocks_x * self.nblocks_y


def create_blocks(nx_global, ny_global, block_size_x, block_size_y, ew="cyclic", ns="closed"):
    """ice_blocks.F90:121-330 (create_blocks)."""
    ew_i = abi.BNDY_NAMES[ew] if isinstance(ew, str) else ew
    ns_i = abi.BNDY_NAMES[ns] if isinstance(ns, str) else ns
    ng = NGHOST
    nx_block = block_size_x + 2 * ng
    ny_block = block_size_y + 2 * ng
    nbx = (nx_global - 1) // block_size_x + 1
    nby = (ny_global - 1) // block_size_y + 1
    ntot = nbx * nby
    ilo = np.full(ntot, ng + 1, np.int32)
    jlo = np.full(ntot, ng + 1, np.int32)
    ihi = np.full(ntot, nx_block - ng, np.int32)
    jhi = np.full(ntot, ny_block - ng, np.int32)
    ibk = np.zeros(ntot, np.int32)
    jbk = np.zeros(ntot, np.int32)
    i_glob = np.zeros((ntot, nx_block), np.int32)
    j_glob = np.zeros((ntot, ny_block), np.int32)
    n = 0
    for jb in range(1, nby + 1):
        js = (jb - 1) * block_size_y + 1
        for ib in range(1, nbx + 1):
            is_ = (ib - 1) * block_size_x + 1
            ibk[n], jbk[n] = ib, jb
            for j in range(1, ny_block + 1):
                gj = js - ng + j - 1
                if gj < 1 and ns_i == abi.BNDY_CYCLIC:
                    gj += ny_global
                if gj > ny_global + ng:
                    gj = 0  # padding
                elif gj > ny_global:
                    if ns_i == abi.BNDY_CYCLIC:
                        gj -= ny_global
                    elif ns_i == abi.BNDY_TRIPOLE:
                        gj = -gj
                elif gj == ny_global and jlo[n] <= j < jhi[n]:
                    jhi[n] = j
                j_glob[n, j - 1] = gj
            for i in range(1, nx_block + 1):
                gi = is_ - ng + i - 1
                if gi < 1 and ew_i == abi.BNDY_CYCLIC:
                    gi += nx_global
                if gi > nx_global + ng:
                    gi = 0
                elif gi > nx_global:
                    if ew_i == abi.BNDY_CYCLIC:
                        gi -= nx_global
                elif gi == nx_global and ilo[n] <= i < ihi[n]:
                    ihi[n] = i
                i_glob[n, i - 1] = gi
            n += 1
    return Blocks(nx_global, ny_global, block_size_x, block_size_y, ew_i, ns_i, nx_block, ny_block, nbx, nby,
                  ilo, ihi, jlo, jhi, ibk, jbk, i_glob, j_glob)


def proc_grid(nranks, nblocks_x, nblocks_y):
    """Minimum-perimeter 2-D processor grid with more ranks along i (the reference's
    'square-ice' advice, doc/source/user_guide/ug_implementation.rst:760-768)."""
    best = None
    for py in range(1, nranks + 1):
        if nranks % py:
            continue
        px = nranks // py
        if nblocks_x % px or nblocks_y % py:
            continue
        key = (abs(px - py), -px)
        if px >= py and (best is None or key < best[0]):
            best = (key, px, py)
    if best is None:
        for py in range(1, nranks + 1):
            if nranks % py == 0 and nblocks_x % (nranks // py) == 0 and nblocks_y % py == 0:
                return nranks // py, py
        raise ValueError(f"cannot split {nblocks_x}x{nblocks_y} blocks over {nranks} ranks as rectangles")
    return best[1], best[2]


def cartesian_owner(blocks, nranks):
    """rank owning each global block; every rank gets a (nblocks_x/px) x (nblocks_y/py) rectangle."""
    px, py = proc_grid(nranks, blocks.nblocks_x, blocks.nblocks_y)
    bx, by = blocks.nblocks_x // px, blocks.nblocks_y // py
    owner = np.zeros(blocks.nblocks_tot, np.int32)
    for n in range(blocks.nblocks_tot):
        owner[n] = ((blocks.jblock[n] - 1) // by) * px + (blocks.iblock[n] - 1) // bx
    return owner, (px, py)
