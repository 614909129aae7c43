"""N>1 host-side logic on CPU: two real processes over gloo execute the library's halo plan (the
host-only planning hook evp_b200_halo_plan, the very enumeration the GPU exchange is built from) on
numpy sub-domains and must reproduce the oracle's halo update of the undecomposed domain -- cyclic,
closed and tripole (fold, top-row symmetrisation, pole points).  Also covers the rank_view /
cartesian_owner partition each rank derives its blocks from."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _rects(case, owner, world):
    rects = []
    b = case.blocks
    for r in range(world):
        ids = np.nonzero(owner == r)[0]
        gi0 = min(b.i_glob[n][b.ilo[n] - 1] for n in ids)
        gi1 = max(b.i_glob[n][b.ihi[n] - 1] for n in ids)
        gj0 = min(b.j_glob[n][b.jlo[n] - 1] for n in ids)
        gj1 = max(b.j_glob[n][b.jhi[n] - 1] for n in ids)
        rects.append([gi0, gj0, gi1 - gi0 + 1, gj1 - gj0 + 1])
    return np.array(rects, np.int32)


def _worker(rank, world, port, ew, ns, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from cice_b200 import abi, decomp, dyn_evp, synth
    from oracle import oracle

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        kmt = "boxislands" if ns == "tripole" else "none"  # land on the fold: zeros whose sign the fold must keep
        case = synth.make_case("tiny", block_size=(12, 10), seed=21, ew=ew, ns=ns, kmt=kmt)
        owner, _ = decomp.cartesian_owner(case.blocks, world)
        rects = _rects(case, owner, world)
        ewi, nsi = case.grid["ew_boundary_type"], case.grid["ns_boundary_type"]
        nxg, nyg = case.grid["nx_global"], case.grid["ny_global"]

        # truth: the oracle's halo update on the undecomposed domain
        whole = synth.make_case("tiny", seed=21, ew=ew, ns=ns, kmt=kmt)
        tu, tv = whole.fields["uvel"].copy(), whole.fields["vvel"].copy()
        tu[0, 0, :] = tu[0, -1, :] = np.nan  # ghost rows/cols must come from the update, not from the input
        tu[0, :, 0] = tu[0, :, -1] = np.nan
        tv[0, 0, :] = tv[0, -1, :] = np.nan
        tv[0, :, 0] = tv[0, :, -1] = np.nan
        oracle.halo_update(whole.grid, [tu, tv], field_loc=1, field_type=1)

        # my sub-domain (ny+2, ld), interior from the case, ring unknown
        gi0, gj0, nx, ny = rects[rank]
        ld = dyn_evp.dom_pitch(nx)
        X = whole.X
        dom = {}
        for name in ("uvel", "vvel"):
            a = np.full((ny + 2, ld), np.nan)
            a[1:ny + 1, 1:nx + 1] = X[name][gj0:gj0 + ny, gi0:gi0 + nx]
            dom[name] = a.reshape(-1)
        wrap_ew = ewi == abi.BNDY_CYCLIC and nx == nxg
        wrap_ns = nsi == abi.BNDY_CYCLIC and ny == nyg
        for name in ("uvel", "vvel"):  # what the compute kernels' wrap stores do
            a = dom[name].reshape(ny + 2, ld)
            if wrap_ew:
                a[1:ny + 1, 0], a[1:ny + 1, nx + 1] = a[1:ny + 1, nx].copy(), a[1:ny + 1, 1].copy()
            if wrap_ns:
                a[0, 1:nx + 1], a[ny + 1, 1:nx + 1] = a[ny, 1:nx + 1].copy(), a[1, 1:nx + 1].copy()
            if wrap_ew and wrap_ns:
                a[0, 0], a[0, nx + 1], a[ny + 1, 0], a[ny + 1, nx + 1] = a[ny, nx], a[ny, 1], a[1, nx], a[1, 1]

        plans = [dyn_evp.halo_plan(rects, r, nxg, nyg, ewi, nsi) for r in range(world)]
        # what every rank needs from me, in that rank's entry order (the send lists of the GPU plan)
        outbox = {}
        for r in range(world):
            if r == rank:
                continue
            idx = []
            for e in plans[r]:
                if e[1] == rank:
                    idx.append(e[2])
                if e[3] == rank:
                    idx.append(e[4])
            outbox[r] = np.array([[dom["uvel"][c], dom["vvel"][c]] for c in idx], dtype=np.float64).reshape(-1, 2)
        boxes = [None] * world
        dist.all_gather_object(boxes, outbox)
        cursor = {r: 0 for r in range(world)}

        def take(r, c):
            if r == rank:
                return dom["uvel"][c], dom["vvel"][c]
            v = boxes[r][rank][cursor[r]]
            cursor[r] += 1
            return v[0], v[1]

        staged = []
        for e in plans[rank]:  # all sources are read before any destination is written
            a = take(e[1], e[2])
            b = take(e[3], e[4]) if e[3] >= 0 else None
            staged.append((e[0], e[5], a, b))
        for dst, op, a, b in staged:
            if op == 0:
                u, v = a
            elif op == 1:
                u, v = -a[0], -a[1]
            elif op == 2:
                u, v = 0.5 * (a[0] - b[0]), 0.5 * (a[1] - b[1])
            else:
                u, v = -(0.5 * (a[0] - b[0])), -(0.5 * (a[1] - b[1]))
            dom["uvel"][dst], dom["vvel"][dst] = u, v

        bad = 0
        for name, truth in (("uvel", tu[0]), ("vvel", tv[0])):
            a = dom[name].reshape(ny + 2, ld)
            for dj in range(ny + 2):
                for di in range(nx + 2):
                    gi, gj = gi0 + di - 1, gj0 + dj - 1  # unwrapped global position
                    if ewi == abi.BNDY_CYCLIC:
                        gi = (gi - 1) % nxg + 1
                    if nsi == abi.BNDY_CYCLIC:
                        gj = (gj - 1) % nyg + 1
                    if 0 <= gi <= nxg + 1 and 0 <= gj <= nyg + 1:
                        want = truth[gj, gi]
                        got = a[dj, di]
                        if np.isnan(want):
                            ok = np.isnan(got)  # outside a closed edge: untouched on both sides
                        else:
                            ok = (got == want) and (np.signbit(got) == np.signbit(want))
                        bad += 0 if ok else 1
        q.put((rank, bad, len(plans[rank])))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("ew,ns", [("cyclic", "closed"), ("cyclic", "cyclic"), ("closed", "closed"), ("cyclic", "tripole")])
@pytest.mark.parametrize("world", [2, 4])
def test_halo_plan_two_processes_gloo(ew, ns, world):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() * 7 + hash((ew, ns, world))) % 2000
    procs = [ctx.Process(target=_worker, args=(r, world, port, ew, ns, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, bad, n in res:
        assert bad == 0, f"rank {rank}: {bad} cells differ from the oracle halo ({n} plan entries)"


def _worker_push(rank, world, port, ew, ns, cfg, bs, q):
    """the in-kernel NVLink form of the halo on numpy arrays: every rank executes ITS push list (stores into other ranks' arrays,
    here delivered through gloo), then its fold list (all sources read before any destination is written)."""
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from cice_b200 import abi, decomp, dyn_evp, synth
    from oracle import oracle

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        kmt = "boxislands" if ns == "tripole" else "none"
        case = synth.make_case(cfg, block_size=bs, seed=21, ew=ew, ns=ns, kmt=kmt)
        owner, _ = decomp.cartesian_owner(case.blocks, world)
        rects = _rects(case, owner, world)
        ewi, nsi = case.grid["ew_boundary_type"], case.grid["ns_boundary_type"]
        nxg, nyg = case.grid["nx_global"], case.grid["ny_global"]
        whole = synth.make_case(cfg, seed=21, ew=ew, ns=ns, kmt=kmt)
        tu, tv = whole.fields["uvel"].copy(), whole.fields["vvel"].copy()
        for a in (tu, tv):
            a[0, 0, :] = a[0, -1, :] = np.nan
            a[0, :, 0] = a[0, :, -1] = np.nan
        oracle.halo_update(whole.grid, [tu, tv], field_loc=1, field_type=1)

        gi0, gj0, nx, ny = rects[rank]
        ld = dyn_evp.dom_pitch(nx)
        assert dyn_evp.dom_cells(nx, ny) == ld * (ny + 4)
        X = whole.X
        dom = {}
        for name in ("uvel", "vvel"):
            a = np.full((ny + 4, ld), np.nan)      # ghost ring + the two staging rows
            a[1:ny + 1, 1:nx + 1] = X[name][gj0:gj0 + ny, gi0:gi0 + nx]
            dom[name] = a.reshape(-1)
        wrap_ew = ewi == abi.BNDY_CYCLIC and nx == nxg
        wrap_ns = nsi == abi.BNDY_CYCLIC and ny == nyg
        for name in ("uvel", "vvel"):
            a = dom[name].reshape(ny + 4, ld)
            if wrap_ew:
                a[1:ny + 1, 0], a[1:ny + 1, nx + 1] = a[1:ny + 1, nx].copy(), a[1:ny + 1, 1].copy()
            if wrap_ns:
                a[0, 1:nx + 1], a[ny + 1, 1:nx + 1] = a[ny, 1:nx + 1].copy(), a[1, 1:nx + 1].copy()
            if wrap_ew and wrap_ns:
                a[0, 0], a[0, nx + 1], a[ny + 1, 0], a[ny + 1, nx + 1] = a[ny, nx], a[ny, 1], a[1, nx], a[1, 1]

        push, fold = dyn_evp.p2p_plan(rects, rank, nxg, nyg, ewi, nsi)
        tripole_top = nsi == abi.BNDY_TRIPOLE and gj0 + ny - 1 == nyg
        outbox = {r: [] for r in range(world)}
        for src, dr, dst, neg in push:
            i, j = src % ld, src // ld
            # the kernel's push condition: boundary points, and row ny-1 below a tripole fold
            assert i in (1, nx) or j in (1, ny) or (tripole_top and j == ny - 1), (i, j)
            assert dr != rank
            sgn = -1.0 if neg else 1.0
            outbox[int(dr)].append((int(dst), sgn * dom["uvel"][src], sgn * dom["vvel"][src]))
        boxes = [None] * world
        dist.all_gather_object(boxes, outbox)
        for r in range(world):
            for dst, u, v in boxes[r][rank]:
                dom["uvel"][dst], dom["vvel"][dst] = u, v
        staged = []
        for dst, c1, c2, op in fold:
            staged.append((dst, op, (dom["uvel"][c1], dom["vvel"][c1]), (dom["uvel"][c2], dom["vvel"][c2])))
        for dst, op, a, b in staged:
            if op == 0:
                u, v = a
            elif op == 1:
                u, v = -a[0], -a[1]
            elif op == 2:
                u, v = 0.5 * (a[0] - b[0]), 0.5 * (a[1] - b[1])
            else:
                u, v = -(0.5 * (a[0] - b[0])), -(0.5 * (a[1] - b[1]))
            dom["uvel"][dst], dom["vvel"][dst] = u, v

        bad = 0
        for name, truth in (("uvel", tu[0]), ("vvel", tv[0])):
            a = dom[name].reshape(ny + 4, ld)
            for dj in range(ny + 2):
                for di in range(nx + 2):
                    gi, gj = gi0 + di - 1, gj0 + dj - 1
                    if ewi == abi.BNDY_CYCLIC:
                        gi = (gi - 1) % nxg + 1
                    if nsi == abi.BNDY_CYCLIC:
                        gj = (gj - 1) % nyg + 1
                    if 0 <= gi <= nxg + 1 and 0 <= gj <= nyg + 1:
                        want, got = truth[gj, gi], a[dj, di]
                        ok = np.isnan(got) if np.isnan(want) else ((got == want) and (np.signbit(got) == np.signbit(want)))
                        bad += 0 if ok else 1
        q.put((rank, bad, len(push) + len(fold)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("ew,ns,cfg,bs", [("cyclic", "closed", "tiny", (12, 10)), ("cyclic", "cyclic", "tiny", (12, 10)),
                                          ("cyclic", "tripole", "tiny", (12, 10)), ("cyclic", "tripole", "tiny", (6, 5)),
                                          ("cyclic", "tripole", "tx1", (90, 60))],
                         ids=["cyclic-closed", "cyclic-cyclic", "tripole", "tripole-small-blocks", "tx1-tripole"])
@pytest.mark.parametrize("world", [2, 4])
def test_push_and_fold_plan_gloo(ew, ns, cfg, bs, world):
    """the halo without a staged exchange (in-kernel NVLink stores + fold kernel, configs[3]): push lists and fold lists of
    evp_b200_p2p_plan executed by 2 and 4 gloo processes reproduce the oracle's halo update, signed zeros included."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() * 7 + hash((ew, ns, cfg, bs, world))) % 2000
    procs = [ctx.Process(target=_worker_push, args=(r, world, port, ew, ns, cfg, bs, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, bad, n in res:
        assert bad == 0, f"rank {rank}: {bad} cells differ from the oracle halo ({n} plan entries)"
        assert n > 0


def test_halo_plan_single_rank_is_empty_without_tripole():
    from cice_b200 import abi, dyn_evp
    for ew, ns in ((abi.BNDY_CYCLIC, abi.BNDY_CLOSED), (abi.BNDY_CYCLIC, abi.BNDY_CYCLIC), (abi.BNDY_CLOSED, abi.BNDY_OPEN)):
        assert len(dyn_evp.halo_plan([[1, 1, 24, 20]], 0, 24, 20, ew, ns)) == 0
    pl = dyn_evp.halo_plan([[1, 1, 24, 20]], 0, 24, 20, abi.BNDY_CYCLIC, abi.BNDY_TRIPOLE)
    # ghost row (26 cells incl. corners, negated copies) + top row (24 interior + 2 ghost columns):
    # two pole columns (i = 12, 24; the ghost column i=0 aliases 24) are negated, the rest symmetrised
    assert len(pl) == 52
    assert (pl[:, 5] == 1).sum() == 26 + 3 and ((pl[:, 5] == 2) | (pl[:, 5] == 3)).sum() == 23 and (pl[:, 5] == 0).sum() == 0


def test_halo_plan_with_eliminated_land_blocks():
    """rectangles that do not cover the domain (land-block elimination): a ghost cell whose source no rank owns has no
    plan entry (it keeps the zeros of the host's halo fill), every other ghost cell is still served."""
    from cice_b200 import abi, dyn_evp
    nxg, nyg = 24, 20
    # rank 0 owns the west half above row 5 (rows 1..4 were an all-land cap), rank 1 the east half in full
    rects = [[1, 5, 12, 16], [13, 1, 12, 20]]
    p0 = dyn_evp.halo_plan(rects, 0, nxg, nyg, abi.BNDY_CYCLIC, abi.BNDY_CLOSED)
    p1 = dyn_evp.halo_plan(rects, 1, nxg, nyg, abi.BNDY_CYCLIC, abi.BNDY_CLOSED)
    ld = dyn_evp.dom_pitch(12)
    # rank 0 (global rows 5..20): E ghost column (global i = 13) and, cyclically, W ghost column (global i = 24) come from
    # rank 1 for its 16 interior rows plus the S ghost row (global row 4, which rank 1 owns); the N ghost row is outside
    assert all(e[1] == 1 and e[5] == 0 for e in p0)
    assert sorted(set(int(e[0]) // ld for e in p0)) == list(range(0, 17)) and len(p0) == 2 * 17
    # rank 1 (global rows 1..20): its ghost columns have a source only where rank 0 exists, global rows 5..20
    assert all(e[1] == 0 and e[5] == 0 for e in p1)
    assert sorted(set(int(e[0]) // ld for e in p1)) == list(range(5, 21)) and len(p1) == 2 * 16


def _worker_stress_fold(rank, world, port, q):
    """the stress symmetrisation across a tripole fold between ranks on numpy arrays: every rank of the top row swaps the top physical
    row of its twelve stress arrays with the ranks evp_b200_stress_fold_plan names (here through gloo, on the GPU one NCCL group) and
    mirrors the assembled row into its north ghost row cell by cell as the plan says."""
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from cice_b200 import abi, decomp, dyn_evp, synth
    from oracle import oracle

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        case = synth.make_case("tiny", block_size=(12, 10), seed=23, ew="cyclic", ns="tripole", kmt="none")
        owner, _ = decomp.cartesian_owner(case.blocks, world)
        rects = _rects(case, owner, world)
        nxg, nyg = case.grid["nx_global"], case.grid["ny_global"]
        whole = synth.make_case("tiny", seed=23, ew="cyclic", ns="tripole", kmt="none")
        rng = np.random.default_rng(3)
        G = {n: rng.normal(size=(nyg + 2, nxg + 2)) for n in abi.STRESS}          # the same on every rank
        truth = {n: G[n][None].copy() for n in abi.STRESS}
        oracle.stress_symmetrise(whole.grid, truth)

        gi0, gj0, nx, ny = (int(v) for v in rects[rank])
        dom = {n: G[n][gj0 - 1:gj0 + ny + 1, gi0 - 1:gi0 + nx + 1].copy() for n in abi.STRESS}   # (ny+2, nx+2) with its ring
        seg, cell = dyn_evp.stress_fold_plan(rects, rank, nxg, nyg, whole.grid["ns_boundary_type"])
        top = gj0 + ny - 1 == nyg
        ntop = sum(1 for r in rects if r[1] + r[3] - 1 == nyg)
        assert len(cell) == (nx + 2 if top else 0) and len(seg) == (ntop - 1 if top else 0)
        rows = {rank: np.stack([dom[n][ny, 1:nx + 1] for n in abi.STRESS])}        # [12][nx]
        reqs, bufs = [], {}
        for t, ti0, tnx in seg:                                                    # swap with every other rank of the top row
            bufs[int(t)] = torch.zeros(12, int(tnx), dtype=torch.float64)
            reqs.append(dist.isend(torch.from_numpy(rows[rank].copy()), dst=int(t)))
            reqs.append(dist.irecv(bufs[int(t)], src=int(t)))
        for r in reqs:
            r.wait()
        for t, b in bufs.items():
            rows[t] = b.numpy()
        partner = lambda k: (k // 4) * 4 + (k % 4 + 2) % 4
        for k, n in enumerate(abi.STRESS):
            for i, src, col in cell:
                dom[n][ny + 1, i] = rows[int(src)][partner(k), col - 1] if src >= 0 else 0.0
        bad = 0
        for n in abi.STRESS:
            want = truth[n][0][gj0 - 1:gj0 + ny + 1, gi0 - 1:gi0 + nx + 1]
            bad += int(np.count_nonzero(dom[n].view(np.int64) != want.view(np.int64)))
        q.put((rank, bad, len(seg), len(cell)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_stress_fold_plan_processes_gloo(world):
    """SURVEY 8f rank 2 between ranks: executing the host-side plan of the stress symmetrisation on numpy sub-domains over gloo gives the
    oracle's result (ice_dyn_evp.F90:1321-1388) on every rank, ghost corners included; ranks below the top row have nothing to do."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() * 11 + 17 * world) % 2000
    procs = [ctx.Process(target=_worker_stress_fold, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, bad, nseg, ncell in res:
        assert bad == 0, f"rank {rank}: {bad} cells differ from the oracle ({nseg} segments, {ncell} ghost cells)"
    assert sum(1 for r in res if r[3] > 0) == 2    # 2x1 and 2x2 processor grids: two ranks hold the top row
