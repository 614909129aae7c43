"""Multi-GPU parity check, one process per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tests/mgpu_check.py [config] [bsx] [bsy] [ndte] [kernel] [ns]

Every rank owns a rectangle of blocks of one synthetic case, runs the EVP loop through the C ABI with the
NCCL halo exchange, and rank 0 compares the gathered block arrays bit for bit with the CPU oracle run on
the undecomposed set of blocks.  Prints MGPU PASS / MGPU FAIL."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    from cice_b200 import abi, decomp, dyn_evp, synth

    cfg = sys.argv[1] if len(sys.argv) > 1 else "gx3"
    bsx = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    bsy = int(sys.argv[3]) if len(sys.argv) > 3 else 29
    ndte = int(sys.argv[4]) if len(sys.argv) > 4 else 20
    kernel = sys.argv[5] if len(sys.argv) > 5 else "fused"
    ns = sys.argv[6] if len(sys.argv) > 6 and sys.argv[6] not in ("-", "none") else None
    elim = len(sys.argv) > 7 and sys.argv[7] == "elim"   # land-block elimination: all-land blocks are owned by nobody
    world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dyn_evp.set_device(local)
    ids = [dyn_evp.get_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    dyn_evp.comm_init(rank, world, ids[0])

    # (the step-resident mode starts from rest with no ice mask, like the first step of a run: set S1, no random state on top)
    case = synth.make_case(cfg, block_size=(bsx, bsy), seed=None if kernel == "step" else 31, ndte=ndte, ns=ns,
                           kmt=(None if cfg == "tx1" else "none") if ns == "tripole" else ("continents" if elim else None))
    owner, pg = decomp.cartesian_owner(case.blocks, world)
    if elim:
        for n in range(case.blocks.nblocks_tot):
            if not case.fields["iceTmask"][n][1:-1, 1:-1].any() and not case.fields["iceUmask"][n].any():
                owner[n] = -1
    g, f, bids = case.rank_view(owner, rank)
    step = kernel == "step"   # the step preparation on the device, state resident (evp_b200_step_resident), two consecutive steps
    resident = kernel == "resident"   # stresses resident over two steps (evp_b200_run_bgrid_resident); on a tripole grid with the
    #                                   symmetrisation across the fold on the device, the top-row segments exchanged between ranks
    p = dict(case.params, mode=abi.MODE_EXACT, kernel=abi.KERNEL_NAMES["auto" if step or resident else kernel])
    dyn_evp.dyn_evp_b200_init(g)
    desc = dyn_evp.describe()
    if step:
        static, prep = synth.step_inputs(case)
        sel = lambda dct: {k: (np.ascontiguousarray(v[bids]) if isinstance(v, np.ndarray) else v) for k, v in dct.items()}
        static, prep = sel(static), sel(prep)
        f = {n: (f[n] if n in abi.STRESS else np.zeros_like(f["uvel"])) for n in abi.FIELDS_INOUT}
        f["iceUmask"] = np.zeros(f["uvel"].shape, dtype=np.int32)
        dyn_evp.dyn_evp_b200_prep_init(static)
        dyn_evp.dyn_evp_b200_step_resident(p, prep, f, init_state=True)
        dyn_evp.dyn_evp_b200_step_resident(p, prep, f, fetch_diag=True, fetch_state=True)
    elif resident:
        dyn_evp.dyn_evp_b200_run_resident(p, f, keep_stress=True, fetch_stress=False)
        dyn_evp.dyn_evp_b200_run_resident(p, f, keep_stress=True, fetch_stress=True)
    else:
        dyn_evp.dyn_evp_b200_run(p, f)
    nl = dyn_evp.last_launches()
    dyn_evp.dyn_evp_b200_finalize()

    out = [None] * world
    dist.gather_object((bids, {n: f[n] for n in abi.FIELDS_INOUT}, desc, nl), out if rank == 0 else None, dst=0)
    ok = True
    if rank == 0:
        from oracle import oracle
        ref = case.copy_fields()
        oracle.evp_run_bgrid(case.grid, case.params, ref)   # all blocks: elimination does not change the kept ones (test_oracle.py)
        if resident:   # after each loop the stresses are symmetrised across the fold (ice_dyn_evp.F90:1321-1388; no-op without a fold)
            oracle.stress_symmetrise(case.grid, ref)
            oracle.evp_run_bgrid(case.grid, case.params, ref)
            oracle.stress_symmetrise(case.grid, ref)
        if step:   # second step: dyn_prep2 clears taubx/tauby, everything else carries over
            for n in ("taubxU", "taubyU"):
                ref[n][...] = 0.0
            oracle.evp_run_bgrid(case.grid, case.params, ref)
        nbad = 0
        for bids_r, fr, desc_r, nl_r in out:
            for n in abi.FIELDS_INOUT:
                a, b = fr[n], ref[n][bids_r]
                if not np.array_equal(a.view(np.int64), b.view(np.int64)):
                    nbad += 1
                    print(f"  differs: {n} on blocks {list(bids_r)}: {np.count_nonzero(a != b)} cells, max {np.nanmax(np.abs(a - b)):.3e}")
        ok = nbad == 0
        print(f"[{cfg} {bsx}x{bsy} ndte={ndte} kernel={kernel} ns={ns} procs={pg}] launches/rank={out[0][3]}  {out[0][2]}")
        for o in out:
            print("   ", o[2])
        print("MGPU PASS" if ok else "MGPU FAIL")
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
