#!/usr/bin/env python
"""bench.py -- EVP grid-cells x subcycles / s (fp64) and the HBM-roofline fraction of the subcycle kernel.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--kernel auto|split|fused|persistent] [--mode fast|exact] [--workload gx1|gx3|p1deg]

A "step" is one dynamics step of the hot path: the whole `do ksub = 1,ndte` loop of
ice_dyn_evp.F90:859-913 over one synthetic box2001 state.
  N = 1 : configs[1] of BASELINE.json -- gx1 320x384 B grid, ndte = 240, one block, one GPU.
  N > 1 : weak scaling -- every GPU owns one gx1-sized sub-domain (320x384) of a (px*320)x(py*384)
          cyclic/closed domain; (uvel,vvel) halo exchanged every subcycle.
`value`   device-resident: fields are uploaded once, the timed region is K subcycle loops
          (CUDA events on the library's stream, per step, L2 flushed between steps).
`e2e`     the same metric through evp_b200_run_bgrid with HOST buffers (pinned), H2D + loop + D2H timed.
`roofline` algorithmic bytes (360 B per cell-subcycle) / measured duration of the
          subcycle kernel vs the measured HBM peak of MEASURED_PEAKS.json.
`cpu_baseline` / `--impl reference`: the CPU restatement of the reference loops (oracle, "port": the
          Fortran reference cannot be built in this image) on the box's host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
os.environ.setdefault("NCCL_DEBUG", "WARN")  # keep stdout to the one JSON line (the image sets NCCL_DEBUG=VERSION)

ALGO_BYTES_PER_CELL_SUBCYCLE = 360.0  # 31 doubles read + 14 written
METRIC = "EVP grid-cells*subcycles/sec at gx1 (fp64)"
UNIT = "cell-subcycles/s"


def measured_traffic(kernel):
    """DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture (profiles/traffic.json)."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
            t = json.load(fh)
        e = t.get(kernel) or t.get("fused")
        return e
    except Exception:
        return None


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu=0):
        self.gpu, self.proc, self.lines = gpu, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=lambda: self.lines.extend(self.proc.stdout), daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        self.t.join(timeout=2)
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            p = [x.strip() for x in ln.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1])); mx.append(float(p[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def build_case(workload, world, rank, ndte=None):
    from cice_b200 import decomp, synth
    if world == 1:
        c = synth.make_case(workload, ndte=ndte)
        return c, c.grid, c.fields, (1, 1)
    base = synth.CONFIGS[workload]
    px, py = decomp.proc_grid(world, world, world)
    nx, ny = base["nx"] * px, base["ny"] * py
    c = synth.make_case(workload, nx=nx, ny=ny, block_size=(base["nx"], base["ny"]), ndte=ndte)
    owner, _ = decomp.cartesian_owner(c.blocks, world)
    g, f, ids = c.rank_view(owner, rank)
    return c, g, f, (px, py)


def pin(fields):
    """numpy views over pinned host memory (torch is only the allocator here)."""
    import torch
    out = {}
    for k, v in fields.items():
        t = torch.from_numpy(v.copy()).pin_memory()
        out[k] = t.numpy()
        out["_keep_" + k] = t
    return out


def cpu_baseline(case, steps, warmup, nthreads=0):
    """time the CPU restatement (oracle, fast build, OpenMP over blocks like ice_dyn_evp.F90:861)."""
    from oracle import oracle
    oracle.build()
    nthreads = nthreads or (os.cpu_count() or 1)
    ts = []
    for it in range(warmup + steps):
        f = case.copy_fields()
        t0 = time.perf_counter()
        oracle.evp_run_bgrid(case.grid, case.params, f, nthreads=nthreads, variant="fast")
        if it >= warmup:
            ts.append(time.perf_counter() - t0)
    return float(np.mean(ts)), nthreads


def run_reference(args):
    """--impl reference: the reference's CPU path (restated; see module docstring) on the host cores, on the
    same global grid as our arm at this GPU count (weak scaling: px*320 x py*384)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from cice_b200 import decomp, synth
    wl = args.workload
    base = synth.CONFIGS[wl]
    world = max(int(os.environ.get("WORLD_SIZE", "1")), args.gpus, 1)
    px, py = decomp.proc_grid(world, world, world)
    nx, ny = base["nx"] * px, base["ny"] * py
    # the reference's own block choice for <=16 PEs at gx1 is 40x48 (configuration/scripts/cice_decomp.csh:93-96)
    bs = (40, 48) if wl == "gx1" else (max(base["nx"] // 8, 8), max(base["ny"] // 8, 8))
    case = synth.make_case(wl, nx=nx, ny=ny, block_size=bs)
    sec, nth = cpu_baseline(case, args.steps, max(args.warmup, 1))
    cells = nx * ny * case.params["ndte"]
    val = cells / sec
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{wl} {base['nx']}x{base['ny']} per GPU, B-grid EVP ndte={case.params['ndte']}, "
                                   f"{px}x{py} GPUs, global {nx}x{ny}, box2001 synthetic",
                       "cpu": f"CPU restatement of the reference loops (oracle port, -O3 AVX2/FMA, OpenMP over {bs[0]}x{bs[1]} blocks); "
                              "the Fortran reference cannot be built in this image"},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": nth, "kind": "port",
                             "sample": f"{args.steps} full dynamics step(s) of the {nx}x{ny} grid, ndte={case.params['ndte']}"},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def run_ours(args):
    import torch
    import torch.distributed as dist
    from cice_b200 import abi, dyn_evp, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dyn_evp.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        ids = [dyn_evp.get_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        dyn_evp.comm_init(rank, world, ids[0])

    case, grid, fields, (px, py) = build_case(args.workload, world, rank)
    base = synth.CONFIGS[args.workload]
    params = dict(case.params, mode=abi.MODE_FAST if args.mode == "fast" else abi.MODE_EXACT,
                  kernel=abi.KERNEL_NAMES[args.kernel])
    ndte = params["ndte"]
    cells_global = base["nx"] * px * base["ny"] * py
    dyn_evp.dyn_evp_b200_init(grid)
    if os.environ.get("EVP_B200_FUSED_VARIANT") in ("59", "63"):
        # derived-geometry kernels (round-2 candidate): hand over HTN, HTE; the library checks them bit for bit on the device
        from cice_b200 import decomp
        sel = slice(None)
        if world > 1:
            owner, _ = decomp.cartesian_owner(case.blocks, world)
            sel = case.rank_view(owner, rank)[2]
        bad = dyn_evp.set_metric(synth.scatter(case.X["HTN"], case.blocks)[sel], synth.scatter(case.X["HTE"], case.blocks)[sel], 1e-11)
        if rank == 0:
            print(f"# set_metric: {bad} cells differ", file=sys.stderr)
    desc = dyn_evp.describe()

    hf = pin(fields)
    stream = torch.cuda.ExternalStream(dyn_evp.stream_handle())
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident loop ------------------------------------------------------------------------
    dyn_evp.upload(hf)
    for _ in range(max(args.warmup, 3)):
        dyn_evp.subcycle(params)
    sampler = ClockSampler(local)
    barrier()
    if rank == 0:
        sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    loop_ms = []
    t_wall0 = time.perf_counter()
    for k in range(args.steps):
        flush.zero_()               # L2 flush between timed iterations (torch stream)
        torch.cuda.synchronize()
        ev[k][0].record(stream)
        dyn_evp.subcycle(params)    # syncs its stream before returning
        ev[k][1].record(stream)
        loop_ms.append(dyn_evp.last_loop_ms())
    barrier()
    t_wall = time.perf_counter() - t_wall0
    clocks = sampler.stop() if rank == 0 else None
    step_ms = [a.elapsed_time(b) for a, b in ev]
    total_ms = float(sum(step_ms))
    launches = dyn_evp.last_launches() * args.steps
    kernel_ms = float(np.mean(loop_ms))  # events inside the library, around the launches only

    # ---- end to end through the C ABI with host buffers ----------------------------------------------
    # (a) every field crosses both ways every step (evp_b200_run_bgrid, the plain drop-in);
    # (b) the 12 carried stress arrays stay on the device (evp_b200_run_bgrid_resident, EVP_B200_KEEP_STRESS): velocities,
    #     the 12 per-step inputs, masks and diagnostics still cross every step.  (b) is what INTEGRATION.md wires up for
    #     every step that does not write a restart/history file; not available on tripole grids.
    e2e_steps = max(2, min(args.steps, 5))
    nblk = int(np.prod(fields["uvel"].shape))

    def time_e2e(fn):
        for _ in range(2):
            fn()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            fn()
        barrier()
        return (time.perf_counter() - t0) / e2e_steps

    e2e_full_s = time_e2e(lambda: dyn_evp.dyn_evp_b200_run(params, hf))
    resident_ok = grid["ns_boundary_type"] != abi.BNDY_NAMES["tripole"]
    if resident_ok:
        e2e_s = time_e2e(lambda: dyn_evp.dyn_evp_b200_run_resident(params, hf, keep_stress=True))
        h2d = 18 * nblk * 8 + 2 * nblk * 4
        d2h = 6 * nblk * 8
        e2e_how = "evp_b200_run_bgrid_resident(EVP_B200_KEEP_STRESS): stresses stay on the device, everything else crosses"
    else:
        e2e_s = e2e_full_s
        h2d = 30 * nblk * 8 + 2 * nblk * 4
        d2h = 18 * nblk * 8
        e2e_how = "evp_b200_run_bgrid: every field crosses both ways (tripole grid)"

    if world > 1:
        t = torch.tensor([total_ms, e2e_s, kernel_ms, e2e_full_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms, e2e_s, kernel_ms, e2e_full_s = (float(x) for x in t.tolist())
    dyn_evp.dyn_evp_b200_finalize()

    if rank == 0:
        value = cells_global * ndte * args.steps / (total_ms * 1e-3)
        peak, peak_src = measured_peak()
        per_gpu_cells = base["nx"] * base["ny"]
        nl_step = launches / args.steps
        traffic = measured_traffic("fused" if args.kernel in ("auto", "fused") else args.kernel)
        # dominant kernel: the subcycle kernel; one launch advances the rank's sub-domain by ndte/launches subcycles
        sub_per_launch = ndte / max(nl_step, 1) if args.kernel != "split" else 0.5
        ach = per_gpu_cells * ndte * ALGO_BYTES_PER_CELL_SUBCYCLE / (kernel_ms * 1e-3) / 1e9
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": {"workload": f"{args.workload} {base['nx']}x{base['ny']} per GPU, B-grid EVP ndte={ndte}, "
                                       f"{px}x{py} GPUs, global {base['nx']*px}x{base['ny']*py}, box2001 synthetic",
                           "kernel": args.kernel, "mode": args.mode, "l2": "flushed between timed steps (256 MiB memset)",
                           "layout": desc},
                "clocks": clocks,
                "e2e": {"value": cells_global * ndte / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "ms_per_step": e2e_s * 1e3, "how": e2e_how},
                "e2e_full_copy": {"value": cells_global * ndte / e2e_full_s, "unit": UNIT, "h2d_bytes_per_step": 30 * nblk * 8 + 2 * nblk * 4,
                                  "d2h_bytes_per_step": 18 * nblk * 8, "ms_per_step": e2e_full_s * 1e3,
                                  "how": "evp_b200_run_bgrid: every field crosses both ways"},
                "gpu_launches": int(launches),
                "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                             "traffic": (traffic or {}).get("dram_bytes_per_launch_cold"), "traffic_detail": traffic,
                             "peak_source": peak_src,
                             "algorithmic_bytes_per_cell_subcycle": ALGO_BYTES_PER_CELL_SUBCYCLE,
                             "kernel_ms_per_step": kernel_ms, "subcycles_per_launch": sub_per_launch},
                "wall_s": t_wall}
        # the active-cell rate beside R (T cells that carry ice; rank 0's sub-domain, the same on every rank up to the
        # synthetic ice edge)
        act = int(np.count_nonzero(np.asarray(fields["iceTmask"])[:, 1:-1, 1:-1]))
        line["active_cells"] = {"icellT_rank0": act, "fraction_rank0": act / float(per_gpu_cells),
                                "active_cell_subcycles_per_s": value * act / float(per_gpu_cells)}
        if world == 1 and not args.no_cpu:
            bs = (40, 48) if args.workload == "gx1" else (max(base["nx"] // 8, 8), max(base["ny"] // 8, 8))
            ccase = synth.make_case(args.workload, block_size=bs)
            ncpu = 12 if args.workload in ("gx1", "gx3", "tx1") else 2   # ~3 s of wall time on 16 threads at gx1 (~50 core-seconds)
            sec, nth = cpu_baseline(ccase, ncpu, 1)
            line["cpu_baseline"] = {"value": per_gpu_cells * ndte / sec, "unit": UNIT, "cores": nth, "kind": "port",
                                    "sample": f"{ncpu} full dynamics steps of {args.workload} (ndte={ndte}), blocks {bs[0]}x{bs[1]}, "
                                              f"{sec:.3f} s each"}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


C_ALGO_BYTES = 504.0  # C grid: 53 doubles read (6 velocities, 4 stresses, strength, 22 geometry/mask, 20 momentum) + 10 written


def run_cgrid(args):
    """configs[2]: gx1 C-grid EVP (standard_2d), ndte = 600, one GPU; same JSON shape, metric at the C grid."""
    import torch
    from cice_b200 import abi, dyn_evp, synth
    ndte = 600
    base = synth.CONFIGS[args.workload]
    cells = base["nx"] * base["ny"]
    if args.impl == "reference":
        from oracle import oracle
        oracle.build()
        c = synth.make_ccase(args.workload, ndte=ndte, block_size=(40, 48))  # OpenMP over blocks like ice_dyn_evp.F90:940
        ts = []
        for it in range(1 + args.steps):
            f = c.copy_fields()
            t0 = time.perf_counter()
            oracle.evp_run_cgrid(c.grid, c.cgrid, c.params, f, nthreads=0, variant="fast")
            if it:
                ts.append(time.perf_counter() - t0)
        sec = float(np.mean(ts))
        val = cells * ndte / sec
        print(json.dumps({"impl": "reference", "metric": METRIC.replace("gx1", "gx1 C-grid"), "value": val, "unit": UNIT, "n_gpus": 1,
                          "steps": args.steps, "warmup": 1, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
                          "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                          "config": {"workload": f"{args.workload} C-grid EVP ndte={ndte}, 40x48 blocks, CPU restatement of the reference loops"},
                          "cpu_baseline": {"value": val, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": f"{args.steps} full steps"},
                          "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}))
        return
    c = synth.make_ccase(args.workload, ndte=ndte)
    torch.cuda.set_device(0)
    dyn_evp.set_device(0)
    dyn_evp.dyn_evp_b200_init(c.grid)
    dyn_evp.dyn_evp_b200_init_cgrid(c.cgrid)
    p = dict(c.params, mode=abi.MODE_FAST if args.mode == "fast" else abi.MODE_EXACT)
    f = pin(c.copy_fields())
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    loop, call = [], []
    for it in range(max(args.warmup, 3) + args.steps):
        flush.zero_()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        dyn_evp.dyn_evp_b200_run_cgrid(p, f)
        call.append(time.perf_counter() - t0)
        loop.append(dyn_evp.last_loop_ms())
    loop, call = loop[-args.steps:], call[-args.steps:]
    ms = float(np.mean(loop))
    peak, peak_src = measured_peak()
    ach = cells * ndte * C_ALGO_BYTES / (ms * 1e-3) / 1e9
    nblk = int(np.prod(c.fields["uvel"].shape))
    print(json.dumps({"metric": METRIC.replace("gx1", "gx1 C-grid"), "value": cells * ndte / (ms * 1e-3), "unit": UNIT, "n_gpus": 1,
                      "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
                      "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                      "config": {"workload": f"{args.workload} {base['nx']}x{base['ny']} C-grid EVP (standard_2d) ndte={ndte}, 1 GPU, box2001 synthetic",
                                 "mode": args.mode, "l2": "flushed between timed steps (256 MiB memset)", "layout": dyn_evp.describe()},
                      "e2e": {"value": cells * ndte / float(np.mean(call)), "unit": UNIT, "h2d_bytes_per_step": 37 * nblk * 8 + 4 * nblk * 4,
                              "d2h_bytes_per_step": 21 * nblk * 8, "ms_per_step": float(np.mean(call)) * 1e3},
                      "gpu_launches": int(dyn_evp.last_launches()) * args.steps,
                      "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": None,
                                   "peak_source": peak_src, "algorithmic_bytes_per_cell_subcycle": C_ALGO_BYTES,
                                   "kernels_per_subcycle": 5}}))
    dyn_evp.dyn_evp_b200_finalize()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--kernel", default="auto", choices=["auto", "split", "fused", "persistent", "queue"])
    ap.add_argument("--mode", default="exact", choices=["fast", "exact"])
    ap.add_argument("--workload", default="gx1")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--grid", default="B", choices=["B", "C"], help="C: configs[2], gx1 C-grid EVP, ndte=600 (one GPU)")
    args = ap.parse_args()
    if args.grid == "C":
        run_cgrid(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
