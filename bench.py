#!/usr/bin/env python
"""bench.py -- EVP grid-cells x subcycles / s (fp64) and the roofline fraction of the subcycle kernel.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload gx1|gx3|tx1|p1deg] [--layout weak|strong] [--sub NXxNY]
                    [--kernel auto|split|fused|stream|resident|persistent|tstream] [--mode exact|fast] [--grid B|C]

A "step" is one dynamics step of the hot path: the whole `do ksub = 1,ndte` loop of
ice_dyn_evp.F90:859-913 over one synthetic box2001 state.
  N = 1 : configs[1] of BASELINE.json -- gx1 320x384 B grid, ndte = 240, one block, one GPU.
  N > 1 : --layout weak (default for gx1/gx3): every GPU owns one sub-domain of --sub cells (default: the workload's own size) of a
          (px*nx) x (py*ny) domain;  --layout strong (default for tx1, p1deg): the workload's global grid is cut into px x py
          rectangles (configs[3]: tx1 360x240 tripole on 2x2; configs[4]: 3600x2400 on 4x2).  (uvel,vvel) halo every subcycle.
`parity`  untimed pre-check, every N: the bench's own decomposition at reduced ndte through the C ABI against the CPU oracle
          (exact build) run by rank 0 on the undecomposed grid; sha256 per field per rank, i.e. bit for bit.
`value`   device-resident: fields are uploaded once, the timed region is K subcycle loops
          (CUDA events on the library's stream, per step, L2 flushed and all ranks aligned before each step).
`e2e`     the same metric through the C ABI with HOST buffers (pinned), H2D + loop + D2H timed;
          `e2e_pageable`: the same call from plain pageable numpy arrays, and after evp_b200_pin_host (what the Fortran shim does).
`roofline` algorithmic bytes (360 B per cell-subcycle) / measured duration of the
          subcycle kernel vs the measured HBM peak of MEASURED_PEAKS.json; beside it the fp64-pipe ceiling that binds at gx1.
`cpu_baseline` / `--impl reference`: the CPU restatement of the reference loops (oracle, "port": the
          Fortran reference cannot be built in this image) on the box's host cores.
"""
import argparse
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ALGO_BYTES_PER_CELL_SUBCYCLE = 360.0  # 31 doubles read + 14 written
METRIC = "EVP grid-cells*subcycles/sec at gx1 (fp64)"
UNIT = "cell-subcycles/s"
PARITY_NDTE = 24


def kernel_counters(kernel):
    """ncu counters of the dominant kernel per launch (profiles/traffic.json, from the committed --set full captures)."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
            t = json.load(fh)
        return t.get(kernel) or t.get("fused")
    except Exception:
        return None


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """SM clock and clock-event reasons DURING the timed region, sampled every 2 ms through NVML (nvidia_ml_py) -- the timed region of
    the default run is ~20 ms of GPU time, shorter than one period of `nvidia-smi -lms 100` -- with nvidia-smi as the fallback
    (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu=0):
        self.gpu, self.proc, self.lines, self.nvml, self.samples, self.stop_flag = gpu, None, [], None, [], False

    def _nvml_loop(self):
        import pynvml as N
        h = self.nvml
        names = (("hw_slowdown", N.nvmlClocksEventReasonHwSlowdown), ("hw_thermal_slowdown", N.nvmlClocksEventReasonHwThermalSlowdown),
                 ("sw_thermal_slowdown", N.nvmlClocksEventReasonSwThermalSlowdown), ("sw_power_cap", N.nvmlClocksEventReasonSwPowerCap),
                 ("hw_power_brake", N.nvmlClocksEventReasonHwPowerBrakeSlowdown))
        while not self.stop_flag:
            try:
                mhz = N.nvmlDeviceGetClockInfo(h, N.NVML_CLOCK_SM)
                r = N.nvmlDeviceGetCurrentClocksEventReasons(h)
                self.samples.append((mhz, tuple(n for n, bit in names if r & bit)))
            except Exception:
                break
            time.sleep(0.002)

    def start(self):
        try:
            import pynvml as N
            N.nvmlInit()
            # NVML enumerates physical devices: honour CUDA_VISIBLE_DEVICES when it is a plain list of indices
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            idx = self.gpu
            if vis and all(t.strip().isdigit() for t in vis.split(",")):
                idx = int(vis.split(",")[self.gpu])
            self.nvml = N.nvmlDeviceGetHandleByIndex(idx)
            self.max_mhz = N.nvmlDeviceGetMaxClockInfo(self.nvml, N.NVML_CLOCK_SM)
            self.t = threading.Thread(target=self._nvml_loop, daemon=True)
            self.t.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=lambda: self.lines.extend(self.proc.stdout), daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def stop(self):
        if self.nvml is not None:
            self.stop_flag = True
            self.t.join(timeout=2)
            sm = [m for m, _ in self.samples]
            reasons = sorted({r for _, rs in self.samples for r in rs})
            return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_min_mhz": float(min(sm)) if sm else None,
                    "sm_max_mhz": float(self.max_mhz), "reasons": reasons, "samples": len(sm), "source": "NVML, every 2 ms during the timed region"}
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
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi -lms 100"}


def layout_of(args):
    return args.layout or ("strong" if args.workload in ("tx1", "p1deg") else "weak")


def geometry(args, world):
    """(global nx, ny), (block nx, ny) = one rank's rectangle, processor grid"""
    from cice_b200 import decomp, synth
    base = synth.CONFIGS[args.workload]
    px, py = decomp.proc_grid(world, world, world) if world > 1 else (1, 1)
    if layout_of(args) == "weak" or world == 1:
        sx, sy = (int(v) for v in args.sub.lower().split("x")) if args.sub else (base["nx"], base["ny"])
        return (sx * px, sy * py), (sx, sy), (px, py)
    nx, ny = base["nx"], base["ny"]
    if nx % px or ny % py:
        raise SystemExit(f"--layout strong: {nx}x{ny} does not split into {px}x{py} equal rectangles")
    return (nx, ny), (nx // px, ny // py), (px, py)


def build_case(args, world, rank, ndte=None):
    from cice_b200 import decomp, synth
    (nx, ny), (sx, sy), (px, py) = geometry(args, world)
    c = synth.make_case(args.workload, nx=nx, ny=ny, block_size=(sx, sy), ndte=ndte)
    if world == 1:
        return c, c.grid, c.fields, np.arange(c.blocks.nblocks_tot), (px, py)
    owner, _ = decomp.cartesian_owner(c.blocks, world)
    g, f, ids = c.rank_view(owner, rank)
    return c, g, f, ids, (px, py)


def workload_string(args, world, ndte):
    (nx, ny), (sx, sy), (px, py) = geometry(args, world)
    bnd = "tripole" if args.workload == "tx1" else "cyclic/closed"
    return (f"{args.workload} {sx}x{sy} per GPU, B-grid EVP ndte={ndte}, {px}x{py} GPUs, global {nx}x{ny} ({bnd}), "
            f"box2001 synthetic, {layout_of(args) if world > 1 else 'single'} layout")


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


def cpu_blocks(args, nx, ny):
    """the reference's own block choice for <=16 PEs at gx1 is 40x48 (configuration/scripts/cice_decomp.csh:93-96)"""
    if args.workload == "gx1":
        return (40, 48)
    if args.workload == "tx1":
        return (45, 40)
    if args.workload == "p1deg":
        return (100, 100)
    return (max(nx // 8, 8), max(ny // 8, 8))


def run_reference(args):
    """--impl reference: the reference's CPU path (restated; see module docstring) on the host cores, on the
    same global grid as our arm at this GPU count."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from cice_b200 import synth
    world = max(int(os.environ.get("WORLD_SIZE", "1")), args.gpus, 1)
    (nx, ny), _, _ = geometry(args, world)
    bs = cpu_blocks(args, nx, ny)
    # a bounded sample: full dynamics steps while the grid is small, fewer subcycles of the same step on the 0.1-degree grid
    ndte_full = synth.CONFIGS[args.workload]["ndte"]
    ndte = ndte_full if nx * ny <= 1280 * 768 else 12
    case = synth.make_case(args.workload, nx=nx, ny=ny, block_size=bs, ndte=ndte)
    case.params.update(synth.evp_params(ndte_full), ndte=ndte)   # constants of the full loop, fewer subcycles
    sec, nth = cpu_baseline(case, args.steps, max(args.warmup, 1))
    val = nx * ny * ndte / sec
    sample = (f"{args.steps} full dynamics step(s) of the {nx}x{ny} grid, ndte={ndte}" if ndte == ndte_full else
              f"{args.steps} x the first {ndte} of {ndte_full} subcycles of one dynamics step of the {nx}x{ny} grid")
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": sec * 1e3 * ndte_full / ndte, "higher_is_better": True,
            "scaling": layout_of(args) if world > 1 else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_string(args, world, ndte_full),
                       "cpu": f"CPU restatement of the reference loops (oracle port, -O3 AVX2/FMA, OpenMP over {bs[0]}x{bs[1]} blocks); "
                              "the Fortran reference cannot be built in this image"},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": nth, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def parity_precheck(args, dyn_evp, abi, dist, world, rank, params, case, grid, fields):
    """untimed: the bench's own decomposition, PARITY_NDTE subcycles with the constants of the full loop, through evp_b200_run_bgrid;
    rank 0 runs the oracle (exact build) on the undecomposed grid and every rank compares sha256 digests of its 18 inout arrays."""
    import torch
    t0 = time.perf_counter()
    p = dict(params, ndte=PARITY_NDTE)
    got = {k: v.copy() for k, v in fields.items()}
    dyn_evp.dyn_evp_b200_init(grid)   # not finalized here: evp_b200_finalize also ends the NCCL communicator; the timed part's
    dyn_evp.dyn_evp_b200_run(p, got)  # own evp_b200_init releases this context
    desc = dyn_evp.describe()
    want = [None]
    if rank == 0:
        from cice_b200 import decomp
        from oracle import oracle
        oracle.build()
        ref = case.copy_fields()
        oracle.evp_run_bgrid(case.grid, dict(case.params, ndte=PARITY_NDTE), ref, nthreads=os.cpu_count() or 1, variant="exact")
        if world > 1:
            owner, _ = decomp.cartesian_owner(case.blocks, world)
            want[0] = [{n: digest(ref[n][np.nonzero(owner == r)[0]]) for n in abi.FIELDS_INOUT} for r in range(world)]
        else:
            want[0] = [{n: digest(ref[n]) for n in abi.FIELDS_INOUT}]
    if world > 1:
        dist.broadcast_object_list(want, src=0)
    bad = [n for n in abi.FIELDS_INOUT if digest(got[n]) != want[0][rank][n]]
    nbad = len(bad)
    if world > 1:
        t = torch.tensor([nbad], dtype=torch.int64, device="cuda")
        dist.all_reduce(t)
        nbad = int(t.item())
    (nx, ny), (sx, sy), (px, py) = geometry(args, world)
    return {"ok": nbad == 0, "bitwise": True, "mismatching_arrays": nbad, "arrays_per_rank": len(abi.FIELDS_INOUT), "ranks": world,
            "case": f"{args.workload} global {nx}x{ny} as {px}x{py} rectangles of {sx}x{sy}, first {PARITY_NDTE} subcycles of the step, "
                    f"mode {args.mode}", "checker": "CPU oracle (exact build) on the undecomposed grid, sha256 per array per rank",
            "first_bad": bad[:3], "halo": desc.split("p2p: ")[-1], "seconds": round(time.perf_counter() - t0, 2)}


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

    case, grid, fields, bids, (px, py) = build_case(args, world, rank)
    (nxg, nyg), (sx, sy), _ = geometry(args, world)
    params = dict(case.params, mode=abi.MODE_FAST if args.mode == "fast" else abi.MODE_EXACT,
                  kernel=abi.KERNEL_NAMES[args.kernel])
    ndte = params["ndte"]
    cells_global = nxg * nyg

    parity = None
    if not args.no_parity:
        parity = parity_precheck(args, dyn_evp, abi, dist, world, rank, params, case, grid, fields) if args.mode == "exact" else \
            {"ok": None, "note": "fast mode is not bit-identical by construction (FMA contraction); see tests for its tolerance checks"}

    dyn_evp.dyn_evp_b200_init(grid)
    # the metric arrays behind the derived geometry (evp_b200_set_metric, what the Fortran shim hands over in its init): the
    # library verifies them bit for bit on the device; the HBM-streaming form then reads two arrays instead of seven
    bad = dyn_evp.set_metric(synth.scatter(case.X["HTN"], case.blocks)[bids], synth.scatter(case.X["HTE"], case.blocks)[bids],
                             case.params["deltaminEVP"])
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
        barrier()                   # every rank enters the step together: no peer lateness inside the timed region
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
    desc = dyn_evp.describe()            # (names the tile-streaming plan once its tensor maps exist)
    tstream = "; tstream:" in desc

    # ---- end to end through the C ABI with host buffers ----------------------------------------------
    # (a) every field crosses both ways every step (evp_b200_run_bgrid, the plain drop-in);
    # (b) the 12 carried stress arrays stay on the device (evp_b200_run_bgrid_resident, EVP_B200_KEEP_STRESS): velocities,
    #     the 12 per-step inputs, masks and diagnostics still cross every step.  (b) is what INTEGRATION.md wires up for
    #     every step that does not write a restart/history file.  On tripole grids the library then symmetrises the stresses across
    #     the fold itself after the loop (evp_b200_stress_symmetrise; between the ranks of the top row by one NCCL swap per step).
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

    hfc = {k: v for k, v in hf.items() if not k.startswith("_keep_")}
    e2e_full_s = time_e2e(dyn_evp.bind("run", params, hfc))   # C structs built once, like a compiled caller: the loop times the C ABI
    resident_ok = True
    step_ok = grid["ns_boundary_type"] != abi.BNDY_NAMES["tripole"]   # the device-side step preparation is not built for tripole grids
    if resident_ok:
        e2e_s = time_e2e(dyn_evp.bind("run_resident", params, hfc, keep_stress=True))
        h2d = 18 * nblk * 8 + 2 * nblk * 4
        d2h = 6 * nblk * 8
        e2e_how = "evp_b200_run_bgrid_resident(EVP_B200_KEEP_STRESS): stresses stay on the device, everything else crosses" + \
                  ("" if step_ok else "; tripole grid: the stresses are symmetrised across the fold on the device after the loop")
    else:
        e2e_s = e2e_full_s
        h2d = 30 * nblk * 8 + 2 * nblk * 4
        d2h = 18 * nblk * 8
        e2e_how = "evp_b200_run_bgrid: every field crosses both ways (tripole grid)"
    # (d) the step preparation on the device too (evp_b200_step_resident): per step the nine T-point
    #     inputs + strength + iceTmask go in, the velocities come out; velocities, stresses and iceUmask stay on the device.
    #     Not for tripole grids in this version; between ranks the velocity halo after dyn_prep2 is one staged exchange per step.
    e2e_step = None
    if step_ok:
        static, prep = synth.step_inputs(case)
        if world > 1:   # this rank's blocks
            static = {k: (np.ascontiguousarray(v[bids]) if isinstance(v, np.ndarray) else v) for k, v in static.items()}
            prep = {k: (np.ascontiguousarray(v[bids]) if isinstance(v, np.ndarray) else v) for k, v in prep.items()}
        pst = pin({k: v for k, v in static.items() if k != "umask"})
        pst["umask"] = static["umask"]
        ppr = pin({k: v for k, v in prep.items() if isinstance(v, np.ndarray) and v.dtype == np.float64})
        for k, v in prep.items():
            if k not in ppr:
                ppr[k] = v
        import torch as _t
        tm = _t.from_numpy(prep["iceTmask"].copy()).pin_memory()
        ppr["iceTmask"] = tm.numpy()
        dyn_evp.dyn_evp_b200_prep_init({k: v for k, v in pst.items() if not k.startswith("_keep_")})
        sf = {k: v for k, v in hf.items() if not k.startswith("_keep_")}
        prep_clean = {k: v for k, v in ppr.items() if not k.startswith("_keep_")}
        dyn_evp.dyn_evp_b200_step_resident(params, prep_clean, sf, init_state=True)
        out = {"uvel": hf["uvel"], "vvel": hf["vvel"]}
        s_step = time_e2e(dyn_evp.bind("step_resident", params, out, prep=prep_clean))
        e2e_step = {"value": cells_global * ndte / s_step, "unit": UNIT, "ms_per_step": s_step * 1e3,
                    "h2d_bytes_per_step": 8 * nblk * 8 + nblk * 4, "d2h_bytes_per_step": 2 * nblk * 8,
                    "how": "evp_b200_step_resident: T-point inputs of the step in (tmass, aice_init, cdn_ocn, uocn, vocn, strairxT, strairyT, "
                           "strength, iceTmask; pinned), T->U averages + dyn_prep2 + the whole subcycle loop on the device, uvel and vvel out; "
                           "velocities, stresses and iceUmask resident"}
    # (c) what a Fortran caller hands over: pageable arrays; then the same arrays page-locked once (evp_b200_pin_host)
    e2e_pageable = None
    if world == 1 and not args.no_pageable:
        pf = {k: v.copy() for k, v in fields.items()}
        run_pf = dyn_evp.bind("run_resident", params, pf, keep_stress=True) if resident_ok else dyn_evp.bind("run", params, pf)
        s_page = time_e2e(run_pf)
        arrays = [v for v in pf.values() if isinstance(v, np.ndarray)]
        for a in arrays:
            dyn_evp.pin_host(a)
        s_reg = time_e2e(run_pf)
        for a in arrays:
            dyn_evp.unpin_host(a)
        e2e_pageable = {"pageable": {"value": cells_global * ndte / s_page, "ms_per_step": s_page * 1e3},
                        "after_evp_b200_pin_host": {"value": cells_global * ndte / s_reg, "ms_per_step": s_reg * 1e3},
                        "unit": UNIT, "how": "the e2e call from plain numpy arrays (what Fortran allocatables are), then from the same "
                                             "arrays page-locked once with evp_b200_pin_host as the Fortran shim does on its first call"}

    if world > 1:
        t = torch.tensor([total_ms, e2e_s, kernel_ms, e2e_full_s, e2e_step["ms_per_step"] if e2e_step else 0.0], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms, e2e_s, kernel_ms, e2e_full_s, step_ms = (float(x) for x in t.tolist())
        if e2e_step:
            e2e_step.update(ms_per_step=step_ms, value=cells_global * ndte / (step_ms * 1e-3))
    dyn_evp.dyn_evp_b200_finalize()

    if rank == 0:
        value = cells_global * ndte * args.steps / (total_ms * 1e-3)
        peak, peak_src = measured_peak()
        per_gpu_cells = sx * sy
        nl_step = launches / args.steps
        persistent = (nl_step == 1 and ndte > 1)   # the library ran the whole loop as one cooperative launch (KERNEL_PERSISTENT, AUTO's choice when it fits)
        kname = "persistent" if persistent else "fused"
        ctr = kernel_counters(kname) or {}
        # dominant kernel: the subcycle kernel; one launch advances the rank's sub-domain by ndte/launches subcycles
        sub_per_launch = ndte if persistent else (0.5 if args.kernel == "split" else 1.0)
        ach = per_gpu_cells * ndte * ALGO_BYTES_PER_CELL_SUBCYCLE / (kernel_ms * 1e-3) / 1e9
        us_per_subcycle = kernel_ms * 1e3 / ndte
        cold, warm = ctr.get("dram_bytes_per_launch_cold"), ctr.get("dram_bytes_per_launch_warm")
        same_shape = (args.workload == "gx1" and (sx, sy) == (320, 384))
        ts_ctr = kernel_counters("tstream_p1deg") if (tstream and (sx, sy) == (3600, 2400)) else None
        roof = {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                # the loop is 1 cold launch + (ndte-1) launches that find the working set in L2: the launch-weighted mean
                # (persistent kernel: ONE launch per step on a flushed L2, so the cold capture is the state of every timed launch)
                "traffic": (None if not (same_shape and cold) else cold if persistent else None if not warm else (cold + (ndte - 1) * warm) / ndte),
                "traffic_cold": cold if same_shape else None, "traffic_warm": warm if same_shape else None,
                "traffic_how": "dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full captures of this "
                               "kernel at gx1 (cold: caches flushed by ncu; warm: --cache-control none, the state of 239 of 240 launches); "
                               "not re-measured in this run" if same_shape else "no ncu capture for this workload shape",
                "traffic_detail": ctr if same_shape else None,
                "peak_source": peak_src, "algorithmic_bytes_per_cell_subcycle": ALGO_BYTES_PER_CELL_SUBCYCLE,
                "kernel_ms_per_step": kernel_ms, "us_per_subcycle": us_per_subcycle, "subcycles_per_launch": sub_per_launch}
        if ts_ctr and ts_ctr.get("shape") == [sx, sy]:
            # the tile-streaming kernel at this very shape: every launch streams the whole working set, one ncu capture is the state of all
            roof.update({"traffic": ts_ctr["dram_bytes_per_launch"], "traffic_cold": ts_ctr["dram_bytes_per_launch"],
                         "traffic_warm": ts_ctr["dram_bytes_per_launch"], "traffic_detail": ts_ctr,
                         "traffic_how": "dram__bytes_read.sum + dram__bytes_write.sum of one launch of this kernel at this shape from the committed "
                                        "ncu --set full capture (" + ts_ctr["source"] + "); not re-measured in this run"})
        if same_shape and ctr.get("fp64_floor_us_per_subcycle"):
            # the ceiling that binds when the working set is L2 resident: the fp64 pipe (0.5 warp-instructions / clk / sub-partition)
            fl = ctr["fp64_floor_us_per_subcycle"]
            roof["fp64_pipe"] = {"floor_us_per_subcycle": fl, "frac": fl / us_per_subcycle,
                                 "how": ctr.get("fp64_floor_how", "fp64 warp-instructions per launch (ncu) x 2 cycles / (592 sub-partitions x SM clock)")}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": total_ms / args.steps, "higher_is_better": True,
                "scaling": layout_of(args) if world > 1 else "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": {"workload": workload_string(args, world, ndte),
                           "kernel": args.kernel + (" -> persistent (one cooperative launch per step)" if persistent and args.kernel == "auto" else
                                                    " -> tstream (TMA tile-streaming kernel, one launch per subcycle)" if tstream and args.kernel == "auto" else ""),
                           "mode": args.mode,
                           "l2": "flushed between timed steps (256 MiB memset); all ranks barrier before every timed step",
                           "layout": desc, "derived_geometry_mismatches": bad},
                "parity": parity,
                "clocks": clocks,
                "e2e": {"value": cells_global * ndte / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "ms_per_step": e2e_s * 1e3, "how": e2e_how},
                "e2e_full_copy": {"value": cells_global * ndte / e2e_full_s, "unit": UNIT, "h2d_bytes_per_step": 30 * nblk * 8 + 2 * nblk * 4,
                                  "d2h_bytes_per_step": 18 * nblk * 8, "ms_per_step": e2e_full_s * 1e3,
                                  "how": "evp_b200_run_bgrid: every field crosses both ways"},
                "gpu_launches": int(launches),
                "roofline": roof,
                "wall_s": t_wall}
        if e2e_pageable:
            line["e2e_pageable"] = e2e_pageable
        if e2e_step:
            # the headline end-to-end number is the call that does the most on the device; the two older entry points stay beside it
            line["e2e_resident_stress"] = line["e2e"]
            line["e2e"] = e2e_step
        # the active-cell rate beside R (T cells that carry ice; rank 0's sub-domain)
        act = int(np.count_nonzero(np.asarray(fields["iceTmask"])[:, 1:-1, 1:-1]))
        line["active_cells"] = {"icellT_rank0": act, "fraction_rank0": act / float(per_gpu_cells),
                                "active_cell_subcycles_per_s": value * act / float(per_gpu_cells)}
        if world == 1 and not args.no_cpu:
            bs = cpu_blocks(args, nxg, nyg)
            big = nxg * nyg > 1280 * 768
            cnd = 12 if big else ndte
            ccase = synth.make_case(args.workload, nx=nxg, ny=nyg, block_size=bs, ndte=cnd)
            ccase.params.update(synth.evp_params(ndte), ndte=cnd)
            ncpu = 2 if big else 12   # ~3 s of wall time on 16 threads at gx1 (~50 core-seconds)
            sec, nth = cpu_baseline(ccase, ncpu, 1)
            line["cpu_baseline"] = {"value": nxg * nyg * cnd / sec, "unit": UNIT, "cores": nth, "kind": "port",
                                    "sample": f"{ncpu} x {cnd} subcycles of {args.workload} {nxg}x{nyg} (full loop: ndte={ndte}), "
                                              f"blocks {bs[0]}x{bs[1]}, {sec:.3f} s each"}
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
    p = dict(c.params, mode=abi.MODE_FAST if args.mode == "fast" else abi.MODE_EXACT,
             kernel=abi.KERNEL_SPLIT if args.kernel == "split" else abi.KERNEL_AUTO)
    parity = None
    if not args.no_parity and args.mode == "exact":
        from oracle import oracle
        oracle.build()
        pc = synth.make_ccase(args.workload, ndte=PARITY_NDTE)
        pc.params.update(synth.evp_params(ndte), ndte=PARITY_NDTE, visc_method=c.params["visc_method"], deltaminEVP=c.params["deltaminEVP"])
        ref, got = pc.copy_fields(), pc.copy_fields()
        oracle.evp_run_cgrid(pc.grid, pc.cgrid, pc.params, ref, nthreads=os.cpu_count() or 1, variant="exact")
        dyn_evp.dyn_evp_b200_init(pc.grid)
        try:
            dyn_evp.dyn_evp_b200_init_cgrid(pc.cgrid)
            dyn_evp.dyn_evp_b200_run_cgrid(dict(pc.params, mode=p["mode"], kernel=p["kernel"]), got)
        finally:
            dyn_evp.dyn_evp_b200_finalize()
        names = [n for n in abi.CFIELDS_INOUT + abi.CFIELDS_OUT if n != "strengthU"]
        badn = [n for n in names if digest(got[n]) != digest(ref[n])]
        parity = {"ok": not badn, "bitwise": True, "mismatching_arrays": len(badn), "arrays_per_rank": len(names), "ranks": 1,
                  "case": f"{args.workload} C grid, first {PARITY_NDTE} subcycles of the step", "first_bad": badn[:3],
                  "checker": "CPU oracle (exact build), sha256 per array"}
    dyn_evp.dyn_evp_b200_init(c.grid)
    dyn_evp.dyn_evp_b200_init_cgrid(c.cgrid)
    f = pin(c.copy_fields())
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    sampler = ClockSampler(0)
    loop, call = [], []
    for it in range(max(args.warmup, 3) + args.steps):
        if it == max(args.warmup, 3):
            sampler.start()
        flush.zero_()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        dyn_evp.dyn_evp_b200_run_cgrid(p, f)
        call.append(time.perf_counter() - t0)
        loop.append(dyn_evp.last_loop_ms())
    clocks = sampler.stop()
    loop, call = loop[-args.steps:], call[-args.steps:]
    ms = float(np.mean(loop))
    peak, peak_src = measured_peak()
    ach = cells * ndte * C_ALGO_BYTES / (ms * 1e-3) / 1e9
    nblk = int(np.prod(c.fields["uvel"].shape))
    nl = int(dyn_evp.last_launches())
    print(json.dumps({"metric": METRIC.replace("gx1", "gx1 C-grid"), "value": cells * ndte / (ms * 1e-3), "unit": UNIT, "n_gpus": 1,
                      "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
                      "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                      "config": {"workload": f"{args.workload} {base['nx']}x{base['ny']} C-grid EVP (standard_2d) ndte={ndte}, 1 GPU, box2001 synthetic",
                                 "mode": args.mode, "l2": "flushed between timed steps (256 MiB memset)", "layout": dyn_evp.describe()},
                      "parity": parity, "clocks": clocks,
                      "e2e": {"value": cells * ndte / float(np.mean(call)), "unit": UNIT, "h2d_bytes_per_step": 37 * nblk * 8 + 4 * nblk * 4,
                              "d2h_bytes_per_step": 21 * nblk * 8, "ms_per_step": float(np.mean(call)) * 1e3},
                      "gpu_launches": nl * args.steps,
                      "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": None,
                                   "peak_source": peak_src, "algorithmic_bytes_per_cell_subcycle": C_ALGO_BYTES,
                                   "kernels_per_subcycle": nl // ndte, "us_per_subcycle": ms * 1e3 / ndte}}))
    dyn_evp.dyn_evp_b200_finalize()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--kernel", default="auto", choices=["auto", "split", "fused", "stream", "resident", "persistent", "tstream"])
    ap.add_argument("--mode", default="exact", choices=["fast", "exact"])
    ap.add_argument("--workload", default="gx1")
    ap.add_argument("--layout", default=None, choices=["weak", "strong"],
                    help="N > 1: weak = one --sub sized sub-domain per GPU (default for gx1, gx3); strong = the workload's global grid cut "
                         "into px x py rectangles (default for tx1, p1deg: configs[3], configs[4])")
    ap.add_argument("--sub", default=None, help="weak layout: per-GPU sub-domain NXxNY (the weak-scaling halo sweep), default the workload's size")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-pageable", action="store_true")
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
