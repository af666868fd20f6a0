#!/usr/bin/env python
"""bench.py - OBCA-MPC solves/sec on B200 (BASELINE.json metric), one JSON line on stdout.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--cfg 2|3|5] [--batch B] [--start warm|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

Default workload (SURVEY.md 8(d) cfg 3, the configuration the metric is quoted on; --cfg 2 / 5 select the other
single-launch BASELINE configurations - cfg 5 is the 65,536-instance FIXED_SET case when run on 8 GPUs): per GPU 8,192 random ego start
poses in one shared scene of 4 rotated-rectangle obstacles (R = 16 half-space rows), horizon N = 20, free-time
mode (obca_mpc4), A* reference window as xref, A* warm start.  A "step" = one launch of the batched solver over
the rank's 8,192 instances (+ for N > 1 the single NCCL gather of the packed results).  Weak scaling: every
rank owns its own 8,192 instances, nothing is exchanged during the solve.

value      whole-job solves/s, inputs resident in HBM, CUDA events on the launch stream around the K steps (L2 flush
           and, for N > 1, the gathers included), max over ranks
e2e        the same with HOST buffers in and out every step: obca_b200_solve_host on one GPU; per rank H2D + solve +
           gather + rank 0's D2H of all results on several
roofline   algorithmic bytes per solve (7,160 B at cfg 3) x solves/s against the measured HBM peak
roofline_fp64  fp64 flop of the launch (ncu count per iteration x iterations) against the DFMA rate measured now
cpu_baseline  the C oracle (oracle/obca_oracle.c, a restatement - CasADi/IPOPT cannot be installed here) on the
           same batch on all host cores
--impl reference  the CPU arm alone, same metric/config/instances (rank 0 only); never loads the CUDA library
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "OBCA-MPC solves/sec (N=20, 4 obstacles)"
UNIT = "solves/s"
# the BASELINE configurations that are one batched launch (cfg 4, the closed loop, has its own driver:
# tools/bench_closed_loop.py).  batch = instances per GPU.
CFGS = {2: dict(batch=1024, N=10, n_obs=2, rows=8, dyn_rows=0, mode="FREE(obca_mpc4)",
                what="2 static polytope obstacles (R=8), N=10, free-time obca_mpc4"),
        3: dict(batch=8192, N=20, n_obs=4, rows=16, dyn_rows=0, mode="FREE(obca_mpc4)",
                what="4 static polytope obstacles (R=16), N=20, free-time obca_mpc4"),
        5: dict(batch=8192, N=20, n_obs=6, rows=24, dyn_rows=8, mode="FIXED_SET(obca_mpc6)",
                what="4 static + 2 dynamic polytope obstacles (R=24), N=20, fixed-time obca_mpc6 with terminal set, Ts=2.0")}


def bytes_per_solve(N, n_obs, rows, free=True, n_dyn_rows=0):
    """compulsory fp64 bytes of one solve (obstacle rows counted per instance)."""
    b_in = 8 * (3 + 2 + 3 * (N + 1) + (1 if free else 3) + 3 * (rows - n_dyn_rows) + 4 * n_dyn_rows)
    b_out = 8 * (3 * (N + 1) + 2 * N + rows * (N + 1) + 4 * n_obs * (N + 1) + 2) + 8
    return b_in, b_out


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons while the timed region runs (B200_PROFILING.md): one streaming
    `nvidia-smi -lms 100` process, lines stamped on arrival; summary over the samples inside [mark_start, mark_stop]."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.rows = []
        self.t0 = self.t1 = None
        self.proc = None

    def run_nvml(self):
        """the same readings straight from NVML in this process (OBCA_BENCH_SAMPLER=nvml): no second process polling the
        driver while the timed region runs"""
        import pynvml as nv
        nv.nvmlInit()
        h = nv.nvmlDeviceGetHandleByIndex(self.index)
        mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
        bits = [(nv.nvmlClocksEventReasonHwSlowdown, 3), (nv.nvmlClocksEventReasonHwThermalSlowdown, 4),
                (nv.nvmlClocksEventReasonSwThermalSlowdown, 5), (nv.nvmlClocksEventReasonSwPowerCap, 6)]
        while not self.quit:
            r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
            row = [str(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)), str(mx), "%.2f" % (nv.nvmlDeviceGetPowerUsage(h) / 1000.0),
                   "", "", "", ""]
            for bit, col in bits:
                row[col] = "Active" if (r & bit) else "Not Active"
            self.rows.append((time.time(), row))
            time.sleep(0.1)

    quit = False

    def run(self):
        mode = os.environ.get("OBCA_BENCH_SAMPLER", "smi")
        if mode == "off":
            return
        if mode == "nvml":
            try:
                return self.run_nvml()
            except Exception:
                pass                       # no NVML binding on this box: the nvidia-smi line below
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
            for line in self.proc.stdout:
                self.rows.append((time.time(), [c.strip() for c in line.strip().split(",")]))
        except Exception:
            pass

    def mark_start(self):
        self.t0 = time.time()

    def mark_stop(self):
        self.t1 = time.time()

    def stop(self):
        time.sleep(0.15)                       # let the last in-window sample arrive
        self.quit = True
        if self.proc is not None:
            self.proc.terminate()
        self.join(timeout=6)
        t0 = self.t0 or 0.0; t1 = (self.t1 or time.time()) + 0.12
        inside = [r for (t, r) in self.rows if t0 <= t <= t1] or [r for (_, r) in self.rows[-3:]]
        sm = []; mx = 0.0; reasons = set(); pw = 0.0
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in inside:
            try:
                sm.append(float(r[0])); mx = max(mx, float(r[1])); pw = max(pw, float(r[2]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None, "power_w_max": pw or None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_arm(b, prm, a, sample, threads, steps=1, warmup=0):
    """The C oracle on the first ``sample`` instances of the batch with ``threads`` pthreads."""
    from oracle import c_oracle
    sl = lambda v: None if v is None else v[:sample]
    args = (prm, sl(a["x0"]), sl(a["u0"]), sl(a["xref"]), a["edge_ptr"], a["A"], a["b0"], a["db"])
    kw = dict(T_max=sl(a["T_max"]), term=sl(a["term"]), nthreads=threads)
    for _ in range(warmup):
        c_oracle.solve(*args, **kw)
    t0 = time.perf_counter()
    for _ in range(steps):
        c = c_oracle.solve(*args, **kw)
    dt = (time.perf_counter() - t0) / steps
    return sample / dt, dt, c


def histograms(status, iters):
    """status histogram (0 converged, 1 acceptable level, 2 noise floor, negative: failures - include/obca_b200.h) and
    iteration histogram of one batch"""
    edges = [0, 10, 15, 20, 25, 30, 40, 50, 75, 100, 150, 200, 300, 500, 1000, 1 << 30]
    h, _ = np.histogram(iters, bins=edges)
    return ({str(int(k)): int(v) for k, v in zip(*np.unique(status, return_counts=True))},
            {("%d-%d" % (edges[i], edges[i + 1] - 1)) if i + 2 < len(edges) else ">=%d" % edges[i]: int(h[i])
             for i in range(len(h)) if h[i]})


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cfg", type=int, default=3, choices=sorted(CFGS), help="BASELINE configuration (SURVEY 8(d))")
    ap.add_argument("--batch", type=int, default=0, help="instances per GPU (0 = the configuration's)")
    ap.add_argument("--start", default="warm", choices=["warm", "reference"],
                    help="warm: A* warm start, batch defaults; reference: the reference's start (zeros, T=1) with IPOPT's "
                         "mu_init 0.1 / bound_push 1e-2")
    ap.add_argument("--cpu-sample", type=int, default=0, help="instances of the CPU sample (0 = the whole batch)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--gather", default="overlapped", choices=["overlapped", "serial"],
                    help="N > 1: the peer-to-peer copies of a step's results run under the next launch (overlapped) or land "
                         "before it starts (serial: a 4-byte all-reduce after the copies orders the ranks).  Measured at N = 4: "
                         "the copies arriving under rank 0's launch slow it by 0.2 ms (they sweep its L2, where the solver's "
                         "per-thread state lives), waiting for them costs 3 ms per step - overlapped is the default")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    cores = os.cpu_count() or 1

    from vehicle_motion_planning_with_obstacles_avoidance_using_mpc_b200 import _abi, scenario as sc

    CFG = args.cfg
    cf = CFGS[CFG]
    B = args.batch or cf["batch"]
    init_kw = dict(init=_abi.INIT_WARM) if args.start == "warm" else dict(init=_abi.INIT_ZERO, mu_init=0.1, bound_push=1e-2)
    config = {"workload": "cfg%d: batch=%d random start poses per GPU, %s, A* reference window%s" % (
                  CFG, B, cf["what"], " + warm start" if args.start == "warm" else ", the reference's start (zeros)"),
              "cfg": CFG, "batch_per_gpu": B, "N": cf["N"], "n_obs": cf["n_obs"], "rows": cf["rows"], "mode": cf["mode"],
              "init": "A* warm start" if args.start == "warm" else "reference start (zeros, T=1; mu_init 0.1, bound_push 1e-2)",
              "parallelism": "batch-sharded x%d, one gather%s" % (world, " (overlapped with the next launch)" if world > 1 else ""),
              "l2": "flushed before every timed step (256 MiB memset, inside the timed region: ~0.05 ms)"}
    metric = METRIC if CFG == 3 else "OBCA-MPC solves/sec (cfg %d: N=%d, %d obstacles)" % (CFG, cf["N"], cf["n_obs"])

    # ------------------------------------------------------------------ reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return
        # the SAME instances as rank 0 of the B200 arm, planned with the Python A* (identical paths) so that this process
        # never loads the CUDA library
        sample = args.cpu_sample or B
        b = sc.make_batch(CFG, B, pose_seed=None if world == 1 else 977, native_planner=False)
        prm, a = sc.batch_arrays(b, **init_kw)
        v, dt, c = cpu_arm(b, prm, a, sample, cores, steps=max(1, args.steps), warmup=min(args.warmup, 1))
        sh, ih = histograms(c["status"], c["iters"])
        cb = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
              "sample": "the first %d of the %d instances of the B200 arm's batch per step, %d pthreads, C oracle "
                        "(restatement of the reference NLP + IPM; CasADi/IPOPT absent)" % (sample, B, cores)}
        print(json.dumps({"impl": "reference", "metric": metric, "value": v, "unit": UNIT, "n_gpus": args.gpus,
                          "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
                          "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
                          "cpu_baseline": cb, "gpu_launches": 0,
                          "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                          "success_rate": float((c["status"] >= 0).mean()), "status_histogram": sh, "iters_histogram": ih,
                          "native_libraries_loaded": sorted(set(os.path.basename(l.split()[-1]) for l in open("/proc/self/maps")
                                                            if "libobca" in l and l.rstrip().endswith(".so")))}))
        return

    # ------------------------------------------------------------------ B200 arm
    import torch
    import torch.distributed as dist

    from vehicle_motion_planning_with_obstacles_avoidance_using_mpc_b200 import _lib, obca as om, sharding

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    # same scene on every rank, own start poses per rank (weak scaling)
    b = sc.make_batch(CFG, B, pose_seed=None if world == 1 else 977 * (rank + 1))
    prm, a = sc.batch_arrays(b, **init_kw)
    solver = om.BatchSolver(prm, a["edge_ptr"], B, device=local)
    t = lambda v: None if v is None else torch.as_tensor(v, dtype=torch.float64, device=dev).contiguous()
    KEYS = ("x0", "u0", "xref", "A", "b0", "db", "T_max", "term")
    d = {k: t(a[k]) for k in KEYS}
    # two result buffers: the gather of launch k (side stream) runs under launch k+1
    packed = [sharding.PackedOutputs(B, prm.N, prm.rows, prm.n_obs, device=dev) for _ in range(2 if world > 1 else 1)]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream(dev)
    # the gather: peer-to-peer copies into rank 0's buffer on a side stream (copy engines, overlapped with the next
    # launch); if the node refuses peer access, ONE NCCL gather per step on the launch stream
    peer = None
    full = [None, None]
    if world > 1:
        try:
            peer = sharding.PeerGather(packed[0].words, 2, dev)
            if rank == 0:
                full = [peer.full[0], peer.full[1]]
                if os.environ.get("OBCA_BENCH_ROOT_COPY", "0") != "1":
                    # rank 0 solves straight into its row of the gathered buffer: one 51 MB copy less through its L2
                    packed = [sharding.PackedOutputs(B, prm.N, prm.rows, prm.n_obs, device=dev, buf=peer.full[j, 0]) for j in range(2)]
        except Exception as e:                                     # noqa: BLE001
            peer = None
            if rank == 0:
                print("bench: peer-to-peer gather unavailable (%s); NCCL gather on the launch stream" % e, file=sys.stderr)
                full = [torch.empty((world, packed[0].words), dtype=torch.float64, device=dev) for _ in packed]
        config["parallelism"] = "batch-sharded x%d, %s" % (world, "one peer-to-peer copy of the packed results per rank into rank 0's "
                                                          "buffer (NVLink, copy engines), %s" % ("landed before the next launch (4-byte all-reduce)" if args.gather == "serial" else "overlapped with the next launch") if peer else
                                                          "one NCCL gather of the packed results")

    def solve_into(pk, src=d):
        solver.solve(src["x0"], src["u0"], src["xref"], src["A"], src["b0"], src["db"], T_max=src["T_max"], term=src["term"],
                     out=pk.views)

    sync_word = torch.zeros(1, dtype=torch.int32, device=dev)

    def run_steps(K, src=d, before=None, after=None):
        """K steps back to back on the launch stream; the gather of step k is issued on the side stream and overlaps
        step k+1.  Returns the per-step (start, end) events of the solver launches."""
        evs, done = [], [None, None]
        for k in range(K):
            j = k % len(packed)
            if done[j] is not None:
                stream.wait_event(done[j])          # the buffer's previous gather has read it
            flush.zero_()                           # L2 flush
            if before:
                before(k)
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            solve_into(packed[j], src)
            e1.record(stream)
            evs.append((e0, e1))
            if peer is not None:
                done[j] = peer.push(packed[j], j, e1)
                if args.gather == "serial":
                    stream.wait_event(done[j])                    # this rank's copy, then everybody's, before the next step
                    with torch.cuda.stream(stream):
                        dist.all_reduce(sync_word)
            elif world > 1:
                sharding.gather_packed(packed[j], dst=0, out=full[j])
            if after:
                after(k, j)
        if peer is not None:
            stream.wait_stream(peer.side)
        return evs

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()                    # streaming nvidia-smi needs a moment to come up: start before the warm-up
    run_steps(max(3, args.warmup))
    barrier()
    l0 = solver.launches
    E0 = torch.cuda.Event(enable_timing=True); E1 = torch.cuda.Event(enable_timing=True)
    barrier()
    if sampler:
        sampler.mark_start()
    w0 = time.perf_counter()
    E0.record(stream)
    evs = run_steps(args.steps)
    E1.record(stream)
    barrier()
    wall = time.perf_counter() - w0
    if sampler:
        sampler.mark_stop()
    launches = solver.launches - l0
    step_ms = E0.elapsed_time(E1)                                # this rank, all K steps (flushes and gathers included)
    kernel_ms = [e0.elapsed_time(e1) for e0, e1 in evs]          # the solver launches alone, on their stream
    tt = torch.tensor([step_ms, sum(kernel_ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    total_ms, kern_total_ms = float(tt[0]), float(tt[1])
    per_rank_kernel_ms = [sum(kernel_ms) / len(kernel_ms)]
    if world > 1:                                                 # which rank sets the step: the mean solver time of each
        allk = [torch.zeros(1, dtype=torch.float64, device=dev) for _ in range(world)]
        dist.all_gather(allk, torch.tensor([per_rank_kernel_ms[0]], dtype=torch.float64, device=dev))
        per_rank_kernel_ms = [float(v[0]) for v in allk]
    value = world * B * args.steps / (total_ms * 1e-3)
    status = packed[(args.steps - 1) % len(packed)].views["status"].cpu().numpy()
    iters = packed[(args.steps - 1) % len(packed)].views["iters"].cpu().numpy()

    # ---- e2e: HOST buffers in, HOST results out, every step.  One GPU: the C-ABI host entry (obca_b200_solve_host: H2D +
    # solve + D2H, chunked over streams).  Several GPUs: per rank H2D of its inputs from pinned memory, the solve, the
    # gather of the packed results to rank 0 and rank 0's D2H of ALL ranks' results (flush included, as above).
    pin = lambda v: None if v is None else torch.as_tensor(np.ascontiguousarray(v), dtype=torch.float64).pin_memory()
    hp = {k: pin(a[k]) for k in KEYS}
    h2d = sum(v.numel() * 8 for v in hp.values() if v is not None)
    e_steps = max(2, min(args.steps, 5))
    if world == 1:
        h = {k: (None if v is None else v.numpy()) for k, v in hp.items()}
        hout = solver.alloc_host_outputs(B, pinned=True)
        d2h = sum(v.nbytes for v in hout.values())
        host_step = lambda: solver.solve_host(h["x0"], h["u0"], h["xref"], h["A"], h["b0"], h["db"], T_max=h["T_max"],
                                              term=h["term"], out=hout)
        host_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e_steps):
            flush.zero_()
            host_step()
        torch.cuda.synchronize(dev)
        et = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    else:
        dd = {k: (None if v is None else torch.empty_like(v, device=dev)) for k, v in hp.items()}
        hfull = torch.empty((world, packed[0].words), dtype=torch.float64).pin_memory() if rank == 0 else None
        d2h = world * packed[0].nbytes if rank == 0 else 0

        def h2d_copy(k):
            for key, v in hp.items():
                if v is not None:
                    dd[key].copy_(v, non_blocking=True)

        def d2h_copy(k, j):
            if peer is not None:
                peer.wait_all()                        # every rank's row of step k has landed in rank 0's buffer
            if rank == 0:
                hfull.copy_(full[j], non_blocking=True)
        run_steps(1, src=dd, before=h2d_copy, after=d2h_copy)
        barrier()
        t0 = time.perf_counter()
        run_steps(e_steps, src=dd, before=h2d_copy, after=d2h_copy)
        torch.cuda.synchronize(dev)
        et = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        dist.all_reduce(et, op=dist.ReduceOp.MAX)
    e2e_value = world * B * e_steps / float(et[0])
    clocks = sampler.stop() if sampler else None

    stats = torch.tensor([float((status >= 0).sum()), float(iters.sum()), float(iters.max())], dtype=torch.float64, device=dev)
    if world > 1:
        s2 = stats.clone(); dist.all_reduce(s2, op=dist.ReduceOp.SUM)
        mx = stats[2:].clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        stats = torch.stack([s2[0], s2[1], mx[0]])
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the (dominant) kernel: algorithmic bytes per launch / mean launch duration
    b_in, b_out = bytes_per_solve(prm.N, prm.n_obs, prm.rows, free=_abi.is_free(prm.mode), n_dyn_rows=cf["dyn_rows"])
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0)); peak_src = "measured" if "hbm_gbs" in peaks else "fallback"
    kern_ms = kern_total_ms / args.steps
    achieved = (b_in + b_out) * B / (kern_ms * 1e-3) / 1e9
    prof = {}
    try:
        prof = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json")))
    except Exception:
        pass
    traffic = prof.get("dram_bytes_per_launch") if CFG == 3 and B == 8192 else None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peak_src + " (MEASURED_PEAKS.json hbm_gbs)",
                "kernel": "obca_solve_kernel", "kernel_ms": kern_ms, "kernel_ms_per_rank": [round(v, 3) for v in per_rank_kernel_ms],
                "bytes_per_solve": b_in + b_out,
                "note": "fp64 interior-point iterations run on-chip/L2; compulsory HBM traffic is ~7 KB per solve, so the "
                        "kernel is fp64-latency bound, not HBM bound (see roofline_fp64, DESIGN.md, profiles/)"}
    # ---- second roofline, the one that binds: fp64 instructions of the launch (counted by ncu per interior-point
    # iteration, profiles/roofline_traffic.json) against the DFMA rate of this device measured by a probe kernel now
    import ctypes as C
    pk = C.c_double(0.0)
    fp64 = None
    if _lib.lib().obca_b200_fp64_peak(local, C.byref(pk)) == 0 and pk.value > 0:
        fpi = prof.get("fp64_flop_per_iteration_cfg%d" % CFG)
        it_sum = float(iters.sum())
        fp64 = {"bound": "fp64", "peak": pk.value, "unit": "TFLOP/s", "peak_source": "measured now (obca_b200_fp64_peak: "
                "8 independent DFMA chains per thread, 8 blocks of 256 threads per SM)",
                "flop_per_iteration": fpi, "iterations_per_launch": it_sum,
                "achieved": (fpi * it_sum / (kern_ms * 1e-3) / 1e12) if fpi else None}
        fp64["frac"] = (fp64["achieved"] / pk.value) if fp64["achieved"] else None

    # ---- CPU baseline beside it (the same batch, all host cores; bounded: one pass)
    cb = None
    if not args.no_cpu_baseline:
        sample = args.cpu_sample or B
        v, dt, c = cpu_arm(b, prm, a, sample, cores, steps=1, warmup=0)
        cb = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
              "sample": "first %d instances of the timed batch, %d pthreads, %.1f s, C oracle (restatement of the "
                        "reference NLP + IPM; CasADi/IPOPT not installable)" % (sample, cores, dt)}

    sh, ih = histograms(status, iters)
    line = {"metric": metric, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": config,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "steps": e_steps, "path": "obca_b200_solve_host" if world == 1 else
                    "per rank H2D + solve + NCCL gather, rank 0 D2H of all ranks' results"},
            "gpu_launches": int(launches), "roofline": roofline, "roofline_fp64": fp64, "cpu_baseline": cb, "clocks": clocks,
            "success_rate": float(stats[0]) / (world * B), "iters_mean": float(stats[1]) / (world * B),
            "iters_max": int(stats[2]), "status_histogram": sh, "iters_histogram": ih, "wall_s": wall,
            # SURVEY 8(d) secondary figure: structural size of the compact primal-dual system, 8*(nnz(H lower) + nnz(J) +
            # n + m) ~ 72 KB per interior-point iteration at cfg 3, times the iterations actually run.  It is the
            # on-chip data rate of the KKT work, NOT HBM traffic (ncu DRAM bytes per launch are two orders below it).
            "kkt": {"bytes_per_iteration": 72000, "equivalent_gb_s": value / world * (float(stats[1]) / (world * B)) * 72000 / 1e9,
                    "per": "GPU", "note": "KKT data never streams through HBM (registers / shared memory / L1-L2 per block); "
                    "reported for the metric's 'KKT GB/s', not a roofline numerator"}}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
