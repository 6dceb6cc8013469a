#!/usr/bin/env python
"""bench.py -- headline benchmark of the collision hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload NAME]

Workloads
  0d_n32   (default) BASELINE config 3: 0D hard spheres at N=32, 1.07e9 precomputed weights (8.59 GB);
           one step = one Q(f,f) evaluation = ComputeQ: forward transform, N^6 weighted convolution
           (one pass over the weight tensor), inverse transform.  At N GPUs: N independent replicas
           (0D does not shard; "replicas only"), metric = total evals/s.
           The same line carries, as sub-records:
             oned       the SHARDED 1D-3V cases at this N (spectralbte_b200/bench1d.py): shock_strong (640 cells in
                        total), shock_strong_601, shock_weak (640 cells per GPU), heattrans_strong (N=24, 250 cells),
                        heattrans22_strong -- cells*steps/s, per-step kernel / non-kernel time, halo mode, e2e -- and,
                        for N > 1, halo_parity: both halo modes on both uneven partitions against the one-GPU run,
                        bit for bit
             sustained  >= 2 s of back-to-back ComputeQ and ComputeQ_maxPreserve with clock samples
             dropin     (N = 1) wall time of the reference's own driver linked against this library
                        (oracle/_ref/boltz_gpu) next to the reference executable (oracle/_ref/boltz_)
  bkw16    BASELINE config 2; shock_strong | shock_strong_601 | shock_weak (alias shock1p2) | heattrans_strong
           (alias heattrans) | heattrans22_strong (alias heattrans22): one 1D case as the headline line.

Every GPU number is timed with CUDA events on the library's stream, max over ranks.  The weights
(8.59 GB per pass) are larger than L2 for 0d_n32; for the 1D cases the per-step working set
(slabs + spectra + weights) exceeds L2 as well -- no explicit flush is needed (config.l2 says which).

--impl reference times the reference's own CPU implementation (oracle/_ref/libref.so: its C sources
compiled unmodified; else the oracle port) on the host cores for the same metric/config.
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
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC_0D = "Q(f,f) evals/s at N=32"
SEED = 20261017


def config_0d():
    """`config` of the default workload -- the same dict in both arms (the driver compares them)."""
    return {"workload": "0d_n32: 0D hard spheres N=32, ComputeQ(f,f), 1.07e9 precomputed weights (8.59 GB)",
            "N": 32, "L_v": 5.0, "init_field": 0,
            "l2": "inputs (8.59 GB weight stream per evaluation) larger than L2; no flush"}


def config_bkw16():
    return {"workload": "bkw16: input_examples/BKW16.in, 0D BKW relaxation N=16 lambda=0 RK2 (6 compute_Qhat per step)",
            "N": 16, "L_v": 5.0, "dt": 0.01,
            "l2": "the 134 MB weight tensor (69 MB symmetrised) is L2-resident between passes; stated, not flushed"}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, copy)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons (B200_PROFILING.md recipe).  The sampler runs from before the
    warm-up to the end of the run; every sample is time-stamped on arrival so the ones inside a timed
    window can be told apart (digest(t0, t1))."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, device):
        self.device, self.proc, self.lines = device, None, []
        self.t0 = self.t1 = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thr = threading.Thread(target=self._pump, daemon=True)
            self.thr.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def mark(self, t0, t1):
        self.t0, self.t1 = t0, t1

    def wait_first_sample(self, timeout=3.0):
        t = time.perf_counter()
        while self.proc and not self.lines and time.perf_counter() - t < timeout:
            time.sleep(0.01)

    def _rows(self, rows):
        sm, mx, pw, reasons = [], [], [], set()
        for _, ln in rows:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                mx.append(float(parts[2]))
                pw.append(float(parts[3]))
            except ValueError:
                continue
            for nm, val in zip(self.NAMES, parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return sm, mx, pw, reasons

    def digest(self, t0, t1):
        """Samples that arrived inside [t0, t1] (perf_counter times)."""
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"], "samples": 0}
        time.sleep(0.05)
        sm, mx, pw, reasons = self._rows([r for r in list(self.lines) if t0 - 0.02 <= r[0] <= t1 + 0.02])
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_min_mhz": min(sm) if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "power_w_max": max(pw) if pw else None,
                "power_w_median": float(np.median(pw)) if pw else None, "reasons": sorted(reasons), "samples": len(sm)}

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        inside = [r for r in self.lines if self.t0 is not None and self.t0 - 0.02 <= r[0] <= self.t1 + 0.02]
        window, rows = "timed region", inside
        if len(inside) < 2:   # the timed region is shorter than nvidia-smi's sampling period
            rows, window = self.lines, "whole run (warm-up, timed region and the legs after it: same kernels)"
        sm, mx, pw, reasons = self._rows(rows)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "reasons": sorted(reasons), "samples": len(sm),
                "samples_in_timed_region": len(inside), "window": window}


# --------------------------------------------------------------------------------------------
# CPU legs (the only place the oracle / oracle/_ref may be executed from bench.py)
# --------------------------------------------------------------------------------------------
def _use_all_cores():
    import ctypes as C
    cores = int(os.environ.get("SBTE_CPU_THREADS", os.cpu_count() or 1))
    os.environ["OMP_NUM_THREADS"] = str(cores)   # torchrun exports OMP_NUM_THREADS=1; the baseline uses every core
    try:
        C.CDLL("libgomp.so.1").omp_set_num_threads(cores)   # also covers an already-initialised libgomp
    except OSError:
        pass
    return cores


def cpu_computeq_n32(steps, warmup, distinct_rows=512):
    """Times full N=32 ComputeQ evaluations (all N^6 = 1.07e9 weight/operand pairs each) with the
    reference's own code when oracle/_ref/libref.so exists, else with the oracle port.  To bound host
    memory the N^3 weight-row pointers alias `distinct_rows` distinct rows (134 MB) of the synthetic
    tensor; the loop, its index arithmetic and its memory stream are unchanged."""
    import ctypes as C
    from oracle import oracle as orc
    from spectralbte_b200 import initial
    from spectralbte_b200.api import velocity_grids
    N, L_v = 32, 5.0
    n3 = N ** 3
    cores = _use_all_cores()
    v, eta = velocity_grids(N, L_v)
    f = initial.init_hom(v, L_v, 0)
    rng = np.random.default_rng(SEED)
    Wsmall = rng.random((distinct_rows, n3)) - 0.5
    dp = C.POINTER(C.c_double)
    Q = np.empty(n3)
    if orc.have_ref():
        kind = "reference"
        R = orc.Reference(N, L_v, 0)
        rows = (dp * n3)()
        base = Wsmall.ctypes.data
        for i in range(n3):
            rows[i] = C.cast(base + (i % distinct_rows) * n3 * 8, dp)
        fp, Qp = f.ctypes.data_as(dp), Q.ctypes.data_as(dp)
        call = lambda: R.R.ComputeQ(fp, fp, Qp, rows)  # noqa: E731
    else:
        kind = "port"
        o = orc.Oracle(N, L_v, 0)
        fp, Qp, Wp = f.ctypes.data_as(dp), Q.ctypes.data_as(dp), Wsmall.ctypes.data_as(dp)
        call = lambda: o.L.orc_compute_q_rowmod(o.h, Wp, C.c_long(distinct_rows), fp, fp, Qp)  # noqa: E731
    for _ in range(warmup):
        call()
    t0 = time.perf_counter()
    for _ in range(steps):
        call()
    dt = time.perf_counter() - t0
    return {"value": steps / dt, "unit": "evals/s", "cores": cores, "kind": kind, "seconds": dt,
            "sample": "%d full N=32 ComputeQ evaluations (1.07e9 pairs each, FFTs included) after %d warm-up, weight rows "
                      "aliased to %d distinct synthetic rows, %d OpenMP threads" % (steps, warmup, distinct_rows, cores)}


def cpu_cell_leg(N, L_v, W_host, f_cell, stages, evals=6):
    """1D cpu_baseline: seconds per ComputeQ + conserveAllMoments on one cell with the reference's own code
    (all host threads); the reference processes cells sequentially (exec/boltz.c:285-345)."""
    from oracle import oracle as orc
    cores = _use_all_cores()
    if orc.have_ref():
        R = orc.Reference(N, L_v, 1)
        rows = R.rows(W_host)
        call = lambda: R.conserve(R.compute_q(rows, f_cell, f_cell))  # noqa: E731
        kind = "reference"
    else:
        o = orc.Oracle(N, L_v, 1)
        call = lambda: o.conserve(o.compute_q(W_host, f_cell, f_cell))  # noqa: E731
        kind = "port"
    call()
    t0 = time.perf_counter()
    for _ in range(evals):
        call()
    sec = (time.perf_counter() - t0) / evals
    return {"value": 1.0 / (stages * sec), "unit": "cells*steps/s", "cores": cores, "kind": kind,
            "sample": "%d x (ComputeQ + conserveAllMoments) on one cell at N=%d with %d OpenMP threads; cells are "
                      "sequential in the reference, %d evaluation(s) per cell per step, transport excluded"
                      % (evals, N, cores, stages)}


def cpu_bkw16(steps, warmup):
    """BKW16 on the host: RK2 steps of exec/boltz.c:189-241 = 2 x (ComputeQ_maxPreserve + conserveAllMoments) + updates,
    i.e. 6 compute_Qhat per step, with the reference's own code (synthetic weights: timing is value-independent)."""
    from oracle import oracle as orc
    from spectralbte_b200 import initial
    from spectralbte_b200.api import velocity_grids
    N, L_v, dt = 16, 5.0, 0.01
    cores = _use_all_cores()
    v, _ = velocity_grids(N, L_v)
    f = initial.init_hom(v, L_v, 2)
    W = orc.synthetic_weights(N) * 1e-3
    if orc.have_ref():
        kind = "reference"
        R = orc.Reference(N, L_v, 0)
        rows = R.rows(W)
        q = lambda x: R.conserve(R.compute_q_maxpreserve(rows, x, x))  # noqa: E731
    else:
        kind = "port"
        o = orc.Oracle(N, L_v, 0)
        q = lambda x: o.conserve(o.compute_q_maxpreserve(W, x, x))  # noqa: E731

    def step(x):
        f1 = x + dt * q(x)
        return 0.5 * (x + f1) + 0.5 * dt * q(f1)

    for _ in range(warmup):
        step(f)
    t0 = time.perf_counter()
    for _ in range(steps):
        step(f)
    sec = time.perf_counter() - t0
    return {"value": 6.0 * steps / sec, "unit": "evals/s", "cores": cores, "kind": kind, "seconds": sec,
            "sample": "%d RK2 steps (6 compute_Qhat each: 2 x ComputeQ_maxPreserve + conserveAllMoments) at N=16 after %d warm-up, "
                      "%d OpenMP threads" % (steps, warmup, cores)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    base = {"n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "higher_is_better": True, "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "impl": "reference", "gpu_launches": 0}
    if args.workload == "bkw16":
        r = cpu_bkw16(args.steps, args.warmup)
        line = dict(base, metric="Q(f,f) evals/s at N=16 (0D BKW, 6 compute_Qhat per RK2 step)", value=r["value"], unit="evals/s",
                    ms_per_step=1e3 * r["seconds"] / args.steps, scaling="weak", config=config_bkw16(),
                    cpu_baseline={k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
                    e2e={"value": r["value"], "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0})
        print(json.dumps(line))
        return
    if args.workload != "0d_n32":
        # 1D: the reference runs ComputeQ + conserveAllMoments cell after cell (exec/boltz.c:285-345); a step of
        # this arm is one cell advanced by one time step (order evaluations), timed value-independently on
        # synthetic weights, transport excluded (it is < 3 % of the reference's step)
        from oracle import oracle as orc
        from spectralbte_b200 import bench1d, initial
        from spectralbte_b200.api import velocity_grids
        cfg = bench1d.WORKLOADS[args.workload]
        N, L_v, stages = cfg["N"], cfg["L_v"], cfg["order"]
        v, _ = velocity_grids(N, L_v, True)
        f_cell = initial.init_inhom(v, cfg["ic"], 4, cfg["order"], 0, 4)[cfg["order"]].copy()
        W = orc.synthetic_weights(N) * 1e-3
        cb = cpu_cell_leg(N, L_v, W, f_cell, stages, evals=max(1, args.steps * stages))
        line = dict(base, metric="cells*steps/s (1D)", value=cb["value"], unit="cells*steps/s", ms_per_step=1e3 / cb["value"],
                    scaling=cfg["scaling"], config={"workload": cfg["desc"], "N": N, "L_v": L_v}, cpu_baseline=cb,
                    e2e={"value": cb["value"], "unit": "cells*steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0})
        print(json.dumps(line))
        return
    r = cpu_computeq_n32(args.steps, args.warmup)
    line = dict(base, metric=METRIC_0D, value=r["value"], unit="evals/s", ms_per_step=1e3 * r["seconds"] / args.steps,
                scaling="weak", config=config_0d(),
                cpu_baseline={"value": r["value"], "unit": "evals/s", "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]},
                e2e={"value": r["value"], "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0})
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------
def dist_setup():
    import torch
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        if not dist.is_initialized():
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    return world, rank, local


def max_over_ranks(x, world):
    if world == 1:
        return x
    import torch
    import torch.distributed as dist
    t = torch.tensor([x], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def barrier(world):
    import torch
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    torch.cuda.synchronize()


def run_0d_n32(args):
    import ctypes as C
    import torch
    import spectralbte_b200 as sb
    from spectralbte_b200 import initial
    world, rank, local = dist_setup()
    N, L_v = 32, 5.0
    n3 = N ** 3
    c = sb.Collisions(N, L_v, device=local)
    if args.weights == "synthetic":
        c.synthetic_weights(SEED)
        wdesc = "synthetic splitmix64(seed=%d)" % SEED
    else:
        c.generate_weights(1.0)   # hard spheres; N32_isotropic_L_v5_lambda1.wts content, generated on the device
        wdesc = "isotropic hard-sphere weights (lambda=1) generated on the device, src/weights.c:265-281"
    f = initial.init_hom(c.v, L_v, 0)
    df, dQ = c.array(n3).put(f), c.array(n3)
    stream = torch.cuda.ExternalStream(c.stream, device=torch.device("cuda", local))
    k2 = {"auto": sb.K2_AUTO, "stream": sb.K2_STREAM, "deep": sb.K2_STREAM_DEEP, "generic": sb.K2_GENERIC}[args.k2]

    def step():
        sb._lib.check(c.L.sbte_compute_q(c.h, df.ptr, df.ptr, dQ.ptr, 1, k2))

    def step_mp():
        sb._lib.check(c.L.sbte_compute_q_maxpreserve(c.h, df.ptr, df.ptr, dQ.ptr, k2))

    def timed(fn, n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for _ in range(n):
            fn()
        b.record(stream)
        c.sync()
        return a.elapsed_time(b)

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        sampler.wait_first_sample()
    for _ in range(args.warmup):
        step()
    c.sync()
    barrier(world)
    c.k2_profile(True)
    l0 = c.launches
    tw0 = time.perf_counter()
    ms = timed(step, args.steps)
    tw1 = time.perf_counter()
    barrier(world)
    sampler.mark(tw0, tw1)
    k2_ms, k2_n = c.k2_profile_read()
    c.k2_profile(False)
    launches = c.launches - l0
    ms = max_over_ranks(ms, world)

    # the same evaluation (i) without the transposed pairing: the symmetrised stream over every zeta column, and
    # (ii) with both switched off: the reference's own 8*N^6-byte formulation
    from spectralbte_b200 import bench1d
    xy_state, xy_dev = c.xy_pairing_state()
    plain = sym_only = None

    def variant(nbytes):
        for _ in range(3):
            step()
        c.sync()
        c.k2_profile(True)
        v_ms = timed(step, args.steps)
        vk_ms, vk_n = c.k2_profile_read()
        c.k2_profile(False)
        rec = {"evals_per_s": args.steps / (v_ms * 1e-3), "kernel_ms": vk_ms / max(1, vk_n), "bytes_per_launch": nbytes}
        rec["achieved_GBs"] = nbytes / (rec["kernel_ms"] * 1e-3) / 1e9
        return rec

    if not os.environ.get("SBTE_NO_SYM"):
        c.set_xy_pairing(False)
        if xy_state == 1 and not os.environ.get("SBTE_NO_XYSYM"):
            sym_only = variant(8.0 * float(N) ** 4 * bench1d.nrep_sum(N))
        c.set_symmetrize(False)
        plain = variant(8.0 * float(N) ** 6)
        c.set_symmetrize(True)
        c.set_xy_pairing(True)

    # the 0D driver's call: ComputeQ_maxPreserve = three reference evaluations folded into one weight pass
    for _ in range(3):
        step_mp()
    c.sync()
    mp_ms = timed(step_mp, args.steps) / args.steps

    # sustained: >= 2 s of each call back to back, clocks sampled over exactly that window (rank 0)
    sustained = None
    if rank == 0 and not args.no_sustained:
        sustained = {}
        for key, fn, per in (("computeq", step, 1.0), ("maxpreserve", step_mp, 3.0)):
            n, tot_ms, t0 = 0, 0.0, time.perf_counter()
            while tot_ms < 2000.0:
                tot_ms += timed(fn, 250)
                n += 250
            t1 = time.perf_counter()
            sustained[key] = {"calls": n, "seconds": tot_ms * 1e-3, "ms_per_call": tot_ms / n,
                              "reference_evals_per_s": per * n / (tot_ms * 1e-3), "clocks": sampler.digest(t0, t1)}
    barrier(world)

    # end to end: host buffers through the reference-facing ComputeQ entry (H2D f, D2H Q every step);
    # pinned buffers for the contract's e2e, and pageable ones as the reference driver's malloc'd arrays are
    dp = C.POINTER(C.c_double)
    fh = torch.from_numpy(f).pin_memory()
    Qh = torch.empty(n3, dtype=torch.float64).pin_memory()
    fp, Qp = C.cast(fh.data_ptr(), dp), C.cast(Qh.data_ptr(), dp)

    def e2e_loop(fptr, Qptr):
        for _ in range(max(2, args.warmup // 2)):
            sb._lib.check(c.L.sbte_compute_q_host(c.h, fptr, fptr, Qptr, k2))
        barrier(world)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            sb._lib.check(c.L.sbte_compute_q_host(c.h, fptr, fptr, Qptr, k2))
        c.sync()
        return max_over_ranks(time.perf_counter() - t0, world)

    e2e_s = e2e_loop(fp, Qp)
    checksum = float(Qh.sum().item())
    f_pg, Q_pg = f.copy(), np.empty(n3)
    e2e_pageable_s = e2e_loop(f_pg.ctypes.data_as(dp), Q_pg.ctypes.data_as(dp))
    clocks = sampler.stop() if rank == 0 else None

    peak, peak_src = peaks()
    sym = not os.environ.get("SBTE_NO_SYM")
    ref_bytes = 8.0 * float(N) ** 6
    # symmetrised stream (f == g): rows zeta read nrep(zeta_x) of their N xi_x planes, nrep = N/2+1 | N/2;
    # transposed pairing: only the zeta columns zx >= zy are streamed ((zx + 1) columns per zx)
    pairing = xy_state == 1 and not os.environ.get("SBTE_NO_XYSYM") and args.k2 in ("auto", "stream")
    planes = [(bench1d.sym_nrep(N, zx) if sym else N) for zx in range(N)]
    wbytes = 8.0 * float(N) ** 3 * sum((zx + 1 if pairing else N) * planes[zx] for zx in range(N))
    k2_avg_ms = k2_ms / max(1, k2_n)
    achieved = wbytes / (k2_avg_ms * 1e-3) / 1e9
    traffic, traffic_src = None, None
    tp = os.path.join(ROOT, "profiles", "k2_stream_traffic.json")
    if os.path.exists(tp):
        key = "dram_bytes_per_launch_pairing" if pairing else ("dram_bytes_per_launch_sym" if sym else "dram_bytes_per_launch")
        traffic = json.load(open(tp)).get(key)
        traffic_src = "static: profiles/k2_stream_traffic.json (one ncu --set full capture, dram__bytes_read.sum + dram__bytes_write.sum per launch; not measured in this run)"
    value = world * args.steps / (ms * 1e-3)
    line = None
    if rank == 0:
        line = {
            "metric": METRIC_0D, "value": value, "unit": "evals/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config_0d(),
            "detail": {"weights": wdesc, "k2": args.k2, "replicas": world,
                       "note": "0D does not shard: N GPUs = N independent replicas; the sharded 1D cases are under `oned`"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_source": traffic_src,
                         "kernel": ("qhat_stream_kernel<32,2,2,%s,16 warps,transposed pairing>" if pairing else
                                    "qhat_stream_kernel<32,1,2,%s>") % ("sym" if sym else "plain"),
                         "kernel_ms": k2_avg_ms, "kernel_share_of_step": k2_ms / ms, "algorithmic_bytes_per_launch": wbytes,
                         "reference_formulation_bytes": ref_bytes,
                         "note": ("f == g: the summand is symmetric under xi <-> zeta-xi, so the kernel streams the symmetrised "
                                  "tensor Ws = W + W o sigma over nrep(zeta_x) of N xi_x planes (4.43 GB at N=32) instead of the "
                                  "reference's 8*N^6 = 8.59 GB" +
                                  ("; and the bound tensor was checked to be invariant under x <-> y of both indices (relative "
                                   "deviation %.1e), so only the zeta columns zx >= zy are streamed (2.29 GB) and every weight "
                                   "serves column (zx,zy) with the spectrum and column (zy,zx) with the transposed spectrum" % xy_dev
                                   if pairing else "") +
                                  "; `sym_kernel` times the symmetrised stream over all columns, `plain_kernel` the unsymmetrised one")
                                 if sym else "unsymmetrised stream (SBTE_NO_SYM): 8*N^6 bytes per evaluation as in the reference",
                         "xy_pairing": {"state": xy_state, "relative_deviation": xy_dev, "in_use": bool(pairing)},
                         "sym_kernel": (dict(sym_only, frac=sym_only["achieved_GBs"] / peak) if sym_only else None),
                         "plain_kernel": (dict(plain, frac=plain["achieved_GBs"] / peak) if plain else None),
                         "peak_source": peak_src},
            "e2e": {"value": world * args.steps / e2e_s, "unit": "evals/s", "h2d_bytes_per_step": n3 * 8,
                    "d2h_bytes_per_step": n3 * 8, "api": "sbte_compute_q_host (the body of the drop-in ComputeQ), pinned host buffers",
                    "pageable_value": world * args.steps / e2e_pageable_s,
                    "pageable_note": "same call with pageable (malloc) buffers, as exec/boltz.c allocates f and Q",
                    "checksum": checksum},
            "gpu_launches": int(launches), "clocks": clocks,
            "maxpreserve": {"ms_per_call": mp_ms, "calls_per_s": 1e3 / mp_ms, "reference_evals_per_s": 3e3 / mp_ms,
                            "note": "ComputeQ_maxPreserve (src/collisions.c:178-210): 3 compute_Qhat of the reference in 1 weight pass"},
        }
        if sustained:
            line["sustained"] = sustained
    n32_dir = n32_info = None
    if rank == 0 and world == 1 and not args.no_dropin and not args.no_dropin_n32:
        import bench_dropin
        n32_dir, n32_info = bench_dropin.stage_n32(c)   # the bound weights as the reference's own .wts file
    # free the 17 GB of N=32 tensors before the 1D cases build theirs
    df.free()
    dQ.free()
    c.close()

    if not args.no_oned:
        R = bench1d.Runner()
        oned = bench1d.oned_records(R, max(10, min(args.steps, 40)), args.warmup,
                                    cpu_leg=None if (args.no_cpu or world > 1) else cpu_cell_leg, parity=not args.no_parity)
        if rank == 0:
            line["oned"] = oned
            line["halo_parity"] = oned["halo_parity"]
    if rank != 0:
        return
    if world == 1 and not args.no_dropin:
        import bench_dropin
        line["dropin"] = bench_dropin.run(ROOT, local, n32_dir=n32_dir, n32_info=n32_info)
    if world == 1 and not args.no_cpu:
        cb = cpu_computeq_n32(args.cpu_steps, 1)
        line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
    print(json.dumps(line))


def finalize():
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            dist.destroy_process_group()
    except Exception:
        pass


def run_bkw16(args):
    """BASELINE config 2 (input_examples/BKW16.in): 0D BKW relaxation, N=16, L_v=5, lambda=0, dt=0.01, RK2.
    One step = one time step of exec/boltz.c:189-241 = 2 x (ComputeQ_maxPreserve + conserveAllMoments) + updates
    = 6 compute_Qhat evaluations of the reference; f stays on the device."""
    import torch
    import spectralbte_b200 as sb
    from spectralbte_b200 import initial
    world, rank, local = dist_setup()
    N, L_v, lam, dt = 16, 5.0, 0.0, 0.01
    n3 = N ** 3
    c = sb.Collisions(N, L_v, device=local)
    c.generate_weights(lam)
    f0 = initial.init_hom(c.v, L_v, 2)
    df = c.array(n3).put(f0)
    stream = torch.cuda.ExternalStream(c.stream, device=torch.device("cuda", local))
    for _ in range(args.warmup):
        c.step_0d(df, dt, 1.0, 2)       # first call direct, second captured into a CUDA graph, then replays
    c.sync()
    barrier(world)
    l0 = c.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        c.step_0d(df, dt, 1.0, 2)
    e1.record(stream)
    c.sync()
    barrier(world)
    ms = max_over_ranks(e0.elapsed_time(e1), world)
    launches = c.launches - l0
    # the convolution kernel alone (profiling brackets every K2 launch with events, which disables the graph)
    c.k2_profile(True)
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record(stream)
    for _ in range(args.steps):
        c.step_0d(df, dt, 1.0, 2)
    p1.record(stream)
    c.sync()
    direct_ms = p0.elapsed_time(p1)
    k2_ms, k2_n = c.k2_profile_read()
    c.k2_profile(False)
    # end to end: f uploaded from pinned host memory and the output row downloaded every step
    fh = torch.from_numpy(f0).pin_memory()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        sb._lib.check(c.L.sbte_h2d(c.h, df.ptr, fh.data_ptr(), n3 * 8))
        c.step_0d(df, dt, 1.0, 2)
        row = c.row_0d(df)
    e2e_s = max_over_ranks(time.perf_counter() - t0, world)
    if rank != 0:
        return
    from spectralbte_b200 import bench1d
    peak, peak_src = peaks()
    wbytes = 8.0 * float(N) ** 4 * bench1d.nrep_sum(N)
    k2_avg = k2_ms / max(1, k2_n)
    l2_gbs = wbytes / (k2_avg * 1e-3) / 1e9
    line = {"metric": "Q(f,f) evals/s at N=16 (0D BKW, 6 compute_Qhat per RK2 step)", "value": 6.0 * world * args.steps / (ms * 1e-3),
            "unit": "evals/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_bkw16(), "detail": {"weights": "generated on the device (lambda = 0)", "replicas": world},
            # The step is neither HBM- nor FP64-bound: the 69 MB symmetrised tensor stays in L2 between its two passes and the
            # 24 launches of a step are latency-bound, so the figure that explains it is launches per step and the replay time,
            # not a DRAM fraction.  The contract's roofline object reports the weight stream against the HBM peak for what it is.
            "roofline": {"bound": "hbm", "achieved": l2_gbs, "peak": peak, "unit": "GB/s", "frac": l2_gbs / peak, "traffic": None,
                         "kernel": "qhat_stream_kernel<16,2,4,sym>", "kernel_ms": k2_avg,
                         "kernel_share_of_step": k2_ms / direct_ms, "algorithmic_bytes_per_launch": wbytes,
                         "note": "NOT a DRAM-bound step: the weight tensor is L2-resident (frac is the L2-fed stream rate over the HBM "
                                 "peak, given for the contract only); the step is launch/latency-bound -- see `latency`",
                         "peak_source": peak_src},
            "latency": {"launches_per_step": launches / args.steps, "graph_replay_ms_per_step": ms / args.steps,
                        "launch_by_launch_ms_per_step": direct_ms / args.steps, "convolution_ms_per_step": k2_ms / args.steps,
                        "non_convolution_ms_per_step": (direct_ms - k2_ms) / args.steps,
                        "weight_bytes_per_pass": wbytes, "l2_capacity_bytes": 126e6,
                        "note": "one RK2 step = 2 x (Maxwellian split, 3 forward transforms, 1 two-pair convolution, inverse transform, "
                                "conservation, update) replayed as one CUDA graph"},
            "e2e": {"value": 6.0 * world * args.steps / e2e_s, "unit": "evals/s", "h2d_bytes_per_step": n3 * 8,
                    "d2h_bytes_per_step": n3 * 8 + 64, "checksum": float(row[0])},
            "gpu_launches": int(launches)}
    if world == 1 and not args.no_cpu:
        cb = cpu_bkw16(max(3, args.cpu_steps), 1)
        line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
    print(json.dumps(line))


def main():
    from spectralbte_b200 import bench1d
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="0d_n32", choices=["0d_n32", "bkw16"] + sorted(bench1d.WORKLOADS))
    ap.add_argument("--k2", default="auto", choices=["auto", "stream", "deep", "generic"])
    ap.add_argument("--weights", default="generated", choices=["generated", "synthetic"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline legs")
    ap.add_argument("--no-oned", action="store_true", help="skip the sharded 1D sub-record of the default workload")
    ap.add_argument("--no-parity", action="store_true", help="skip the cross-GPU halo parity check (N > 1)")
    ap.add_argument("--no-sustained", action="store_true", help="skip the 2 x 2 s sustained legs")
    ap.add_argument("--no-dropin", action="store_true", help="skip the drop-in executable timing (N = 1)")
    ap.add_argument("--no-dropin-n32", action="store_true", help="drop-in timing without the N=32 run (8.59 GB weight file)")
    ap.add_argument("--cpu-steps", type=int, default=20)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        run_reference(args)
        return
    if args.workload == "0d_n32":
        run_0d_n32(args)
    elif args.workload == "bkw16":
        run_bkw16(args)
    else:
        bench1d.run(args, ROOT, cpu_leg=cpu_cell_leg, sampler_cls=ClockSampler)
    finalize()


if __name__ == "__main__":
    main()
