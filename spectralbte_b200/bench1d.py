"""1D-3V benchmark cases (bench.py: the `oned` sub-record of the default line and --workload NAME), one process
per GPU, spatial cells block-partitioned over the ranks (uneven blocks allowed), weights replicated, ghost
cells read from the neighbour GPU's memory over NVLink (peer-memory halo) or exchanged over NCCL
(spectralbte_b200/halo.py).  One step = one full time step of exec/boltz.c:264-353.

Cases (the shipped Shock1p2.in is unstable and its 601-cell mesh is prime):
  shock_strong       Shock1p2-derived: N=16, L_v=9, Kn=1.52, lambda=1, Init_field 6, Space_order 2 (minmod),
                     dt=1e-3, 640 cells on [0,6] IN TOTAL, sharded over the ranks (strong scaling)
  shock_strong_601   the same on the shipped 601-cell mesh (uneven blocks)
  shock_weak         the same with 640 cells PER GPU (the domain grows with the ranks)
  heattrans_strong   input_examples/heatTrans.long.in at BASELINE's N=24: Kn=0.3, Init_field 3 (diffuse walls),
                     Space_order 1, dt=1e-4, 250 cells in total (31-32 per GPU at 8)
  heattrans22_strong the same file as shipped (N=22)
"""
import ctypes as C
import hashlib
import json
import os
import time

import numpy as np

FP64_PEAK_TFLOPS = 36.5
FP64_PEAK_SOURCE = ("builder-measured with tools/micro/dfma_rf.cu on this pool's B200 (DMUL stream 36.3-37.0 TFLOP/s-equivalent, "
                    "profiles/r01_micro_dfma_register_reads.txt; datasheet 37); not in MEASURED_PEAKS.json")

_SHOCK = dict(N=16, L_v=9.0, Kn=1.52, lam=1.0, order=2, ic=6, dt=1e-3)
_HEAT = dict(L_v=9.0, Kn=0.3, lam=1.0, order=1, ic=3, dt=1e-4)
CASES = {
    "shock_strong": dict(_SHOCK, cells_per_gpu=None, total_cells=640, length=6.0, scaling="strong",
                         desc="shock1p2-derived: 1D-3V Mach 1.2 shock, N=16, Space_order 2 (minmod), 640 cells in total"),
    "shock_strong_601": dict(_SHOCK, cells_per_gpu=None, total_cells=601, length=6.0, scaling="strong",
                             desc="shock1p2-derived on the shipped 601-cell mesh (uneven blocks), N=16, Space_order 2"),
    "shock_weak": dict(_SHOCK, cells_per_gpu=640, total_cells=None, length=None, scaling="weak",
                       desc="shock1p2-derived: 1D-3V Mach 1.2 shock, N=16, Space_order 2 (minmod), 640 cells per GPU"),
    "heattrans_strong": dict(_HEAT, N=24, cells_per_gpu=None, total_cells=250, length=1.0, scaling="strong",
                             desc="heatTrans.long: 1D-3V heat transfer between diffuse walls, N=24, Space_order 1, 250 cells in total"),
    "heattrans22_strong": dict(_HEAT, N=22, cells_per_gpu=None, total_cells=250, length=1.0, scaling="strong",
                               desc="heatTrans.long as shipped: N=22, Space_order 1, 250 cells in total"),
}
ALIASES = {"shock1p2": "shock_weak", "heattrans": "heattrans_strong", "heattrans22": "heattrans22_strong"}
WORKLOADS = dict(CASES, **{a: CASES[c] for a, c in ALIASES.items()})   # bench.py --workload / --impl reference


def k2_name(N):
    """The convolution kernel csrc/qhat_batch.cu picks for N (launch_qhat_batch2 / launch_qhat_batch_any)."""
    if N in (8, 16):
        return "qhat_batch2_kernel<%d>" % N
    if N == 24 or (N in (20, 22) and not os.environ.get("SBTE_NO_BATCH3G")):
        return "qhat_batch3_kernel<%d>" % N
    return "qhat_batch_any_kernel"


def sym_nrep(N, zx):
    """Representative xi_x planes of the symmetrised tensor for rows with this zeta_x (csrc/common.cuh sym_nrep)."""
    a = (zx + N // 2) % N
    return a // 2 + 1 + (a + N) // 2 - a


def nrep_sum(N):
    """Sum over zeta_x of the representative xi_x planes the symmetrised tensor keeps (csrc/common.cuh sym_nrep)."""
    return sum(((zx + N // 2) % N) // 2 + 1 + (((zx + N // 2) % N) + N) // 2 - ((zx + N // 2) % N) for zx in range(N))


def _digest(cell):
    return hashlib.blake2b(np.ascontiguousarray(cell).tobytes(), digest_size=16).digest()


class Runner:
    """The process-group view of this rank and one Collisions context (velocity grid + replicated weights) per N."""

    def __init__(self):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1 and not dist.is_initialized():
            dist.init_process_group("nccl", device_id=self.dev)
        self.ctx = {}

    def collisions(self, N, L_v, lam):
        import spectralbte_b200 as sb
        key = (N, L_v, lam)
        if key not in self.ctx:
            c = sb.Collisions(N, L_v, inhomogeneous=True, device=self.local)
            wfile = os.environ.get("SBTE_WEIGHTS")
            if wfile:
                c.load_weights(wfile)
                wdesc = wfile
            elif os.environ.get("SBTE_SYNTHETIC_WEIGHTS"):
                c.synthetic_weights(20261017)
                wdesc = "synthetic splitmix64"
            else:
                c.generate_weights(lam)   # generated on the device (src/weights.c:265-281)
                wdesc = "isotropic lambda=%g, generated on device (adaptive GK21)" % lam
            self.ctx[key] = (c, wdesc)
        return self.ctx[key]

    def release(self, N=None):
        for key in [k for k in self.ctx if N is None or k[0] == N]:
            self.ctx.pop(key)[0].close()

    def sync_all(self, c):
        c.sync()
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def _reduce(self, x, op):
        if self.world == 1:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=op)
        return float(t.item())

    def max_over_ranks(self, x):
        return self._reduce(x, self.dist.ReduceOp.MAX)

    def min_over_ranks(self, x):
        return self._reduce(x, self.dist.ReduceOp.MIN)

    # ------------------------------------------------------------------------------------------------------
    def build(self, cfg, halo_mode=None, alone=False):
        """Mesh, block partition, slab holding the reference's initial data, halo object.  alone=True builds the whole
        mesh on this rank (the single-rank run of the parity check)."""
        import spectralbte_b200 as sb
        from spectralbte_b200 import halo as H
        from spectralbte_b200 import initial
        world, rank = (1, 0) if alone else (self.world, self.rank)
        N, order, ic = cfg["N"], cfg["order"], cfg["ic"]
        if cfg["total_cells"] is None:
            per = int(os.environ.get("SBTE_CELLS_PER_GPU", cfg["cells_per_gpu"]))
            nX = per * self.world
            length = 6.0 / 640.0 * nX
        else:
            nX = int(os.environ.get("SBTE_TOTAL_CELLS", cfg["total_cells"]))
            length = cfg["length"]
        _, x, dx = initial.make_mesh([nX], [length], order)
        part = initial.partition(nX, world)
        lo, hi = part[rank]
        c, wdesc = self.collisions(N, cfg["L_v"], cfg["lam"])
        s = sb.Slab(c, hi - lo, order, x[lo:hi + 2 * order].copy(), dx[lo:hi + 2 * order].copy(), ic, cfg["dt"], rank, world)
        s.upload(initial.init_inhom(c.v, ic, nX, order, lo, hi))
        halo = H.SlabHalo(s, self.dev, mode=halo_mode)
        return dict(c=c, s=s, halo=halo, nX=nX, lo=lo, hi=hi, part=part, wdesc=wdesc)

    def run_case(self, name, steps, warmup, halo_mode=None, e2e=True, cpu_leg=None, sampler=None):
        """Times `steps` steps of one case on all ranks; returns the record on rank 0 (None elsewhere)."""
        import spectralbte_b200 as sb
        from spectralbte_b200 import halo as H
        torch = self.torch
        cfg = WORKLOADS[name]
        b = self.build(cfg, halo_mode)
        c, s, halo, nX, lo, hi = b["c"], b["s"], b["halo"], b["nX"], b["lo"], b["hi"]
        N, order, ic, Kn = cfg["N"], cfg["order"], cfg["ic"], cfg["Kn"]
        stream = halo.stream
        for _ in range(warmup):
            H.step(s, halo, Kn, ic)
        self.sync_all(c)
        l0 = c.launches
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        tw0 = time.perf_counter()
        e0.record(stream)
        for _ in range(steps):
            H.step(s, halo, Kn, ic)
        e1.record(stream)
        self.sync_all(c)
        tw1 = time.perf_counter()
        if sampler:
            sampler.mark(tw0, tw1)
        ms = self.max_over_ranks(e0.elapsed_time(e1))
        launches = c.launches - l0
        # The timed steps replay a CUDA graph where they can (one rank, or peer-memory halos); CUDA events cannot sit
        # between the nodes of a replayed graph, so the convolution kernel is timed over the same number of identical
        # steps issued launch by launch right after.
        c.k2_profile(True)
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        p0.record(stream)
        for _ in range(steps):
            H.step(s, halo, Kn, ic)
        p1.record(stream)
        self.sync_all(c)
        prof_ms = p0.elapsed_time(p1)
        k2_ms, k2_n = c.k2_profile_read()
        c.k2_profile(False)
        k2_ms_max = self.max_over_ranks(k2_ms)

        e2e_rec = None
        if e2e:   # end to end: the step's input slab comes from pinned host memory and the moments go back
            ncell = hi - lo + 2 * order
            host = torch.from_numpy(s.download()).pin_memory()
            mom = np.empty((hi - lo, 8))
            reps = max(3, steps // 4)
            self.sync_all(c)
            t0 = time.perf_counter()
            for _ in range(reps):
                sb._lib.check(c.L.sbte_slab_upload(s.h, C.cast(host.data_ptr(), C.POINTER(C.c_double))))
                H.step(s, halo, Kn, ic)
                mom = s.moments()
            c.sync()
            e2e_s = self.max_over_ranks((time.perf_counter() - t0) / reps)
            e2e_rec = {"value": nX / e2e_s, "unit": "cells*steps/s", "h2d_bytes_per_step": ncell * N ** 3 * 8,
                       "d2h_bytes_per_step": int(mom.nbytes), "steps": reps, "checksum": float(mom[:, 0].sum()),
                       "api": "sbte_slab_upload + sbte_slab_step + sbte_slab_moments on every rank (bytes are rank 0's)"}
        want_cpu = cpu_leg is not None and self.rank == 0
        f_cell = s.download()[order + (hi - lo) // 2].copy() if want_cpu else None
        W_host = c.weights_to_host() if want_cpu else None
        mode = "none (one rank)" if self.world == 1 else ("peer memory: stencils read neighbour cells over NVLink" if halo.mode == "p2p"
                                                           else "NCCL batch_isend_irecv")
        halo_state = s.halo_state() if (self.world > 1 and halo.mode == "p2p") else None
        self.sync_all(c)
        halo.close()
        s.close()
        if self.rank != 0:
            return None
        stages = order                               # collision evaluations per cell per step (Euler / Heun)
        cells_local = hi - lo
        ref_flops = 10.0 * float(N) ** 6 * stages * cells_local * steps       # the reference's N^6 pair sum, rank 0's cells
        sym = not os.environ.get("SBTE_NO_SYM")
        flops = ref_flops * (nrep_sum(N) / float(N * N)) if sym else ref_flops  # pairs actually visited (f == g symmetry)
        ach = flops / (k2_ms * 1e-3) / 1e12
        rec = {
            "value": nX * steps / (ms * 1e-3), "unit": "cells*steps/s", "scaling": cfg["scaling"], "n_gpus": self.world,
            "cells_total": nX, "cells_per_gpu": [h - l for l, h in b["part"]], "steps": steps, "warmup": warmup,
            "ms_per_step": ms / steps, "halo": mode,
            "kernel_ms": k2_ms_max / steps, "non_kernel_ms": ms / steps - k2_ms_max / steps,
            "config": {"workload": cfg["desc"], "N": N, "L_v": cfg["L_v"], "Kn": Kn, "dt": cfg["dt"], "order": order,
                       "init_field": ic, "weights": b["wdesc"],
                       "l2": "per-step working set (slabs + spectra + weights) larger than L2; no flush"},
            "roofline": {"bound": "fp64", "achieved": ach, "peak": FP64_PEAK_TFLOPS, "unit": "TFLOP/s", "frac": ach / FP64_PEAK_TFLOPS,
                         "traffic": None, "kernel": k2_name(N), "kernel_ms": k2_ms / max(1, k2_n),
                         "kernel_share_of_step": k2_ms_max / ms, "launch_by_launch_ms_per_step": prof_ms / steps,
                         "reference_equivalent_tflops": ref_flops / (k2_ms * 1e-3) / 1e12,
                         "note": "rank 0's kernel: achieved counts 10 flops per (weight, cell) pair actually visited; with f == g only "
                                 "nrep(zeta_x)/N of the reference's N^6 pairs are visited (symmetrised weights); "
                                 "reference_equivalent_tflops counts all N^6 pairs; 6 FP64 instructions are issued per 10 counted flops",
                         "peak_source": FP64_PEAK_SOURCE},
            "gpu_launches": int(launches),
        }
        if halo_state is not None:
            rec["halo_flags_rank0"] = dict(zip(("ready", "done", "epoch", "error"), halo_state))
        if e2e_rec:
            rec["e2e"] = e2e_rec
        if want_cpu:
            rec["cpu_baseline"] = cpu_leg(N, cfg["L_v"], W_host, f_cell, stages)
        return rec

    # ------------------------------------------------------------------------------------------------------
    def parity_case(self, name, steps, halo_mode):
        """Rank-count invariance across real GPUs: the sharded run (this process group, `halo_mode`) against the same
        mesh stepped on rank 0 alone, compared cell by cell through digests of the raw bytes."""
        from spectralbte_b200 import halo as H
        cfg = WORKLOADS[name]
        b = self.build(cfg, halo_mode)
        c, s, halo, order = b["c"], b["s"], b["halo"], cfg["order"]
        for _ in range(steps):
            H.step(s, halo, cfg["Kn"], cfg["ic"])
        self.sync_all(c)
        mine = s.download()[order:order + b["hi"] - b["lo"]]
        mode = halo.mode
        err = s.halo_state()[3] if mode == "p2p" else 0
        self.sync_all(c)
        halo.close()
        s.close()
        digests = [None]
        if self.rank == 0:
            one = self.build(cfg, None, alone=True)
            for _ in range(steps):
                one["s"].step(cfg["Kn"])
            c.sync()
            whole = one["s"].download()[order:order + one["nX"]]
            one["s"].close()
            digests = [[_digest(whole[l]) for l in range(one["nX"])]]
        self.dist.broadcast_object_list(digests, src=0)
        ok = all(_digest(mine[l]) == digests[0][b["lo"] + l] for l in range(b["hi"] - b["lo"]))
        ok = ok and err == 0 and bool(np.isfinite(mine).all())
        ok = self.min_over_ranks(1.0 if ok else 0.0) == 1.0
        return {"case": name, "halo": mode, "requested": halo_mode, "cells_total": b["nX"],
                "cells_per_gpu": [h - l for l, h in b["part"]], "steps": steps, "equal": bool(ok)}

    def halo_parity(self, steps=3):
        """Both halo modes on both uneven partitions (601-cell shock at order 2, 250-cell heat transfer at order 1)."""
        if self.world == 1:
            return "n/a (one rank)", []
        out = [self.parity_case(name, steps, mode) for name in ("shock_strong_601", "heattrans_strong") for mode in ("p2p", "nccl")]
        return ("bit-identical" if all(r["equal"] for r in out) else "MISMATCH"), out


def run(args, root, cpu_leg=None, sampler_cls=None):
    """bench.py --workload NAME: one 1D case as the headline line."""
    R = Runner()
    name = ALIASES.get(args.workload, args.workload)
    sampler = sampler_cls(R.local) if (sampler_cls is not None and R.rank == 0) else None
    if sampler:
        sampler.start()
        sampler.wait_first_sample()
    rec = R.run_case(name, args.steps, args.warmup, cpu_leg=None if (R.world > 1 or args.no_cpu) else cpu_leg, sampler=sampler)
    parity, detail = R.halo_parity() if (R.world > 1 and not args.no_parity) else ("not run", [])
    clocks = sampler.stop() if sampler else None
    R.release()
    if R.rank != 0:
        return
    line = {"metric": "cells*steps/s (1D)", "value": rec["value"], "unit": "cells*steps/s", "n_gpus": R.world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": rec["ms_per_step"], "higher_is_better": True, "scaling": rec["scaling"],
            "vs_baseline": None, "dtype": "f64",
            "data": "synthetic (the reference's own initial data, src/initializer.c:330-351,391-420)",
            "config": dict(rec["config"], case=name, cells_total=rec["cells_total"], cells_per_gpu=rec["cells_per_gpu"], halo=rec["halo"]),
            "roofline": rec["roofline"], "e2e": rec["e2e"], "gpu_launches": rec["gpu_launches"], "clocks": clocks,
            "kernel_ms": rec["kernel_ms"], "non_kernel_ms": rec["non_kernel_ms"], "halo_parity": parity, "halo_parity_detail": detail}
    if "cpu_baseline" in rec:
        line["cpu_baseline"] = rec["cpu_baseline"]
    print(json.dumps(line))


def oned_records(R, steps, warmup, cpu_leg=None, parity=True):
    """The `oned` sub-record of the default bench line: every case at this world size, then the cross-GPU parity check.
    Returns the dict on rank 0, None elsewhere."""
    out = {"unit": "cells*steps/s"}
    for name in ("shock_strong", "shock_strong_601", "shock_weak", "heattrans_strong", "heattrans22_strong"):
        if name == "shock_weak" and R.world == 1:
            continue   # on one GPU it is the same 640-cell run as shock_strong
        cfg = CASES[name]
        st = steps if cfg["N"] == 16 else max(10, steps // 2)
        leg = cpu_leg if (R.world == 1 and name != "shock_strong_601") else None
        rec = R.run_case(name, st, warmup, cpu_leg=leg)
        if R.rank == 0:
            out[name] = rec
    if R.world == 1 and R.rank == 0:
        out["shock_weak"] = dict(out["shock_strong"], scaling="weak", note="one GPU: the same 640-cell run as shock_strong")
    status, detail = R.halo_parity() if parity else ("not run", [])
    out["halo_parity"] = status
    out["halo_parity_detail"] = detail
    R.release()
    return out if R.rank == 0 else None
