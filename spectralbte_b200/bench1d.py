"""1D-3V benchmark workload (bench.py --workload shock1p2): Mach-1.2 shock derived from
/root/reference/input_examples/Shock1p2.in (the shipped file is unstable and its 601-cell
mesh is prime).  N=16, L_v=9, Kn=1.52, lambda=1, Init_field 6, Space_order 2, dt=1e-3, dx=6/640;
640 cells per GPU (weak scaling: the domain grows with the rank count), one process per GPU, ghost
cells exchanged over NCCL."""
import json
import os
import time

import numpy as np


def run(args, root):
    import torch
    import torch.distributed as dist
    import spectralbte_b200 as sb
    from spectralbte_b200 import halo as H
    from spectralbte_b200 import initial

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    N, L_v, Kn, order, ic, dt = 16, 9.0, 1.52, 2, 6, 1e-3
    cells_per_gpu = int(os.environ.get("SBTE_CELLS_PER_GPU", "640"))
    nX = cells_per_gpu * world
    _, x, dx = initial.make_mesh([nX], [6.0 * nX / 640.0], order)
    lo, hi = initial.partition(nX, world)[rank]
    c = sb.Collisions(N, L_v, inhomogeneous=True, device=local)
    wfile = os.environ.get("SBTE_WEIGHTS")
    if wfile:
        c.load_weights(wfile)
        wdesc = wfile
    elif os.environ.get("SBTE_SYNTHETIC_WEIGHTS"):
        c.synthetic_weights(20261017)
        wdesc = "synthetic splitmix64"
    else:
        c.generate_weights(1.0)   # hard spheres, generated on the device (src/weights.c:265-281)
        wdesc = "isotropic lambda=1, generated on device (adaptive GK21)"
    s = sb.Slab(c, hi - lo, order, x[lo:hi + 2 * order].copy(), dx[lo:hi + 2 * order].copy(), ic, dt, rank, world)
    s.upload(initial.init_inhom(c.v, ic, nX, order, lo, hi))
    halo = H.SlabHalo(s, dev)
    stream = halo.stream

    def sync_all():
        c.sync()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        H.step(s, halo, Kn, ic)
    sync_all()
    c.k2_profile(True)
    l0 = c.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        H.step(s, halo, Kn, ic)
    e1.record(stream)
    sync_all()
    ms = e0.elapsed_time(e1)
    k2_ms, k2_n = c.k2_profile_read()
    c.k2_profile(False)
    launches = c.launches - l0
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())

    # end to end: the step's input slab comes from pinned host memory and the moments go back
    ncell = hi - lo + 2 * order
    host = torch.from_numpy(s.download()).pin_memory()
    mom = np.empty((hi - lo, 8))
    import ctypes as C
    t0 = time.perf_counter()
    for _ in range(max(1, args.steps // 4)):
        sb._lib.check(c.L.sbte_slab_upload(s.h, C.cast(host.data_ptr(), C.POINTER(C.c_double))))
        H.step(s, halo, Kn, ic)
        mom = s.moments()
    c.sync()
    e2e_s = (time.perf_counter() - t0) / max(1, args.steps // 4)
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    if rank != 0:
        return
    evals = 2.0 * (hi - lo) * args.steps          # cell evaluations on this rank (2 RK stages)
    flops = 10.0 * float(N) ** 6 * evals
    line = {
        "metric": "cells*steps/s (1D)", "value": nX * args.steps / (ms * 1e-3), "unit": "cells*steps/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic (Mach-1.2 shock initial data of the reference, src/initializer.c:341-351,405-420)",
        "config": {"workload": "shock1p2-derived: 1D-3V Mach 1.2 shock, N=16, Space_order 2, %d cells/GPU" % cells_per_gpu,
                   "N": N, "L_v": L_v, "Kn": Kn, "dt": dt, "cells_total": nX, "weights": wdesc,
                   "l2": "per-step working set (slabs+spectra+weights > 400 MB) larger than L2; no flush"},
        "roofline": {"bound": "fp64", "achieved": flops / (k2_ms * 1e-3) / 1e12, "peak": 37.0, "unit": "TFLOP/s",
                     "frac": flops / (k2_ms * 1e-3) / 1e12 / 37.0, "traffic": None, "kernel": "qhat_batch_kernel<16>",
                     "kernel_ms": k2_ms / max(1, k2_n), "kernel_share_of_step": k2_ms / ms,
                     "peak_source": "datasheet FP64 vector (not in MEASURED_PEAKS.json); 10 counted flops per 6 FP64 instructions"},
        "e2e": {"value": nX / e2e_s, "unit": "cells*steps/s", "h2d_bytes_per_step": ncell * N ** 3 * 8,
                "d2h_bytes_per_step": int(mom.nbytes), "checksum": float(mom[:, 0].sum())},
        "gpu_launches": int(launches),
    }
    print(json.dumps(line))
