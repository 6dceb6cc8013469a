"""Drop-in (Level 1) end-to-end timing: the reference's OWN driver, exec/boltz.c compiled unmodified, linked against
libsbte_b200.so (oracle/_ref/boltz_gpu) next to the reference executable (oracle/_ref/boltz_), both started exactly as
tests/run_test.sh starts boltz_ (same working directory layout, same input and weight files).

What is measured is what a maintainer gets by relinking and nothing else: per-cell ComputeQ with batch 1
(exec/boltz.c:288), pageable host buffers (malloc'd f / Q, :196-232), the host gather/scatter of advectOne/advectTwo,
process start, CUDA context creation and the one-time upload of the weight rows.  `wall_s` is the whole process;
`step_s_median` (0D runs only) is the median of the driver's own "Time elapsed" prints (:203-204, wall clock of one
ComputeQ_maxPreserve + conserveAllMoments; the 1D branch prints clock() differences per cell, which are not comparable).

Both executables are test infrastructure built by oracle/build_ref.sh; this module is part of bench.py (its reference /
cpu_baseline legs are the one place outside tests/ allowed to execute oracle/) and lives beside it, not in the package."""
import lzma
import os
import re
import shutil
import subprocess
import tempfile
import time

N32_INPUT = """N
32
L_v
5.0
Knudsen
1.0
Lambda
1.0
Time_step
0.01
Number_of_time_steps
%d
Space_order
2
Data_writing_frequency
%d
Restart
0
Restart_time
0
Init_field
0
SpaceInhom
0
Recompute_weights
0
Anisotropic
0
num_species
1
default
Stop
"""


N32_OUTPUT = "density\n1\nvelocity\n1\ntemperature\n1\npressure\n1\nmarginal\n0\nslice\n1\nentropy\n1\nStop\n"


def _time_exe(exe, cwd, args, env, timeout, parse_steps=True):
    t0 = time.perf_counter()
    r = subprocess.run([exe] + args, cwd=cwd, capture_output=True, text=True, timeout=timeout, env=env)
    wall = time.perf_counter() - t0
    if r.returncode != 0:
        return {"error": "exit %d: %s" % (r.returncode, (r.stdout + r.stderr)[-300:])}
    el = [float(x) for x in re.findall(r"Time elapsed: ([0-9.eE+-]+)", r.stdout)]
    out = {"wall_s": wall}
    tally = re.findall(r"libsbte_b200 timing: (.+?)\s+(\d+) calls\s+([0-9.]+) s", r.stdout)
    if tally:
        out["library_time"] = {name.strip(): {"calls": int(n), "seconds": float(sec)} for name, n, sec in tally}
    if el and parse_steps:
        out["steps"] = len(el)
        s = sorted(el)
        out["step_s_median"] = s[len(s) // 2]
        out["step_s_first"] = el[0]
    return out


def _prepare(tmp, golden, name, wts):
    for d in ("input", "Data", "Weights", "Restart"):
        os.makedirs(os.path.join(tmp, d), exist_ok=True)
    for fn in os.listdir(os.path.join(golden, "inputs")):
        if fn.startswith(name):
            shutil.copy(os.path.join(golden, "inputs", fn), os.path.join(tmp, "input", fn))
    raw = lzma.decompress(open(os.path.join(golden, wts + ".xz"), "rb").read())
    with open(os.path.join(tmp, "Weights", wts), "wb") as fh:
        fh.write(raw)


def stage_n32(coll):
    """Scratch run directory holding the bound N=32 weights of `coll` as Weights/N32_isotropic_L_v5_lambda1.wts (the
    reference's format, src/weights.c:100-103); returns (dir, seconds) or (None, reason)."""
    from spectralbte_b200.api import weights_filename
    tmp = tempfile.mkdtemp(prefix="sbte_dropin_n32_", dir=os.environ.get("SBTE_SCRATCH", "/tmp"))
    if shutil.disk_usage(tmp).free < 10 * 2 ** 30:
        shutil.rmtree(tmp, ignore_errors=True)
        return None, "less than 10 GiB free under %s for the 8.59 GB weight file" % os.path.dirname(tmp)
    for d in ("input", "Data", "Weights", "Restart"):
        os.makedirs(os.path.join(tmp, d), exist_ok=True)
    t0 = time.perf_counter()
    coll.save_weights(os.path.join(tmp, weights_filename(32, 5.0, 1.0)))
    return tmp, time.perf_counter() - t0


def run(root, device=0, n32_dir=None, n32_info=None, n32_steps=5):
    ref = os.path.join(root, "oracle", "_ref", "boltz_")
    gpu = os.path.join(root, "oracle", "_ref", "boltz_gpu")
    golden = os.path.join(root, "tests", "golden")
    if not (os.path.exists(ref) and os.path.exists(gpu)):
        return {"unavailable": "oracle/_ref/boltz_ or boltz_gpu not built (oracle/build_ref.sh)"}
    cores = os.cpu_count() or 1
    env = dict(os.environ, OMP_NUM_THREADS=str(cores), SBTE_DEVICE=str(device), SBTE_DROPIN_TIMING="1")
    for k in ("WORLD_SIZE", "RANK", "LOCAL_RANK"):
        env.pop(k, None)
    out = {"cores": cores,
           "note": "whole-process wall time of the reference's own exec/boltz.c: `reference` = built from its own collision / "
                   "conservation / transport sources, `b200` = the same driver objects linked against libsbte_b200.so; "
                   "step_s_median (0D) = the driver's own per-step print: wall seconds of ComputeQ_maxPreserve + conserveAllMoments"}
    for name, wts in (("BKW8", "N8_isotropic_L_v5_lambda0.wts"), ("heat_transport", "N8_isotropic_L_v9_lambda1.wts")):
        rec = {}
        for key, exe in (("reference", ref), ("b200", gpu)):
            tmp = tempfile.mkdtemp(prefix="sbte_dropin_")
            try:
                _prepare(tmp, golden, name, wts)
                rec[key] = _time_exe(exe, tmp, [name + ".test.in", name + ".test.out"], env, 900, parse_steps=(name == "BKW8"))
            finally:
                shutil.rmtree(tmp, ignore_errors=True)
        if "wall_s" in rec["reference"] and "wall_s" in rec["b200"]:
            rec["wall_ratio"] = rec["reference"]["wall_s"] / rec["b200"]["wall_s"]
        out[name] = rec
    if n32_dir:
        out["0d_n32"] = _run_n32(n32_dir, ref, gpu, env, n32_steps, n32_info)
    elif n32_info:
        out["0d_n32"] = {"unavailable": n32_info}
    return out


def _run_n32(tmp, ref, gpu, env, steps, save_s):
    """0D hard spheres at N=32 (BASELINE config 3), `steps` RK2 steps = 2*steps ComputeQ_maxPreserve calls, both
    executables reading the SAME 8.59 GB weight file staged by stage_n32()."""
    rec = {"steps": steps, "computeq_maxpreserve_calls": 2 * steps, "weights_save_s": save_s}
    try:
        with open(os.path.join(tmp, "input", "n32.in"), "w") as fh:
            fh.write(N32_INPUT % (steps, steps))
        with open(os.path.join(tmp, "input", "n32.out"), "w") as fh:   # the output-flags file (src/output.c)
            fh.write(N32_OUTPUT)
        for key, exe in (("b200", gpu), ("reference", ref)):
            rec[key] = _time_exe(exe, tmp, ["n32.in", "n32.out"], env, 1200)
        a, b = rec["reference"], rec["b200"]
        if "wall_s" in a and "wall_s" in b:
            rec["wall_ratio"] = a["wall_s"] / b["wall_s"]
            if "step_s_median" in a and "step_s_median" in b and b["step_s_median"] > 0:
                rec["step_ratio"] = a["step_s_median"] / b["step_s_median"]
            rec["note"] = ("wall_s includes reading the 8.59 GB weight file (both) and its one-time upload (b200); "
                           "step_s_* is one ComputeQ_maxPreserve + conserveAllMoments through the drop-in symbols with "
                           "pageable host buffers")
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return rec
