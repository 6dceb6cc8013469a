"""Builds libsbte_b200.so (CUDA kernels + C ABI) in-tree with nvcc for sm_100a.

    python -m spectralbte_b200.build [--force] [--verbose]

The shared library lands next to this file (spectralbte_b200/libsbte_b200.so) so that it travels to
the GPU box with the repository snapshot.  nvcc cross-compiles without a GPU.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "csrc", "_obj")
LIB = os.path.join(HERE, "libsbte_b200.so")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
          "--expt-relaxed-constexpr"]
# elementwise / reduction kernels keep the reference's operation-by-operation rounding (no FMA
# contraction); the FP64-bound convolution and the DFTs use FMAs.
PER_FILE = {"transport.cu": ["-fmad=false"], "conserve.cu": ["-fmad=false"], "weightgen.cu": ["-fmad=false"]}
SOURCES = ["capi.cu", "dropin.cu", "slab.cu", "fft.cu", "qhat.cu", "qhat_batch.cu", "conserve.cu", "transport.cu", "weightgen.cu"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    headers = [os.path.join(CSRC, h) for h in os.listdir(CSRC) if h.endswith((".h", ".cuh"))]
    headers.append(os.path.join(os.path.dirname(HERE), "include", "sbte_b200.h"))
    nvcc = _nvcc()
    jobs = []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ, src.replace(".cu", ".o"))
        if force or _stale(o, [s] + headers):
            cmd = [nvcc] + ARCH + COMMON + PER_FILE.get(src, []) + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o]
            jobs.append(cmd)

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed: %s\n%s\n%s" % (" ".join(cmd), r.stdout, r.stderr))
        return r.stderr

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            for log in ex.map(run, jobs):
                if verbose and log:
                    sys.stderr.write(log)
    objs = [os.path.join(OBJ, s.replace(".cu", ".o")) for s in SOURCES]
    if force or jobs or _stale(LIB, objs):
        cmd = [nvcc] + ARCH + ["-shared", "-o", LIB] + objs + ["-Xcompiler", "-fvisibility=hidden"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed: %s\n%s" % (r.stdout, r.stderr))
    build_host(force or bool(jobs))
    return LIB


HOST_SRC = os.path.join(HERE, "host", "boltz_b200.c")
HOST_BIN = os.path.join(HERE, "host", "boltz_b200")


def build_host(force=False):
    """The C host driver (reference-style command line, device-resident time loop)."""
    inc = os.path.join(os.path.dirname(HERE), "include")
    if force or _stale(HOST_BIN, [HOST_SRC, os.path.join(inc, "sbte_b200.h"), LIB]):
        cmd = ["gcc", "-std=gnu99", "-O2", "-Wall", "-I", inc, HOST_SRC, "-L", HERE, "-lsbte_b200",
               "-Wl,-rpath,$ORIGIN/..", "-lm", "-lpthread", "-o", HOST_BIN]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("host driver build failed:\n%s\n%s" % (r.stdout, r.stderr))
    return HOST_BIN


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
