"""CPU-side checks of the C host driver: it builds, parses the reference's input files and fails
loudly (exit 1, no CPU fallback) when no CUDA device is present."""
import os
import shutil
import subprocess

import pytest

from conftest import GOLDEN, ROOT


def test_host_driver_builds_and_refuses_without_gpu(tmp_path):
    import torch
    from spectralbte_b200 import build
    build.build()
    host = build.HOST_BIN
    assert os.path.exists(host)
    for d in ("input", "Data", "Weights"):
        os.makedirs(tmp_path / d, exist_ok=True)
    for fn in ("BKW8.test.in", "BKW8.test.out"):
        shutil.copy(os.path.join(GOLDEN, "inputs", fn), tmp_path / "input" / fn)
    r = subprocess.run([host, "BKW8.test.in", "BKW8.test.out"], cwd=tmp_path, capture_output=True, text=True, timeout=120)
    assert "Opening input file ./input/BKW8.test.in" in r.stdout and "done with input file" in r.stdout
    if not torch.cuda.is_available():
        assert r.returncode == 1
        assert "no CUDA device available" in r.stdout
    r2 = subprocess.run([host, "missing.in", "BKW8.test.out"], cwd=tmp_path, capture_output=True, text=True, timeout=120)
    assert r2.returncode == 1 and "Error - input file not found" in r2.stdout
