"""bench.py contract on CPU: the reference arm runs without a GPU and prints one JSON line with the keys the
driver reads (it times the reference's own ComputeQ on the host cores)."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT
from oracle import oracle as orc


@pytest.mark.skipif(not orc.have_ref(), reason="oracle/_ref not built")
def test_reference_arm_1d_prints_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "shock1p2",
                        "--steps", "1", "--warmup", "3"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "impl", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["vs_baseline"] is None and line["dtype"] == "f64"
    assert line["cpu_baseline"]["kind"] == "reference" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["value"] > 0


def test_bench_refuses_gpu_arm_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--no-cpu"], capture_output=True,
                       text=True, timeout=600, cwd=ROOT)
    assert r.returncode != 0      # no CPU fallback: the product arm fails loudly
