"""Run the reference executable and boltz_b200 on a patched heat_transport input and print the first mismatching rows."""
import lzma, os, shutil, subprocess, sys, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_gpu_dropin_driver import _patch_input, GOLDEN, REF_EXE, HOST
name, wts = "heat_transport", "N8_isotropic_L_v9_lambda1.wts"
raw = lzma.decompress(open(os.path.join(GOLDEN, wts + ".xz"), "rb").read())
for ic, order in [(2, 1), (6, 1), (6, 2)]:
    out = {}
    tmp = tempfile.mkdtemp()
    for tag, exe in (("ref", REF_EXE), ("gpu", HOST)):
        d = os.path.join(tmp, tag)
        for sub in ("input", "Data", "Weights", "Restart"):
            os.makedirs(os.path.join(d, sub), exist_ok=True)
        for fn in os.listdir(os.path.join(GOLDEN, "inputs")):
            if fn.startswith(name):
                shutil.copy(os.path.join(GOLDEN, "inputs", fn), os.path.join(d, "input", fn))
        open(os.path.join(d, "Weights", wts), "wb").write(raw)
        _patch_input(os.path.join(d, "input", name + ".test.in"), Init_field=ic, Space_order=order)
        r = subprocess.run([exe, name + ".test.in", name + ".test.out"], cwd=d, capture_output=True, text=True)
        out[tag] = np.loadtxt(os.path.join(d, "Data", "moments_%s.test.in" % name), comments="#")
    g, w = out["gpu"], out["ref"]
    tol = np.maximum(1e-14, 1e-6 * np.abs(w))
    viol = np.abs(g - w) > tol
    print("violations per column", viol.sum(axis=0))
    bad = np.where(viol.any(axis=1))[0]
    print("ic", ic, "order", order, "bad rows", bad.size, "first", bad[:10])
    for i in bad[:6]:
        print("  row", i, "\n   gpu", g[i], "\n   ref", w[i])
