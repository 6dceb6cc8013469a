import lzma, os, shutil, subprocess, sys, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_gpu_dropin_driver import _patch_input, GOLDEN, HOST
name, wts = "heat_transport", "N8_isotropic_L_v9_lambda1.wts"
raw = lzma.decompress(open(os.path.join(GOLDEN, wts + ".xz"), "rb").read())
for ic, order in [(3, 1), (6, 2)]:
    out = {}
    tmp = tempfile.mkdtemp()
    for g in (1, 2):
        d = os.path.join(tmp, "g%d" % g)
        for sub in ("input", "Data", "Weights", "Restart"):
            os.makedirs(os.path.join(d, sub), exist_ok=True)
        for fn in os.listdir(os.path.join(GOLDEN, "inputs")):
            if fn.startswith(name):
                shutil.copy(os.path.join(GOLDEN, "inputs", fn), os.path.join(d, "input", fn))
        open(os.path.join(d, "Weights", wts), "wb").write(raw)
        _patch_input(os.path.join(d, "input", name + ".test.in"), Init_field=ic, Space_order=order)
        r = subprocess.run([HOST, name + ".test.in", name + ".test.out"], cwd=d, capture_output=True, text=True, env=dict(os.environ, SBTE_GPUS=str(g)))
        print("rc", r.returncode, [l for l in r.stdout.split("\n") if "GPU" in l], r.stdout[-500:] if r.returncode else "", r.stderr[-300:])
        out[g] = np.loadtxt(os.path.join(d, "Data", "moments_%s.test.in" % name), comments="#")
    a, b = out[1], out[2]
    diff = np.abs(a - b)
    rows = np.where(diff.max(axis=1) > 0)[0]
    print("ic", ic, "order", order, "rows differing", rows.size, "of", a.shape[0], "max abs diff", diff.max())
    for i in rows[:12]:
        print("  row", i, "step", i // 250, "cell", i % 250, "\n   1gpu", a[i], "\n   2gpu", b[i])
