import os, sys, lzma
import numpy as np
sys.path.insert(0, os.getcwd())
import spectralbte_b200 as sb
from oracle import oracle as orc
raw = lzma.decompress(open("tests/golden/N8_isotropic_L_v5_lambda0.wts.xz", "rb").read())
W = np.frombuffer(raw, dtype=np.float64).copy()
o = orc.Oracle(8, 5.0, 0)
f = o.init_hom(2)
cells = np.stack([f * (1.0 + 0.01 * b) for b in range(40)])
want = np.stack([o.compute_q(W, cells[b], cells[b]) for b in range(40)])
print("scale |Q|max", np.abs(want).max(), "|Qmp|max", np.abs(o.compute_q_maxpreserve(W, f, f)).max())
def errs(Q):
    return [float(np.abs(Q[b] - want[b]).max() / np.abs(want).max()) for b in range(40)]
# A: fresh context, batch directly
c = sb.Collisions(8, 5.0); c.set_weights(W)
e = errs(c.ComputeQ(cells, k2=sb.K2_BATCH).reshape(40, -1)); print("A fresh batch     max err", max(e), "argmax", int(np.argmax(e)))
e = errs(c.ComputeQ(cells, k2=sb.K2_GENERIC).reshape(40, -1)); print("A fresh generic   max err", max(e))
# B: single first, then batch (capacity regrowth)
c2 = sb.Collisions(8, 5.0); c2.set_weights(W)
c2.ComputeQ_maxPreserve(f)
e = errs(c2.ComputeQ(cells, k2=sb.K2_BATCH).reshape(40, -1)); print("B regrow batch    max err", max(e), "argmax", int(np.argmax(e)), [round(x, 3) for x in e[:8]], [round(x,3) for x in e[30:]])
e = errs(c2.ComputeQ(cells, k2=sb.K2_BATCH).reshape(40, -1)); print("B second call     max err", max(e))
# C: N=32 stream with f == g through Qhat
o32 = orc.Oracle(32, 5.0, 0)
c3 = sb.Collisions(32, 5.0); c3.synthetic_weights(1)
f32 = o32.init_hom(0)
try:
    q = c3.Qhat(f32, None, k2=sb.K2_STREAM); print("C N=32 f==g stream ok", np.abs(q).max())
    qg = c3.Qhat(f32, None, k2=sb.K2_GENERIC); print("C rel diff vs generic", np.abs(q - qg).max() / np.abs(qg).max())
    Q = c3.ComputeQ(f32); print("C ComputeQ host ok", np.abs(Q).max())
except Exception as ex:
    print("C failed:", ex)
