import os, sys, lzma, time
import numpy as np
sys.path.insert(0, os.getcwd())
import spectralbte_b200 as sb
from oracle import oracle as orc
W = np.frombuffer(lzma.decompress(open("tests/golden/N8_isotropic_L_v9_lambda1.wts.xz", "rb").read()), dtype=np.float64).copy()
order, ic = 1, 3
N, nX, dt, Kn = 8, 16, 2e-3, 1.52
o = orc.Oracle(N, 9.0, 1)
_, x, dx = orc.make_mesh([nX], [0.8], order)
f0 = o.init_inhom(ic, nX, order)
ctxs = [sb.Collisions(N, 9.0, inhomogeneous=True) for _ in range(2)]
for c in ctxs: c.set_weights(W)
h = nX // 2
parts = []
for r in range(2):
    lo = r * h
    p = sb.Slab(ctxs[r], h, order, x[lo:lo + h + 2 * order].copy(), dx[lo:lo + h + 2 * order].copy(), ic, dt, rank=r, nranks=2)
    p.upload(f0[lo:lo + h + 2 * order].copy())
    parts.append(p)
parts[0].peer_attach(1, parts[1]); parts[1].peer_attach(0, parts[0])
for p in parts: p.set_peer_halo(True)
for c in ctxs: c.sync()
print("flags before", parts[0].halo_state(), parts[1].halo_state(), flush=True)
if os.environ.get("ONLY_B"):
    parts[1].upwind_stage(0, 0)
    time.sleep(1.0)
    print("flags after B enqueue (A not started)", parts[0].halo_state(), parts[1].halo_state(), flush=True)
    parts[0].upwind_stage(0, 0)
    ctxs[0].sync(); ctxs[1].sync()
    print("flags after both", parts[0].halo_state(), parts[1].halo_state(), flush=True)
def stage(msg, fn):
    t0 = time.time()
    try:
        fn()
        print(msg, "ok %.3fs" % (time.time() - t0), flush=True)
    except Exception as e:
        print(msg, "FAILED %.3fs" % (time.time() - t0), e, flush=True); sys.exit(1)
for it in range(3):
    stage("it%d enqueue upwind A" % it, lambda: parts[0].upwind_stage(0, 0))
    stage("it%d enqueue upwind B" % it, lambda: parts[1].upwind_stage(0, 0))
    stage("it%d sync A" % it, lambda: ctxs[0].sync())
    stage("it%d sync B" % it, lambda: ctxs[1].sync())
    stage("it%d collide A" % it, lambda: parts[0].collide(Kn))
    stage("it%d collide B" % it, lambda: parts[1].collide(Kn))
    stage("it%d sync A" % it, lambda: ctxs[0].sync())
    stage("it%d sync B" % it, lambda: ctxs[1].sync())
print("done")
