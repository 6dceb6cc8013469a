import os, sys, time
sys.path.insert(0, os.getcwd())
import spectralbte_b200 as sb
for N, L_v, lam, inh in ((16, 9.0, 1.0, True), (24, 9.0, 1.0, True), (32, 5.0, 1.0, False)):
    c = sb.Collisions(N, L_v, inhomogeneous=inh)
    t0 = time.perf_counter(); c.generate_weights(lam); c.sync(); dt = time.perf_counter() - t0
    print("generate N=%d lambda=%g: %.2f s for %.3g integrals (%.1f Mintegrals/s)" % (N, lam, dt, N**6, N**6 / dt / 1e6), flush=True)
    c.close()
