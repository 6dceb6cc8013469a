#!/usr/bin/env python
"""tests/golden/make_golden.py -- regenerates the committed fixtures in tests/golden/.

Run in the build container (needs /root/reference and oracle/_ref/libref.so from
oracle/build_ref.sh); the GPU box only ever reads the committed outputs.

1. Copies the reference's own golden vectors (tests/BKW8/target, tests/heat_transport/target):
   the two moments_* text files verbatim, the two 2 MiB .wts files xz-compressed.
2. Calls the reference's own functions (compiled unmodified into libref.so) on seeded inputs and
   stores inputs' seeds + outputs: fft3D, ComputeQ, ComputeQ_maxPreserve, conserveAllMoments,
   moments, setDiffuseReflectionBC, advectOne, advectTwo.
"""
import lzma
import os
import shutil
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import oracle as orc  # noqa: E402

REF = os.environ.get("SBTE_REFERENCE_ROOT", "/root/reference")


def copy_reference_goldens():
    pairs = [("tests/BKW8/target/moments_BKW8.test.in", "moments_BKW8.test.in"),
             ("tests/heat_transport/target/moments_heat_transport.test.in", "moments_heat_transport.test.in")]
    for src, dst in pairs:
        shutil.copyfile(os.path.join(REF, src), os.path.join(HERE, dst))
    os.makedirs(os.path.join(HERE, "inputs"), exist_ok=True)
    for src in ("tests/BKW8/input/BKW8.test.in", "tests/BKW8/input/BKW8.test.out",
                "tests/heat_transport/input/heat_transport.test.in", "tests/heat_transport/input/heat_transport.test.out",
                "tests/heat_transport/input/heat_transport.test.mesh"):
        shutil.copyfile(os.path.join(REF, src), os.path.join(HERE, "inputs", os.path.basename(src)))
    for src in ("tests/BKW8/target/N8_isotropic_L_v5_lambda0.wts",
                "tests/heat_transport/target/N8_isotropic_L_v9_lambda1.wts"):
        raw = open(os.path.join(REF, src), "rb").read()
        with open(os.path.join(HERE, os.path.basename(src) + ".xz"), "wb") as fh:
            fh.write(lzma.compress(raw, preset=9))


def seeded_f(o, seed, noise=0.05):
    """Positive, Maxwellian-like, slightly perturbed distribution on o's grid."""
    rng = np.random.default_rng(seed)
    v = o.v
    vx, vy, vz = np.meshgrid(v, v, v, indexing="ij")
    f = np.exp(-((vx - 0.3) ** 2 + (vy + 0.2) ** 2 + vz ** 2) / 1.7) / 7.0
    f *= 1.0 + noise * rng.standard_normal(f.shape)
    return np.ascontiguousarray(f.reshape(-1))


def hot_path_vectors(tag, N, L_v, grid_rule, W, out):
    o = orc.Oracle(N, L_v, grid_rule)
    R = orc.Reference(N, L_v, grid_rule)
    rows = R.rows(W)
    f = seeded_f(o, 11)
    g = seeded_f(o, 12)
    rng = np.random.default_rng(5)
    z = rng.standard_normal(o.n3) + 1j * rng.standard_normal(o.n3)
    out[f"{tag}_fft_fwd"] = R.fft3d(z, False)
    out[f"{tag}_fft_inv"] = R.fft3d(z, True)
    Qff = R.compute_q(rows, f, f)
    out[f"{tag}_Q_ff"] = Qff
    out[f"{tag}_Q_fg"] = R.compute_q(rows, f, g)
    Qmp = R.compute_q_maxpreserve(rows, f, f)
    out[f"{tag}_Qmp_ff"] = Qmp
    out[f"{tag}_Qmp_fg"] = R.compute_q_maxpreserve(rows, f, g)
    out[f"{tag}_cons_Q_ff"] = R.conserve(Qff)
    out[f"{tag}_cons_Qmp_ff"] = R.conserve(Qmp)
    rho, u, T, e = R.moments(f)
    out[f"{tag}_moments_f"] = np.array([rho, u[0], u[1], u[2], T, e[0], e[1]])


def transport_vectors(tag, N, L_v, nX, ic, dt, out):
    o = orc.Oracle(N, L_v, 1)
    R = orc.Reference(N, L_v, 1)
    rng = np.random.default_rng(77)
    for order in (1, 2):
        _, x, dx = orc.make_mesh([nX // 2, nX - nX // 2], [0.4, 0.6], order)  # two zones: non-uniform dx
        R.init_transport(nX, x, dx, ic, dt)
        f = o.init_inhom(ic, nX, order)
        f[order:nX + order] *= 1.0 + 0.2 * rng.standard_normal((nX, o.n3))
        fin = f.copy()
        fc = R.advect_one(f) if order == 1 else R.advect_two(f)
        out[f"{tag}_o{order}_in"] = fin[order:nX + order]
        out[f"{tag}_o{order}_out"] = fc[order:nX + order]
    fin = seeded_f(o, 3)
    for bdry, TW in ((0, 1.0), (1, 2.0)):
        res = R.diffuse_bc(fin, np.zeros(o.n3), TW, bdry)
        out[f"{tag}_bc{bdry}"] = res


def main():
    copy_reference_goldens()
    out = {}
    W0 = np.frombuffer(lzma.decompress(open(os.path.join(HERE, "N8_isotropic_L_v5_lambda0.wts.xz"), "rb").read()))
    W1 = np.frombuffer(lzma.decompress(open(os.path.join(HERE, "N8_isotropic_L_v9_lambda1.wts.xz"), "rb").read()))
    hot_path_vectors("n8_l0", 8, 5.0, 0, W0.copy(), out)
    hot_path_vectors("n8_l1", 8, 9.0, 1, W1.copy(), out)
    # non-power-of-two and larger N with deterministic synthetic weights (values irrelevant to the
    # algebra: the operator is linear in W)
    hot_path_vectors("n12_syn", 12, 6.0, 0, orc.synthetic_weights(12), out)
    hot_path_vectors("n16_syn", 16, 5.0, 0, orc.synthetic_weights(16), out)
    transport_vectors("tr_ic3", 8, 9.0, 12, 3, 1e-3, out)   # diffuse walls
    transport_vectors("tr_ic6", 8, 9.0, 12, 6, 1e-3, out)   # periodic (order 1) / no-flux (order 2)
    transport_vectors("tr_ic0", 6, 7.0, 10, 0, 2e-3, out)   # copy / no-flux, N=6
    transport_vectors("tr_ic5", 8, 9.0, 12, 5, 1e-3, out)   # Poiseuille: diffuse walls + v_y forcing (order 2)
    transport_vectors("tr_ic1", 8, 9.0, 12, 1, 1e-3, out)   # sudden heating: left wall at 2 T, right copy
    np.savez_compressed(os.path.join(HERE, "ref_vectors.npz"), **out)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()
