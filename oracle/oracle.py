"""oracle/oracle.py -- TEST INFRASTRUCTURE ONLY.

ctypes front-end to oracle/liboracle.so (the CPU restatement, oracle/oracle.c) and, when present,
oracle/_ref/libref.so (the reference's own C sources compiled unmodified, oracle/build_ref.sh).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module; the product package spectralbte_b200 never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


def _p(a):
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_dp)


def build(force=False):
    """Compile liboracle.so (always possible) and _ref (only where /root/reference exists)."""
    so = os.path.join(HERE, "liboracle.so")
    srcs = [os.path.join(HERE, f) for f in ("oracle.c", "oracle.h", "qag21.c", "qag21.h")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
    ref_root = os.environ.get("SBTE_REFERENCE_ROOT", "/root/reference")
    if os.path.isdir(os.path.join(ref_root, "src")):
        ref_so = os.path.join(HERE, "_ref", "libref.so")
        gpu_drv = os.path.join(HERE, "_ref", "boltz_gpu")
        have_lib = os.path.exists(os.path.join(HERE, "..", "spectralbte_b200", "libsbte_b200.so"))
        if force or not os.path.exists(ref_so) or (have_lib and not os.path.exists(gpu_drv)):
            subprocess.check_call([os.path.join(HERE, "build_ref.sh")], stdout=subprocess.DEVNULL)
    return so


_lib = None


def lib():
    global _lib
    if _lib is None:
        so = build()
        L = C.CDLL(so)
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [C.c_int, C.c_double, C.c_int]
        L.orc_destroy.argtypes = [C.c_void_p]
        L.orc_density.restype = C.c_double
        L.orc_temperature.restype = C.c_double
        L.orc_weight_one.restype = C.c_double
        _lib = L
    return _lib


class Oracle:
    """One velocity grid (N, L_v) + scratch; methods mirror oracle/oracle.h."""

    def __init__(self, N, L_v, grid_rule=0):
        self.L = lib()
        self.N = int(N)
        self.n3 = self.N ** 3
        self.L_v = float(L_v)
        self.h = C.c_void_p(self.L.orc_create(self.N, C.c_double(L_v), int(grid_rule)))
        self.v = np.zeros(self.N)
        self.eta = np.zeros(self.N)
        self.L.orc_get_grid(self.h, _p(self.v), _p(self.eta))

    def __del__(self):
        try:
            self.L.orc_destroy(self.h)
        except Exception:
            pass

    # -- transforms / convolution
    def fft3d(self, x, invert=False):
        x = np.ascontiguousarray(x, dtype=np.complex128).reshape(-1)
        out = np.empty_like(x)
        self.L.orc_fft3d(self.h, _p(x.view(np.float64)), _p(out.view(np.float64)), int(bool(invert)))
        return out

    def dft3(self, x, sign):
        x = np.array(x, dtype=np.complex128).reshape(-1)
        self.L.orc_dft3(self.h, _p(x.view(np.float64)), int(sign))
        return x

    def qhat(self, W, fhat, ghat):
        fhat = np.ascontiguousarray(fhat, dtype=np.complex128).reshape(-1)
        ghat = np.ascontiguousarray(ghat, dtype=np.complex128).reshape(-1)
        out = np.empty(self.n3, dtype=np.complex128)
        self.L.orc_qhat(self.h, _p(W), _p(fhat.view(np.float64)), _p(ghat.view(np.float64)),
                        _p(out.view(np.float64)))
        return out

    def compute_q(self, W, f, g, want_qhat=False):
        f = np.ascontiguousarray(f, dtype=np.float64).reshape(-1)
        g = np.ascontiguousarray(g, dtype=np.float64).reshape(-1)
        Q = np.empty(self.n3)
        qh = np.empty(self.n3, dtype=np.complex128) if want_qhat else None
        self.L.orc_compute_q(self.h, _p(W), _p(f), _p(g), _p(Q),
                             _p(qh.view(np.float64)) if want_qhat else None)
        return (Q, qh) if want_qhat else Q

    def compute_q_maxpreserve(self, W, f, g):
        f = np.ascontiguousarray(f, dtype=np.float64).reshape(-1)
        g = np.ascontiguousarray(g, dtype=np.float64).reshape(-1)
        Q = np.empty(self.n3)
        self.L.orc_compute_q_maxpreserve(self.h, _p(W), _p(f), _p(g), _p(Q))
        return Q

    def find_maxwellian(self, f):
        f = np.ascontiguousarray(f, dtype=np.float64).reshape(-1)
        M = np.empty(self.n3)
        ruT = np.empty(5)
        self.L.orc_find_maxwellian(self.h, _p(f), _p(M), _p(ruT))
        return M, ruT

    # -- moments / conservation
    def density(self, f):
        return float(self.L.orc_density(self.h, _p(np.ascontiguousarray(f).reshape(-1))))

    def bulk_velocity(self, f, rho):
        u = np.empty(3)
        self.L.orc_bulk_velocity(self.h, _p(np.ascontiguousarray(f).reshape(-1)), C.c_double(rho), _p(u))
        return u

    def temperature(self, f, u, rho):
        u = np.ascontiguousarray(u, dtype=np.float64)
        return float(self.L.orc_temperature(self.h, _p(np.ascontiguousarray(f).reshape(-1)), _p(u),
                                            C.c_double(rho)))

    def energy(self, f):
        e = np.empty(2)
        self.L.orc_energy(self.h, _p(np.ascontiguousarray(f).reshape(-1)), _p(e))
        return e

    def conserve(self, Q):
        Q = np.array(Q, dtype=np.float64).reshape(-1)
        self.L.orc_conserve(self.h, _p(Q))
        return Q

    def moment_functionals(self, Q):
        b = np.empty(5)
        self.L.orc_moment_functionals(self.h, _p(np.ascontiguousarray(Q, dtype=np.float64).reshape(-1)), _p(b))
        return b

    def conserve_lu(self):
        lu = np.empty(25)
        piv = np.zeros(5, dtype=np.int32)
        self.L.orc_conserve_lu(self.h, _p(lu), piv.ctypes.data_as(_ip))
        return lu.reshape(5, 5), piv

    # -- walls / transport (single rank, contiguous slabs)
    def diffuse_bc(self, fin, out, TW, bdry):
        self.L.orc_diffuse_bc(self.h, _p(fin), _p(out), C.c_double(TW), int(bdry))
        return out

    def upwind_one(self, nX, x, dx, dt, ic, f):
        fc = np.zeros_like(f)
        self.L.orc_upwind_one(self.h, int(nX), _p(x), _p(dx), C.c_double(dt), int(ic), _p(f), _p(fc))
        return fc

    def upwind_two(self, nX, x, dx, dt, ic, f):
        fc = np.zeros_like(f)
        self.L.orc_upwind_two(self.h, int(nX), _p(x), _p(dx), C.c_double(dt), int(ic), _p(f), _p(fc))
        return fc

    def advect_two(self, nX, x, dx, dt, ic, f):
        fc = np.zeros_like(f)
        ft = np.zeros_like(f)
        self.L.orc_advect_two(self.h, int(nX), _p(x), _p(dx), C.c_double(dt), int(ic), _p(f), _p(fc), _p(ft))
        return fc

    # -- time steps
    def step_0d(self, W, f, dt, Kn, order):
        self.L.orc_step_0d(self.h, _p(W), _p(f), C.c_double(dt), C.c_double(Kn), int(order))
        return f

    def step_1d(self, W, nX, x, dx, dt, Kn, order, ic, f, fc, f1, ft):
        self.L.orc_step_1d(self.h, _p(W), int(nX), _p(x), _p(dx), C.c_double(dt), C.c_double(Kn),
                           int(order), int(ic), _p(f), _p(fc), _p(f1), _p(ft))
        return f

    # -- set-up helpers
    def init_hom(self, flag):
        f = np.empty(self.n3)
        self.L.orc_init_hom(self.h, int(flag), _p(f))
        return f

    def init_inhom(self, flag, nX, order):
        f = np.zeros((nX + 2 * order, self.n3))
        self.L.orc_init_inhom(self.h, int(flag), int(nX), int(order), _p(f))
        return f

    def row_0d(self, f):
        r = np.empty(5 + self.N)
        self.L.orc_row_0d(self.h, _p(np.ascontiguousarray(f).reshape(-1)), _p(r))
        return r

    def row_1d(self, f):
        r = np.empty(4)
        self.L.orc_row_1d(self.h, _p(np.ascontiguousarray(f).reshape(-1)), _p(r))
        return r

    def weights_iso(self, lam):
        W = np.empty(self.n3 * self.n3)
        self.L.orc_weights_iso(self.h, C.c_double(lam), _p(W))
        return W

    def weight_one(self, lam, zeta, xi):
        return float(self.L.orc_weight_one(self.h, C.c_double(lam), int(zeta), int(xi)))


def make_mesh(zone_n, zone_len, order):
    zone_n = np.asarray(zone_n, dtype=np.int32)
    zone_len = np.asarray(zone_len, dtype=np.float64)
    nX = int(zone_n.sum())
    x = np.zeros(nX + 2 * order)
    dx = np.zeros(nX + 2 * order)
    lib().orc_make_mesh(len(zone_n), zone_n.ctypes.data_as(_ip), _p(zone_len), int(order), _p(x), _p(dx))
    return nX, x, dx


def synthetic_weights(N, seed=20261017):
    """Deterministic dense stand-in weights (splitmix64 -> uniform[-0.5,0.5)), SURVEY.md 8(d)."""
    n = N ** 6
    z = (np.arange(1, n + 1, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15) + np.uint64(seed))
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    z = z ^ (z >> np.uint64(31))
    return (z >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0) - 0.5


# ---------------------------------------------------------------------------------------------
# The reference's own code (oracle/_ref/libref.so), file-static state and all.
# ---------------------------------------------------------------------------------------------
class _Species(C.Structure):
    # layout of `species`, /root/reference/src/species.h:11-26
    _fields_ = [("id", C.c_size_t), ("num_levels", C.c_size_t), ("lev_id", C.c_void_p),
                ("Rgas", C.c_double), ("mass", C.c_double), ("mm", C.c_double), ("d_ref", C.c_double),
                ("T_ref", C.c_double), ("mu_ref", C.c_double), ("omega", C.c_double), ("E0", C.c_double),
                ("Ei", C.c_void_p), ("gi", C.c_void_p), ("name", C.c_char * 80)]


def have_ref():
    return os.path.exists(os.path.join(HERE, "_ref", "libref.so"))


class Reference:
    """Drives the reference's functions in-process. One instance per process (file-static state)."""

    def __init__(self, N, L_v, grid_rule=0):
        self.R = C.CDLL(os.path.join(HERE, "_ref", "libref.so"))
        self.N, self.n3, self.L_v = int(N), int(N) ** 3, float(L_v)
        o = Oracle(N, L_v, grid_rule)  # grids only (values identical by construction; checked in tests)
        self.v = o.v.copy()
        self.eta = o.eta.copy()
        self.mix = (_Species * 1)()
        self.mix[0].id = 0
        self.mix[0].mass = 1.0
        self.mix[0].d_ref = 2.0
        self.mix[0].Rgas = self.mix[0].mm = self.mix[0].T_ref = self.mix[0].mu_ref = 1.0
        self.mix[0].name = b"default"
        R = self.R
        R.initialize_coll(self.N, C.c_double(L_v), _p(self.v), _p(self.eta))
        R.initialize_moments(self.N, _p(self.v), self.mix)
        R.initialize_conservation(self.N, C.c_double(self.v[1] - self.v[0]), _p(self.v), self.mix, 1)
        R.getDensity.restype = C.c_double
        R.getTemperature.restype = C.c_double
        self._rows = None
        self._keep = []

    def rows(self, W):
        """Row-pointer view (double **) of a contiguous N^3 x N^3 matrix, as weights.c allocates."""
        Wm = W.reshape(self.n3, self.n3)
        arr = (_dp * self.n3)()
        base = Wm.ctypes.data
        for i in range(self.n3):
            arr[i] = C.cast(base + i * self.n3 * 8, _dp)
        self._keep.append(W)
        return arr

    def fft3d(self, x, invert=False):
        x = np.ascontiguousarray(x, dtype=np.complex128).reshape(-1)
        out = np.empty_like(x)
        self.R.fft3D(_p(x.view(np.float64)), _p(out.view(np.float64)), int(bool(invert)))
        return out

    def compute_q(self, rows, f, g):
        Q = np.empty(self.n3)
        self.R.ComputeQ(_p(np.ascontiguousarray(f).reshape(-1)), _p(np.ascontiguousarray(g).reshape(-1)), _p(Q), rows)
        return Q

    def compute_q_maxpreserve(self, rows, f, g):
        Q = np.empty(self.n3)
        self.R.ComputeQ_maxPreserve(_p(np.ascontiguousarray(f).reshape(-1)),
                                    _p(np.ascontiguousarray(g).reshape(-1)), _p(Q), rows)
        return Q

    def conserve(self, Q):
        Q = np.array(Q, dtype=np.float64).reshape(-1)
        qp = (_dp * 1)(_p(Q))
        self.R.conserveAllMoments(qp)
        return Q

    def moments(self, f):
        f = np.ascontiguousarray(f).reshape(-1)
        rho = float(self.R.getDensity(_p(f), 0))
        u = np.empty(3)
        self.R.getBulkVelocity(_p(f), _p(u), C.c_double(rho), 0)
        T = float(self.R.getTemperature(_p(f), _p(u), C.c_double(rho), 0))
        e = np.empty(2)
        self.R.getEnergy(_p(f), _p(e))
        return rho, u, T, e

    def init_transport(self, nX, x, dx, ic, dt):
        self._tx = (x.copy(), dx.copy())
        self.R.initialize_transport(self.N, int(nX), C.c_double(self.L_v), _p(self._tx[0]), _p(self._tx[1]),
                                    _p(self.v), int(ic), C.c_double(dt), C.c_double(1.0), self.mix)

    @staticmethod
    def _cells(slab):
        n = slab.shape[0]
        arr = (_dp * n)()
        for i in range(n):
            arr[i] = C.cast(slab.ctypes.data + i * slab.shape[1] * 8, _dp)
        return arr

    def advect_one(self, f):
        fc = np.zeros_like(f)
        self.R.advectOne(self._cells(f), self._cells(fc), 0)
        return fc

    def advect_two(self, f):
        fc = np.zeros_like(f)
        self.R.advectTwo(self._cells(f), self._cells(fc), 0)
        return fc

    def diffuse_bc(self, fin, out, TW, bdry):
        self.R.setDiffuseReflectionBC(_p(fin), _p(out), C.c_double(TW), int(bdry), 0)
        return out
