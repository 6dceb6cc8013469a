"""Host-side mirror of the reference's collision interface on top of the C ABI.

Names follow the reference (/root/reference/src/collisions.h, conserve.h, transportroutines.h,
exec/boltz.c): ComputeQ, ComputeQ_maxPreserve, conserveAllMoments, advectOne/advectTwo live on
`Collisions` (one velocity grid + one weight tensor on one GPU) and `Slab` (the 1D cell block of
one rank).  Arrays passed in/out are numpy (host); `DeviceArray` keeps data resident in HBM.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import check

K2_AUTO, K2_GENERIC, K2_STREAM, K2_BATCH, K2_STREAM_DEEP = 0, 1, 2, 3, 4
_dp = C.POINTER(C.c_double)


def _p(a):
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_dp)


def velocity_grids(N, L_v, inhomogeneous=False):
    """v and eta grids: src/initializer.c:66-82 (0D) and :256-266 (1D)."""
    dv = 2 * L_v / (N - 1)
    v = np.array([-L_v + i * dv for i in range(N)], dtype=np.float64)
    if not inhomogeneous:
        deta = (2 * np.pi / N) / dv
        L_eta = 0.5 * N * deta if N % 2 == 0 else 0.5 * (N - 1) * deta
    else:
        L_eta = 0.5 * (N - 1) * np.pi / L_v
        deta = np.pi * (N - 1) / (N * L_v)
    eta = np.array([-L_eta + i * deta for i in range(N)], dtype=np.float64)
    return v, eta


def weights_filename(N, L_v, lam):
    """src/weights.c:68 ("old style" name for the default species)."""
    return "Weights/N%d_isotropic_L_v%g_lambda%g.wts" % (N, L_v, lam)


class DeviceArray:
    """n doubles in HBM, owned by this object."""

    def __init__(self, ctx, n):
        self.ctx, self.n = ctx, int(n)
        p = C.c_void_p()
        check(ctx.L.sbte_dev_alloc(C.byref(p), self.n * 8))
        self.ptr = p.value

    def put(self, a):
        a = np.ascontiguousarray(a, dtype=np.float64).reshape(-1)
        assert a.size == self.n
        check(self.ctx.L.sbte_h2d(self.ctx.h, self.ptr, a.ctypes.data, a.nbytes))
        return self

    def get(self):
        out = np.empty(self.n)
        check(self.ctx.L.sbte_d2h(self.ctx.h, out.ctypes.data, self.ptr, out.nbytes))
        return out

    def free(self):
        if self.ptr:
            self.ctx.L.sbte_dev_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Collisions:
    """initialize_coll + initialize_conservation for one grid on one GPU (src/collisions.c:33-74)."""

    def __init__(self, N, L_v, inhomogeneous=False, device=0, v=None, eta=None):
        self.L = _lib.load()
        self.N, self.n3, self.L_v = int(N), int(N) ** 3, float(L_v)
        if v is None:
            v, eta = velocity_grids(N, L_v, inhomogeneous)
        self.v, self.eta = np.ascontiguousarray(v), np.ascontiguousarray(eta)
        h = C.c_void_p()
        check(self.L.sbte_create(C.byref(h), self.N, self.L_v, _p(self.v), _p(self.eta), int(device)))
        self.h = h
        self._keep = None

    def close(self):
        if self.h:
            self.L.sbte_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- weights (src/weights.c:33-125)
    def set_weights(self, W):
        W = np.ascontiguousarray(W, dtype=np.float64).reshape(-1)
        assert W.size == self.n3 * self.n3
        check(self.L.sbte_weights_upload(self.h, _p(W)))

    def set_weights_rows(self, W):
        """Upload through the reference's N^3-row-pointer layout (src/weights.c:61-63)."""
        Wm = np.ascontiguousarray(W, dtype=np.float64).reshape(self.n3, self.n3)
        rows = (_dp * self.n3)()
        for i in range(self.n3):
            rows[i] = C.cast(Wm.ctypes.data + i * self.n3 * 8, _dp)
        self._keep = (Wm, rows)
        check(self.L.sbte_weights_upload_rows(self.h, rows))

    def load_weights(self, path):
        check(self.L.sbte_weights_load_file(self.h, path.encode()))

    def synthetic_weights(self, seed=20261017):
        check(self.L.sbte_weights_fill_synthetic(self.h, int(seed)))

    def generate_weights(self, lam):
        """generate_conv_weights_iso on the device (src/weights.c:265-281)."""
        check(self.L.sbte_weights_generate_iso(self.h, float(lam)))

    def save_weights(self, path):
        check(self.L.sbte_weights_save_file(self.h, path.encode()))

    def weights_to_host(self):
        out = np.empty(self.n3 * self.n3)
        check(self.L.sbte_d2h(self.h, out.ctypes.data, self.L.sbte_weights_device(self.h), out.nbytes))
        return out

    # -- plumbing
    def sync(self):
        check(self.L.sbte_sync(self.h))

    @property
    def stream(self):
        return self.L.sbte_stream(self.h)

    @property
    def launches(self):
        return int(self.L.sbte_launch_count(self.h))

    def array(self, n):
        return DeviceArray(self, n)

    def set_symmetrize(self, enable=True):
        """Toggle the symmetrised weight stream used when f == g (default on)."""
        check(self.L.sbte_set_symmetrize(self.h, int(bool(enable))))

    def set_xy_pairing(self, enable=True):
        """Toggle the transposed pairing of the 0D stream (half of the zeta columns; needs an x<->y invariant tensor)."""
        check(self.L.sbte_set_xy_pairing(self.h, int(bool(enable))))

    def xy_pairing_state(self):
        """(state, deviation): -1 not examined, 0 tensor not invariant, 1 in use; relative deviation found by the check."""
        st, dev = C.c_int(), C.c_double()
        check(self.L.sbte_xy_pairing_state(self.h, C.byref(st), C.byref(dev)))
        return st.value, dev.value

    def k2_profile(self, enable=True):
        check(self.L.sbte_k2_profile(self.h, int(bool(enable))))

    def k2_profile_read(self):
        """(summed device ms, launches) of the convolution kernel since the last read."""
        ms, n = C.c_double(), C.c_int()
        check(self.L.sbte_k2_profile_read(self.h, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    # -- reference-named operations on host arrays
    def fft3D(self, x, invert=False):
        x = np.ascontiguousarray(x, dtype=np.complex128).reshape(-1)
        b = x.size // self.n3
        din, dout = self.array(2 * x.size), self.array(2 * x.size)
        din.put(x.view(np.float64))
        check(self.L.sbte_fft3d(self.h, din.ptr, dout.ptr, int(bool(invert)), b))
        return dout.get().view(np.complex128)

    def Qhat(self, f, g=None, k2=K2_AUTO):
        f = np.ascontiguousarray(f, dtype=np.float64).reshape(-1)
        b = f.size // self.n3
        df = self.array(f.size).put(f)
        dg = df if g is None else self.array(f.size).put(g)
        dq = self.array(2 * f.size)
        check(self.L.sbte_qhat(self.h, df.ptr, dg.ptr, dq.ptr, b, int(k2)))
        return dq.get().view(np.complex128)

    def ComputeQ(self, f, g=None, k2=K2_AUTO):
        f = np.ascontiguousarray(f, dtype=np.float64).reshape(-1)
        b = f.size // self.n3
        if b == 1:
            Q = np.empty(self.n3)
            gg = f if g is None else np.ascontiguousarray(g, dtype=np.float64).reshape(-1)
            check(self.L.sbte_compute_q_host(self.h, _p(f), _p(gg), _p(Q), int(k2)))
            return Q
        df = self.array(f.size).put(f)
        dg = df if g is None else self.array(f.size).put(g)
        dQ = self.array(f.size)
        check(self.L.sbte_compute_q(self.h, df.ptr, dg.ptr, dQ.ptr, b, int(k2)))
        return dQ.get()

    def ComputeQ_maxPreserve(self, f, g=None, k2=K2_AUTO):
        f = np.ascontiguousarray(f, dtype=np.float64).reshape(-1)
        gg = f if g is None else np.ascontiguousarray(g, dtype=np.float64).reshape(-1)
        Q = np.empty(self.n3)
        check(self.L.sbte_compute_q_maxpreserve_host(self.h, _p(f), _p(gg), _p(Q), int(k2)))
        return Q

    def conserveAllMoments(self, Q):
        Q = np.ascontiguousarray(Q, dtype=np.float64).reshape(-1)
        b = Q.size // self.n3
        d = self.array(Q.size).put(Q)
        check(self.L.sbte_conserve(self.h, d.ptr, b))
        return d.get()

    def moment_functionals(self, Q):
        Q = np.ascontiguousarray(Q, dtype=np.float64).reshape(-1)
        b = Q.size // self.n3
        d, o = self.array(Q.size).put(Q), self.array(5 * b)
        check(self.L.sbte_moment_functionals(self.h, d.ptr, o.ptr, b))
        return o.get().reshape(b, 5)

    def moments(self, f):
        """rho, u_x, u_y, u_z, T, E_pos, E_neg, p per cell (src/momentRoutines.c:58-183)."""
        f = np.ascontiguousarray(f, dtype=np.float64).reshape(-1)
        b = f.size // self.n3
        d, o = self.array(f.size).put(f), self.array(8 * b)
        check(self.L.sbte_moments(self.h, d.ptr, o.ptr, b))
        return o.get().reshape(b, 8)

    def step_0d(self, dev_f, dt, Kn, order, k2=K2_AUTO):
        """exec/boltz.c:189-241 on a DeviceArray (f never leaves HBM)."""
        check(self.L.sbte_step_0d(self.h, dev_f.ptr, float(dt), float(Kn), int(order), int(k2)))

    def row_0d(self, dev_f):
        """Output row of src/output.c:180-209 without the time column."""
        m = self.array(8)
        check(self.L.sbte_moments(self.h, dev_f.ptr, m.ptr, 1))
        mom = m.get()
        f = dev_f.get()
        N = self.N
        sl = [f[N // 2 + N * (N // 2 + N * l)] for l in range(N)]
        return np.concatenate([[mom[0], mom[1], mom[4], mom[7], mom[6] / mom[5]], sl])


class Slab:
    """The cells of one rank of a 1D-3V run, resident on the device (exec/boltz.c:264-353)."""

    def __init__(self, coll, cells_local, order, x, dx, init_field, dt, rank=0, nranks=1):
        self.coll, self.L = coll, coll.L
        self.nX, self.order = int(cells_local), int(order)
        self.ncell = self.nX + 2 * self.order
        self.rank, self.nranks = int(rank), int(nranks)
        self.cells_local, self.init_field = self.nX, int(init_field)
        x = np.ascontiguousarray(x, dtype=np.float64)
        dx = np.ascontiguousarray(dx, dtype=np.float64)
        assert x.size == self.ncell and dx.size == self.ncell
        h = C.c_void_p()
        check(self.L.sbte_slab_create(coll.h, C.byref(h), self.nX, self.order, _p(x), _p(dx), int(init_field),
                                      float(dt), self.rank, self.nranks))
        self.h = h

    def close(self):
        if self.h:
            self.L.sbte_slab_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def upload(self, f):
        f = np.ascontiguousarray(f, dtype=np.float64).reshape(-1)
        assert f.size == self.ncell * self.coll.n3
        check(self.L.sbte_slab_upload(self.h, _p(f)))

    def download(self):
        out = np.empty((self.ncell, self.coll.n3))
        check(self.L.sbte_slab_download(self.h, _p(out)))
        return out

    def download_fconv(self):
        out = np.empty((self.ncell, self.coll.n3))
        check(self.L.sbte_d2h(self.coll.h, out.ctypes.data, self.L.sbte_slab_fconv(self.h), out.nbytes))
        return out

    def advect(self, which=0):
        """advectOne / advectTwo (src/transportroutines.c:473-492), single rank."""
        check(self.L.sbte_slab_advect(self.h, int(which)))

    def upwind_stage(self, which, stage):
        check(self.L.sbte_slab_upwind_stage(self.h, int(which), int(stage)))

    def advect_finish(self, which):
        check(self.L.sbte_slab_advect_finish(self.h, int(which)))

    def halo_regions(self, which, stage, side):
        s, r, n = C.c_void_p(), C.c_void_p(), C.c_size_t()
        check(self.L.sbte_slab_halo_regions(self.h, int(which), int(stage), int(side), C.byref(s), C.byref(r), C.byref(n)))
        return s.value, r.value, n.value

    def ipc_export(self):
        """256 bytes of CUDA IPC handles (f, f_conv, f_tmp, flags) for the neighbouring ranks."""
        buf = C.create_string_buffer(256)
        check(self.L.sbte_slab_ipc_export(self.h, buf))
        return buf.raw

    def ipc_import(self, side, handles, neighbour_cells):
        check(self.L.sbte_slab_ipc_import(self.h, int(side), handles, int(neighbour_cells)))

    def peer_attach(self, side, other):
        """Same-process neighbour (another context / GPU): no IPC mapping needed."""
        check(self.L.sbte_slab_peer_attach(self.h, int(side), other.h))

    def halo_state(self):
        """(ready, done, epoch, error) words of the peer-memory halo; error != 0: a wait for a neighbour timed out."""
        st = (C.c_int * 4)()
        check(self.L.sbte_slab_halo_state(self.h, st))
        return tuple(st)

    def set_halo_timeout(self, seconds):
        """Bound of the device-side waits for a neighbouring rank (<= 0: wait for ever)."""
        check(self.L.sbte_slab_set_halo_timeout(self.h, float(seconds)))

    def set_peer_halo(self, enable=True):
        check(self.L.sbte_slab_set_peer_halo(self.h, int(bool(enable))))

    def peer_detach(self):
        check(self.L.sbte_slab_peer_detach(self.h))

    def collide(self, Kn, k2=K2_AUTO):
        check(self.L.sbte_slab_collide(self.h, float(Kn), int(k2)))

    def step(self, Kn, k2=K2_AUTO):
        check(self.L.sbte_slab_step(self.h, float(Kn), int(k2)))

    def moments(self):
        out = np.empty((self.nX, 8))
        check(self.L.sbte_slab_moments(self.h, _p(out)))
        return out
