"""Initial data and meshes of the reference driver, restated for the host side
(/root/reference/src/initializer.c:91-198 0D, :297-421 1D; src/mesh_setup.c:65-146,165-175)."""
import numpy as np


def init_hom(v, L_v, init_field):
    """0D initial distribution on the N^3 grid, flat index k + N*(j + N*i)."""
    vx, vy, vz = np.meshgrid(v, v, v, indexing="ij")
    r2 = vx * vx + vy * vy + vz * vz
    if init_field == 0:      # shifted isotropic
        sigma, S = 0.3 * L_v, 10.0
        f = np.exp(-1 * S * (np.sqrt(r2) - sigma) * (np.sqrt(r2) - sigma) / (sigma * sigma)) / (S * S)
    elif init_field == 2:    # BKW
        K, T = 1 - np.exp(-5.5 / 6.0), 1.0
        f = (np.exp(-r2 / (2 * K * T * T))) / (2.0 * (2 * np.pi * K * T * T) ** 1.5) * \
            ((5 * K - 3) / K + (1 - K) * r2 / (K * K * T * T))
    elif init_field == 4:    # Maxwellian rho = 1, T = 1
        f = (0.5 / np.pi) ** 1.5 * np.exp(-0.5 * r2)
    elif init_field == 5:    # perturbed Maxwellian
        f = (1 + 0.1 * np.sin(r2)) * np.exp(-r2) / (np.pi * np.sqrt(np.pi))
    else:
        raise ValueError("Init_field %d not implemented for the 0D case" % init_field)
    return np.ascontiguousarray(f.reshape(-1))


def init_inhom(v, init_field, nX_global, order, lo, hi):
    """Cells lo..hi-1 (global, 0-based) of the 1D initial data as a slab with 2*order ghost cells.
    The left/right split of the shock cases is at slab index l < nX/2 with l counted from `order`
    (src/initializer.c:362,406)."""
    vx, vy, vz = np.meshgrid(v, v, v, indexing="ij")

    def maxw(rho, ux, T):
        return (rho * np.exp(-((vx - ux) ** 2 + vy * vy + vz * vz) / T) / ((T * np.pi) * np.sqrt(T * np.pi))).reshape(-1)

    n3 = v.size ** 3
    slab = np.zeros((hi - lo + 2 * order, n3))
    if init_field == 3:
        cell = {None: maxw(1.0, 0.0, 1.5)}
        pick = lambda l: cell[None]  # noqa: E731
    elif init_field == 1:    # sudden heating: uniform gas at T = 1 (the k_B T = m <v^2>/3 convention of
        r2 = (vx * vx + vy * vy + vz * vz).reshape(-1)   # src/initializer.c:381), left wall at 2 T
        cell = {None: (0.5 / np.pi) ** 1.5 * np.exp(-0.5 * r2)}
        pick = lambda l: cell[None]  # noqa: E731
    elif init_field == 5:    # Poiseuille: gas at rest, rho 1, T 1, between diffuse walls (src/initializer.c:336-340,398-403)
        cell = {None: maxw(1.0, 0.0, 1.0)}
        pick = lambda l: cell[None]  # noqa: E731
    elif init_field == 2:    # uniform shifted Maxwellian (rho 1, u_x -1, T 1)
        cell = {None: maxw(1.0, -1.0, 1.0)}
        pick = lambda l: cell[None]  # noqa: E731
    elif init_field == 6:
        left, right = maxw(1.0, 1.2972, 1.0), maxw(1.297, 1.0, 1.195)
        pick = lambda l: left if l < nX_global // 2 else right  # noqa: E731
    elif init_field == 0:
        Ma = 1.0
        rho_l = 4.0 * Ma * Ma / (Ma * Ma + 3.0)
        T_l = (5.0 * Ma * Ma - 1.0) * (Ma * Ma + 3.0) / (16.0 * Ma * Ma)
        left, right = maxw(rho_l, 0.0, T_l), maxw(1.0, 0.0, 1.0)
        pick = lambda l: left if l < nX_global // 2 else right  # noqa: E731
    else:
        raise ValueError("Init_field %d not implemented for the 1D case" % init_field)
    for g in range(lo, hi):
        slab[g - lo + order] = pick(g + order)
    return slab


def make_mesh(zone_n, zone_len, order):
    """Global cell centres/widths with `order` ghost entries at both ends."""
    dxs = np.concatenate([np.full(n, L / float(n)) for n, L in zip(zone_n, zone_len)])
    nX = dxs.size
    x = np.zeros(nX + 2 * order)
    dx = np.zeros(nX + 2 * order)
    edge = 0.0
    for i, d in enumerate(dxs):
        dx[order + i] = d
        x[order + i] = edge + 0.5 * d
        edge += d
    for g in range(order - 1, -1, -1):
        dx[g] = dx[g + 1]
        x[g] = x[g + 1] - dx[g + 1]
    for g in range(nX + order, nX + 2 * order):
        dx[g] = dx[g - 1]
        x[g] = x[g - 1] + dx[g - 1]
    return nX, x, dx


def partition(nX, nranks):
    """Contiguous block partition of the cells; uneven blocks allowed (the reference requires
    nX % nranks == 0, src/mesh_setup.c:46-53)."""
    base, rem = divmod(nX, nranks)
    sizes = [base + (1 if r < rem else 0) for r in range(nranks)]
    starts = np.concatenate([[0], np.cumsum(sizes)])
    return [(int(starts[r]), int(starts[r + 1])) for r in range(nranks)]
