"""CPU restatement (numpy, fp64) of Ewald / smooth particle-mesh Ewald electrostatics -- SURVEY 8f row 1.

TEST INFRASTRUCTURE ONLY: imported by tests/, never by the product package.
Parity unpinned: the reference's reciprocal space lives in the un-vendored crate `ewald` 0.1.15 (Cargo.toml:30);
this file restates the published algorithm (Essmann et al., J. Chem. Phys. 103, 8577 (1995), order-4 B-splines)
and is anchored by known answers: the exact Ewald sum below reproduces the Madelung constant of rock salt, and
SPME converges to the exact sum.  Charges carry sqrt(332.0522) (Coulomb constant 1, energies in kcal/mol).
"""
import numpy as np
from math import erf, erfc


def ewald_recip_exact(xyzq, ext, alpha, kmax=12):
    """Reciprocal-space Ewald sum by direct summation over |m_a| <= kmax: (energy, forces (n,3))."""
    x = xyzq[:, :3].astype(np.float64)
    q = xyzq[:, 3].astype(np.float64)
    L = np.asarray(ext, np.float64)
    V = L.prod()
    rng = np.arange(-kmax, kmax + 1)
    m = np.stack(np.meshgrid(rng, rng, rng, indexing="ij"), -1).reshape(-1, 3)
    m = m[(m != 0).any(1)]
    h = m / L                                     # reciprocal vectors m / L
    h2 = (h * h).sum(1)
    coef = np.exp(-np.pi ** 2 * h2 / alpha ** 2) / (np.pi * V * h2)
    keep = coef > 1e-18 * coef.max()
    h, coef = h[keep], coef[keep]
    e, f = 0.0, np.zeros_like(x)
    for s in range(0, len(h), 4096):
        hh, cc = h[s:s + 4096], coef[s:s + 4096]
        ph = 2 * np.pi * x @ hh.T                 # (n, k)
        c, sn = np.cos(ph), np.sin(ph)
        Sr, Si = q @ c, q @ sn                    # structure factor
        e += 0.5 * (cc * (Sr * Sr + Si * Si)).sum()
        # F_i = q_i sum_m coef 2 pi h (sin(ph_i) Sr - cos(ph_i) Si)
        w = cc[None, :] * (sn * Sr[None, :] - c * Si[None, :])
        f += 2 * np.pi * q[:, None] * (w @ hh)
    return e, f


def self_energy(xyzq, alpha):
    return -alpha / np.sqrt(np.pi) * float((xyzq[:, 3].astype(np.float64) ** 2).sum())


def real_space_brute(xyzq, ext, alpha, rc):
    """erfc(alpha r)/r over minimum-image pairs with r < rc (rc <= L/2): (energy, forces)."""
    x = xyzq[:, :3].astype(np.float64)
    q = xyzq[:, 3].astype(np.float64)
    L = np.asarray(ext, np.float64)
    e, f = 0.0, np.zeros_like(x)
    for i in range(len(x)):
        d = x[i] - x
        d -= np.rint(d / L) * L
        r = np.sqrt((d * d).sum(1))
        ok = (r < rc) & (r > 0)
        rr = r[ok]
        erfc_ = np.array([erfc(alpha * v) for v in rr])
        e += 0.5 * (q[i] * q[ok] * erfc_ / rr).sum()
        mag = q[i] * q[ok] * (erfc_ / rr ** 2 + 2 * alpha / np.sqrt(np.pi) * np.exp(-(alpha * rr) ** 2) / rr) / rr
        f[i] = (d[ok] * mag[:, None]).sum(0)
    return e, f


def bspline4(w):
    """weights and derivatives of grid points floor(u) - 3 + t, t = 0..3, for w = u - floor(u)."""
    v = 1.0 - w
    th = np.stack([v ** 3 / 6, (3 * w ** 3 - 6 * w ** 2 + 4) / 6, (-3 * w ** 3 + 3 * w ** 2 + 3 * w + 1) / 6, w ** 3 / 6], -1)
    dth = np.stack([-0.5 * v ** 2, 0.5 * (3 * w ** 2 - 4 * w), 0.5 * (-3 * w ** 2 + 2 * w + 1), 0.5 * w ** 2], -1)
    return th, dth


def bmod4(K):
    t = 2 * np.pi * np.arange(K) / K
    return (2.0 / 3.0 + np.cos(t) / 3.0) ** 2


def influence(K, ext, alpha):
    """B(m) C(m) on the rfft half grid (K1, K2, K3//2+1)."""
    L = np.asarray(ext, np.float64)
    V = L.prod()
    m = [np.where(np.arange(k) > k // 2, np.arange(k) - k, np.arange(k)) for k in K]
    h1, h2, h3 = np.meshgrid(m[0] / L[0], m[1] / L[1], m[2][:K[2] // 2 + 1] / L[2], indexing="ij")
    msq = h1 ** 2 + h2 ** 2 + h3 ** 2
    b = bmod4(K[0])[:, None, None] * bmod4(K[1])[None, :, None] * bmod4(K[2])[None, None, :K[2] // 2 + 1]
    with np.errstate(divide="ignore", invalid="ignore"):
        bc = np.exp(-np.pi ** 2 * msq / alpha ** 2) / (np.pi * V * msq * b)
    bc[0, 0, 0] = 0.0
    return bc


def _coords(xyzq, lo, ext, K):
    s = (xyzq[:, :3].astype(np.float64) - np.asarray(lo, np.float64)) / np.asarray(ext, np.float64)
    s -= np.floor(s)
    u = s * np.asarray(K, np.float64)
    k0 = np.minimum(np.floor(u).astype(np.int64), np.asarray(K) - 1)
    return k0, u - k0


def spread(xyzq, lo, ext, K):
    k0, w = _coords(xyzq, lo, ext, K)
    q = xyzq[:, 3].astype(np.float64)
    th = [bspline4(w[:, a])[0] for a in range(3)]
    grid = np.zeros(K, np.float64)
    for a in range(4):
        ia = (k0[:, 0] - 3 + a) % K[0]
        for b in range(4):
            ib = (k0[:, 1] - 3 + b) % K[1]
            for c in range(4):
                ic = (k0[:, 2] - 3 + c) % K[2]
                np.add.at(grid, (ia, ib, ic), q * th[0][:, a] * th[1][:, b] * th[2][:, c])
    return grid


def spme(xyzq, lo, ext, alpha, K):
    """SPME reciprocal energy and forces with the algorithm of pme.cu in fp64 (order 4, rfftn / irfftn)."""
    K = tuple(int(k) for k in K)
    grid = spread(xyzq, lo, ext, K)
    fq = np.fft.rfftn(grid)
    bc = influence(K, ext, alpha)
    mult = np.full(K[2] // 2 + 1, 2.0)
    mult[0] = 1.0
    if K[2] % 2 == 0:
        mult[-1] = 1.0
    energy = 0.5 * float((bc * (fq.real ** 2 + fq.imag ** 2) * mult[None, None, :]).sum())
    phi = np.fft.irfftn(fq * bc, s=K, axes=(0, 1, 2)) * np.prod(K)         # unnormalised backward transform
    k0, w = _coords(xyzq, lo, ext, K)
    q = xyzq[:, 3].astype(np.float64)
    td = [bspline4(w[:, a]) for a in range(3)]
    f = np.zeros((len(q), 3), np.float64)
    for a in range(4):
        ia = (k0[:, 0] - 3 + a) % K[0]
        for b in range(4):
            ib = (k0[:, 1] - 3 + b) % K[1]
            for c in range(4):
                ic = (k0[:, 2] - 3 + c) % K[2]
                p = phi[ia, ib, ic]
                f[:, 0] += p * td[0][1][:, a] * td[1][0][:, b] * td[2][0][:, c]
                f[:, 1] += p * td[0][0][:, a] * td[1][1][:, b] * td[2][0][:, c]
                f[:, 2] += p * td[0][0][:, a] * td[1][0][:, b] * td[2][1][:, c]
    f *= -q[:, None] * (np.asarray(K, np.float64) / np.asarray(ext, np.float64))[None, :]
    return energy, f


def excl_correction(xyzq, ext, periodic, excl_start, excl_idx, alpha):
    """-qq erf(alpha r)/r for every excluded pair (rows list both directions): (energy, forces)."""
    x = xyzq[:, :3].astype(np.float64)
    q = xyzq[:, 3].astype(np.float64)
    L = np.asarray(ext, np.float64)
    e, f = 0.0, np.zeros_like(x)
    for i in range(len(x)):
        for t in range(excl_start[i], excl_start[i + 1]):
            j = excl_idx[t]
            if j == i:
                continue
            d = x[i] - x[j]
            if periodic:
                d -= np.rint(d / L) * L
            r = np.sqrt(d @ d)
            er = erf(alpha * r)
            e += 0.5 * (-q[i] * q[j] * er / r)
            f[i] += d * q[i] * q[j] * (2 * alpha / np.sqrt(np.pi) * np.exp(-(alpha * r) ** 2) - er / r) / (r * r)
    return e, f
