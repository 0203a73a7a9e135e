"""Synthetic workloads of the named BASELINE.json shapes.

Inputs only -- no distance arithmetic here.  Every generator is indexed by absolute frame /
row number and seeded per basin segment, so any shard [begin, begin+count) of a trajectory
is identical however the run is partitioned across GPUs.

G_traj   C-alpha traces: `n_basins` freely-rotating chains (bond 0.38 nm, bond angle
         1.5376 rad as in examples/mld/semirigid_helices.py:63-65); frame f belongs to basin
         floor(n_basins*f/n_total) and is  basin + sum_{m<8} a_m(f) V_m  (sinusoidal modes
         along the chain, a_m AR(1) in f with rho=0.99, sigma=0.05 nm) + N(0, 0.02 nm) per
         coordinate, then a uniformly random rigid rotation and a U(-1,1) nm translation,
         rounded to the 0.001 nm XTC grid.  Masses are all 12.0107 (carbon).
G_phipsi rows of interleaved (sin t_j, cos t_j) exactly as angles_to_sincos.cpp:109-110
         writes them: dim/2 angles per row, basin means U(-pi,pi), N(0,0.3 rad) noise with
         AR(1) rho=0.9 inside a basin.
"""
import numpy as np
from scipy.signal import lfilter

CARBON_MASS = 12.0107


def _segments(n_total, n_basins, begin, count):
    """Yield (basin, seg_begin, seg_end, lo, hi): absolute overlap [lo,hi) with the request."""
    for b in range(n_basins):
        s0 = (b * n_total + n_basins - 1) // n_basins  # first f with floor(n_basins*f/n_total) == b
        s1 = ((b + 1) * n_total + n_basins - 1) // n_basins
        lo, hi = max(s0, begin), min(s1, begin + count)
        if lo < hi:
            yield b, s0, s1, lo, hi


def _ar1(rng, n, rho, sigma, width):
    eps = rng.standard_normal((n, width))
    x0 = eps[0] * sigma
    y = lfilter([np.sqrt(1.0 - rho * rho) * sigma], [1.0, -rho], eps[1:], axis=0, zi=(rho * x0)[None, :])[0]
    return np.concatenate([x0[None, :], y], axis=0)


def _chain(rng, n_atoms, bond=0.38, angle=1.5376):
    pos = np.zeros((n_atoms, 3))
    pos[1] = [bond, 0.0, 0.0]
    for i in range(2, n_atoms):
        b = pos[i - 1] - pos[i - 2]
        b /= np.linalg.norm(b)
        # any unit vector perpendicular to b, rotated by a random dihedral
        t = np.cross(b, [1.0, 0.0, 0.0] if abs(b[0]) < 0.9 else [0.0, 1.0, 0.0])
        t /= np.linalg.norm(t)
        u = np.cross(b, t)
        phi = rng.uniform(-np.pi, np.pi)
        perp = np.cos(phi) * t + np.sin(phi) * u
        pos[i] = pos[i - 1] + bond * (-np.cos(angle) * b + np.sin(angle) * perp)
    return pos - pos.mean(axis=0)


def _rotations(rng, n):
    q = rng.standard_normal((n, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    w, x, y, z = q.T
    R = np.empty((n, 3, 3))
    R[:, 0, 0] = 1 - 2 * (y * y + z * z); R[:, 0, 1] = 2 * (x * y - z * w); R[:, 0, 2] = 2 * (x * z + y * w)
    R[:, 1, 0] = 2 * (x * y + z * w); R[:, 1, 1] = 1 - 2 * (x * x + z * z); R[:, 1, 2] = 2 * (y * z - x * w)
    R[:, 2, 0] = 2 * (x * z - y * w); R[:, 2, 1] = 2 * (y * z + x * w); R[:, 2, 2] = 1 - 2 * (x * x + y * y)
    return R


def traj_frames(n_total, n_atoms=300, n_basins=16, seed=20260117, begin=0, count=None, out=None):
    """float32 [count, n_atoms, 3] (nm): frames [begin, begin+count) of the synthetic trajectory."""
    count = n_total - begin if count is None else count
    if out is None:
        out = np.empty((count, n_atoms, 3), dtype=np.float32)
    n_modes = 8
    s = np.arange(n_atoms) / max(1, n_atoms - 1)
    for b, s0, s1, lo, hi in _segments(n_total, n_basins, begin, count):
        rng = np.random.default_rng([seed, b])
        basin = _chain(rng, n_atoms)
        # mode m displaces the chain along a fixed random direction with a sinusoidal profile
        dirs = rng.standard_normal((n_modes, 3))
        dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
        modes = np.stack([np.sin(np.pi * (m + 1) * s)[:, None] * dirs[m][None, :] for m in range(n_modes)])
        amps = _ar1(rng, s1 - s0, 0.99, 0.05, n_modes)  # whole segment: continuity across shards
        step = 8192
        for k in range((lo - s0) // step, (hi - s0 + step - 1) // step):
            a0 = s0 + k * step
            a1 = min(s1, a0 + step)
            c0, c1 = max(a0, lo), min(a1, hi)
            # one stream per aligned chunk of the segment: a shard never depends on where it starts
            crng = np.random.default_rng([seed, b, k, 1])
            n = a1 - a0
            noise = crng.standard_normal((n, n_atoms, 3)) * 0.02
            R = _rotations(crng, n)
            T = crng.uniform(-1.0, 1.0, (n, 3))
            sl = slice(c0 - a0, c1 - a0)
            x = basin[None] + np.einsum("fm,mad->fad", amps[c0 - s0:c1 - s0], modes) + noise[sl]
            x = np.einsum("fij,faj->fai", R[sl], x) + T[sl][:, None, :]
            out[c0 - begin:c1 - begin] = (np.round(x * 1000.0) * np.float32(0.001)).astype(np.float32)
    return out


def traj_masses(n_atoms=300):
    return np.full(n_atoms, CARBON_MASS, dtype=np.float32)


def phipsi_rows(n_total, dim=512, n_basins=64, seed=20260119, begin=0, count=None):
    """float64 [count, dim]: rows [begin, begin+count) of the sin/cos feature matrix."""
    count = n_total - begin if count is None else count
    out = np.empty((count, dim), dtype=np.float64)
    n_ang = dim // 2
    for b, s0, s1, lo, hi in _segments(n_total, n_basins, begin, count):
        rng = np.random.default_rng([seed, b])
        mean = rng.uniform(-np.pi, np.pi, n_ang)
        dev = _ar1(rng, s1 - s0, 0.9, 0.3, n_ang)
        th = mean[None, :] + dev[lo - s0:hi - s0]
        out[lo - begin:hi - begin, 0::2] = np.sin(th)
        out[lo - begin:hi - begin, 1::2] = np.cos(th)
    return out
