"""Seeded synthetic LArTPC-like events (tracks + showers) in the reference's blob layout.

The reference reads LArCV/ROOT files (R/uresnet/iotools/iotools_sparse.py:235-331) that do
not exist here; this generator produces the same *shape* of data the hot path consumes:

    data  [N, D+2] float64 : columns x, y, (z), batch_id, value
                             (R/uresnet/iotools/iotools_sparse.py:138-141,160 builds it with
                             np.pad(voxel, batch_id) ++ feature via np.concatenate -> float64)
    label [N, 1]   float32 : class id 0..4 = (HIP, MIP, Shower, Delta, Michel)
                             (R/bin/relabel_showers.py:87)

Recipe: per event 6 tracks + 3 showers, scale-free in the spatial
size ss; coordinates are unique within an event.  numpy only; no torch import here so the
generator can be used by the oracle, the tests and the bench alike.
"""
import numpy as np

N_TRACKS = 6
N_SHOWERS = 3
BASE_SEED = 1234


def _unit(rng, n=1):
    v = rng.normal(size=(n, 3))
    v /= np.linalg.norm(v, axis=1, keepdims=True) + 1e-12
    return v


def _track(rng, ss):
    start = rng.uniform(0.1 * ss, 0.9 * ss, size=3)
    d = _unit(rng)[0]
    length = rng.uniform(0.2 * ss, 0.9 * ss)
    nstep = max(2, int(length / 0.5))
    # direction random walk (multiple scattering), 0.5-voxel steps
    kicks = rng.normal(scale=0.004, size=(nstep, 3))
    dirs = d[None, :] + np.cumsum(kicks, axis=0)
    dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    pts = start[None, :] + np.cumsum(0.5 * dirs, axis=0)
    jitter = pts + rng.normal(scale=0.6, size=pts.shape)  # diffusion thickness
    return np.concatenate([pts, jitter], axis=0), pts


def _shower(rng, ss):
    vtx = rng.uniform(0.2 * ss, 0.8 * ss, size=3)
    axis = _unit(rng)[0]
    n = int(rng.uniform(3000, 12000) * ss / 512.0)
    n = max(n, 16)
    t = rng.gamma(3.0, ss / 18.0, size=n)
    r = rng.exponential(ss / 90.0, size=n) * (1.0 + 6.0 * t / ss)
    # orthonormal basis of the plane perpendicular to the axis
    a = np.cross(axis, [1.0, 0.0, 0.0])
    if np.linalg.norm(a) < 1e-3:
        a = np.cross(axis, [0.0, 1.0, 0.0])
    a /= np.linalg.norm(a)
    b = np.cross(axis, a)
    phi = rng.uniform(0, 2 * np.pi, size=n)
    return (vtx[None, :] + t[:, None] * axis[None, :]
            + (r * np.cos(phi))[:, None] * a[None, :]
            + (r * np.sin(phi))[:, None] * b[None, :])


def make_event(rng, ss, dim=3):
    """One event -> (coords int32 [n, dim], value float32 [n], label int64 [n]), coords unique."""
    pts, labs = [], []
    lengths, cores = [], []
    for _ in range(N_TRACKS):
        p, core = _track(rng, ss)
        pts.append(p)
        labs.append(np.full(len(p), 1, np.int64))
        lengths.append(len(core))
        cores.append(core)
    labs[int(np.argmin(lengths))][:] = 0  # shortest track -> HIP
    for _ in range(N_SHOWERS):
        p = _shower(rng, ss)
        pts.append(p)
        labs.append(np.full(len(p), 2, np.int64))
    # 20-voxel stubs at two track ends -> Delta (3) and Michel (4)
    for cls, core in ((3, cores[0]), (4, cores[1])):
        end = core[-1]
        d = _unit(rng)[0]
        stub = end[None, :] + np.arange(1, 21)[:, None] * 0.7 * d[None, :]
        pts.append(stub)
        labs.append(np.full(len(stub), cls, np.int64))
    p = np.concatenate(pts, axis=0)
    l = np.concatenate(labs, axis=0)
    c = np.floor(p).astype(np.int64)
    c = np.clip(c, 0, ss - 1)
    if dim == 2:
        c = c[:, :2]
    # unique coordinates (first occurrence keeps its label); lexicographic order like a
    # voxel-id sorted LArCV sparse tensor
    key = np.zeros(len(c), np.int64)
    for d_ in range(dim):
        key = key * ss + c[:, d_]
    _, first = np.unique(key, return_index=True)
    c = c[first].astype(np.int32)
    l = l[first]
    v = rng.lognormal(mean=-1.0, sigma=0.8, size=len(c)).astype(np.float32)
    return c, v, l


def make_batch(n_events, ss, dim=3, seed=BASE_SEED, rank=0, first_batch_id=0):
    """Concatenate events into the reference blob.

    Returns (data float64 [N, dim+2], label float32 [N, 1], n_per_event list).
    Event e of rank r is drawn sequentially from default_rng(seed + 1000 * r).
    """
    rng = np.random.default_rng(seed + 1000 * rank)
    rows, labs, counts = [], [], []
    for e in range(n_events):
        c, v, l = make_event(rng, ss, dim)
        bid = np.full((len(c), 1), first_batch_id + e, np.float64)
        rows.append(np.concatenate([c.astype(np.float64), bid, v[:, None].astype(np.float64)], axis=1))
        labs.append(l[:, None].astype(np.float32))
        counts.append(len(c))
    return np.concatenate(rows, 0), np.concatenate(labs, 0), counts


class EventStream(object):
    """Endless stream of batches with a persistent RNG (different events every step)."""

    def __init__(self, n_events, ss, dim=3, seed=BASE_SEED, rank=0):
        self.n_events, self.ss, self.dim = n_events, ss, dim
        self.rng = np.random.default_rng(seed + 1000 * rank)

    def next(self):
        rows, labs = [], []
        for e in range(self.n_events):
            c, v, l = make_event(self.rng, self.ss, self.dim)
            bid = np.full((len(c), 1), e, np.float64)
            rows.append(np.concatenate([c.astype(np.float64), bid, v[:, None].astype(np.float64)], axis=1))
            labs.append(l[:, None].astype(np.float32))
        return np.concatenate(rows, 0), np.concatenate(labs, 0)
