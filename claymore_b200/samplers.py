"""Synthetic particle-cloud generators (deterministic; no RNG unless jitter is asked for).

`uniform_box` follows the reference's sample_uniform_box (Library/MnBase/Geometry/GeometrySampler.h:11-37):
8 particles per cell at +-0.25 dx around the cell "centre" i*dx.
"""
import numpy as np

_OFFS = np.array([[a, b, c] for a in (-1, 1) for b in (-1, 1) for c in (-1, 1)], dtype=np.float64) * 0.25


def uniform_box(dx, minc, maxc, dtype=np.float32):
    i, j, k = np.meshgrid(np.arange(minc[0], maxc[0]), np.arange(minc[1], maxc[1]), np.arange(minc[2], maxc[2]), indexing="ij")
    cells = np.stack([i, j, k], -1).reshape(-1, 1, 3).astype(np.float64)
    return ((cells + _OFFS[None]) * dx).reshape(-1, 3).astype(dtype)


def sphere(dx, center, radius, dtype=np.float32):
    """Lattice sampler (8 per cell) masked by |x - c| <= r."""
    c = np.asarray(center, dtype=np.float64)
    lo = np.floor((c - radius) / dx).astype(int) - 1
    hi = np.ceil((c + radius) / dx).astype(int) + 2
    out = []
    # slab by slab to bound memory for 20 M-particle spheres
    for i in range(lo[0], hi[0]):
        j, k = np.meshgrid(np.arange(lo[1], hi[1]), np.arange(lo[2], hi[2]), indexing="ij")
        cells = np.stack([np.full_like(j, i), j, k], -1).reshape(-1, 1, 3).astype(np.float64)
        p = ((cells + _OFFS[None]) * dx).reshape(-1, 3)
        m = ((p - c) ** 2).sum(-1) <= radius * radius
        if m.any():
            out.append(p[m].astype(dtype))
    return np.concatenate(out, 0) if out else np.zeros((0, 3), dtype)


def jitter(pos, dx, amount=0.2, seed=0):
    rng = np.random.default_rng(seed)
    return (pos + rng.uniform(-amount * dx, amount * dx, size=pos.shape)).astype(pos.dtype)


def split_slabs(pos, parts, axis=0):
    """MGSP static partition: equal-count particle sets along `axis` (Projects/MGSP/mgsp.cu:34-81 assigns one model per device)."""
    order = np.argsort(pos[:, axis], kind="stable")
    return [pos[idx] for idx in np.array_split(order, parts)]
