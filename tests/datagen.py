"""Seeded synthetic data shared by tests and bench.py (no files, no network)."""
import numpy as np


def uniform(n, dim, seed):
    """north_star: 'synthetic uniform-random f32 vectors' in [0,1)."""
    return np.random.default_rng(seed).random((n, dim), dtype=np.float32)


def sift_shaped(n, dim, seed, latent=32, noise=0.1, centers=1):
    """BASELINE.json configs[1]: 'synthetic f32 (SIFT-shaped)'.

    Low intrinsic dimension like real descriptor data: a `latent`-dim Gaussian (mixture if centers > 1) pushed
    through a fixed random linear map into `dim` dims, plus small isotropic noise, shifted/clipped to be >= 0.
    Calibrated with the oracle (100k points, M=32, ef=100): recall@10 = 0.995 with ~4.4k distance evaluations per
    query, i.e. as bandwidth-hungry as uniform data but reaching the metric's recall bar at the config's ef_search=100
    (uniform 128-d data needs ef in the high hundreds at 1M points: 20k points already give only 0.92).
    The map and the mixture centres depend only on (dim, latent, centers), never on `seed`, so points and
    queries drawn with different seeds come from the same distribution.
    """
    g = np.random.default_rng(0xC0FFEE + dim * 131 + latent)
    A = (g.standard_normal((latent, dim)) / np.sqrt(latent)).astype(np.float32)
    C = (2.0 * g.standard_normal((centers, latent))).astype(np.float32)
    r = np.random.default_rng(seed)
    out = np.empty((n, dim), dtype=np.float32)
    step = 1 << 18
    for s in range(0, n, step):
        m = min(step, n - s)
        z = r.standard_normal((m, latent), dtype=np.float32)
        if centers > 1:
            z = z + C[r.integers(0, centers, m)]
        x = z @ A + noise * r.standard_normal((m, dim), dtype=np.float32)
        out[s:s + m] = np.maximum(x + 4.0, 0.0)
    return out


def grid_ties(n, dim, seed, side=12):
    """Integer-grid points: many exact distance ties and duplicate vectors (SURVEY §7 hard part 1)."""
    return np.random.default_rng(seed).integers(0, side, (n, dim)).astype(np.float32)
