"""Seeded synthetic clouds shared by the tests: uniform [0,1)^3 like the reference's own
scripts (unit_test.py:15-16, emd_module.py:91-92), a sphere-surface variant, a variant with exact
duplicates, and an integer lattice (maximal tie stress for FPS / argmin)."""
import numpy as np


def uniform(b, n, seed):
    return np.random.default_rng(seed).random((b, n, 3), dtype=np.float32)


def sphere(b, n, seed):
    g = np.random.default_rng(seed).standard_normal((b, n, 3)).astype(np.float32)
    g /= np.linalg.norm(g, axis=2, keepdims=True) + 1e-12
    return (0.5 + 0.5 * g).astype(np.float32)


def duplicates(b, n, seed, frac=0.05):
    rng = np.random.default_rng(seed)
    x = rng.random((b, n, 3), dtype=np.float32)
    k = max(1, int(n * frac))
    for i in range(b):
        src = rng.integers(0, n, k)
        dst = rng.integers(0, n, k)
        x[i, dst] = x[i, src]
    return x


def lattice(b, n, seed, side=8):
    rng = np.random.default_rng(seed)
    return (rng.integers(0, side, (b, n, 3)).astype(np.float32) / side).astype(np.float32)


def clustered(b, n, seed):
    """Half of the points in a ball of radius 1e-3, the rest uniform: very dense grid cells next to sparse ones."""
    rng = np.random.default_rng(seed)
    x = rng.random((b, n, 3), dtype=np.float32)
    k = n // 2
    x[:, :k] = (0.3 + 1e-3 * rng.standard_normal((b, k, 3))).astype(np.float32)
    return x


def planar(b, n, seed):
    """z is constant: a zero-extent axis for the grid."""
    x = np.random.default_rng(seed).random((b, n, 3), dtype=np.float32)
    x[..., 2] = 0.25
    return x


def outliers(b, n, seed):
    """A unit cloud plus a few points very far away (the bounding box is mostly empty)."""
    rng = np.random.default_rng(seed)
    x = rng.random((b, n, 3), dtype=np.float32)
    x[:, :3] = (rng.random((b, 3, 3), dtype=np.float32) * 1000).astype(np.float32)
    return x


def shifted(b, n, seed):
    """A unit cloud displaced by a seed-dependent offset: the two clouds of a pair do not overlap."""
    return (np.random.default_rng(seed).random((b, n, 3), dtype=np.float32) + np.float32(3 * (seed % 5))).astype(np.float32)


def tiny(b, n, seed):
    """Coordinates around 1e-20: squared distances underflow to zero / denormals."""
    return (np.random.default_rng(seed).random((b, n, 3), dtype=np.float32) * np.float32(1e-20)).astype(np.float32)


def constant(b, n, seed):
    """Every point of a cloud identical (zero extent on all axes)."""
    p = np.random.default_rng(seed).random((b, 1, 3), dtype=np.float32)
    return np.repeat(p, n, axis=1)


def blob(b, n, seed):
    """A small Gaussian blob inside the unit cube: what PCN's decoder emits at random initialisation (BASELINE config
    C2) — against a ground-truth cloud that fills the cube, nearly every ground-truth point is far outside it."""
    rng = np.random.default_rng(seed)
    return (0.5 + 0.02 * rng.standard_normal((b, n, 3))).astype(np.float32)


KINDS = {"blob": blob, "clustered": clustered, "planar": planar, "outliers": outliers, "shifted": shifted, "tiny": tiny,
         "constant": constant, "uniform": uniform, "sphere": sphere, "duplicates": duplicates, "lattice": lattice}


def cloud(kind, b, n, seed):
    return KINDS[kind](b, n, seed)
