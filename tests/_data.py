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


KINDS = {"uniform": uniform, "sphere": sphere, "duplicates": duplicates, "lattice": lattice}


def cloud(kind, b, n, seed):
    return KINDS[kind](b, n, seed)
