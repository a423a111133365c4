"""Deterministic synthetic inputs for measurement and tests.

Water at 0.0334 molecules/A^3 on a jittered simple-cubic lattice with random orientations;
numpy `default_rng(seed)` draws in the order jitter(n_mol x 3) -> quaternions(n_mol x 4).
Atom order O,H,H per molecule; types H = 1, O = 2.
"""
from __future__ import annotations

import math
from typing import Tuple

import numpy as np

BOHR_PER_ANGSTROM = 1.0 / 0.529177249
FROM_ATOMIC_MASS = 1.0 / 5.48579957163e-4
KB = 3.166811563e-6
_MASS_U = {1: 1.008, 2: 15.999}


def water_box(n_atoms: int, seed: int = 2024) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """-> (positions [N,3] wrapped into the box, types [N] int32, box lengths [3]) in Bohr."""
    assert n_atoms % 3 == 0
    n_mol = n_atoms // 3
    rho = 0.0334 / BOHR_PER_ANGSTROM**3
    L = (n_mol / rho) ** (1.0 / 3.0)
    m = int(math.ceil(n_mol ** (1.0 / 3.0) - 1e-9))
    rng = np.random.default_rng(seed)
    jitter = rng.uniform(-0.15, 0.15, size=(n_mol, 3))
    quat = rng.standard_normal(size=(n_mol, 4))
    quat /= np.linalg.norm(quat, axis=1, keepdims=True)
    idx = np.stack(np.unravel_index(np.arange(n_mol), (m, m, m)), axis=1).astype(np.float64)
    o = (idx + 0.5 + jitter) * (L / m)
    w, x, y, z = quat.T
    rot = np.stack([
        np.stack([1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)], axis=1),
        np.stack([2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)], axis=1),
        np.stack([2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)], axis=1),
    ], axis=1)
    half = math.radians(52.26)
    h_local = np.array([[math.sin(half), 0.0, math.cos(half)], [-math.sin(half), 0.0, math.cos(half)]]) * 1.80885
    pos = np.stack([o, o + rot @ h_local[0], o + rot @ h_local[1]], axis=1).reshape(-1, 3)
    types = np.tile(np.array([2, 1, 1], dtype=np.int32), n_mol)
    box = np.array([L, L, L])
    return np.remainder(pos, box), types, box


def water_masses(types: np.ndarray) -> np.ndarray:
    return np.asarray([_MASS_U[int(t)] * FROM_ATOMIC_MASS for t in types], dtype=np.float64)


def md_velocities(types: np.ndarray, temperature: float = 300.0, seed: int = 2025) -> np.ndarray:
    """Normal draw, rescaled to `temperature`, COM velocity removed (reference system.py:91-96)."""
    n = len(types)
    m = water_masses(types)[:, None]
    v = np.random.default_rng(seed).standard_normal((n, 3))
    t_now = 2 * (0.5 * np.sum(m * v * v)) / (3 * n * KB)
    v = v * math.sqrt(temperature / t_now)
    return v - np.sum(m * v, axis=0) / np.sum(m)
