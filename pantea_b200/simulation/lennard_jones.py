"""Lennard-Jones potential (API of reference `pantea/simulation/lennard_jones.py:15-123`).

Energy and "forces" come from `pantea_lj_energy_forces` on the CUDA neighbour rows.  Like the reference,
`compute_forces` returns +dE/dr (the gradient, `lennard_jones.py:100-123`), for both gradient methods.
"""
from __future__ import annotations

from typing import Literal

import torch

from pantea_b200 import _lib, engine
from pantea_b200.atoms.neighbor import _workspace
from pantea_b200.types import Array


class LJPotential:
    def __init__(self, sigma: float, epsilon: float, r_cutoff: float,
                 gradient_method: Literal["direct", "autodiff"] = "direct") -> None:
        if gradient_method not in ("direct", "autodiff"):
            raise ValueError("Unknown gradient method")
        self.sigma = float(sigma)
        self.epsilon = float(epsilon)
        self.r_cutoff = float(r_cutoff)
        self.gradient_method = gradient_method

    def _evaluate(self, structure, want_energy: bool, want_forces: bool):
        ws = _workspace(structure)
        dev = structure.positions.device
        types = torch.ones(structure.natoms, dtype=torch.int32, device=dev)
        ws.bind(structure.positions, types, engine.box_lengths(structure), self.r_cutoff)
        energy = torch.zeros((), dtype=structure.dtype, device=dev) if want_energy else None
        forces = torch.zeros((structure.natoms, 3), dtype=structure.dtype, device=dev) if want_forces else None
        _lib.check(_lib.load().pantea_lj_energy_forces(ws.handle, self.sigma, self.epsilon, None, _lib.ptr(forces),
                                                       _lib.ptr(energy), _lib.stream_ptr()))
        return energy, forces

    def __call__(self, structure) -> Array:
        return self._evaluate(structure, True, False)[0]

    def compute_forces(self, structure) -> Array:
        return self._evaluate(structure, False, True)[1]
