"""Metropolis Monte-Carlo driver (reference `pantea/simulation/monte_carlo.py:24-95`).

Host-driven like the reference; the order of the numpy RNG calls (displacements, atom indices,
acceptance draw) is preserved so that a given seed reproduces the same Markov chain.
"""
from __future__ import annotations

import math

import numpy as np
import torch

from pantea_b200.simulation.system import System
from pantea_b200.units import units

KB: float = units.BOLTZMANN_CONSTANT


class MCSimulator:
    def __init__(self, translate_step: float, target_temperature: float, movements_per_step: int = 1,
                 seed: int = 12345) -> None:
        self.translate_step = translate_step
        self.target_temperature = target_temperature
        self.movements_per_step = movements_per_step
        self.step: int = 0
        np.random.seed(seed)

    def simulate_one_step(self, system: System) -> None:
        self.metropolis_algorithm(system)
        self.step += 1

    def metropolis_algorithm(self, system: System) -> None:
        displacements = np.random.uniform(low=-self.translate_step, high=self.translate_step,
                                          size=(self.movements_per_step, 3))
        atom_indices = np.random.randint(low=0, high=system.natoms, size=(self.movements_per_step,))
        pos = system.positions
        new_positions = pos.clone()
        # `.at[indices].add(...)` accumulates repeated indices
        new_positions.index_add_(0, torch.as_tensor(atom_indices, device=pos.device, dtype=torch.long),
                                 torch.as_tensor(displacements, device=pos.device, dtype=pos.dtype))
        new_structure = system.structure.replace(positions=new_positions)  # re-wraps, like dataclasses.replace
        new_energy = system.potential(new_structure)
        energy = system.structure.total_energy
        if float(new_energy) <= float(energy):
            accept = True
        else:
            prob = math.exp(-(float(new_energy) - float(energy)) / (KB * self.target_temperature))
            accept = prob >= np.random.uniform(0.0, 1.0)
        if accept:
            system.structure.total_energy = new_energy
            system.structure.positions = new_structure.positions

    def repr_physical_params(self, system: System) -> str:
        return f"{self.step:<10} Epot[Ha]:{float(system.get_potential_energy()):<15.10f} "
