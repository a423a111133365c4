"""Simulation system: structure + potential + velocities + masses (reference `pantea/simulation/system.py:20-164`).

Thermodynamic observables: KE = 1/2 sum m v^2, T = 2 KE / (3 N kB), pressure = (2 KE + sum x.F) / (3 V).
"""
from __future__ import annotations

import ctypes as C
from copy import copy
from typing import Optional, Protocol, Tuple

import torch

from pantea_b200 import _lib
from pantea_b200.atoms.box import Box
from pantea_b200.atoms.element import ElementMap
from pantea_b200.atoms.structure import Structure
from pantea_b200.types import Array, Element
from pantea_b200.units import units

KB: float = units.BOLTZMANN_CONSTANT


def _get_kinetic_energy(velocities: Array, masses: Array) -> Array:
    """0.5 * sum(m v^2) through the fixed-order CUDA reduction (`pantea_md_kinetic_energy`)."""
    n = velocities.shape[0]
    out = torch.zeros(1, dtype=torch.float64, device=velocities.device)
    vel = velocities.contiguous()
    mass = masses.reshape(-1).to(velocities.dtype).contiguous()
    _lib.check(_lib.load().pantea_md_kinetic_energy(_lib.ptr(vel), _lib.ptr(mass), 0, n, _lib.ptr(out),
                                                    _lib.dtype_code(vel.dtype), _lib.stream_ptr()))
    return out[0].to(velocities.dtype)


def _get_temperature(velocities: Array, masses: Array) -> Array:
    return 2 * _get_kinetic_energy(velocities, masses) / (3 * velocities.shape[0] * KB)


def _get_virial(velocities: Array, masses: Array, positions: Array, forces: Array) -> Array:
    return 2 * _get_kinetic_energy(velocities, masses) + torch.sum(positions * forces)


def _calculate_center_of_mass(array: Array, masses: Array) -> Array:
    return torch.sum(masses * array, dim=0) / torch.sum(masses)


class PotentialInterface(Protocol):
    def __call__(self, structure: Structure) -> Array: ...

    def compute_forces(self, structure: Structure) -> Array: ...


class System:
    def __init__(self, potential: PotentialInterface, structure: Structure, velocities: Array, masses: Array) -> None:
        self.potential = potential
        self.structure = structure
        self.velocities = velocities
        self.masses = masses
        self.update_forces_from_positions()
        self.update_total_potential_energy_from_positions()

    @classmethod
    def from_structure(cls, structure: Structure, potential: PotentialInterface, temperature: float = 300.0,
                       seed: int = 2024) -> "System":
        masses = ElementMap.get_masses_from_structure(structure).reshape(-1, 1)
        velocities = cls.generate_random_velocities(temperature, masses, seed)
        return cls(potential, copy(structure), velocities, masses)

    @classmethod
    def generate_random_velocities(cls, temperature: float, masses: Array, seed: int) -> Array:
        """Maxwell-Boltzmann velocities rescaled to `temperature` with the COM velocity removed
        (`system.py:83-96`).  The reference draws from `jax.random.PRNGKey(seed)`; that stream needs JAX,
        so the normal draw here comes from a seeded torch CPU generator -- pass explicit velocities to
        `System(...)` when a bit-identical start is required."""
        gen = torch.Generator(device="cpu").manual_seed(int(seed))
        v = torch.randn((masses.shape[0], 3), generator=gen, dtype=torch.float64).to(device=masses.device, dtype=masses.dtype)
        v = v * torch.sqrt(torch.as_tensor(temperature, dtype=v.dtype, device=v.device) / _get_temperature(v, masses))
        return v - _calculate_center_of_mass(v, masses)

    def update_forces_from_positions(self) -> None:
        self.structure.forces = self.potential.compute_forces(self.structure)

    def update_total_potential_energy_from_positions(self) -> None:
        self.structure.total_energy = self.potential(self.structure)

    @classmethod
    def compute_forces(cls, potential: PotentialInterface, structure: Structure) -> Array:
        return potential.compute_forces(structure)

    def get_elements(self) -> Tuple[Element, ...]:
        return self.structure.get_elements()

    def get_pressure(self) -> Array:
        box = self.structure.box
        assert box is not None, "Calculating pressure... input structure must have PBC box"
        return _get_virial(self.velocities, self.masses, self.positions, self.forces) / (3.0 * box.volume)

    def get_temperature(self) -> Array:
        return _get_temperature(self.velocities, self.masses)

    def get_center_of_mass_velocity(self) -> Array:
        return _calculate_center_of_mass(self.velocities, self.masses)

    def get_center_of_mass_position(self) -> Array:
        return _calculate_center_of_mass(self.positions, self.masses)

    def get_potential_energy(self) -> Array:
        return self.potential(self.structure)

    def get_kinetic_energy(self) -> Array:
        return _get_kinetic_energy(self.velocities, self.masses)

    def get_total_energy(self) -> Array:
        return self.get_potential_energy() + self.get_kinetic_energy()

    @property
    def positions(self) -> Array:
        return self.structure.positions

    @property
    def forces(self) -> Array:
        return self.structure.forces

    @property
    def box(self) -> Optional[Box]:
        return self.structure.box

    @property
    def natoms(self) -> int:
        return self.structure.natoms

    def __repr__(self) -> str:
        return (f"{self.__class__.__name__}(potential={self.potential.__class__.__name__}, "
                f"structure={self.structure}, temperature={float(self.get_temperature()):.2f})")
