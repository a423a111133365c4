"""Atom-centred symmetry function descriptor (API of reference `pantea/descriptors/acsf/acsf.py:26-160`).

`ACSF(...)(structure, atom_index=None) -> [n, n_sf]` and `ACSF.grad(structure, atom_index=None) ->
[n, n_sf, 3]` run the fused CUDA kernel through `pantea_acsf_compute`.  Semantics kept from the
reference: `__call__` without `atom_index` evaluates the atoms of the central element
(`acsf.py:66-67`) and validates explicit indices against it (`:70-80`); `grad` without `atom_index`
evaluates *all* atoms (`:103-104`) and returns dG_i/dr_i in the central role only (`:215-218`).
"""
from __future__ import annotations

import itertools
from typing import List, Optional, Tuple

import torch

from pantea_b200 import engine
from pantea_b200.atoms.structure import Structure
from pantea_b200.descriptors.acsf.symmetry import (AngularSymmetryFunction, NeighborElements,
                                                   RadialSymmetryFunction)
from pantea_b200.logger import logger
from pantea_b200.types import Array

AssignedRadialSymmetryFunction = Tuple[RadialSymmetryFunction, NeighborElements]
AssignedAngularSymmetryFunction = Tuple[AngularSymmetryFunction, NeighborElements]


class AtomCenteredSymmetryFunction:
    def __init__(
        self,
        central_element: str,
        radial_symmetry_functions: Tuple[AssignedRadialSymmetryFunction, ...] = (),
        angular_symmetry_functions: Tuple[AssignedAngularSymmetryFunction, ...] = (),
    ) -> None:
        self.central_element = central_element
        self.radial_symmetry_functions = tuple(radial_symmetry_functions)
        self.angular_symmetry_functions = tuple(angular_symmetry_functions)
        self._device_potential: Optional[engine.DevicePotential] = None

    # ------------------------------------------------------------------ host description
    def symfunc_records(self) -> List[engine.SymFuncRecord]:
        """Radial functions first, then angular, in declaration order (`acsf.py:175-201`)."""
        records = []
        for sf, nb in self.radial_symmetry_functions:
            records.append(engine.SymFuncRecord(sf.kind, sf.cfn.code, sf.r_cutoff, nb.neighbor_j, None,
                                                getattr(sf, "eta", 0.0), getattr(sf, "r_shift", 0.0)))
        for sf, nb in self.angular_symmetry_functions:
            records.append(engine.SymFuncRecord(sf.kind, sf.cfn.code, sf.r_cutoff, nb.neighbor_j, nb.neighbor_k,
                                                sf.eta, sf.r_shift, sf.lambda0, sf.zeta))
        return records

    def _potential(self) -> engine.DevicePotential:
        if self._device_potential is None:
            record = engine.ElementRecord(self.central_element, self.symfunc_records())
            self._device_potential = engine.DevicePotential([record])
        return self._device_potential

    def _bind(self, structure: Structure) -> engine.Workspace:
        pot = self._potential()
        for sf, nb in itertools.chain(self.radial_symmetry_functions, self.angular_symmetry_functions):
            for el in (nb.neighbor_j, nb.neighbor_k):
                if el is not None and el not in structure.element_map.element_to_atom_type:
                    raise KeyError(el)  # reference: structure.element_map[element] (acsf.py:183,198-199)
        ws = pot.workspace(structure.natoms, structure.dtype, engine.number_density(structure))
        ws.bind(structure.positions, engine.remap_types(structure, pot.type_of), engine.box_lengths(structure), pot.r_cutoff)
        return ws

    # ------------------------------------------------------------------ evaluation
    def __call__(self, structure: Structure, atom_index: Optional[Array] = None) -> Array:
        if self.num_symmetry_functions == 0:
            logger.warning("No symmetry function was found")
        if atom_index is None:
            index = structure.select(self.central_element)
        else:
            index = torch.atleast_1d(torch.as_tensor(atom_index, device=structure.atom_types.device))
            central_type = structure.element_map.element_to_atom_type[self.central_element]
            if not bool(torch.all(structure.atom_types[index.long()] == central_type)):
                logger.error(
                    f"Inconsistent central element '{self.central_element}':  input atom index={atom_index}",
                    exception=ValueError,
                )
        ws = self._bind(structure)
        pot = self._potential()
        values, _ = ws.acsf(pot.slot(self.central_element), self.num_symmetry_functions, index, True, False)
        return values

    def grad(self, structure: Structure, atom_index: Optional[Array] = None) -> Array:
        index = None
        if atom_index is not None:
            index = torch.atleast_1d(torch.as_tensor(atom_index, device=structure.atom_types.device))
            if not bool(torch.all((0 <= index) & (index < structure.natoms))):
                logger.error(
                    f"unexpected {atom_index=}.Input index must be between [0, {structure.natoms})",
                    exception=ValueError,
                )
        ws = self._bind(structure)
        pot = self._potential()
        _, grads = ws.acsf(pot.slot(self.central_element), self.num_symmetry_functions, index, False, True)
        return grads

    # ------------------------------------------------------------------ properties
    @property
    def num_radial_symmetry_functions(self) -> int:
        return len(self.radial_symmetry_functions)

    @property
    def num_angular_symmetry_functions(self) -> int:
        return len(self.angular_symmetry_functions)

    @property
    def num_symmetry_functions(self) -> int:
        return self.num_radial_symmetry_functions + self.num_angular_symmetry_functions

    @property
    def r_cutoff(self) -> float:
        return max(sf.r_cutoff for sf, _ in itertools.chain(self.radial_symmetry_functions,
                                                            self.angular_symmetry_functions))

    def __repr__(self) -> str:
        return (f"{self.__class__.__name__}(central_element='{self.central_element}'"
                f", num_symmetry_functions={self.num_symmetry_functions})")


ACSF = AtomCenteredSymmetryFunction
