"""Atomic structure container (reference `pantea/atoms/structure.py:21-387`).

Holds per-atom arrays resident in HBM (torch CUDA tensors) plus the static element map and
box.  Positions are wrapped into the box at construction exactly as the reference does
(`structure.py:80-81`).  Attributes are mutated by rebinding, like the reference.
"""
from __future__ import annotations

from collections import defaultdict
from typing import Any, Dict, Iterator, List, NamedTuple, Optional, Tuple

import numpy as np
import torch

from pantea_b200.atoms.box import Box, _wrap_into_box
from pantea_b200.atoms.element import ElementMap
from pantea_b200.logger import logger
from pantea_b200.types import Array, Dtype, Element, as_torch_dtype, asarray, default_dtype
from pantea_b200.units import units

_ATOM_ATTRIBUTES: Tuple[str, ...] = (
    "positions", "forces", "energies", "charges", "total_energy", "total_charge", "atom_types",
)


class StructureAsKernelArgs(NamedTuple):
    """Plain-array view handed to the compute kernels (reference `structure.py:377-387`)."""

    positions: Array
    atom_types: Array
    lattice: Optional[Array]
    total_energy: Array
    element_map: Dict[Element, int]


class Structure:
    def __init__(
        self,
        positions: Array,
        forces: Array,
        energies: Array,
        charges: Array,
        total_energy: Array,
        total_charge: Array,
        atom_types: Array,
        element_map: ElementMap,
        box: Optional[Box] = None,
    ) -> None:
        self.positions = positions
        self.forces = forces
        self.energies = energies
        self.charges = charges
        self.total_energy = total_energy
        self.total_charge = total_charge
        self.atom_types = atom_types
        self.element_map = element_map
        self.box = box
        self._select_cache: Dict[Element, Array] = {}
        if self.box is not None:
            self.positions = _wrap_into_box(self.positions, self.lattice)

    # ------------------------------------------------------------------ construction
    @classmethod
    def from_dict(cls, data: Dict[str, Any], dtype: Optional[Dtype] = None) -> "Structure":
        dt = as_torch_dtype(dtype) if dtype is not None else default_dtype.FLOATX
        input_data: Dict[str, List] = defaultdict(list, data)
        kwargs: Dict[str, Any] = {}
        try:
            element_map = ElementMap.from_list(input_data["elements"])
            kwargs.update(cls._init_arrays(input_data, element_map, dt))
            kwargs["element_map"] = element_map
            kwargs["box"] = cls._init_box(input_data["lattice"], dt)
        except KeyError:
            logger.error(
                "Cannot find at least one of the expected keyword in the input data dictionary.",
                exception=KeyError,
            )
        return cls(**kwargs)

    @classmethod
    def from_ase(cls, atoms: Any, dtype: Optional[Dtype] = None) -> "Structure":
        """Build from an `ase.Atoms` (ase is an optional dependency; reference `structure.py:83-137`)."""
        data: Dict[str, Any] = {
            "elements": [ElementMap.get_element_from_atomic_number(int(n)) for n in atoms.get_atomic_numbers()],
            "lattice": np.asarray(atoms.get_cell()) * units.FROM_ANGSTROM,
            "positions": np.asarray(atoms.get_positions()) * units.FROM_ANGSTROM,
        }
        for attr, ase_attr in (("energies", "potential_energies"), ("charges", "charges")):
            try:
                data[attr] = getattr(atoms, f"get_{ase_attr}")()
            except RuntimeError:
                continue
        for attr in ("energies", "charges"):
            if attr in data:
                data[f"total_{attr}"] = sum(data[attr])
        return cls.from_dict(data, dtype=dtype)

    @classmethod
    def _init_arrays(cls, data: Dict[str, Any], element_map: ElementMap, dtype: Dtype) -> Dict[str, Array]:
        arrays: Dict[str, Array] = {}
        for attr in _ATOM_ATTRIBUTES:
            if attr == "atom_types":
                arr = asarray(
                    [element_map.get_atom_type_from_element(name) for name in data["elements"]],
                    dtype=default_dtype.INDEX,
                )
            else:
                arr = asarray(np.asarray(data[attr], dtype=np.float64), dtype=dtype)
            arrays[attr] = torch.squeeze(arr)
        return arrays

    @classmethod
    def _init_box(cls, lattice: Any, dtype: Dtype) -> Optional[Box]:
        if len(lattice) > 0:
            return Box.from_list(np.asarray(lattice, dtype=np.float64), dtype=dtype)
        return None

    @classmethod
    def _get_atom_attributes(cls) -> Tuple[str, ...]:
        return _ATOM_ATTRIBUTES

    # ------------------------------------------------------------------ properties
    @property
    def natoms(self) -> int:
        return int(self.positions.shape[0])

    @property
    def dtype(self) -> Dtype:
        return self.positions.dtype

    @property
    def lattice(self) -> Optional[Array]:
        return self.box.lattice if self.box is not None else None

    def get_unique_elements(self) -> Tuple[Element, ...]:
        return self.element_map.unique_elements

    def get_elements(self) -> Tuple[Element, ...]:
        to_element = self.element_map.atom_type_to_element
        return tuple(str(to_element[t]) for t in self.atom_types.detach().cpu().tolist())

    def select(self, element: Element) -> Array:
        """Ascending indices of all atoms of `element` (reference `structure.py:252-261`)."""
        cached = self._select_cache.get(element)
        if cached is None or cached.device != self.atom_types.device:
            atom_type = self.element_map.element_to_atom_type[element]
            cached = torch.nonzero(self.atom_types == atom_type, as_tuple=True)[0]
            self._select_cache[element] = cached
        return cached

    # ------------------------------------------------------------------ conversion
    def to_dict(self) -> Dict[str, Any]:
        data: Dict[str, Any] = {a: getattr(self, a).detach().cpu().numpy() for a in _ATOM_ATTRIBUTES}
        data["lattice"] = self.box.lattice.detach().cpu().numpy() if self.box else []
        data["elements"] = [self.element_map.get_element_from_atom_type(int(n)) for n in data["atom_types"]]
        return data

    def to_ase(self) -> Any:
        from ase import Atoms as AseAtoms  # optional dependency

        to_element = self.element_map.atom_type_to_element
        cell = units.TO_ANGSTROM * self.box.lattice.detach().cpu().numpy() if self.box is not None else None
        return AseAtoms(
            symbols=[to_element[int(t)] for t in self.atom_types.detach().cpu().tolist()],
            positions=units.TO_ANGSTROM * self.positions.detach().cpu().numpy(),
            cell=cell,
            pbc=bool(self.box),
            charges=self.charges.detach().cpu().numpy(),
        )

    # ------------------------------------------------------------------ energy offsets
    def _get_energy_offset(self, atom_energy: Dict[Element, float]) -> Array:
        offset = torch.empty_like(self.energies)
        for element in self.get_unique_elements():
            offset[self.select(element)] = atom_energy[element]
        return offset

    def remove_energy_offset(self, atom_energy: Dict[Element, float]) -> None:
        offset = self._get_energy_offset(atom_energy)
        self.energies = self.energies - offset
        self.total_energy = self.total_energy - offset.sum()

    def add_energy_offset(self, atom_energy: Dict[Element, float]) -> None:
        offset = self._get_energy_offset(atom_energy)
        self.energies = self.energies + offset
        self.total_energy = self.total_energy + offset.sum()

    # ------------------------------------------------------------------ kernel views
    def as_kernel_args(self) -> StructureAsKernelArgs:
        return StructureAsKernelArgs(
            self.positions, self.atom_types, self.lattice, self.total_energy,
            self.element_map.element_to_atom_type,
        )

    def _get_positions_per_element(self) -> Iterator[Tuple[Element, Array]]:
        for element in self.get_unique_elements():
            yield element, self.positions[self.select(element)]

    def get_positions_per_element(self) -> Dict[Element, Array]:
        return dict(self._get_positions_per_element())

    def get_forces_per_element(self) -> Dict[Element, Array]:
        return {el: self.forces[self.select(el)] for el in self.get_unique_elements()}

    def replace(self, **changes: Any) -> "Structure":
        """Shallow copy with some attributes replaced (re-wraps positions, like `dataclasses.replace`)."""
        kwargs = {a: getattr(self, a) for a in _ATOM_ATTRIBUTES}
        kwargs.update(element_map=self.element_map, box=self.box)
        kwargs.update(changes)
        return Structure(**kwargs)

    def __repr__(self) -> str:
        return (
            f"{self.__class__.__name__}(natoms={self.natoms}, "
            f"elements={self.get_unique_elements()}, dtype={self.dtype})"
        )
