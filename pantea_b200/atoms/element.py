"""Element <-> atom-type <-> mass tables.

Semantics follow reference `pantea/atoms/element.py:73-147`: atom types are 1-based and
ordered by ascending atomic number; masses are tabulated in atomic mass units and
converted to Hartree atomic units (electron masses).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Sequence, Tuple

import torch

from pantea_b200.types import Array, Element, asarray
from pantea_b200.units import units

_SYMBOLS: Tuple[str, ...] = tuple(
    """H He Li Be B C N O F Ne Na Mg Al Si P S Cl Ar K Ca Sc Ti V Cr Mn Fe Co Ni Cu Zn
    Ga Ge As Se Br Kr Rb Sr Y Zr Nb Mo Tc Ru Rh Pd Ag Cd In Sn Sb Te I Xe Cs Ba La Ce
    Pr Nd Pm Sm Eu Gd Tb Dy Ho Er Tm Yb Lu Hf Ta W Re Os Ir Pt Au Hg Tl Pb Bi Po At Rn
    Fr Ra Ac Th Pa U Np Pu Am Cm Bk Cf Es Fm Md""".split()
)
_ATOMIC_NUMBER: Dict[str, int] = {s: z for z, s in enumerate(_SYMBOLS, start=1)}

# atomic masses [u] in order of atomic number (same tabulated values as the reference table,
# `element.py:29-60`, so that masses in Hartree units agree to the last digit)
_MASS_U: Tuple[float, ...] = (
    1.008, 4.003, 6.941, 9.012, 10.811, 12.011, 14.007, 15.999, 18.998, 20.180,
    22.990, 24.305, 26.982, 28.086, 30.974, 32.066, 35.453, 39.948, 39.098, 40.078,
    44.956, 47.867, 50.942, 51.996, 54.938, 55.845, 58.933, 58.693, 63.546, 65.38,
    69.723, 72.631, 74.922, 78.971, 79.904, 84.798, 84.468, 87.62, 88.906, 91.224,
    92.906, 95.95, 98.907, 101.07, 102.906, 106.42, 107.868, 112.414, 114.818, 118.711,
    121.760, 126.7, 126.904, 131.294, 132.905, 137.328, 138.905, 140.116, 140.908, 144.243,
    144.913, 150.36, 151.964, 157.25, 158.925, 162.500, 164.930, 167.259, 168.934, 173.055,
    174.967, 178.49, 180.948, 183.84, 186.207, 190.23, 192.217, 195.085, 196.967, 200.592,
    204.383, 207.2, 208.980, 208.982, 209.987, 222.081, 223.020, 226.025, 227.028, 232.038,
    231.036, 238.029, 237, 244, 243, 247, 247, 251, 252, 257, 258,
)
_MASS: Dict[str, float] = dict(zip(_SYMBOLS, _MASS_U))


@dataclass
class ElementMap:
    """Maps element names to integer atom types (and back) for array processing."""

    unique_elements: Tuple[Element, ...]
    element_to_atomic_number: Dict[Element, int]
    element_to_atom_type: Dict[Element, int]
    atom_type_to_element: Dict[int, Element]

    @classmethod
    def from_list(cls, elements: Sequence[Element]) -> "ElementMap":
        uniq = tuple(sorted(set(elements)))
        to_z = {el: _ATOMIC_NUMBER[el] for el in uniq}
        by_z = sorted(to_z, key=to_z.__getitem__)
        to_type = {el: t for t, el in enumerate(by_z, start=1)}
        return cls(uniq, to_z, to_type, {t: el for el, t in to_type.items()})

    def __getitem__(self, element: Element) -> int:
        return self.element_to_atom_type[element]

    def get_atom_type_from_element(self, name: Element) -> int:
        return self.element_to_atom_type[name]

    def get_element_from_atom_type(self, value: int) -> Element:
        return self.atom_type_to_element[int(value)]

    @classmethod
    def get_element_from_atomic_number(cls, value: int) -> Element:
        return _SYMBOLS[value - 1]

    @classmethod
    def get_atomic_number_from_element(cls, name: Element) -> int:
        return _ATOMIC_NUMBER[name]

    @classmethod
    def get_atomic_mass_from_element(cls, name: Element) -> float:
        return _MASS[name] * units.FROM_ATOMIC_MASS

    @classmethod
    def get_masses_from_structure(cls, structure) -> Array:
        to_element = structure.element_map.atom_type_to_element
        types_host = structure.atom_types.detach().cpu().tolist()
        masses = [cls.get_atomic_mass_from_element(to_element[t]) for t in types_host]
        return asarray(masses, dtype=torch.float64).to(structure.positions.dtype)

    def __repr__(self) -> str:
        return f"{self.__class__.__name__}(element_to_atom_type={self.element_to_atom_type})"
