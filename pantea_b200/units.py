"""Hartree atomic units and conversion factors (values: reference `pantea/units.py:64-71`)."""
from __future__ import annotations


class PhysicalUnits:
    def __init__(self, *, boltzmann: float, angstrom: float, pico_second: float, bar: float,
                 electron_volt: float, atomic_mass: float) -> None:
        self.BOLTZMANN_CONSTANT = boltzmann
        self.TO_ANGSTROM = angstrom
        self.TO_PICO_SECOND = pico_second
        self.TO_BAR = bar
        self.TO_ELECTRON_VOLT = electron_volt
        self.TO_ATOMIC_MASS = atomic_mass
        # derived conversions, same derivation order as the reference (`units.py:33-52`)
        self.TO_NANO_METER = self.TO_ANGSTROM * 0.1
        self.TO_FEMTO_SECOND = self.TO_PICO_SECOND * 1000
        self.TO_NANO_SECOND = self.TO_PICO_SECOND * 0.001
        self.TO_GIGA_PASCAL = self.TO_BAR * 0.0001
        self.TO_PASCAL = self.TO_BAR * 100000
        self.TO_ATMOSPHERE = self.TO_BAR * 0.986923
        self.TO_KILO_BAR = self.TO_BAR * 0.001
        self.TO_KCAL_PER_MOL = self.TO_ELECTRON_VOLT * 23.0609
        for name in ("ANGSTROM", "NANO_METER", "FEMTO_SECOND", "PICO_SECOND", "NANO_SECOND",
                     "GIGA_PASCAL", "KILO_BAR", "BAR", "ELECTRON_VOLT", "ATOMIC_MASS", "KCAL_PER_MOL"):
            setattr(self, f"FROM_{name}", 1 / getattr(self, f"TO_{name}"))


hartree_units = PhysicalUnits(
    boltzmann=3.166811563e-6,
    angstrom=5.29177249e-01,
    pico_second=2.418884326e-05,
    bar=2.942102648e08,
    electron_volt=2.7211407953e01,
    atomic_mass=5.48579957163e-4,
)
units = hartree_units
