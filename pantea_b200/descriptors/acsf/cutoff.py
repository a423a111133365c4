"""Cutoff functions of the ACSF descriptor (reference `pantea/descriptors/acsf/cutoff.py:45-110`).

Host-side description only: `fc(r) = [r < rc] * f(r)` with the seven RuNNer types is evaluated
inside the CUDA kernels (value and derivative).  Integer codes follow the RuNNer `cutoff_type`
numbering (`settings.py:43-51`).  `poly1`/`poly2` act on the raw distance, as in the reference.
"""
from __future__ import annotations

from dataclasses import dataclass

CUTOFF_CODES = {"hard": 0, "cos": 1, "tanhu": 2, "tanh": 3, "exp": 4, "poly1": 5, "poly2": 6}


@dataclass(frozen=True)
class CutoffFunction:
    r_cutoff: float
    cutoff_type: str = "tanh"

    @classmethod
    def from_type(cls, cutoff_type: str, r_cutoff: float) -> "CutoffFunction":
        if cutoff_type not in CUTOFF_CODES:
            raise KeyError(cutoff_type)
        return cls(float(r_cutoff), cutoff_type)

    @property
    def code(self) -> int:
        return CUTOFF_CODES[self.cutoff_type]

    def __repr__(self) -> str:
        return f"{self.__class__.__name__}(r_cutoff={self.r_cutoff})"
