"""Cutoff functions of the ACSF descriptor (reference `pantea/descriptors/acsf/cutoff.py:45-110`).

Host-side description only: `fc(r) = [r < rc] * f(r)` with the seven RuNNer types is evaluated
inside the CUDA kernels (value and derivative).  Integer codes follow the RuNNer `cutoff_type`
numbering (`settings.py:43-51`).  `poly1`/`poly2` act on the raw distance, as in the reference.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import torch

CUTOFF_CODES = {"hard": 0, "cos": 1, "tanhu": 2, "tanh": 3, "exp": 4, "poly1": 5, "poly2": 6}


_TANH_PRE: float = ((math.e + 1 / math.e) / (math.e - 1 / math.e)) ** 3


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

    def __call__(self, r):
        """`fc(r) = where(r < r_cutoff, f(r), 0)` on a tensor (reference `cutoff.py:64-110`).  Plain torch expression
        for inspection and plotting; the kernels evaluate the same formulas themselves."""
        r = torch.as_tensor(r)
        rc = self.r_cutoff
        kind = self.cutoff_type
        if kind == "hard":
            f = torch.ones_like(r)
        elif kind == "tanhu":
            f = torch.tanh(1.0 - r / rc) ** 3
        elif kind == "tanh":
            f = _TANH_PRE * torch.tanh(1.0 - r / rc) ** 3
        elif kind == "cos":
            f = 0.5 * (torch.cos(math.pi * r / rc) + 1.0)
        elif kind == "exp":
            f = torch.exp(1.0 - 1.0 / (1.0 - (r / rc) ** 2))
        elif kind == "poly1":  # on the raw distance, as the reference
            f = (2.0 * r - 3.0) * r**2 + 1.0
        else:  # poly2
            f = ((15.0 - 6.0 * r) * r - 10) * r**3 + 1.0
        return torch.where(r < rc, f, torch.zeros_like(r))

    def __repr__(self) -> str:
        return f"{self.__class__.__name__}(r_cutoff={self.r_cutoff})"
