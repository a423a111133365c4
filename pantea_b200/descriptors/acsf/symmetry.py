"""Symmetry-function parameter records (reference `pantea/descriptors/acsf/{symmetry,radial,angular}.py`).

Positional constructor orders are the reference's: `G2(cfn, r_shift, eta)`,
`G3(cfn, eta, zeta, lambda0, r_shift)`, `G9(cfn, eta, zeta, lambda0, r_shift)`.
`r_shift` is stored but ignored by G3/G9, exactly as in the reference (`angular.py:58-65,100-107`).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import NamedTuple, Optional

from pantea_b200.descriptors.acsf.cutoff import CutoffFunction
from pantea_b200.types import Element


class NeighborElements(NamedTuple):
    neighbor_j: Element
    neighbor_k: Optional[Element] = None


class BaseSymmetryFunction:
    cfn: CutoffFunction
    kind: int = 0

    @property
    def r_cutoff(self) -> float:
        return self.cfn.r_cutoff


class RadialSymmetryFunction(BaseSymmetryFunction):
    pass


class AngularSymmetryFunction(BaseSymmetryFunction):
    pass


@dataclass(frozen=True)
class G1(RadialSymmetryFunction):
    cfn: CutoffFunction
    kind = 1


@dataclass(frozen=True)
class G2(RadialSymmetryFunction):
    cfn: CutoffFunction
    r_shift: float
    eta: float
    kind = 2


@dataclass(frozen=True)
class G3(AngularSymmetryFunction):
    cfn: CutoffFunction
    eta: float
    zeta: float
    lambda0: float
    r_shift: float
    kind = 3


@dataclass(frozen=True)
class G9(AngularSymmetryFunction):
    cfn: CutoffFunction
    eta: float
    zeta: float
    lambda0: float
    r_shift: float
    kind = 9
