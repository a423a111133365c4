"""Symmetry-function parameter records (reference `pantea/descriptors/acsf/{symmetry,radial,angular}.py`).

Positional constructor orders are the reference's: `G2(cfn, r_shift, eta)`,
`G3(cfn, eta, zeta, lambda0, r_shift)`, `G9(cfn, eta, zeta, lambda0, r_shift)`.
`r_shift` is stored but ignored by G3/G9, exactly as in the reference (`angular.py:58-65,100-107`).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import NamedTuple, Optional

import torch

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

    def __call__(self, rij):
        """Reference `radial.py:39-41`."""
        return self.cfn(rij)


@dataclass(frozen=True)
class G2(RadialSymmetryFunction):
    cfn: CutoffFunction
    r_shift: float
    eta: float
    kind = 2

    def __call__(self, rij):
        """Reference `radial.py:59-61`."""
        rij = torch.as_tensor(rij)
        return torch.exp(-self.eta * (rij - self.r_shift) ** 2) * self.cfn(rij)


@dataclass(frozen=True)
class G3(AngularSymmetryFunction):
    cfn: CutoffFunction
    eta: float
    zeta: float
    lambda0: float
    r_shift: float
    kind = 3

    def __call__(self, rij, rik, rjk, cost):
        """Reference `angular.py:51-66` (`r_shift` is ignored there too)."""
        rij, rik, rjk, cost = (torch.as_tensor(t) for t in (rij, rik, rjk, cost))
        return (2.0 ** (1.0 - self.zeta) * torch.pow(1 + self.lambda0 * cost, self.zeta)
                * torch.exp(-self.eta * (rij**2 + rik**2 + rjk**2)) * self.cfn(rij) * self.cfn(rik) * self.cfn(rjk))


@dataclass(frozen=True)
class G9(AngularSymmetryFunction):
    cfn: CutoffFunction
    eta: float
    zeta: float
    lambda0: float
    r_shift: float
    kind = 9

    def __call__(self, rij, rik, rjk, cost):
        """Reference `angular.py:92-107`: no r_jk terms, `r_shift` ignored."""
        rij, rik, cost = (torch.as_tensor(t) for t in (rij, rik, cost))
        return (2.0 ** (1.0 - self.zeta) * torch.pow(1 + self.lambda0 * cost, self.zeta)
                * torch.exp(-self.eta * (rij**2 + rik**2)) * self.cfn(rij) * self.cfn(rik))
