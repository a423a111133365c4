"""Per-element chain descriptor -> scaler -> network (reference `pantea/potentials/nnp/atomic_potential.py:12-60`)."""
from __future__ import annotations

from dataclasses import dataclass

from pantea_b200.descriptors.acsf.acsf import ACSF
from pantea_b200.descriptors.scaler import DescriptorScaler
from pantea_b200.models.nn.model import NeuralNetworkModel


@dataclass(frozen=True)
class AtomicPotential:
    descriptor: ACSF
    scaler: DescriptorScaler
    model: NeuralNetworkModel

    @property
    def model_input_size(self) -> int:
        return self.descriptor.num_symmetry_functions

    def __repr__(self) -> str:
        return (f"{self.__class__.__name__}(\n  descriptor={self.descriptor},\n  scaler={self.scaler},"
                f"\n  model={self.model},\n)")
