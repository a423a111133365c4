"""Scaler fitting over a dataset (reference `pantea/potentials/nnp/trainer.py:68-88`, `fit_scaler`): the step that follows
descriptor preprocessing.  Descriptors come from the CUDA ACSF kernel, the per-feature statistics from
`pantea_scaler_stats`, batches are merged with the reference's `partial_fit` rule, and with several ranks the structures
are split `index mod world` and the statistics merged across ranks (`distributed.merge_scaler_params`).

Model training (`fit_model`, Kalman filter / gradient-descent updaters) is outside the hot path this package rebuilds."""
from __future__ import annotations

from typing import Dict, Optional, Sequence

from pantea_b200.descriptors.scaler import ScalerParams
from pantea_b200.logger import logger
from pantea_b200.potentials.nnp.potential import NeuralNetworkPotential
from pantea_b200.types import Element


class NeuralNetworkPotentialTrainer:
    def __init__(self, potential: NeuralNetworkPotential) -> None:
        self.potential = potential

    @classmethod
    def from_runner(cls, potential: NeuralNetworkPotential, filename: str = "input.nn") -> "NeuralNetworkPotentialTrainer":
        return cls(potential)

    def fit_scaler(self, dataset: Sequence, rank: int = 0, world: int = 1) -> Dict[Element, Optional[ScalerParams]]:
        """Fit the scaler parameters of every element over `dataset` (indexable, yields `Structure`).  Resets nothing:
        like the reference, existing parameters are continued with `partial_fit`."""
        pot = self.potential
        for index in range(rank, len(dataset), world):
            structure = dataset[index]
            for element in structure.get_unique_elements():
                if element not in pot.atomic_potentials:
                    continue
                x = pot.atomic_potentials[element].descriptor(structure)
                scaler = pot.atomic_potentials[element].scaler
                params = pot.scalers_params[element]
                pot.scalers_params[element] = scaler.fit(x) if params is None else scaler.partial_fit(params, x)
        if world > 1:
            from pantea_b200.distributed import merge_scaler_params
            for element in pot.elements:
                pot.scalers_params[element] = merge_scaler_params(pot.scalers_params[element])
        return pot.scalers_params

    def fit_model(self, dataset) -> None:
        logger.error("fit_model: training is outside the scope of pantea_b200 (energy/force hot path only)",
                     exception=NotImplementedError)
