from pantea_b200.potentials.nnp.atomic_potential import AtomicPotential
from pantea_b200.potentials.nnp.potential import NNP, NeuralNetworkPotential
from pantea_b200.potentials.nnp.settings import NeuralNetworkPotentialSettings
from pantea_b200.potentials.nnp.trainer import NeuralNetworkPotentialTrainer

__all__ = ["AtomicPotential", "NeuralNetworkPotential", "NNP", "NeuralNetworkPotentialSettings",
           "NeuralNetworkPotentialTrainer"]
