from pantea_b200.potentials.nnp.potential import NNP, NeuralNetworkPotential

__all__ = ["NeuralNetworkPotential", "NNP"]
