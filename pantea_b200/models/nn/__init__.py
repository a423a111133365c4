from pantea_b200.models.nn.initializer import UniformInitializer
from pantea_b200.models.nn.model import NeuralNetworkModel

__all__ = ["NeuralNetworkModel", "UniformInitializer"]
