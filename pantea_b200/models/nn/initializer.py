"""Uniform weight range holder (reference `pantea/models/nn/initializer.py:10-20`)."""
from typing import Tuple


class UniformInitializer:
    def __init__(self, weights_range: Tuple[float, float] = (-1.0, 1.0)) -> None:
        self.weights_range = (float(weights_range[0]), float(weights_range[1]))

    def __repr__(self) -> str:
        return f"{self.__class__.__name__}(weights_range={self.weights_range})"
