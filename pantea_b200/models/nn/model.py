"""Per-element feed-forward network description and weight files
(reference `pantea/models/nn/model.py:16-81`).

Parameters keep the reference's tree layout `{"layers_{2l}": {"kernel": [in,out], "bias": [out]}}`
(`tests/test_nn.py:97-138`).  Weight pickles written by the reference contain `jax.Array`
objects; `load` reads them without JAX through a restricted unpickler that rebuilds the
embedded numpy arrays.  Evaluation (forward + input gradient) happens inside the fused CUDA
energy/force kernel.
"""
from __future__ import annotations

import io
import pickle
from pathlib import Path
from typing import Any, Dict, List, Tuple

import numpy as np

from pantea_b200.models.nn.activation import ACTIVATION_CODES
from pantea_b200.types import Dtype, default_dtype

ModelParams = Dict[str, Dict[str, np.ndarray]]


def _rebuild_jax_array(fun: Any, args: Any, arr_state: Any, aval_state: Any = None) -> np.ndarray:
    """Stand-in for `jax._src.array._reconstruct_array`: rebuild the wrapped numpy array."""
    arr = fun(*args)
    arr.__setstate__(arr_state)
    return np.asarray(arr)


class _WeightsUnpickler(pickle.Unpickler):
    def find_class(self, module: str, name: str) -> Any:
        if module == "jax._src.array" and name == "_reconstruct_array":
            return _rebuild_jax_array
        if module in ("numpy.core.multiarray", "numpy._core.multiarray") and name in ("_reconstruct", "scalar"):
            import numpy._core.multiarray as ma
            return getattr(ma, name)
        if module == "numpy" and name in ("ndarray", "dtype"):
            return getattr(np, name)
        if module in ("builtins", "collections") and name in ("dict", "OrderedDict", "list", "tuple"):
            return super().find_class(module, name)
        if module in ("frozendict", "frozendict.core", "flax.core.frozen_dict") and name in ("frozendict", "FrozenDict"):
            return dict
        raise pickle.UnpicklingError(f"refusing to load {module}.{name} from a weights file")


def _to_numpy_tree(tree: Any) -> Any:
    if isinstance(tree, dict):
        return {str(k): _to_numpy_tree(v) for k, v in tree.items()}
    return np.asarray(tree)


class NeuralNetworkModel:
    def __init__(
        self,
        hidden_layers: Tuple[Tuple[int, str], ...],
        output_layer: Tuple[int, str] = (1, "identity"),
        params_dtype: Dtype = None,
        kernel_initializer: Any = None,
    ) -> None:
        self.hidden_layers = tuple((int(n), str(a)) for n, a in hidden_layers)
        self.output_layer = (int(output_layer[0]), str(output_layer[1]))
        self.params_dtype = params_dtype if params_dtype is not None else default_dtype.FLOATX
        self.kernel_initializer = kernel_initializer
        for _, act in (*self.hidden_layers, self.output_layer):
            if act not in ACTIVATION_CODES:
                raise KeyError(act)

    @property
    def layer_spec(self) -> List[Tuple[int, str]]:
        return [*self.hidden_layers, self.output_layer]

    def param_shapes(self, input_size: int) -> Dict[str, Dict[str, Tuple[int, ...]]]:
        shapes, n_in = {}, int(input_size)
        for l, (n_out, _) in enumerate(self.layer_spec):
            shapes[f"layers_{2 * l}"] = {"kernel": (n_in, n_out), "bias": (n_out,)}
            n_in = n_out
        return shapes

    def init_params(self, input_size: int, seed: int = 0, weights_range: Tuple[float, float] = (-1.0, 1.0)) -> ModelParams:
        """Uniform kernels / zero biases.  (The reference draws from the JAX PRNG, `potential.py:162-186`;
        that stream cannot be reproduced without JAX -- load real weights for parity work.)"""
        rng = np.random.default_rng(seed)
        lo, hi = weights_range
        return {
            name: {"kernel": rng.uniform(lo, hi, size=s["kernel"]), "bias": np.zeros(s["bias"])}
            for name, s in self.param_shapes(input_size).items()
        }

    def flatten(self, params: ModelParams, input_size: int) -> Tuple[List[int], List[int], np.ndarray]:
        """-> (layer sizes [L+1], activation codes [L], packed float64 weights: per layer kernel then bias)."""
        sizes, acts, chunks = [int(input_size)], [], []
        for l, (n_out, act) in enumerate(self.layer_spec):
            layer = params[f"layers_{2 * l}"]
            kernel = np.asarray(layer["kernel"], dtype=np.float64)
            bias = np.asarray(layer["bias"], dtype=np.float64)
            if kernel.shape != (sizes[-1], n_out) or bias.shape != (n_out,):
                raise ValueError(
                    f"layers_{2 * l}: kernel {kernel.shape} / bias {bias.shape} do not match ({sizes[-1]}, {n_out})")
            chunks += [kernel.ravel(), bias.ravel()]
            sizes.append(n_out)
            acts.append(ACTIVATION_CODES[act])
        return sizes, acts, np.concatenate(chunks)

    def save(self, filename: Path, params: ModelParams) -> None:
        with open(str(Path(filename)), "wb") as handle:
            pickle.dump(_to_numpy_tree(params), handle)

    def load(self, filename: Path) -> ModelParams:
        with open(str(Path(filename)), "rb") as handle:
            tree = _WeightsUnpickler(io.BytesIO(handle.read())).load()
        return _to_numpy_tree(tree)

    def __repr__(self) -> str:
        return f"{self.__class__.__name__}(hidden_layers={self.hidden_layers}, dtype={self.params_dtype})"
