"""Convert the reference's weight pickles (tests/weights.{001,008}.pkl, which embed jax.Array
objects) into plain .npz files so that the oracle can read them without JAX and without the
product package.  Run once; outputs are committed:  python tests/golden/convert_weights.py
"""
import io
import pickle
from pathlib import Path

import numpy as np

HERE = Path(__file__).parent


def _rebuild(fun, args, state, aval=None):
    arr = fun(*args)
    arr.__setstate__(state)
    return np.asarray(arr)


class Unpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if (module, name) == ("jax._src.array", "_reconstruct_array"):
            return _rebuild
        if module.startswith("numpy"):
            import numpy._core.multiarray as ma
            return getattr(ma, name) if hasattr(ma, name) and name == "_reconstruct" else getattr(np, name)
        return super().find_class(module, name)


for z in (1, 8):
    tree = Unpickler(io.BytesIO((HERE / f"weights.{z:03d}.pkl").read_bytes())).load()
    flat = {f"{layer}.{key}": np.asarray(val) for layer, d in tree.items() for key, val in d.items()}
    np.savez(HERE / f"weights.{z:03d}.npz", **flat)
    print(z, {k: (v.shape, str(v.dtype)) for k, v in flat.items()})
