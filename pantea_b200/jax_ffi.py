"""Optional JAX binding: registers the XLA FFI custom-call targets of `libpantea_b200_ffi.so` (csrc/xla_ffi_shim.cc).

The north star asks for a JAX FFI custom call so that `jax.Array`s pass zero-copy on XLA's own stream.  jax / jaxlib are
not installed in the image this round was built in, so this module is NOT exercised by the test-suite beyond its error
path: importing it without jax raises ImportError with the reason, and `csrc/build.py --ffi` refuses to build the shim
without jaxlib's headers.  The tested binding is ctypes + torch (`_lib.py`, `engine.py`); both wrap the same extern "C"
symbols.  Usage, once jax is available:

    from pantea_b200 import jax_ffi
    jax_ffi.register()
    forces, e_atom, e_total = jax_ffi.energy_forces(workspace_handle, positions, types, box_diag, r_cutoff)
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

try:
    import jax
    import jax.numpy as jnp
    from jax import ffi as _ffi
except ImportError as exc:  # pragma: no cover - depends on the environment
    raise ImportError("pantea_b200.jax_ffi needs jax with jax.ffi (jax >= 0.4.38); the tested binding of this package is "
                      "ctypes + torch (pantea_b200._lib)") from exc

SHIM_PATH = Path(__file__).resolve().parent / "libpantea_b200_ffi.so"
_registered = False


def register() -> None:
    """Register `pantea_energy_forces` and `pantea_acsf` as CUDA FFI targets (idempotent)."""
    global _registered
    if _registered:
        return
    if not SHIM_PATH.exists():
        raise RuntimeError(f"{SHIM_PATH} is missing: build it with `python -m pantea_b200.csrc.build --ffi`")
    lib = C.CDLL(str(SHIM_PATH))
    _ffi.register_ffi_target("pantea_energy_forces", _ffi.pycapsule(lib.PanteaEnergyForces), platform="CUDA")
    _ffi.register_ffi_target("pantea_acsf", _ffi.pycapsule(lib.PanteaAcsf), platform="CUDA")
    _registered = True


def _box_attrs(box):
    if box is None:
        return dict(has_box=0, lx=0.0, ly=0.0, lz=0.0)
    return dict(has_box=1, lx=float(box[0]), ly=float(box[1]), lz=float(box[2]))


def energy_forces(workspace: int, positions, types, box, r_cutoff: float, force_mode: int = 0):
    """(forces [n,3], e_atom [n], e_total [1]) of the structure; replaces `_jitted_grad_compute_energy` / `_jitted_compute_energy`."""
    n = positions.shape[0]
    out = (jax.ShapeDtypeStruct((n, 3), positions.dtype), jax.ShapeDtypeStruct((n,), positions.dtype),
           jax.ShapeDtypeStruct((1,), positions.dtype))
    return _ffi.ffi_call("pantea_energy_forces", out)(positions, jnp.asarray(types, dtype=jnp.int32), workspace=int(workspace),
                                                      r_cutoff=float(r_cutoff), force_mode=int(force_mode), **_box_attrs(box))


def acsf(workspace: int, element: int, n_symfunc: int, positions, types, centres, box, r_cutoff: float):
    """(G [n_c, n_sf], dG [n_c, n_sf, 3]); replaces `_jitted_calculate_acsf_descriptor` and its gradient."""
    n_c = centres.shape[0]
    out = (jax.ShapeDtypeStruct((n_c, n_symfunc), positions.dtype), jax.ShapeDtypeStruct((n_c, n_symfunc, 3), positions.dtype))
    return _ffi.ffi_call("pantea_acsf", out)(positions, jnp.asarray(types, dtype=jnp.int32), jnp.asarray(centres, dtype=jnp.int32),
                                             workspace=int(workspace), element=int(element), r_cutoff=float(r_cutoff),
                                             **_box_attrs(box))
