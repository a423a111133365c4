"""ctypes binding of `libpantea_b200.so` (the C ABI declared in `include/pantea_b200.h`).

Device arrays are torch CUDA tensors; their `data_ptr()` goes straight into the C ABI (zero copy)
and every call is queued on torch's current CUDA stream.  There is no CPU fallback: if the
library is missing, or no CUDA device is visible, the compute entry points raise.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path
from typing import Optional

import torch

_PKG = Path(__file__).resolve().parent
# PANTEA_B200_LIB selects another build of the same library (kernel-variant experiments, tools/build_variant.py)
LIB_PATH = Path(os.environ["PANTEA_B200_LIB"]) if os.environ.get("PANTEA_B200_LIB") else _PKG / "libpantea_b200.so"

PANTEA_OK, PANTEA_EINVAL, PANTEA_ECUDA, PANTEA_ECAPACITY, PANTEA_ENOMEM = 0, -1, -2, -3, -4
PANTEA_F64, PANTEA_F32 = 64, 32
FORCE_REFERENCE, FORCE_FULL = 0, 1
MAX_TYPES, MAX_SYMFUNC, MAX_LAYERS, MAX_CUTOFFS = 8, 128, 8, 4


class CapacityError(RuntimeError):
    """A neighbour row overflowed the workspace capacity (`PANTEA_ECAPACITY`)."""


class SymFuncDesc(C.Structure):
    _fields_ = [("kind", C.c_int32), ("cutoff_type", C.c_int32), ("type_j", C.c_int32), ("type_k", C.c_int32),
                ("r_cutoff", C.c_double), ("eta", C.c_double), ("r_shift", C.c_double), ("lambda0", C.c_double),
                ("zeta", C.c_double)]


class ElementDesc(C.Structure):
    _fields_ = [("n_symfunc", C.c_int32), ("symfunc", C.POINTER(SymFuncDesc)),
                ("scale_shift", C.POINTER(C.c_double)), ("scale_slope", C.POINTER(C.c_double)),
                ("scale_offset", C.POINTER(C.c_double)),
                ("n_layers", C.c_int32), ("layer_sizes", C.POINTER(C.c_int32)), ("activations", C.POINTER(C.c_int32)),
                ("weights", C.POINTER(C.c_double))]


class PotentialDesc(C.Structure):
    _fields_ = [("n_elements", C.c_int32), ("elements", C.POINTER(ElementDesc))]


class MDParams(C.Structure):
    _fields_ = [("dt", C.c_double), ("t_target", C.c_double), ("tau", C.c_double), ("kb", C.c_double),
                ("record", C.c_int32), ("use_graph", C.c_int32), ("mass_scaled", C.c_int32), ("force_mode", C.c_int32)]


# name -> (restype, argtypes); must list every symbol declared in include/pantea_b200.h
_VP, _I32, _I64, _DBL = C.c_void_p, C.c_int32, C.c_int64, C.c_double
PROTOTYPES = {
    "pantea_last_error": (C.c_char_p, []),
    "pantea_version": (C.c_char_p, []),
    "pantea_device_count": (C.c_int, []),
    "pantea_potential_create": (C.c_int, [C.POINTER(PotentialDesc), C.POINTER(_VP)]),
    "pantea_potential_destroy": (C.c_int, [_VP]),
    "pantea_potential_cutoff": (_DBL, [_VP]),
    "pantea_workspace_create": (C.c_int, [_VP, _I64, _I32, _I32, C.POINTER(_VP)]),
    "pantea_workspace_destroy": (C.c_int, [_VP]),
    "pantea_neighbor_build": (C.c_int, [_VP, _VP, _VP, _I64, C.POINTER(_DBL), _DBL, _VP]),
    "pantea_neighbor_build_batch": (C.c_int, [_VP, _VP, _VP, _I64, _VP, _VP, _I64, _DBL, _VP]),
    "pantea_workspace_set_owned_range": (C.c_int, [_VP, _I64, _I64]),
    "pantea_neighbor_status": (C.c_int, [_VP, C.POINTER(_I32), _VP]),
    "pantea_neighbor_counts": (C.c_int, [_VP, _VP, _VP]),
    "pantea_neighbor_export": (C.c_int, [_VP, _VP, _VP, _VP]),
    "pantea_distances": (C.c_int, [_VP, _VP, _I64, _VP, _I64, _VP, _VP, _VP]),
    "pantea_acsf_compute": (C.c_int, [_VP, _I32, _VP, _I64, _VP, _VP, _VP]),
    "pantea_energy_forces": (C.c_int, [_VP, _VP, _VP, _VP, _I32, _VP]),
    "pantea_md_update_positions": (C.c_int, [_VP, _VP, _VP, _I64, _I64, C.POINTER(_DBL), _DBL, _I32, _VP]),
    "pantea_md_update_velocities": (C.c_int, [_VP, _VP, _VP, _I64, _I64, _DBL, _I32, _VP]),
    "pantea_md_update_positions_mass": (C.c_int, [_VP, _VP, _VP, _VP, _I64, _I64, C.POINTER(_DBL), _DBL, _I32, _VP]),
    "pantea_md_update_velocities_mass": (C.c_int, [_VP, _VP, _VP, _VP, _I64, _I64, _DBL, _I32, _VP]),
    "pantea_md_kinetic_energy": (C.c_int, [_VP, _VP, _I64, _I64, _VP, _I32, _VP]),
    "pantea_md_rescale_velocities": (C.c_int, [_VP, _I64, _I64, _VP, _I64, _DBL, _DBL, _DBL, _DBL, _I32, _VP]),
    "pantea_md_run": (C.c_int, [_VP, _VP, _VP, _VP, _VP, _VP, _I64, C.POINTER(_DBL), _I64, C.POINTER(MDParams), _VP, _VP]),
    "pantea_lj_energy_forces": (C.c_int, [_VP, _DBL, _DBL, _VP, _VP, _VP, _VP]),
    "pantea_launch_count": (_I64, []),
    "pantea_bench_fma": (C.c_int, [_I32, _I32, _I32, _I32, _VP, C.POINTER(_DBL), _VP]),
    "pantea_l2_flush": (C.c_int, [_VP, _I64, _VP]),
    "pantea_workspace_set_counters": (C.c_int, [_VP, _VP]),
    "pantea_eval_timing": (C.c_int, [_I32, C.POINTER(C.c_float)]),
    "pantea_workspace_set_compute_precision": (C.c_int, [_VP, _I32]),
    "pantea_set_fast_path": (C.c_int, [_I32]),
    "pantea_set_gauss_screen": (_DBL, [_DBL]),
    "pantea_workspace_set_skin": (C.c_int, [_VP, _DBL]),
    "pantea_neighbor_rebuilds": (C.c_int, [_VP, C.POINTER(_I64), _VP]),
    "pantea_scaler_stats": (C.c_int, [_VP, _I64, _I64, _I64, _I32, _VP, _VP]),
    "pantea_halo_pack": (C.c_int, [_VP, _VP, _I64, _VP, _VP, _I64, C.POINTER(_DBL), _DBL, _VP, _I32, _VP]),
    "pantea_halo_unpack_add": (C.c_int, [_VP, _VP, _VP, _VP, _I64, _I32, _VP]),
    "pantea_mgpu_create": (C.c_int, [_VP, _I32, _I32, _I64, C.POINTER(_DBL), C.POINTER(_I32), C.POINTER(_DBL),
                                     C.POINTER(_DBL), C.POINTER(_DBL), _DBL, _DBL, _I64, C.POINTER(_VP)]),
    "pantea_mgpu_handle_bytes": (_I64, []),
    "pantea_mgpu_export_handle": (C.c_int, [_VP, _VP]),
    "pantea_mgpu_connect": (C.c_int, [_VP, _VP]),
    "pantea_mgpu_set_state": (C.c_int, [_VP, _VP, _VP, _VP, _VP]),
    "pantea_mgpu_run": (C.c_int, [_VP, _I64, _I32, _VP]),
    "pantea_mgpu_read": (C.c_int, [_VP, _VP, _VP, _VP, _VP, C.POINTER(_I32), _VP]),
    "pantea_mgpu_destroy": (C.c_int, [_VP]),
}

_lib: Optional[C.CDLL] = None


def load() -> C.CDLL:
    """Load the shared library (never builds or falls back: a missing library is an error)."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m pantea_b200.csrc.build` "
                "(pantea_b200 has no CPU fallback)")
        lib = C.CDLL(str(LIB_PATH))
        for name, (restype, argtypes) in PROTOTYPES.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = restype, argtypes
        _lib = lib
    return _lib


def require_cuda() -> None:
    if not torch.cuda.is_available() or load().pantea_device_count() < 1:
        raise RuntimeError("pantea_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")


def last_error() -> str:
    return load().pantea_last_error().decode()


def check(code: int) -> None:
    if code == PANTEA_OK:
        return
    msg = last_error()
    if code == PANTEA_ECAPACITY:
        raise CapacityError(msg)
    if code == PANTEA_EINVAL:
        raise ValueError(msg)
    if code == PANTEA_ENOMEM:
        raise MemoryError(msg)
    raise RuntimeError(msg)


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("pantea_b200 kernels need CUDA tensors (no CPU fallback)")
    if not t.is_contiguous():
        raise ValueError("tensor must be contiguous")
    return t.data_ptr()


def dtype_code(dtype: torch.dtype) -> int:
    if dtype == torch.float64:
        return PANTEA_F64
    if dtype == torch.float32:
        return PANTEA_F32
    raise TypeError(f"unsupported floating type {dtype}")


def box_arg(box) -> Optional[C.Array]:
    """Host double[3] lattice diagonal (or None)."""
    if box is None:
        return None
    return (C.c_double * 3)(float(box[0]), float(box[1]), float(box[2]))
