"""Global array conventions (mirrors reference `pantea/types.py:13-30`).

Arrays are torch tensors living on the B200 (`cuda:LOCAL_RANK`) when a GPU is
visible, otherwise on the host (host arrays only serve the CPU-side container
logic; every compute entry point requires the CUDA library and a GPU).
`default_dtype.FLOATX` is the FP64/FP32 mode switch exactly as in the reference.
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Any

import torch

Array = torch.Tensor
Dtype = torch.dtype
Element = str
Scalar = torch.Tensor


@dataclass
class DataType:
    FLOATX: Dtype = torch.float64
    INT: Dtype = torch.int32
    UINT: Dtype = torch.int32  # torch has no general uint32 arithmetic; int32 is used
    INDEX: Dtype = torch.int32


default_dtype = DataType()

_NP_ALIASES = {"float64": torch.float64, "float32": torch.float32, "int32": torch.int32, "int64": torch.int64}


def as_torch_dtype(dtype: Any) -> torch.dtype:
    """Accept torch dtypes, numpy dtypes / scalar types and strings ('float32')."""
    if dtype is None:
        return default_dtype.FLOATX
    if isinstance(dtype, torch.dtype):
        return dtype
    name = getattr(dtype, "__name__", None) or getattr(dtype, "name", None) or str(dtype)
    name = name.replace("torch.", "")
    if name in _NP_ALIASES:
        return _NP_ALIASES[name]
    raise TypeError(f"Unsupported dtype {dtype!r}")


def device() -> torch.device:
    """Device on which structure arrays are resident (one process per GPU)."""
    if torch.cuda.is_available():
        return torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")) % max(torch.cuda.device_count(), 1))
    return torch.device("cpu")


def asarray(data: Any, dtype: Any = None) -> torch.Tensor:
    dt = as_torch_dtype(dtype) if dtype is not None else None
    if isinstance(data, torch.Tensor):
        return data.to(device=device(), dtype=dt or data.dtype)
    return torch.as_tensor(data, dtype=dt, device=device())
