"""Descriptor scaler (reference `pantea/descriptors/scaler.py:14-299`).

The four transforms are affine per feature; inside the fused energy/force kernel they are
applied as `offset + slope * (G - shift)` (see `DescriptorScaler.affine`).  The standalone
`__call__`, `fit` and `partial_fit` are thin elementwise/reduction expressions on the resident
device arrays.  `scale_center_sigma` keeps the reference's sign, slope `(smin - smax)/sigma`
(`scaler.py:241-246`).
"""
from __future__ import annotations

import json
from pathlib import Path
from typing import Dict, NamedTuple, Tuple

import numpy as np
import torch

from pantea_b200.logger import logger
from pantea_b200.types import Array, asarray, default_dtype

SCALE_TYPES: Tuple[str, ...] = ("center", "scale", "scale_center", "scale_center_sigma")


class ScalerParams(NamedTuple):
    dimension: Array
    nsamples: Array
    mean: Array
    sigma: Array
    minval: Array
    maxval: Array


class ScalerWarnings(NamedTuple):
    number_of_warnings: int
    max_number_of_warnings: int


class ScaleRange(NamedTuple):
    min_value: float
    max_value: float


class DescriptorScaler:
    def __init__(self, scale_range: ScaleRange, scale_type: str) -> None:
        self.scale_range = scale_range
        self.scale_type = scale_type

    @classmethod
    def from_type(cls, scale_type: str, scale_min: float = 0.0, scale_max: float = 1.0) -> "DescriptorScaler":
        if not (scale_min < scale_max):
            logger.error("Unexpected scale range values", exception=ValueError)
        if scale_type not in SCALE_TYPES:
            raise KeyError(scale_type)
        return cls(ScaleRange(float(scale_min), float(scale_max)), scale_type)

    # ---------------------------------------------------------------- transform
    def affine(self, params: ScalerParams) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
        """Host float64 `(shift, slope, offset)` with `x_scaled = offset + slope * (x - shift)`."""
        smin, smax = self.scale_range
        mean = params.mean.detach().cpu().double().numpy()
        ones = np.ones_like(mean)
        if self.scale_type == "center":
            return mean, ones, 0.0 * ones
        lo = params.minval.detach().cpu().double().numpy()
        hi = params.maxval.detach().cpu().double().numpy()
        if self.scale_type == "scale":
            return lo, (smax - smin) / (hi - lo), smin * ones
        if self.scale_type == "scale_center":
            return mean, (smax - smin) / (hi - lo), smin * ones
        sigma = params.sigma.detach().cpu().double().numpy()
        return mean, (smin - smax) / sigma, smin * ones

    def __call__(self, params: ScalerParams, data: Array) -> Array:
        data = torch.atleast_2d(data)
        smin, smax = self.scale_range
        if self.scale_type == "center":
            return data - params.mean
        if self.scale_type == "scale":
            return smin + (smax - smin) * (data - params.minval) / (params.maxval - params.minval)
        if self.scale_type == "scale_center":
            return smin + (smax - smin) * (data - params.mean) / (params.maxval - params.minval)
        return smin + (smin - smax) * (data - params.mean) / params.sigma

    # ---------------------------------------------------------------- fitting (scaler.py:249-283)
    @classmethod
    def fit(cls, data: Array) -> ScalerParams:
        """Per-feature mean / sigma / min / max of a descriptor batch (`scaler.py:250-259`).  Device-resident batches
        go through the library's two-pass reduction (`pantea_scaler_stats`); host tensors (unit tests of the host
        logic, tiny inputs) use the same definitions in torch."""
        data = torch.atleast_2d(data)
        dimension = asarray(data.shape[1], dtype=default_dtype.INT)
        nsamples = asarray(data.shape[0], dtype=default_dtype.INT)
        if data.is_cuda and data.dtype in (torch.float32, torch.float64) and data.numel() > 0:
            from pantea_b200 import _lib
            data = data if data.stride(1) == 1 else data.contiguous()
            n, d = data.shape
            stats = torch.empty((4, d), dtype=torch.float64, device=data.device)
            import ctypes
            _lib.check(_lib.load().pantea_scaler_stats(ctypes.c_void_p(data.data_ptr()), n, d, data.stride(0), _lib.dtype_code(data.dtype),
                                                       _lib.ptr(stats), _lib.stream_ptr()))
            stats = stats.to(data.dtype)
            return ScalerParams(dimension, nsamples, stats[0], stats[1], stats[2], stats[3])
        return ScalerParams(
            dimension=dimension,
            nsamples=nsamples,
            mean=data.mean(dim=0),
            sigma=data.std(dim=0, unbiased=False),
            minval=data.min(dim=0).values,
            maxval=data.max(dim=0).values,
        )

    @classmethod
    def merge(cls, params: ScalerParams, new: ScalerParams) -> ScalerParams:
        """Statistics of the union of two batches from their separate statistics: the combination rule of
        `_partial_fit` (`scaler.py:262-283`), coefficients split exactly as there."""
        dtype = params.mean.dtype
        m, n = params.nsamples.to(dtype), new.nsamples.to(dtype)  # fractions in the data precision
        fm, fn = m / (m + n), n / (m + n)
        diff = params.mean - new.mean
        mean = fm * params.mean + fn * new.mean
        sigma = torch.sqrt((fm * params.sigma) * params.sigma + (fn * new.sigma) * new.sigma
                           + (fm * diff) * (fn * diff))
        return ScalerParams(params.dimension, params.nsamples + new.nsamples.to(params.nsamples.dtype), mean, sigma,
                            torch.minimum(params.minval, new.minval), torch.maximum(params.maxval, new.maxval))

    @classmethod
    def partial_fit(cls, params: ScalerParams, data: Array) -> ScalerParams:
        data = torch.atleast_2d(data)
        new = cls.fit(data)
        new = ScalerParams(new.dimension, new.nsamples.to(params.mean.device), *(t.to(params.mean.device) for t in new[2:]))
        return cls.merge(params, new)

    @classmethod
    def initialize_warnings(cls, number_of_warnings: int = 0, max_number_of_warnings: int = -1) -> ScalerWarnings:
        return ScalerWarnings(number_of_warnings, max_number_of_warnings)

    @classmethod
    def check_warnings(cls, params: ScalerParams, data: Array, warnings: ScalerWarnings) -> ScalerWarnings:
        if warnings.max_number_of_warnings < 0:
            return warnings
        out_of_range = bool(((data > params.maxval) | (params.minval > data)).any())
        new = ScalerWarnings(warnings.number_of_warnings + int(out_of_range), warnings.max_number_of_warnings)
        if new.number_of_warnings >= new.max_number_of_warnings:
            logger.warning(
                "Exceeding maximum number scaler extrapolation warnings: "
                f"{new.number_of_warnings} (max={new.max_number_of_warnings})"
            )
        return new

    # ---------------------------------------------------------------- persistence (scaler.py:145-178)
    @classmethod
    def save(cls, params: ScalerParams, filename: Path) -> None:
        with open(str(filename), "w") as file:
            json.dump({k: v.detach().cpu().tolist() for k, v in params._asdict().items()}, file, indent=4)

    @classmethod
    def load(cls, filename: Path, integer_type_keys: Tuple[str, ...] = ("dimension", "nsamples")) -> ScalerParams:
        with open(str(filename), "r") as file:
            raw: Dict = json.load(file)
        out = {}
        for key, value in raw.items():
            dtype = default_dtype.INT if key in integer_type_keys else default_dtype.FLOATX
            out[key] = asarray(np.asarray(value), dtype=dtype)
        return ScalerParams(**out)

    @classmethod
    def _check_dimension(cls, params: ScalerParams, data: Array) -> Array:
        data = torch.atleast_2d(data)
        if data.shape[1] != int(params.dimension):
            logger.error(
                f"Data dimension doesn't match: {data.shape[1]} (expected {int(params.dimension)})",
                exception=ValueError,
            )
        return data

    @property
    def scale_min(self) -> float:
        return float(self.scale_range.min_value)

    @property
    def scale_max(self) -> float:
        return float(self.scale_range.max_value)

    def __repr__(self) -> str:
        return (f"{self.__class__.__name__}(transform='_{self.scale_type}', "
                f"scale_range=({self.scale_min}, {self.scale_max}))")
