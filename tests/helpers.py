"""Shared helpers for the parity tests: build product-side objects from oracle specs."""
from __future__ import annotations

from typing import Dict, List, Sequence

import numpy as np
import torch

from oracle.spec import ACT_CODES, CUTOFF_CODES, ElementSpec

TYPE_TO_ELEMENT: Dict[int, str] = {1: "H", 2: "O"}


def device_potential_from_specs(specs: Sequence[ElementSpec], type_to_element=TYPE_TO_ELEMENT):
    """oracle ElementSpec list -> pantea_b200.engine.DevicePotential (through the C ABI)."""
    from pantea_b200 import engine

    records = []
    for spec in specs:
        sfs = [engine.SymFuncRecord(s.kind, CUTOFF_CODES[s.cutoff_type], s.r_cutoff, type_to_element[s.type_j],
                                    type_to_element[s.type_k] if s.type_k else None, s.eta, s.r_shift, s.lambda0, s.zeta)
               for s in spec.symfuncs]
        affine = spec.affine() if spec.scale_type is not None else None
        sizes: List[int] = []
        acts: List[int] = []
        weights = None
        if spec.layers:
            sizes = [len(spec.symfuncs)] + [k.shape[1] for k, _, _ in spec.layers]
            acts = [ACT_CODES[a] for _, _, a in spec.layers]
            weights = np.concatenate([np.concatenate([np.asarray(k, dtype=np.float64).ravel(),
                                                      np.asarray(b, dtype=np.float64).ravel()]) for k, b, _ in spec.layers])
        records.append(engine.ElementRecord(type_to_element[spec.atom_type], sfs, affine, sizes, acts, weights))
    return engine.DevicePotential(records, elements=sorted(set(type_to_element.values()),
                                                          key=lambda e: {"H": 1, "O": 8, "Ne": 10}[e]))


def cuda(a, dtype=torch.float64):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=dtype, device="cuda")


def csr_rows(row_ptr, col):
    row_ptr = np.asarray(row_ptr)
    col = np.asarray(col)
    return [col[row_ptr[i]:row_ptr[i + 1]] for i in range(len(row_ptr) - 1)]


def rel_err(a, b, floor=0.0):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    scale = max(np.abs(b).max(), floor, 1e-300)
    return float(np.abs(a - b).max() / scale)
