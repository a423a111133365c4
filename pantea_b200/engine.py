"""Host-side driver of the CUDA library: device potentials, workspaces and structure binding.

This is the thin layer between pantea's Python API classes and the C ABI.  It owns
  * `DevicePotential`  -- symmetry-function tables, scaler affine maps and MLP weights on the GPU
                          (`pantea_potential_create`);
  * `Workspace`        -- neighbour-search scratch sized for a number of atoms and a neighbour-row
                          capacity, grown automatically when a row overflows;
  * structure binding  -- remapping of a structure's atom types onto the potential's element list
                          (reference semantics: `structure.element_map[element]`, acsf.py:183,198-199).
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from pantea_b200 import _lib
from pantea_b200.atoms.element import ElementMap
from pantea_b200.types import Element


@dataclass
class SymFuncRecord:
    kind: int
    cutoff_code: int
    r_cutoff: float
    neighbor_j: Element
    neighbor_k: Optional[Element] = None
    eta: float = 0.0
    r_shift: float = 0.0
    lambda0: float = 0.0
    zeta: float = 0.0


@dataclass
class ElementRecord:
    element: Element
    symfuncs: List[SymFuncRecord]
    affine: Optional[Tuple[np.ndarray, np.ndarray, np.ndarray]] = None  # (shift, slope, offset)
    layer_sizes: List[int] = field(default_factory=list)
    activations: List[int] = field(default_factory=list)
    weights: Optional[np.ndarray] = None


def _dptr(a: Optional[np.ndarray]):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


def _iptr(a: Optional[np.ndarray]):
    return a.ctypes.data_as(C.POINTER(C.c_int32)) if a is not None else None


class DevicePotential:
    """GPU-resident potential tables for an ordered element list (ascending atomic number)."""

    def __init__(self, records: Sequence[ElementRecord], elements: Optional[Sequence[Element]] = None) -> None:
        _lib.require_cuda()
        names = list(elements) if elements is not None else [r.element for r in records]
        for r in records:
            for sf in r.symfuncs:
                for el in (sf.neighbor_j, sf.neighbor_k):
                    if el is not None and el not in names:
                        names.append(el)
        self.elements: Tuple[Element, ...] = tuple(sorted(set(names), key=ElementMap.get_atomic_number_from_element))
        self.type_of: Dict[Element, int] = {el: t for t, el in enumerate(self.elements, start=1)}
        by_name = {r.element: r for r in records}
        keep: List[object] = []
        el_descs = (_lib.ElementDesc * len(self.elements))()
        self.n_symfunc: Dict[Element, int] = {}
        self.r_cutoff = 0.0
        for e, name in enumerate(self.elements):
            rec = by_name.get(name)
            d = el_descs[e]
            if rec is None:  # element only appears as a neighbour: empty descriptor, no network
                d.n_symfunc, d.n_layers = 0, 0
                self.n_symfunc[name] = 0
                continue
            sfs = (_lib.SymFuncDesc * max(len(rec.symfuncs), 1))()
            for s, sf in enumerate(rec.symfuncs):
                sfs[s] = _lib.SymFuncDesc(sf.kind, sf.cutoff_code, self.type_of[sf.neighbor_j],
                                          self.type_of[sf.neighbor_k] if sf.neighbor_k is not None else 0,
                                          sf.r_cutoff, sf.eta, sf.r_shift, sf.lambda0, sf.zeta)
                self.r_cutoff = max(self.r_cutoff, float(sf.r_cutoff))
            keep.append(sfs)
            d.n_symfunc, d.symfunc = len(rec.symfuncs), sfs
            self.n_symfunc[name] = len(rec.symfuncs)
            if rec.affine is not None:
                arrs = [np.ascontiguousarray(a, dtype=np.float64) for a in rec.affine]
                keep += arrs
                d.scale_shift, d.scale_slope, d.scale_offset = (_dptr(a) for a in arrs)
            if rec.layer_sizes:
                sizes = np.ascontiguousarray(rec.layer_sizes, dtype=np.int32)
                acts = np.ascontiguousarray(rec.activations, dtype=np.int32)
                weights = np.ascontiguousarray(rec.weights, dtype=np.float64)
                keep += [sizes, acts, weights]
                d.n_layers, d.layer_sizes, d.activations, d.weights = len(acts), _iptr(sizes), _iptr(acts), _dptr(weights)
        desc = _lib.PotentialDesc(len(self.elements), el_descs)
        handle = C.c_void_p()
        _lib.check(_lib.load().pantea_potential_create(C.byref(desc), C.byref(handle)))
        self.handle = handle
        self._workspaces: Dict[Tuple[int, int], "Workspace"] = {}
        del keep

    def slot(self, element: Element) -> int:
        return self.type_of[element] - 1

    def workspace(self, n_atoms: int, dtype: torch.dtype, density_hint: Optional[float] = None) -> "Workspace":
        """A cached workspace large enough for `n_atoms` atoms of this dtype."""
        code = _lib.dtype_code(dtype)
        best = None
        for (c, cap_atoms), ws in self._workspaces.items():
            if c == code and cap_atoms >= n_atoms and (best is None or cap_atoms < best.max_atoms):
                best = ws
        if best is None:
            max_atoms = max(64, int(n_atoms))
            best = Workspace(self, max_atoms, estimate_max_neighbors(self.r_cutoff, density_hint, n_atoms), dtype)
            self._workspaces[(code, max_atoms)] = best
        return best

    def __del__(self) -> None:
        try:
            for ws in self._workspaces.values():
                ws.close()
            if getattr(self, "handle", None):
                _lib.load().pantea_potential_destroy(self.handle)
                self.handle = None
        except Exception:
            pass


def estimate_max_neighbors(r_cutoff: float, density: Optional[float], n_atoms: int) -> int:
    """Neighbour-row capacity: 1.5x the mean count for the given number density + 32, capped by n-1."""
    if density is None or not math.isfinite(density) or density <= 0.0:
        return int(min(max(n_atoms - 1, 32), 512))
    mean = density * 4.0 / 3.0 * math.pi * r_cutoff**3
    cap = int(1.5 * mean) + 32
    return int(max(32, min(cap, max(n_atoms - 1, 32), 4096)))


class Workspace:
    def __init__(self, potential: Optional[DevicePotential], max_atoms: int, max_neighbors: int, dtype: torch.dtype) -> None:
        _lib.require_cuda()
        self.potential = potential
        self.max_atoms = int(max_atoms)
        self.max_neighbors = int(max_neighbors)
        self.dtype = dtype
        self.code = _lib.dtype_code(dtype)
        self.handle = C.c_void_p()
        self.skin = 0.0
        self._create()
        self.n_atoms = 0
        self._bound: Tuple = ()
        self._keep: Tuple = ()

    def _create(self) -> None:
        pot = self.potential.handle if self.potential is not None else None
        _lib.check(_lib.load().pantea_workspace_create(pot, self.max_atoms, self.max_neighbors, self.code, C.byref(self.handle)))
        if self.skin > 0.0:
            _lib.check(_lib.load().pantea_workspace_set_skin(self.handle, float(self.skin)))

    def set_skin(self, skin: float) -> None:
        """Verlet skin (Bohr) for the energy/force path: neighbour rows and pair lists are reused until an atom has moved
        more than skin / 2 (decided on the device).  0 disables; exact neighbour-set queries need 0."""
        self.skin = float(skin)
        _lib.check(_lib.load().pantea_workspace_set_skin(self.handle, self.skin))

    def set_compute_precision(self, bits: int) -> None:
        """32: mixed mode of a float64 workspace -- symmetry functions in single precision on double-precision state
        (difference vectors still formed in double); 64: everything double (default)."""
        _lib.check(_lib.load().pantea_workspace_set_compute_precision(self.handle, int(bits)))
        self.compute_bits = int(bits)

    def rebuild_counts(self) -> Tuple[int, int]:
        """(neighbour builds that ran with a skin, how many of them rebuilt the rows)."""
        out = (C.c_int64 * 2)()
        _lib.check(_lib.load().pantea_neighbor_rebuilds(self.handle, out, _lib.stream_ptr()))
        return int(out[0]), int(out[1])

    def close(self) -> None:
        if self.handle:
            _lib.load().pantea_workspace_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self) -> None:
        try:
            self.close()
        except Exception:
            pass

    def _grow(self, needed: int) -> None:
        self.close()
        self.max_neighbors = int(needed * 1.25) + 32
        self._create()

    # ---------------------------------------------------------------------------------- binding
    def bind(self, positions: torch.Tensor, types: torch.Tensor, box: Optional[Sequence[float]], r_cutoff: float,
             check: bool = True, owned: Optional[Tuple[int, int]] = None) -> None:
        """Neighbour search for one structure; `types` are potential atom types (int32)."""
        lib = _lib.load()
        n = int(positions.shape[0])
        if n > self.max_atoms:
            raise ValueError(f"workspace holds {self.max_atoms} atoms, got {n}")
        positions = positions.contiguous()
        types = types.contiguous()
        assert positions.dtype == self.dtype and types.dtype == torch.int32
        box_c = _lib.box_arg(box)
        while True:
            _lib.check(lib.pantea_workspace_set_owned_range(self.handle, *(owned if owned is not None else (0, -1))))
            _lib.check(lib.pantea_neighbor_build(self.handle, _lib.ptr(positions), _lib.ptr(types), n, box_c,
                                                 float(r_cutoff), _lib.stream_ptr()))
            self.n_atoms = n
            self._keep = (positions, types)
            if not check:
                return
            mx = C.c_int32(0)
            code = lib.pantea_neighbor_status(self.handle, C.byref(mx), _lib.stream_ptr())
            if code == _lib.PANTEA_ECAPACITY:
                self._grow(mx.value)
                continue
            _lib.check(code)
            return

    def bind_batch(self, positions: torch.Tensor, types: torch.Tensor, struct_ptr: torch.Tensor,
                   boxes: Optional[torch.Tensor], r_cutoff: float, check: bool = True) -> None:
        """Neighbour search inside each of many independent structures (dataset preprocessing)."""
        lib = _lib.load()
        n = int(positions.shape[0])
        if n > self.max_atoms:
            raise ValueError(f"workspace holds {self.max_atoms} atoms, got {n}")
        positions, types, struct_ptr = positions.contiguous(), types.contiguous(), struct_ptr.contiguous()
        assert positions.dtype == self.dtype and types.dtype == torch.int32 and struct_ptr.dtype == torch.int32
        if boxes is not None:
            boxes = boxes.to(torch.float64).contiguous()
        while True:
            _lib.check(lib.pantea_workspace_set_owned_range(self.handle, 0, -1))
            _lib.check(lib.pantea_neighbor_build_batch(self.handle, _lib.ptr(positions), _lib.ptr(types), n,
                                                       _lib.ptr(struct_ptr), _lib.ptr(boxes), int(struct_ptr.numel() - 1),
                                                       float(r_cutoff), _lib.stream_ptr()))
            self.n_atoms = n
            self._keep = (positions, types, struct_ptr, boxes)
            if not check:
                return
            mx = C.c_int32(0)
            code = lib.pantea_neighbor_status(self.handle, C.byref(mx), _lib.stream_ptr())
            if code == _lib.PANTEA_ECAPACITY:
                self._grow(mx.value)
                continue
            _lib.check(code)
            return

    # ---------------------------------------------------------------------------------- queries
    def neighbor_lists(self) -> Tuple[torch.Tensor, torch.Tensor]:
        """CSR neighbour lists `(row_ptr int64 [n+1], col int32)` with ascending columns."""
        lib = _lib.load()
        n = self.n_atoms
        dev = self._keep[0].device
        counts = torch.zeros(n, dtype=torch.int32, device=dev)
        _lib.check(lib.pantea_neighbor_counts(self.handle, _lib.ptr(counts), _lib.stream_ptr()))
        row_ptr = torch.zeros(n + 1, dtype=torch.int64, device=dev)
        torch.cumsum(counts, 0, out=row_ptr[1:])
        total = int(row_ptr[-1].item()) if n else 0
        col = torch.empty(max(total, 1), dtype=torch.int32, device=dev)
        _lib.check(lib.pantea_neighbor_export(self.handle, _lib.ptr(row_ptr), _lib.ptr(col), _lib.stream_ptr()))
        return row_ptr, col[:total]

    def distances(self, idx_i: Optional[torch.Tensor], idx_j: Optional[torch.Tensor], with_aux: bool):
        lib = _lib.load()
        dev = self._keep[0].device
        n_i = self.n_atoms if idx_i is None else int(idx_i.numel())
        n_j = self.n_atoms if idx_j is None else int(idx_j.numel())
        r = torch.empty((n_i, n_j), dtype=self.dtype, device=dev)
        d = torch.empty((n_i, n_j, 3), dtype=self.dtype, device=dev) if with_aux else None
        _lib.check(lib.pantea_distances(self.handle, _lib.ptr(idx_i), n_i, _lib.ptr(idx_j), n_j, _lib.ptr(r), _lib.ptr(d),
                                        _lib.stream_ptr()))
        return (r, d) if with_aux else r

    def _checked(self, launch) -> None:
        """Run `launch()` and verify (synchronising) that no device-side capacity was exceeded; the library raises
        its capacities from the observed maxima, so an overflowing evaluation is simply repeated."""
        lib = _lib.load()
        for attempt in range(4):
            launch()
            code = lib.pantea_neighbor_status(self.handle, None, _lib.stream_ptr())
            if code != _lib.PANTEA_ECAPACITY or attempt == 3:
                _lib.check(code)
                return

    def acsf(self, element_slot: int, n_symfunc: int, centres: Optional[torch.Tensor], values: bool, grad: bool):
        lib = _lib.load()
        dev = self._keep[0].device
        n_c = self.n_atoms if centres is None else int(centres.numel())
        if centres is not None:
            centres = centres.to(torch.int32).contiguous()
        G = torch.zeros((n_c, n_symfunc), dtype=self.dtype, device=dev) if values else None
        dG = torch.zeros((n_c, n_symfunc, 3), dtype=self.dtype, device=dev) if grad else None
        if n_c > 0 and n_symfunc > 0:
            self._checked(lambda: _lib.check(lib.pantea_acsf_compute(
                self.handle, element_slot, _lib.ptr(centres), n_c, _lib.ptr(G), _lib.ptr(dG), _lib.stream_ptr())))
        return G, dG

    def energy_forces(self, want_energy: bool = True, want_forces: bool = True, want_atomic: bool = False,
                      out_forces: Optional[torch.Tensor] = None, force_mode: int = 0):
        lib = _lib.load()
        dev = self._keep[0].device
        n = self.n_atoms
        e_total = torch.zeros((), dtype=self.dtype, device=dev) if want_energy else None
        e_atom = torch.zeros(n, dtype=self.dtype, device=dev) if want_atomic else None
        forces = None
        if want_forces:
            forces = out_forces if out_forces is not None else torch.zeros((n, 3), dtype=self.dtype, device=dev)
        self._checked(lambda: _lib.check(lib.pantea_energy_forces(
            self.handle, _lib.ptr(e_atom), _lib.ptr(forces), _lib.ptr(e_total), int(force_mode), _lib.stream_ptr())))
        return e_total, e_atom, forces


def remap_types(structure, type_of: Dict[Element, int]) -> torch.Tensor:
    """Structure atom types (numbered within the structure's own element set) -> potential atom types.

    Elements unknown to the potential map to 0 (an atom that no symmetry function refers to).
    """
    emap = structure.element_map.element_to_atom_type
    if all(type_of.get(el, 0) == t for el, t in emap.items()):
        return structure.atom_types
    lut = torch.zeros(max(emap.values()) + 1, dtype=torch.int32, device=structure.atom_types.device)
    for el, t in emap.items():
        lut[t] = type_of.get(el, 0)
    return lut[structure.atom_types.long()]


def box_lengths(structure) -> Optional[List[float]]:
    """Host lattice diagonal of a structure (cached on the box: the reference also only uses the diagonal)."""
    if structure.box is None:
        return None
    box = structure.box
    cached = getattr(box, "_host_diag", None)
    if cached is None:
        cached = torch.diagonal(box.lattice).detach().double().cpu().tolist()
        box._host_diag = cached
    return cached


def number_density(structure) -> Optional[float]:
    diag = box_lengths(structure)
    if diag is None:
        return None
    return structure.natoms / (diag[0] * diag[1] * diag[2])
