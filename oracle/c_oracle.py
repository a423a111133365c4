"""TEST INFRASTRUCTURE ONLY -- ctypes binding of oracle/liboracle_hdnnp.so (see hdnnp_oracle.c)."""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path
from typing import List, Optional, Sequence, Tuple

import numpy as np

from oracle.spec import ACT_CODES, CUTOFF_CODES, ElementSpec

_HERE = Path(__file__).parent
_LIB: Optional[C.CDLL] = None


class _SymFunc(C.Structure):
    _fields_ = [("kind", C.c_int), ("cutoff_type", C.c_int), ("type_j", C.c_int), ("type_k", C.c_int),
                ("r_cutoff", C.c_double), ("eta", C.c_double), ("r_shift", C.c_double), ("lambda0", C.c_double),
                ("zeta", C.c_double)]


class _Element(C.Structure):
    _fields_ = [("central_type", C.c_int), ("n_sf", C.c_int), ("sf", C.POINTER(_SymFunc)),
                ("shift", C.POINTER(C.c_double)), ("slope", C.POINTER(C.c_double)), ("offset", C.POINTER(C.c_double)),
                ("n_layers", C.c_int), ("sizes", C.POINTER(C.c_int)), ("acts", C.POINTER(C.c_int)),
                ("weights", C.POINTER(C.c_double))]


def build(force: bool = False) -> Path:
    so = _HERE / "liboracle_hdnnp.so"
    src = _HERE / "hdnnp_oracle.c"
    if force or not so.exists() or so.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["make", "-C", str(_HERE), "-B" if force else "-s"], check=True, capture_output=True)
    return so


def lib() -> C.CDLL:
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(str(build()))
        _LIB.orc_neighbors.restype = C.c_long
        _LIB.orc_num_threads.restype = C.c_int
    return _LIB


def _dptr(a: Optional[np.ndarray]):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


def _iptr(a: Optional[np.ndarray]):
    return a.ctypes.data_as(C.POINTER(C.c_int)) if a is not None else None


class PackedPotential:
    """Keeps the numpy buffers alive behind an array of `orc_element`."""

    def __init__(self, specs: Sequence[ElementSpec]) -> None:
        self._keep: List[object] = []
        self.n = len(specs)
        self.array = (_Element * self.n)()
        for e, spec in enumerate(specs):
            sfs = (_SymFunc * max(len(spec.symfuncs), 1))()
            for s, sf in enumerate(spec.symfuncs):
                sfs[s] = _SymFunc(sf.kind, CUTOFF_CODES[sf.cutoff_type], sf.type_j, sf.type_k, sf.r_cutoff, sf.eta,
                                  sf.r_shift, sf.lambda0, sf.zeta)
            el = self.array[e]
            el.central_type, el.n_sf, el.sf = spec.atom_type, len(spec.symfuncs), sfs
            self._keep.append(sfs)
            if spec.scale_type is not None:
                shift, slope, offset = (np.ascontiguousarray(a, dtype=np.float64) for a in spec.affine())
                el.shift, el.slope, el.offset = _dptr(shift), _dptr(slope), _dptr(offset)
                self._keep += [shift, slope, offset]
            sizes = np.asarray([len(spec.symfuncs)] + [k.shape[1] for k, _, _ in spec.layers], dtype=np.int32)
            acts = np.asarray([ACT_CODES[a] for _, _, a in spec.layers], dtype=np.int32)
            chunks = [np.zeros(0)]
            for k, b, _ in spec.layers:
                chunks += [np.asarray(k, dtype=np.float64).ravel(), np.asarray(b, dtype=np.float64).ravel()]
            weights = np.ascontiguousarray(np.concatenate(chunks))
            el.n_layers, el.sizes, el.acts, el.weights = len(spec.layers), _iptr(sizes), _iptr(acts), _dptr(weights)
            self._keep += [sizes, acts, weights]


def _prep(pos, types, box):
    pos = np.ascontiguousarray(pos, dtype=np.float64)
    types = np.ascontiguousarray(types, dtype=np.int32)
    box = None if box is None else np.ascontiguousarray(box, dtype=np.float64)
    return pos, types, box


def neighbors(pos, types, box, rc: float) -> Tuple[np.ndarray, np.ndarray]:
    pos, types, box = _prep(pos, types, box)
    n = len(pos)
    row_ptr = np.zeros(n + 1, dtype=np.int64)
    rp = row_ptr.ctypes.data_as(C.POINTER(C.c_long))
    total = lib().orc_neighbors(_dptr(pos), _iptr(types), C.c_long(n), _dptr(box), C.c_double(rc), rp, None, C.c_long(0))
    col = np.zeros(max(total, 1), dtype=np.int32)
    lib().orc_neighbors(_dptr(pos), _iptr(types), C.c_long(n), _dptr(box), C.c_double(rc), rp, _iptr(col), C.c_long(total))
    return row_ptr, col[:total]


def distances(pos, box) -> np.ndarray:
    pos = np.ascontiguousarray(pos, dtype=np.float64)
    box = None if box is None else np.ascontiguousarray(box, dtype=np.float64)
    out = np.zeros((len(pos), len(pos)))
    lib().orc_distances(_dptr(pos), C.c_long(len(pos)), _dptr(box), _dptr(out))
    return out


def acsf(spec: ElementSpec, pos, types, box, centres=None, grad: bool = True):
    pos, types, box = _prep(pos, types, box)
    packed = PackedPotential([spec])
    n = len(pos)
    centres_arr = None if centres is None else np.ascontiguousarray(centres, dtype=np.int32)
    n_c = n if centres is None else len(centres_arr)
    nsf = len(spec.symfuncs)
    G = np.zeros((n_c, nsf))
    dG = np.zeros((n_c, nsf, 3)) if grad else None
    rc = lib().orc_acsf(C.byref(packed.array[0]), _dptr(pos), _iptr(types), C.c_long(n), _dptr(box),
                        _iptr(centres_arr), C.c_long(n_c), _dptr(G), _dptr(dG))
    assert rc == 0
    return G, dG


def energy_forces(specs: Sequence[ElementSpec], pos, types, box, want_forces: bool = True, begin: int = 0,
                  end: Optional[int] = None):
    pos, types, box = _prep(pos, types, box)
    packed = PackedPotential(specs)
    n = len(pos)
    end = n if end is None else end
    e_atom = np.zeros(n)
    forces = np.zeros((n, 3)) if want_forces else None
    e_sum = C.c_double(0.0)
    rc = lib().orc_energy_forces_range(packed.array, C.c_int(packed.n), _dptr(pos), _iptr(types), C.c_long(n),
                                       _dptr(box), C.c_long(begin), C.c_long(end), _dptr(e_atom), _dptr(forces),
                                       C.byref(e_sum))
    assert rc == 0
    return float(e_atom[begin:end].sum()), e_atom, forces


def energy_full_forces(specs: Sequence[ElementSpec], pos, types, box):
    """(E, e_atom [n], F [n,3]) with the FULL force -dE/dr (extension, not a reference mode; analytic, any size)."""
    pos, types, box = _prep(pos, types, box)
    packed = PackedPotential(specs)
    n = len(pos)
    e_atom, forces, e_total = np.zeros(n), np.zeros((n, 3)), C.c_double(0.0)
    rc = lib().orc_energy_full_forces(packed.array, C.c_int(packed.n), _dptr(pos), _iptr(types), C.c_long(n), _dptr(box),
                                      _dptr(e_atom), _dptr(forces), C.byref(e_total))
    assert rc == 0
    return float(e_total.value), e_atom, forces


def md_run(specs: Sequence[ElementSpec], pos, vel, mass, types, box, dt: float, n_steps: int,
           t_target: float = 0.0, tau: float = 0.0, kb: float = 3.166811563e-6, mass_scaled: bool = False):
    """Returns (pos, vel, forces, scalars[n_steps+1,3] = (Epot, Ekin, T)).  `mass_scaled` (extension, not a reference
    mode): accelerations F/m in place of F."""
    lib().orc_set_mass_scaled(C.c_int(1 if mass_scaled else 0))
    pos, types, box = _prep(np.array(pos, copy=True), types, box)
    vel = np.ascontiguousarray(np.array(vel, copy=True), dtype=np.float64)
    mass = np.ascontiguousarray(np.asarray(mass, dtype=np.float64).reshape(-1))
    _, _, forces = energy_forces(specs, pos, types, box)
    packed = PackedPotential(specs)
    scalars = np.zeros((n_steps + 1, 3))
    rc = lib().orc_md_run(packed.array, C.c_int(packed.n), _dptr(pos), _dptr(vel), _dptr(forces), _dptr(mass),
                          _iptr(types), C.c_long(len(pos)), _dptr(box), C.c_double(dt), C.c_long(n_steps),
                          C.c_double(t_target), C.c_double(tau), C.c_double(kb), _dptr(scalars))
    lib().orc_set_mass_scaled(C.c_int(0))
    assert rc == 0
    return pos, vel, forces, scalars


def set_use_cells(flag: bool) -> None:
    """Toggle the cell-list neighbour gather (results are identical either way; tests check that)."""
    lib().orc_set_use_cells(C.c_int(1 if flag else 0))


def num_threads() -> int:
    return int(lib().orc_num_threads())


def set_num_threads(n: int) -> None:
    lib().orc_set_num_threads(C.c_int(n))
