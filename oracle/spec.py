"""TEST INFRASTRUCTURE ONLY -- potential/structure specs shared by the two oracle tiers.

Builds plain-numpy descriptions of an HDNNP (symmetry functions, scaler statistics, MLP
weights) straight from the fixture files, independently of the product package, and generates
the deterministic synthetic water boxes of SURVEY.md section 8(d).
"""
from __future__ import annotations

import json
import math
from dataclasses import dataclass, field
from pathlib import Path
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

CUTOFF_CODES = {"hard": 0, "cos": 1, "tanhu": 2, "tanh": 3, "exp": 4, "poly1": 5, "poly2": 6}
CUTOFF_NAMES = {v: k for k, v in CUTOFF_CODES.items()}
ACT_CODES = {"identity": 0, "tanh": 1, "logistic": 2, "softplus": 3, "relu": 4, "gaussian": 5, "cos": 6, "exp": 7,
             "harmonic": 8}
ACT_NAMES = {v: k for k, v in ACT_CODES.items()}
ATOMIC_NUMBER = {"H": 1, "He": 2, "C": 6, "N": 7, "O": 8, "Ne": 10}
MASS_U = {"H": 1.008, "He": 4.003, "C": 12.011, "N": 14.007, "O": 15.999, "Ne": 20.180}
FROM_ATOMIC_MASS = 1.0 / 5.48579957163e-4  # pantea/units.py:64-71
KB = 3.166811563e-6
BOHR_PER_ANGSTROM = 1.0 / 0.529177249


@dataclass
class SymFuncSpec:
    kind: int
    cutoff_type: str
    r_cutoff: float
    type_j: int
    type_k: int = 0
    eta: float = 0.0
    r_shift: float = 0.0
    lambda0: float = 0.0
    zeta: float = 0.0


@dataclass
class ElementSpec:
    atom_type: int
    symfuncs: List[SymFuncSpec]
    scale_type: Optional[str] = None           # None -> no scaler (bare descriptor)
    scaler: Dict[str, np.ndarray] = field(default_factory=dict)
    scale_min: float = 0.0
    scale_max: float = 1.0
    layers: List[Tuple[np.ndarray, np.ndarray, str]] = field(default_factory=list)  # (kernel[in,out], bias, act)

    def affine(self) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
        """(shift, slope, offset): x' = offset + slope * (x - shift)  (pantea/descriptors/scaler.py:206-246)."""
        n = len(self.symfuncs)
        one, zero = np.ones(n), np.zeros(n)
        if self.scale_type is None:
            return zero, one, zero
        sc, smin, smax = self.scaler, self.scale_min, self.scale_max
        if self.scale_type == "center":
            return sc["mean"], one, zero
        if self.scale_type == "scale":
            return sc["minval"], (smax - smin) / (sc["maxval"] - sc["minval"]), smin * one
        if self.scale_type == "scale_center":
            return sc["mean"], (smax - smin) / (sc["maxval"] - sc["minval"]), smin * one
        if self.scale_type == "scale_center_sigma":
            return sc["mean"], (smin - smax) / sc["sigma"], smin * one
        raise KeyError(self.scale_type)


def type_map(elements: Sequence[str]) -> Dict[str, int]:
    """1-based atom types by ascending atomic number -- pantea/atoms/element.py:99-108."""
    uniq = sorted(set(elements), key=ATOMIC_NUMBER.__getitem__)
    return {el: t for t, el in enumerate(uniq, start=1)}


def load_potential(json_file: Path, weights_dir: Optional[Path] = None) -> List[ElementSpec]:
    """h2o.json-style settings + scaling.{Z:03d}.json + weights.{Z:03d}.npz -> element specs.

    Follows pantea/potentials/nnp/potential.py:196-317 (radial SFs first, then angular, file
    order within each; hidden layers = zip(nodes, activations[:-1]), output (1, activations[-1])).
    """
    json_file = Path(json_file)
    weights_dir = Path(weights_dir) if weights_dir else json_file.parent
    cfg = json.loads(json_file.read_text())
    elements = sorted(cfg["elements"], key=ATOMIC_NUMBER.__getitem__)
    tmap = type_map(elements)
    specs = []
    for el in elements:
        radial, angular = [], []
        for row in cfg["symfunction_short"]:
            if row[0] != el:
                continue
            if len(row) == 6:
                _, kind, j, eta, rc, rs = row
                radial.append(SymFuncSpec(int(kind), cfg["cutoff_type"], rc, tmap[j], 0, eta, rs))
            else:
                _, kind, j, eta, rc, rs, k, lam, zeta = row
                angular.append(SymFuncSpec(int(kind), cfg["cutoff_type"], rc, tmap[j], tmap[k], eta, rs, lam, zeta))
        z = ATOMIC_NUMBER[el]
        sc_raw = json.loads((weights_dir / f"scaling.{z:03d}.json").read_text())
        scaler = {k: np.asarray(v, dtype=np.float64) for k, v in sc_raw.items() if k not in ("dimension", "nsamples")}
        wz = np.load(weights_dir / f"weights.{z:03d}.npz")
        acts = list(cfg["global_activation_short"])
        layers = []
        for l, act in enumerate(acts):
            layers.append((wz[f"layers_{2 * l}.kernel"].astype(np.float64), wz[f"layers_{2 * l}.bias"].astype(np.float64), act))
        specs.append(ElementSpec(tmap[el], radial + angular, cfg.get("scale_type", "center"), scaler,
                                 cfg.get("scale_min_short", 0.0), cfg.get("scale_max_short", 1.0), layers))
    return specs


def read_runner(file: Path) -> List[Dict[str, np.ndarray]]:
    """Minimal RuNNer input.data reader (pantea/datasets/runner.py:90-124). Positions are NOT wrapped here."""
    frames, cur = [], None
    for line in Path(file).read_text().splitlines():
        tok = line.split()
        if not tok:
            continue
        key = tok[0].lower()
        if key == "begin":
            cur = {"positions": [], "elements": [], "lattice": [], "forces": []}
        elif key == "atom":
            cur["positions"].append([float(t) for t in tok[1:4]])
            cur["elements"].append(tok[4])
            cur["forces"].append([float(t) for t in tok[7:10]])
        elif key == "lattice":
            cur["lattice"].append([float(t) for t in tok[1:4]])
        elif key == "energy":
            cur["total_energy"] = float(tok[1])
        elif key == "end":
            tm = type_map(cur["elements"])
            frames.append({
                "positions": np.asarray(cur["positions"], dtype=np.float64),
                "types": np.asarray([tm[e] for e in cur["elements"]], dtype=np.int32),
                "elements": list(cur["elements"]),
                "box": np.diag(np.asarray(cur["lattice"], dtype=np.float64)).copy() if cur["lattice"] else None,
            })
    return frames


def rune_width_potential(seed: int = 7) -> List[ElementSpec]:
    """Optional stress potential "RuNNer width": 30 SFs per element, MLP 30-25-25-1 tanh (non-reference)."""
    rng = np.random.default_rng(seed)
    specs = []
    for t in (1, 2):
        sfs = []
        for tj in (1, 2):
            for eta in (0.001, 0.01, 0.03, 0.06, 0.15):
                sfs.append(SymFuncSpec(2, "tanhu", 12.0, tj, 0, eta, 0.0))
        for (tj, tk) in ((1, 1), (1, 2), (2, 2)):
            for eta in (0.001, 0.01, 0.03):
                for lam, zeta in ((1.0, 1.0), (-1.0, 2.0)):
                    sfs.append(SymFuncSpec(3, "tanhu", 12.0, tj, tk, eta, 0.0, lam, zeta))
        for eta in (0.005, 0.02):
            sfs.append(SymFuncSpec(9, "tanhu", 12.0, 1, 2, eta, 0.0, 1.0, 4.0))
        sfs.sort(key=lambda s: s.kind >= 3)  # radial first
        n = len(sfs)
        sizes = [n, 25, 25, 1]
        layers = []
        for l in range(3):
            layers.append((rng.uniform(-1, 1, (sizes[l], sizes[l + 1])) / math.sqrt(sizes[l]),
                           rng.uniform(-0.1, 0.1, sizes[l + 1]), "tanh" if l < 2 else "identity"))
        scaler = {"mean": rng.uniform(0, 1, n), "sigma": rng.uniform(0.5, 1, n),
                  "minval": rng.uniform(-1, 0, n), "maxval": rng.uniform(1, 2, n)}
        specs.append(ElementSpec(t, sfs, "scale_center", scaler, 0.0, 1.0, layers))
    return specs


# ----------------------------------------------------------------------------- synthetic water
# the generator is shared with bench.py, which may not import oracle/ on the product path
from pantea_b200.utils.synthetic import md_velocities, water_box, water_masses  # noqa: E402,F401
