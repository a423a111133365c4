"""HDNNP settings (keywords, defaults and file formats of reference
`pantea/potentials/nnp/settings.py:83-308` and `pantea/config.py:10-36`).

Two on-disk formats are read: the JSON dump of the settings object and the RuNNer/n2p2
`input.nn` keyword file.  Quirks preserved on purpose:
JSON `symfunction_short` rows are positional -- radial `(central, type, j, eta, r_cutoff, r_shift)`,
angular `(central, type, j, eta, r_cutoff, r_shift, k, lambda, zeta)` -- whereas `.nn` rows read
`central type j eta r_shift r_cutoff` / `central type j k eta lambda zeta r_cutoff [r_shift]`;
unknown JSON keys are ignored; `.nn` keywords that are not settings fields are dropped, so the
`*_symmetry_functions` scaler switches never take effect for `.nn` files.
"""
from __future__ import annotations

import json
from dataclasses import dataclass, field, fields
from pathlib import Path
from typing import Any, Dict, List, Mapping, NamedTuple, Union

from pantea_b200.atoms.element import ElementMap
from pantea_b200.logger import logger
from pantea_b200.types import Element
from pantea_b200.utils.tokenize import tokenize


class RadialSymFuncArgs(NamedTuple):
    central_element: str
    acsf_type: int
    neighbor_element_j: str
    eta: float
    r_cutoff: float
    r_shift: float


class AngularSymFuncArgs(NamedTuple):
    central_element: str
    acsf_type: int
    neighbor_element_j: str
    eta: float
    r_cutoff: float
    r_shift: float
    neighbor_element_k: str
    lambda0: float
    zeta: float


SymFuncArgs = Union[RadialSymFuncArgs, AngularSymFuncArgs]

cutoff_function_map: Mapping[str, str] = {
    "0": "hard", "1": "cos", "2": "tanhu", "3": "tanh", "4": "exp", "5": "poly1", "6": "poly2",
}
scaler_type_map: Mapping[str, str] = {
    "center_symmetry_functions": "center",
    "scale_symmetry_functions": "scale",
    "scale_center_symmetry_functions": "scale_center",
    "scale_center_symmetry_functions_sigma": "scale_center_sigma",
}
activation_function_map: Mapping[str, str] = {
    "l": "identity", "t": "tanh", "s": "logistic", "p": "softplus", "r": "relu",
    "g": "gaussian", "c": "cos", "e": "exp", "h": "harmonic",
}
updater_type_map: Mapping[str, str] = {"0": "gradient_descent", "1": "kalman_filter"}
gradient_type_map: Mapping[str, str] = {"0": "fixed_step", "1": "adam"}

_MISSING = object()


@dataclass
class NeuralNetworkPotentialSettings:
    # General
    number_of_elements: int = _MISSING  # type: ignore[assignment]
    global_hidden_layers_short: int = _MISSING  # type: ignore[assignment]
    global_nodes_short: List[int] = _MISSING  # type: ignore[assignment]
    global_activation_short: List[str] = _MISSING  # type: ignore[assignment]
    random_seed: int = 2023
    elements: List[Element] = field(default_factory=list)
    atom_energy: Dict[Element, float] = field(default_factory=dict)
    scaler_save_format: str = "scaling.{:03d}.json"
    model_save_format: str = "weights.{:03d}.pkl"
    # Neural network
    weights_min: float = -1.0
    weights_max: float = 1.0
    # Trainer (kept for file compatibility; training is outside this package's scope)
    epochs: int = 1
    updater_type: str = "gradient_descent"
    gradient_type: str = "adam"
    main_error_metric: str = "RMSE"
    force_weight: float = 1.0
    short_force_fraction: float = 0.1
    short_energy_fraction: float = 1.0
    test_fraction: float = 0.1
    save_best_model: bool = True
    gradient_eta: float = 1.0e-5
    gradient_adam_eta: float = 1.0e-3
    gradient_adam_beta1: float = 0.9
    gradient_adam_beta2: float = 0.999
    gradient_adam_epsilon: float = 1.0e-8
    gradient_adam_weight_decay: float = 1.0e-4
    kalman_type: int = 0
    kalman_epsilon: float = 0.01
    kalman_q0: float = 0.01
    kalman_qtau: float = 2.302
    kalman_qmin: float = 1.0e-6
    kalman_eta: float = 0.01
    kalman_etatau: float = 2.302
    kalman_etamax: float = 1.0
    kalman_lambda_short: float = 0.96000
    kalman_neu_short: float = 0.99950
    # Symmetry functions
    cutoff_type: str = "tanh"
    scale_type: str = "center"
    scale_min_short: float = 0.0
    scale_max_short: float = 1.0
    symfunction_short: List[SymFuncArgs] = field(default_factory=list)

    # -------------------------------------------------------------- validation / coercion
    def __post_init__(self) -> None:
        hints = {f.name: f.type for f in fields(self)}
        for name in ("number_of_elements", "global_hidden_layers_short", "global_nodes_short",
                     "global_activation_short"):
            if getattr(self, name) is _MISSING:
                raise ValueError(f"missing required setting '{name}'")
        for name, hint in hints.items():
            value = getattr(self, name)
            hint = str(hint)
            try:
                if hint == "int":
                    value = int(value)
                elif hint == "float":
                    value = float(value)
                elif hint == "bool":
                    value = bool(value)
                elif hint == "List[int]":
                    value = [int(v) for v in value]
                elif hint in ("List[str]", "List[Element]"):
                    value = [str(v) for v in value]
                elif hint == "Dict[Element, float]":
                    value = {str(k): float(v) for k, v in dict(value).items()}
            except (TypeError, ValueError) as exc:
                raise ValueError(f"invalid value for setting '{name}': {value!r}") from exc
            setattr(self, name, value)
        self.symfunction_short = [self._coerce_symfunc(row) for row in self.symfunction_short]

    @staticmethod
    def _coerce_symfunc(row: Any) -> SymFuncArgs:
        if isinstance(row, (RadialSymFuncArgs, AngularSymFuncArgs)):
            return row
        row = list(row)
        if len(row) == 6:
            c, t, j, eta, rc, rs = row
            return RadialSymFuncArgs(str(c), int(t), str(j), float(eta), float(rc), float(rs))
        if len(row) == 9:
            c, t, j, eta, rc, rs, k, lam, zeta = row
            return AngularSymFuncArgs(str(c), int(t), str(j), float(eta), float(rc), float(rs),
                                      str(k), float(lam), float(zeta))
        raise ValueError(f"symfunction_short row must have 6 or 9 entries: {row!r}")

    # -------------------------------------------------------------- dict-style access (config.py:14-26)
    def __getitem__(self, keyword: str) -> Any:
        return getattr(self, keyword)

    def __setitem__(self, name: str, value: Any) -> None:
        setattr(self, name, value)

    def keywords(self) -> List[str]:
        return [f.name for f in fields(self)]

    def dict(self) -> Dict[str, Any]:
        out = {f.name: getattr(self, f.name) for f in fields(self)}
        out["symfunction_short"] = [list(row) for row in self.symfunction_short]
        return out

    def to_json(self, file: Path) -> None:
        with open(str(Path(file)), "w") as fp:
            json.dump(self.dict(), fp, indent=4)

    # -------------------------------------------------------------- readers
    @classmethod
    def from_file(cls, filename: Path) -> "NeuralNetworkPotentialSettings":
        suffix = Path(filename).suffix
        if suffix == ".nn":
            return cls.from_nn(filename)
        if suffix == ".json":
            return cls.from_json(filename)
        logger.error(f"Unknown file format '{suffix}': {str(filename)}", exception=ValueError)
        raise AssertionError  # unreachable

    @classmethod
    def from_json(cls, file: Path) -> "NeuralNetworkPotentialSettings":
        with open(str(Path(file)), "r") as fp:
            raw = json.load(fp)
        known = {f.name for f in fields(cls)}
        return cls(**{k: v for k, v in raw.items() if k in known})

    @classmethod
    def from_nn(cls, filename: Path) -> "NeuralNetworkPotentialSettings":
        known = {f.name for f in fields(cls)}
        kwargs: Dict[str, Any] = {"atom_energy": {}, "symfunction_short": []}
        with open(str(filename), "r") as file:
            for line in file:
                keyword, tokens = tokenize(line, comment="#")
                if keyword is None or keyword not in known:
                    continue  # includes the scaler switch keywords (App. B 12)
                if keyword == "elements":
                    kwargs[keyword] = sorted(set(tokens), key=ElementMap.get_atomic_number_from_element)
                elif keyword == "atom_energy":
                    kwargs[keyword][tokens[0]] = tokens[1]
                elif keyword == "global_nodes_short":
                    kwargs[keyword] = list(tokens)
                elif keyword == "global_activation_short":
                    kwargs[keyword] = [activation_function_map[t] for t in tokens]
                elif keyword == "updater_type":
                    kwargs[keyword] = updater_type_map[tokens[0]]
                elif keyword == "gradient_type":
                    kwargs[keyword] = gradient_type_map[tokens[0]]
                elif keyword == "cutoff_type":
                    kwargs[keyword] = cutoff_function_map[tokens[0]]
                elif keyword == "symfunction_short":
                    kind = int(tokens[1])
                    if kind < 3:
                        kwargs[keyword].append(RadialSymFuncArgs(
                            central_element=tokens[0], acsf_type=kind, neighbor_element_j=tokens[2],
                            eta=float(tokens[3]), r_shift=float(tokens[4]), r_cutoff=float(tokens[5])))
                    else:
                        kwargs[keyword].append(AngularSymFuncArgs(
                            central_element=tokens[0], acsf_type=kind, neighbor_element_j=tokens[2],
                            neighbor_element_k=tokens[3], eta=float(tokens[4]), lambda0=float(tokens[5]),
                            zeta=float(tokens[6]), r_cutoff=float(tokens[7]),
                            r_shift=float(tokens[8]) if len(tokens) == 9 else 0.0))
                elif keyword in ("scaler_save_format", "model_save_format", "main_error_metric", "scale_type"):
                    kwargs[keyword] = tokens[0]
                else:  # scalar keywords: first token
                    kwargs[keyword] = tokens[0]
        try:
            return cls(**kwargs)
        except (TypeError, ValueError) as exc:
            logger.error(str(exc), exception=ValueError)
            raise
