"""RuNNer `input.data` reader (format handled as in reference `pantea/datasets/runner.py:12-140`).

Each `begin ... end` block holds `lattice` rows, `atom x y z element charge energy fx fy fz`
rows and the collective `energy` / `charge` values.
"""
from __future__ import annotations

from collections import defaultdict
from pathlib import Path
from typing import Dict, Iterator, List, Optional, TextIO

from pantea_b200.atoms.structure import Structure
from pantea_b200.types import Dtype, as_torch_dtype, default_dtype
from pantea_b200.utils.tokenize import tokenize


class RunnerDataSource:
    def __init__(self, filename: Path, dtype: Optional[Dtype] = None) -> None:
        self.filename = Path(filename)
        self.dtype = as_torch_dtype(dtype) if dtype is not None else default_dtype.FLOATX

    def __len__(self) -> int:
        count = 0
        with open(self.filename, "r") as file:
            while self._skip_block(file):
                count += 1
        return count

    def __getitem__(self, index: int) -> Structure:
        with open(self.filename, "r") as file:
            for _ in range(index):
                self._skip_block(file)
            data = self._read_block(file)
        if not data:
            raise IndexError(f"The given index {index} is out of bound (len={len(self)})")
        return self._to_structure(data)

    def read_structures(self) -> Iterator[Structure]:
        with open(self.filename, "r") as file:
            while True:
                data = self._read_block(file)
                if not data:
                    return
                yield self._to_structure(data)

    @staticmethod
    def _read_block(file: TextIO) -> Dict[str, List]:
        data: Dict[str, List] = defaultdict(list)
        for line in file:
            if tokenize(line)[0] == "begin":
                break
        else:
            return data
        for line in file:
            keyword, tokens = tokenize(line)
            if keyword == "atom":
                data["positions"].append([float(t) for t in tokens[:3]])
                data["elements"].append(tokens[3])
                data["charges"].append(float(tokens[4]))
                data["energies"].append(float(tokens[5]))
                data["forces"].append([float(t) for t in tokens[6:9]])
            elif keyword == "lattice":
                data["lattice"].append([float(t) for t in tokens[:3]])
            elif keyword == "energy":
                data["total_energy"].append(float(tokens[0]))
            elif keyword == "charge":
                data["total_charge"].append(float(tokens[0]))
            elif keyword == "comment":
                data["comment"].append(" ".join(line.split()[1:]))
            elif keyword == "end":
                break
        return data

    @staticmethod
    def _skip_block(file: TextIO) -> bool:
        for line in file:
            if tokenize(line)[0] == "end":
                return True
        return False

    def _to_structure(self, data: Dict[str, List]) -> Structure:
        data = dict(data)
        data.pop("comment", None)
        return Structure.from_dict(data, dtype=self.dtype)

    def __repr__(self) -> str:
        return f"{self.__class__.__name__}(filename='{self.filename}', dtype={self.dtype})"
