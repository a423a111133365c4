"""Structure dataset with optional in-memory caching (reference `pantea/datasets/dataset.py:25-70`)."""
from __future__ import annotations

from pathlib import Path
from typing import Any, Dict, Optional

from pantea_b200.atoms.structure import Structure
from pantea_b200.datasets.runner import RunnerDataSource
from pantea_b200.types import Dtype


class Dataset:
    def __init__(self, datasource: Any, persist: bool = False) -> None:
        self.datasource = datasource
        self.persist = persist
        self.cache: Dict[int, Structure] = {}

    @classmethod
    def from_runner(cls, filename: Path, persist: bool = False, dtype: Optional[Dtype] = None) -> "Dataset":
        return cls(RunnerDataSource(filename, dtype), persist)

    def __len__(self) -> int:
        return len(self.datasource)

    def __getitem__(self, index: int) -> Structure:
        if self.persist and index in self.cache:
            return self.cache[index]
        structure = self.datasource[index]
        if self.persist:
            self.cache[index] = structure
        return structure

    def __iter__(self):
        for index in range(len(self)):
            yield self[index]

    def preload(self) -> None:
        """Cache every structure (sequential read when the source supports it)."""
        self.persist = True
        reader = getattr(self.datasource, "read_structures", None)
        if reader is not None:
            for index, structure in enumerate(reader()):
                self.cache[index] = structure
        else:
            for index in range(len(self)):
                self.cache[index] = self.datasource[index]

    def __repr__(self) -> str:
        return f"{self.__class__.__name__}(datasource={self.datasource}, persist={self.persist})"
