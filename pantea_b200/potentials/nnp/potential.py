"""High-dimensional neural network potential (API of reference `pantea/potentials/nnp/potential.py:30-389`).

`nnp(structure)` and `nnp.compute_forces(structure)` run the fused CUDA energy/force kernel
(`pantea_energy_forces`).  The force is the reference's: minus the gradient of the total energy
with respect to the *central* positions only (`force.py:16-43`).  No `atom_energy`
offset is added to the energy (`energy.py:54-63`).
"""
from __future__ import annotations

from collections import defaultdict
from pathlib import Path
from typing import Dict, Optional, Tuple

import torch

from pantea_b200 import engine
from pantea_b200.atoms.element import ElementMap
from pantea_b200.atoms.structure import Structure
from pantea_b200.descriptors.acsf.acsf import ACSF
from pantea_b200.descriptors.acsf.cutoff import CutoffFunction
from pantea_b200.descriptors.acsf.symmetry import G1, G2, G3, G9, NeighborElements
from pantea_b200.descriptors.scaler import DescriptorScaler, ScalerParams
from pantea_b200.logger import logger
from pantea_b200.models.nn.initializer import UniformInitializer
from pantea_b200.models.nn.model import ModelParams, NeuralNetworkModel
from pantea_b200.potentials.nnp.atomic_potential import AtomicPotential
from pantea_b200.potentials.nnp.settings import NeuralNetworkPotentialSettings
from pantea_b200.types import Array, Element


class NeuralNetworkPotential:
    def __init__(
        self,
        directory: Path,
        elements: Tuple[Element, ...],
        scaler_save_format: str,
        model_save_format: str,
        atomic_potentials: Dict[Element, AtomicPotential],
        models_params: Dict[Element, ModelParams],
        scalers_params: Dict[Element, Optional[ScalerParams]],
    ) -> None:
        self.directory = Path(directory)
        self.elements = tuple(elements)
        self.scaler_save_format = scaler_save_format
        self.model_save_format = model_save_format
        self.atomic_potentials = atomic_potentials
        self.models_params = models_params
        self.scalers_params = scalers_params
        self._device: Optional[engine.DevicePotential] = None
        self._device_key: Optional[tuple] = None

    # ------------------------------------------------------------------ construction
    @classmethod
    def from_runner(cls, filename: Path) -> "NeuralNetworkPotential":
        potfile = Path(filename)
        settings = NeuralNetworkPotentialSettings.from_file(potfile)
        atomic_potentials = cls._build_atomic_potentials(settings)
        return cls(
            directory=potfile.parent,
            elements=tuple(settings.elements),
            scaler_save_format=settings.scaler_save_format,
            model_save_format=settings.model_save_format,
            atomic_potentials=atomic_potentials,
            models_params=cls._initialize_models_params(settings, atomic_potentials),
            scalers_params={element: None for element in settings.elements},
        )

    @classmethod
    def _build_atomic_potentials(cls, settings: NeuralNetworkPotentialSettings) -> Dict[Element, AtomicPotential]:
        descriptors = cls._build_descriptors(settings)
        scalers = cls._build_scalers(settings)
        models = cls._build_models(settings)
        return {el: AtomicPotential(descriptors[el], scalers[el], models[el]) for el in settings.elements}

    @classmethod
    def _initialize_models_params(cls, settings, atomic_potentials) -> Dict[Element, ModelParams]:
        """Uniform random kernels in [weights_min, weights_max], zero biases.  The reference seeds a JAX
        PRNG (`potential.py:162-186`); that stream is not reproducible without JAX, so freshly
        initialised (unloaded) potentials differ in their random weights -- `load()` for parity."""
        params = {}
        for i, element in enumerate(settings.elements):
            pot = atomic_potentials[element]
            params[element] = pot.model.init_params(pot.model_input_size, seed=settings.random_seed + i,
                                                    weights_range=(settings.weights_min, settings.weights_max))
        return params

    @classmethod
    def _build_descriptors(cls, settings: NeuralNetworkPotentialSettings) -> Dict[Element, ACSF]:
        """Symmetry functions grouped per central element, radial first then angular, file order within
        each; G3/G9 receive r_shift = r_cutoff, which they ignore (`potential.py:196-260`)."""
        radials, angulars = defaultdict(list), defaultdict(list)
        for args in settings.symfunction_short:
            cfn = CutoffFunction.from_type(settings.cutoff_type, args.r_cutoff)
            if args.acsf_type == 1:
                radials[args.central_element].append((G1(cfn), NeighborElements(args.neighbor_element_j)))
            elif args.acsf_type == 2:
                radials[args.central_element].append(
                    (G2(cfn, eta=args.eta, r_shift=args.r_shift), NeighborElements(args.neighbor_element_j)))
            elif args.acsf_type in (3, 9):
                kind = G3 if args.acsf_type == 3 else G9
                angulars[args.central_element].append(
                    (kind(cfn, eta=args.eta, zeta=args.zeta, lambda0=args.lambda0, r_shift=args.r_cutoff),
                     NeighborElements(args.neighbor_element_j, args.neighbor_element_k)))
        return {el: ACSF(el, tuple(radials[el]), tuple(angulars[el])) for el in settings.elements}

    @classmethod
    def _build_scalers(cls, settings: NeuralNetworkPotentialSettings) -> Dict[Element, DescriptorScaler]:
        return {el: DescriptorScaler.from_type(settings.scale_type, settings.scale_min_short, settings.scale_max_short)
                for el in settings.elements}

    @classmethod
    def _build_models(cls, settings: NeuralNetworkPotentialSettings) -> Dict[Element, NeuralNetworkModel]:
        hidden = tuple(zip(settings.global_nodes_short, settings.global_activation_short[:-1]))
        output = (1, settings.global_activation_short[-1])
        init = UniformInitializer((settings.weights_min, settings.weights_max))
        return {el: NeuralNetworkModel(hidden_layers=hidden, output_layer=output, kernel_initializer=init)
                for el in settings.elements}

    # ------------------------------------------------------------------ parameters
    def load_scaler(self) -> None:
        for element in self.elements:
            z = ElementMap.get_atomic_number_from_element(element)
            file = Path(self.directory, self.scaler_save_format.format(z))
            logger.info(f"Loading scaler parameters for element ({element}): {file.name}")
            self.scalers_params[element] = self.atomic_potentials[element].scaler.load(file)

    def load_model(self) -> None:
        for element in self.elements:
            z = ElementMap.get_atomic_number_from_element(element)
            file = Path(self.directory, self.model_save_format.format(z))
            logger.info(f"Loading model weights for element ({element}): {file.name}")
            self.models_params[element] = self.atomic_potentials[element].model.load(file)

    def load(self) -> None:
        self.load_scaler()
        self.load_model()

    def _check_scaler_params_exist(self) -> None:
        if None in self.scalers_params.values():
            logger.error(
                f"Scaler parameters are not set yet for all the elements ({self.scalers_params})."
                "Try loading or fitting the scaler first.",
                exception=ValueError,
            )

    # ------------------------------------------------------------------ device tables
    def device_potential(self) -> engine.DevicePotential:
        """Upload (once per parameter set) the SF tables, scaler affine maps and network weights."""
        key = tuple((id(self.models_params[el]), id(self.scalers_params[el])) for el in self.elements)
        if self._device is None or key != self._device_key:
            records = []
            for element in self.elements:
                pot = self.atomic_potentials[element]
                sizes, acts, weights = pot.model.flatten(self.models_params[element], pot.model_input_size)
                affine = pot.scaler.affine(self.scalers_params[element])
                records.append(engine.ElementRecord(element, pot.descriptor.symfunc_records(), affine, sizes, acts, weights))
            self._device = engine.DevicePotential(records, elements=self.elements)
            self._device_key = key
        return self._device

    def _bind(self, structure: Structure) -> engine.Workspace:
        dev = self.device_potential()
        for element in self.elements:  # reference: positions[element] for every potential element (energy.py:54-60)
            if element not in structure.element_map.element_to_atom_type:
                raise KeyError(element)
        ws = dev.workspace(structure.natoms, structure.dtype, engine.number_density(structure))
        ws.bind(structure.positions, engine.remap_types(structure, dev.type_of), engine.box_lengths(structure), dev.r_cutoff)
        return ws

    # ------------------------------------------------------------------ evaluation
    def __call__(self, structure: Structure) -> Array:
        """Total energy (0-d array)."""
        self._check_scaler_params_exist()
        ws = self._bind(structure)
        energy, _, _ = ws.energy_forces(want_energy=True, want_forces=False)
        return energy

    def compute_forces(self, structure: Structure, forces: str = "reference") -> Array:
        """Force components [N, 3].  `forces="reference"` (default): the reference's central-role gradient
        (`force.py:16-43`, does not sum to zero).  `forces="full"` (extension): -dE/dr of the total
        energy including every atom's neighbour role (Newton's third law holds)."""
        if forces not in ("reference", "full"):
            logger.error(f"Unknown force definition '{forces}'", exception=ValueError)
        self._check_scaler_params_exist()
        ws = self._bind(structure)
        _, _, out = ws.energy_forces(want_energy=False, want_forces=True, force_mode=1 if forces == "full" else 0)
        return out

    def compute_energy_and_forces(self, structure: Structure) -> Tuple[Array, Array]:
        """Extension: both results from the same fused launch."""
        self._check_scaler_params_exist()
        ws = self._bind(structure)
        energy, _, forces = ws.energy_forces(want_energy=True, want_forces=True)
        return energy, forces

    # ------------------------------------------------------------------ accessors
    @property
    def descriptors(self) -> Dict[Element, ACSF]:
        return {el: pot.descriptor for el, pot in self.atomic_potentials.items()}

    @property
    def scalers(self) -> Dict[Element, DescriptorScaler]:
        return {el: pot.scaler for el, pot in self.atomic_potentials.items()}

    @property
    def models(self) -> Dict[Element, NeuralNetworkModel]:
        return {el: pot.model for el, pot in self.atomic_potentials.items()}

    @property
    def r_cutoff(self) -> float:
        return max(pot.descriptor.r_cutoff for pot in self.atomic_potentials.values())

    @property
    def num_elements(self) -> int:
        return len(self.elements)

    def __repr__(self) -> str:
        return f"{self.__class__.__name__}(directory={self.directory}, elements={self.elements})"


NNP = NeuralNetworkPotential
