"""Simulation loop with periodic output (reference `pantea/simulation/simulate.py:31-88`).

Between output events the steps are handed to `simulator.simulate_steps` when the simulator has one
(device-resident MD loop), otherwise stepped one by one as in the reference.  Configurations are
appended to `filename` in extended-xyz form without needing ase.
"""
from __future__ import annotations

from pathlib import Path
from typing import Optional

from pantea_b200.logger import logger
from pantea_b200.units import units


def _append_xyz(filename: Path, structure) -> None:
    pos = (structure.positions.detach().cpu().double() * units.TO_ANGSTROM).tolist()
    elements = structure.get_elements()
    header = ""
    if structure.box is not None:
        lat = (structure.box.lattice.detach().cpu().double() * units.TO_ANGSTROM).reshape(-1).tolist()
        header = 'Lattice="' + " ".join(f"{v:.8f}" for v in lat) + '" Properties=species:S:1:pos:R:3 pbc="T T T"'
    with open(str(filename), "a") as fh:
        fh.write(f"{len(elements)}\n{header}\n")
        for el, (x, y, z) in zip(elements, pos):
            fh.write(f"{el:<2} {x:16.8f} {y:16.8f} {z:16.8f}\n")


def simulate(system, simulator, num_steps: int = 1, output_freq: Optional[int] = None,
             filename: Optional[Path] = None, append: bool = False) -> None:
    logger.info(f"Running {simulator.__class__.__name__} for {num_steps} steps")
    if output_freq is None:
        output_freq = 1 if num_steps < 100 else int(0.01 * num_steps)
    is_output = output_freq > 0
    if filename is not None:
        filename = Path(filename)
        if not append:
            open(str(filename), "w").close()
    init_step = simulator.step
    fast = getattr(simulator, "simulate_steps", None)
    try:
        done = 0
        while done < num_steps:
            if is_output and ((simulator.step - init_step) % output_freq == 0):
                print(simulator.repr_physical_params(system))
                if filename is not None:
                    _append_xyz(filename, system.structure)
            if fast is not None:
                span = min(output_freq, num_steps - done) if is_output else num_steps - done
                fast(system, span)
                done += span
            else:
                simulator.simulate_one_step(system)
                done += 1
    except KeyboardInterrupt:
        print("KeyboardInterrupt")
    if is_output:
        print(simulator.repr_physical_params(system))
