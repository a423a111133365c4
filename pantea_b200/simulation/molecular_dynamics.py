"""Velocity-Verlet MD driver (API of reference `pantea/simulation/molecular_dynamics.py:33-90`).

As in the reference no mass enters the integrator: x += v dt + F dt^2 / 2, v += (F + F') dt / 2
(`molecular_dynamics.py:16-30`), positions are wrapped after the drift, and the Berendsen
thermostat acts after the step with the post-step temperature (`:57-63`).

`simulate_one_step` follows the reference call by call through the C ABI.  `simulate_steps` runs
many steps in a single `pantea_md_run` call -- neighbour build, fused force kernel and integrator
replayed as one CUDA graph per step with no host synchronisation -- when the potential is a
`NeuralNetworkPotential`.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from pantea_b200 import _lib, engine
from pantea_b200.simulation.system import System
from pantea_b200.simulation.thermostat import BrendsenThermostat
from pantea_b200.units import units


class MDSimulator:
    """`mass_scaled` and `forces` are extensions: `mass_scaled=True` integrates with accelerations F/m,
    `forces="full"` uses -dE/dr of the total energy; together they give the usual energy-conserving velocity Verlet.
    The defaults are the reference's integrator (no mass) and force definition (central-role gradient)."""

    def __init__(self, time_step: float, thermostat: Optional[BrendsenThermostat] = None, mass_scaled: bool = False,
                 forces: str = "reference") -> None:
        if forces not in ("reference", "full"):
            raise ValueError(f"Unknown force definition '{forces}'")
        self.time_step: float = float(time_step)
        self.thermostat = thermostat
        self.mass_scaled = bool(mass_scaled)
        self.forces = forces
        self.step: int = 0
        self.elapsed_time: float = 0.0

    def simulate_one_step(self, system: System) -> None:
        self.verlet_integration(system)
        self.step += 1
        self.elapsed_time += float(self.time_step)
        if self.thermostat is not None:
            system.velocities = self.thermostat.get_rescaled_velocities(self, system)

    def _ensure_forces(self, system: System) -> None:
        """`System` fills `structure.forces` with the reference force; the full-force extension needs F(t) in its own
        definition before the first half-kick."""
        if self.forces == "full" and getattr(system, "_forces_kind", "reference") != "full":
            system.structure.forces = system.potential.compute_forces(system.structure, forces="full")
            system._forces_kind = "full"

    def verlet_integration(self, system: System) -> None:
        lib = _lib.load()
        self._ensure_forces(system)
        s = system.structure
        n = s.natoms
        code = _lib.dtype_code(s.dtype)
        pos = s.positions.clone().contiguous()
        vel = system.velocities.clone().contiguous()
        frc = s.forces.clone().contiguous()
        mass = system.masses.reshape(-1).to(s.dtype).contiguous() if self.mass_scaled else None
        _lib.check(lib.pantea_md_update_positions_mass(_lib.ptr(pos), _lib.ptr(vel), _lib.ptr(frc), _lib.ptr(mass), 0, n,
                                                       _lib.box_arg(engine.box_lengths(s)), self.time_step, code,
                                                       _lib.stream_ptr()))
        s.positions = pos
        if self.forces == "full":
            new_forces = system.potential.compute_forces(s, forces="full").contiguous()
        else:
            new_forces = system.potential.compute_forces(s).contiguous()
        _lib.check(lib.pantea_md_update_velocities_mass(_lib.ptr(vel), _lib.ptr(frc), _lib.ptr(new_forces), _lib.ptr(mass),
                                                        0, n, self.time_step, code, _lib.stream_ptr()))
        system.velocities = vel
        s.forces = frc  # == new_forces (the kernel rotates F(t+dt) into place)

    # ------------------------------------------------------------------ device-resident loop
    def simulate_steps(self, system: System, num_steps: int, record: bool = False, use_graph: bool = True):
        """Run `num_steps` steps on the device.  Returns a [num_steps, 2] tensor of (E_pot, E_kin) when `record`."""
        potential = system.potential
        if not hasattr(potential, "device_potential"):
            for _ in range(num_steps):
                self.simulate_one_step(system)
            return None
        if num_steps <= 0:
            return None
        potential._check_scaler_params_exist()
        self._ensure_forces(system)
        s = system.structure
        dev = potential.device_potential()
        ws = dev.workspace(s.natoms, s.dtype, engine.number_density(s))
        types = engine.remap_types(s, dev.type_of).contiguous()
        box = engine.box_lengths(s)
        # capacity check up front; the device loop itself never synchronises, so it runs in chunks with a capacity check
        # between them: a chunk during which a neighbour row, the staged block or a pair list overflowed is repeated
        # from its saved start state after the capacities have been raised (a densifying box needs this on long runs)
        ws.bind(s.positions, types, box, dev.r_cutoff)
        pos = s.positions.clone().contiguous()
        vel = system.velocities.clone().contiguous()
        frc = s.forces.clone().contiguous()
        mass = system.masses.reshape(-1).to(s.dtype).contiguous()
        scalars = torch.zeros((num_steps, 2), dtype=torch.float64, device=pos.device) if record else None
        thermo = self.thermostat
        params = _lib.MDParams(self.time_step, thermo.target_temperature if thermo else 0.0,
                               thermo.time_constant if thermo else 0.0, units.BOLTZMANN_CONSTANT,
                               1 if record else 0, 1 if use_graph else 0, 1 if self.mass_scaled else 0,
                               1 if self.forces == "full" else 0)
        lib = _lib.load()
        chunk = int(self.check_every) if getattr(self, "check_every", 0) else 100
        done = 0
        while done < num_steps:
            todo = min(chunk, num_steps - done)
            saved = (pos.clone(), vel.clone(), frc.clone())
            for attempt in range(6):
                out = scalars[done:done + todo] if record else None
                _lib.check(lib.pantea_md_run(ws.handle, _lib.ptr(pos), _lib.ptr(vel), _lib.ptr(frc), _lib.ptr(mass),
                                             _lib.ptr(types), s.natoms, _lib.box_arg(box), int(todo),
                                             C.byref(params), _lib.ptr(out), _lib.stream_ptr()))
                ws._keep = (pos, types)
                mx = C.c_int32(0)
                code = lib.pantea_neighbor_status(ws.handle, C.byref(mx), _lib.stream_ptr())
                if code != _lib.PANTEA_ECAPACITY:
                    _lib.check(code)
                    break
                if attempt == 5:
                    _lib.check(code)
                if mx.value > (ws.max_neighbors + 31) // 32 * 32:  # a neighbour row overflowed: only a larger workspace helps
                    ws._grow(mx.value)
                pos.copy_(saved[0]); vel.copy_(saved[1]); frc.copy_(saved[2])
            done += todo
        s.positions, system.velocities, s.forces = pos, vel, frc
        self.step += num_steps
        self.elapsed_time += num_steps * float(self.time_step)
        return scalars

    def repr_physical_params(self, system: System) -> str:
        if not system.structure.box:
            return ""
        return (
            f"{self.step:<10} "
            f"time[ps]:{units.TO_PICO_SECOND * self.elapsed_time:<10.5f} "
            f"Temp[K]:{float(system.get_temperature()):<10.5f} "
            f"Etot[Ha]:{float(system.get_total_energy()):<15.10f} "
            f"Epot[Ha]:{float(system.get_potential_energy()):<15.10f} "
            f"Pres[kb]:{float(system.get_pressure()) * units.TO_KILO_BAR:<10.5f}"
        )
