"""Berendsen thermostat (reference `pantea/simulation/thermostat.py:12-66`; class name spelling kept)."""
from __future__ import annotations

import torch

from pantea_b200 import _lib
from pantea_b200.types import Array
from pantea_b200.units import units


class BrendsenThermostat:
    def __init__(self, target_temperature: float, time_constant: float) -> None:
        self.target_temperature = float(target_temperature)
        self.time_constant = float(time_constant)

    def get_rescaled_velocities(self, simulator, system) -> Array:
        """v * 1/sqrt(1 + dt/tau (T/T0 - 1)) with the current temperature, evaluated on the device."""
        from pantea_b200.simulation.system import _get_kinetic_energy

        vel = system.velocities.clone().contiguous()
        ke = _get_kinetic_energy(vel, system.masses).to(torch.float64).reshape(1).contiguous()
        n = vel.shape[0]
        _lib.check(_lib.load().pantea_md_rescale_velocities(
            _lib.ptr(vel), 0, n, _lib.ptr(ke), n, float(simulator.time_step), self.time_constant,
            self.target_temperature, units.BOLTZMANN_CONSTANT, _lib.dtype_code(vel.dtype), _lib.stream_ptr()))
        return vel
