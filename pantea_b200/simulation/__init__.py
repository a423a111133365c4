from pantea_b200.simulation.lennard_jones import LJPotential
from pantea_b200.simulation.molecular_dynamics import MDSimulator
from pantea_b200.simulation.monte_carlo import MCSimulator
from pantea_b200.simulation.simulate import simulate
from pantea_b200.simulation.system import System
from pantea_b200.simulation.thermostat import BrendsenThermostat

__all__ = ["System", "MDSimulator", "BrendsenThermostat", "MCSimulator", "simulate", "LJPotential"]
