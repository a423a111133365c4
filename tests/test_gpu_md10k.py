"""BASELINE.json configs[2] as written: 3 000-atom periodic water box, 10 000 MD steps, NVE and Berendsen NVT
(reference simulation/molecular_dynamics.py:57-77, thermostat.py:54-66), through the public driver
`MDSimulator.simulate_steps` (device-resident CUDA-graph loop, run in chunks with capacity checks: the mass-less
reference dynamics densifies the box, so neighbour rows / pair lists outgrow their first capacities along the run).

Oracle: tests/golden/md10k_3000.json (tests/golden/make_md10k_3000.py: C oracle curves sampled every 100 steps, plus a
run with one coordinate perturbed by 1e-13 Bohr).  The dynamics is chaotic: point-wise agreement exists over the first
few hundred steps; afterwards two correct implementations share the statistics of the curve, and the perturbed oracle
run says how far they may be apart.  Required: <= 1e-6 of the energy scale up to step 500, then within 4x the oracle's
own perturbed-run deviation (floors: 0.1 in log E_kin, 2 % of the E_pot scale)."""
import json

import numpy as np
import pytest
import torch

from oracle.spec import load_potential, md_velocities, water_box, water_masses
from tests.helpers import cuda

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("ensemble", ["nve", "nvt"])
def test_md_3000_atoms_10k_steps_follows_the_oracle_curve(ensemble, golden_dir):
    path = golden_dir / "md10k_3000.json"
    if not path.exists():
        pytest.skip("tests/golden/md10k_3000.json not generated (python tests/golden/make_md10k_3000.py, ~50 min of CPU)")
    fx = json.loads(path.read_text())
    from pantea_b200.atoms import Structure
    from pantea_b200.potentials import NeuralNetworkPotential
    from pantea_b200.simulation import BrendsenThermostat, MDSimulator, System

    n_atoms, n_steps, dt = fx["n_atoms"], fx["n_steps"], fx["dt"]
    steps = np.asarray(fx["steps"])
    ref = np.asarray(fx[ensemble]["e_pot_e_kin"])
    spread = fx[ensemble]["perturbed_1e-13"]
    pos, types, box = water_box(n_atoms)
    vel, mass = md_velocities(types), water_masses(types)
    nnp = NeuralNetworkPotential.from_runner(golden_dir / "h2o.json")
    nnp.load()
    s = Structure.from_dict({"positions": pos, "elements": ["H" if t == 1 else "O" for t in types], "lattice": np.diag(box)})
    system = System(nnp, s, velocities=cuda(vel), masses=cuda(mass).reshape(-1, 1))
    thermo = BrendsenThermostat(target_temperature=fx["t_target"], time_constant=fx["tau"]) if ensemble == "nvt" else None
    sim = MDSimulator(time_step=dt, thermostat=thermo)
    sim.check_every = 100
    scal = sim.simulate_steps(system, n_steps, record=True).cpu().numpy()     # [n_steps, 2]: (E_pot, E_kin) after each step
    assert np.isfinite(scal).all()
    scale_pot, scale_kin = np.abs(ref[:6, 0]).max(), ref[:6, 1].max()
    log_band = max(4.0 * spread["max_abs_log_ratio_e_kin"], 0.1)
    pot_band = max(4.0 * spread["max_abs_dev_e_pot"], 0.02 * scale_pot)
    worst = [0.0, 0.0]
    for k, (e_pot, e_kin) in zip(steps[1:], ref[1:]):
        g_pot, g_kin = scal[k - 1]
        if k <= 500:
            assert abs(g_pot - e_pot) < 1e-6 * scale_pot and abs(g_kin - e_kin) < 1e-6 * scale_kin, (int(k), g_pot, e_pot, g_kin, e_kin)
        else:
            worst = [max(worst[0], abs(np.log(g_kin / e_kin))), max(worst[1], abs(g_pot - e_pot))]
            assert abs(np.log(g_kin / e_kin)) < log_band and abs(g_pot - e_pot) < pot_band, (int(k), g_pot, e_pot, g_kin, e_kin, log_band, pot_band)
    print(f"{ensemble}: worst |log E_kin ratio| {worst[0]:.3e} (band {log_band:.3e}), worst |dE_pot| {worst[1]:.3e} (band {pot_band:.3e})")
