"""Generates tests/golden/md10k_3000.json: BASELINE.json configs[2] as written -- 3 000-atom periodic water box, 10 000
steps of the reference integrator (no mass, dt = 0.25 a.u.), NVE and Berendsen NVT (T0 = 300 K, tau = 100 dt; reference
simulation/molecular_dynamics.py:57-77, thermostat.py:54-66) -- potential / kinetic energy sampled every 100 steps from
the C oracle (oracle/hdnnp_oracle.c, pinned to the reference's golden vectors by tests/test_oracle_golden.py), plus one
NVE and one NVT run with a 1e-13 Bohr perturbation of one coordinate: their deviation from the unperturbed curves is the
chaotic spread a correct implementation may show (see make_md10k_curve.py).  ~20 min on 8 cores.

    python tests/golden/make_md10k_3000.py
"""
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))

from oracle import c_oracle  # noqa: E402
from oracle.spec import KB, load_potential, md_velocities, water_box, water_masses  # noqa: E402

GOLDEN = Path(__file__).resolve().parent
N_ATOMS, N_STEPS, DT, EVERY = 3000, 10000, 0.25, 100


def main() -> None:
    pot = load_potential(GOLDEN / "h2o.json")
    pos, types, box = water_box(N_ATOMS)
    vel, mass = md_velocities(types), water_masses(types)
    out = {"source": "oracle/hdnnp_oracle.c orc_md_run, water_box(3000), md_velocities, dt = 0.25 a.u.; NVT: Berendsen "
                     "T0 = 300 K, tau = 25 a.u.", "n_atoms": N_ATOMS, "n_steps": N_STEPS, "dt": DT, "t_target": 300.0,
           "tau": 100 * DT, "steps": list(range(0, N_STEPS + 1, EVERY))}
    for name, tau in (("nve", 0.0), ("nvt", 100 * DT)):
        curves = []
        for eps in (0.0, 1e-13):
            p = pos.copy()
            p[0, 0] += eps
            t0 = time.time()
            _, _, _, sc = c_oracle.md_run(pot, p, vel, mass, types, box, DT, N_STEPS, 300.0, tau, KB)
            print(name, eps, f"{time.time() - t0:.0f} s", flush=True)
            curves.append(sc[::EVERY, :2].copy())
        ref, per = curves
        out[name] = {"e_pot_e_kin": ref.tolist(),
                     "perturbed_1e-13": {"max_abs_log_ratio_e_kin": float(np.abs(np.log(per[1:, 1] / ref[1:, 1])).max()),
                                         "max_abs_dev_e_pot": float(np.abs(per[:, 0] - ref[:, 0]).max()),
                                         "max_rel_dev_first_500_steps": float((np.abs(per[1:6] - ref[1:6]) / np.abs(ref[1:6])).max()),
                                         "abs_log_ratio_e_kin": np.abs(np.log(per[1:, 1] / ref[1:, 1])).tolist(),
                                         "abs_dev_e_pot": np.abs(per[:, 0] - ref[:, 0]).tolist()}}
    (GOLDEN / "md10k_3000.json").write_text(json.dumps(out))
    print("wrote", GOLDEN / "md10k_3000.json")


if __name__ == "__main__":
    main()
