"""Generates tests/golden/md10k_nve_192.json: potential / kinetic energy of a 10 000-step NVE run of the reference
integrator (no mass, dt = 0.25 a.u.) on the synthetic 192-atom water box, sampled every 250 steps, from the C oracle
(oracle/hdnnp_oracle.c, pinned to the reference's golden vectors by tests/test_oracle_golden.py).  ~1 min on 8 cores.

    python tests/golden/make_md10k_curve.py

North star: "NVE energy drift must match the reference over 10k steps".  The mass-less dynamics is violently chaotic:
a 1e-13 Bohr perturbation of one coordinate changes nothing visible for 500 steps (relative 1e-10), reaches 1e-6 of
the energy scale at step 1 000 and order one by step 2 000 (`spread` below: five such runs).  A trajectory can
therefore be compared point by point only over the first few hundred steps (tests/test_gpu_parity.py does that to 1e-7
over 600 steps); beyond, what two correct implementations share is the statistics of the drift -- E_kin grows from 0.27
to 3e6 Ha -- which the perturbed oracle runs reproduce within |log ratio| <= 0.16 and E_pot within 13 Ha.  The GPU test
uses 0.4 and 30 Ha.
"""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))

from oracle import c_oracle  # noqa: E402
from oracle.spec import KB, load_potential, md_velocities, water_box, water_masses  # noqa: E402

GOLDEN = Path(__file__).resolve().parent
N_ATOMS, N_STEPS, DT, EVERY = 192, 10000, 0.25, 250


def main() -> None:
    pot = load_potential(GOLDEN / "h2o.json")
    pos, types, box = water_box(N_ATOMS)
    vel, mass = md_velocities(types), water_masses(types)
    curves = []
    for k, eps in enumerate((0.0, 1e-13, -1e-13, 3e-13, 1e-12)):
        p = pos.copy()
        p[k, 0] += eps
        _, _, _, sc = c_oracle.md_run(pot, p, vel, mass, types, box, DT, N_STEPS, 300.0, 0.0, KB)
        curves.append(sc[::EVERY, :2].copy())
    curves = np.asarray(curves)
    ref = curves[0]
    log_ratio = np.abs(np.log(curves[1:, 1:, 1] / ref[None, 1:, 1]))
    out = {"source": "oracle/hdnnp_oracle.c orc_md_run, water_box(192), md_velocities, NVE, dt = 0.25 a.u.",
           "n_atoms": N_ATOMS, "n_steps": N_STEPS, "dt": DT,
           "steps": list(range(0, N_STEPS + 1, EVERY)), "e_pot_e_kin": ref.tolist(),
           "spread": {"perturbations_bohr": [1e-13, -1e-13, 3e-13, 1e-12],
                      "max_abs_log_ratio_e_kin": float(log_ratio.max()),
                      "max_abs_dev_e_pot": float(np.abs(curves[1:, :, 0] - ref[None, :, 0]).max()),
                      "max_rel_dev_first_500_steps": float((np.abs(curves[1:, 1:3, :] - ref[None, 1:3, :])
                                                            / np.abs(ref[None, 1:3, :])).max())}}
    (GOLDEN / "md10k_nve_192.json").write_text(json.dumps(out, indent=1))
    print("wrote", GOLDEN / "md10k_nve_192.json", out["spread"])


if __name__ == "__main__":
    main()
