"""Multi-GPU invariance check (run under torchrun on N GPUs; not collected by pytest):

    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tests/mgpu_check.py [atoms] [steps]

Runs the same water box with ReplicatedMD on N ranks and, on rank 0, on a single rank; requires identical
neighbour sets implicitly through bitwise identical positions / velocities / forces after `steps` MD steps
(per-atom results do not depend on the ownership split: same kernels, same per-atom arithmetic and order).
"""
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from pantea_b200.distributed import ReplicatedMD, init_distributed  # noqa: E402
from pantea_b200.potentials import NeuralNetworkPotential  # noqa: E402
from pantea_b200.utils.synthetic import md_velocities, water_box, water_masses  # noqa: E402


def main():
    n_atoms = int(sys.argv[1]) if len(sys.argv) > 1 else 24000
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    rank, world, local = init_distributed()
    dev = torch.device("cuda", local)
    nnp = NeuralNetworkPotential.from_runner(ROOT / "tests" / "golden" / "h2o.json")
    nnp.load()
    pot = nnp.device_potential()
    pos, types, box = water_box(n_atoms)
    vel, mass = md_velocities(types), water_masses(types)
    t = lambda a, dt=torch.float64: torch.as_tensor(np.ascontiguousarray(a), dtype=dt, device=dev)  # noqa: E731

    md = ReplicatedMD(pot, t(pos), t(vel), t(mass), t(types, torch.int32), list(box), 0.25, rank, world)
    for _ in range(steps):
        md.step()
    e_pot = float(md.potential_energy())
    e_kin = float(md.kinetic_energy())
    vel_all = md.gather_owned(md.vel)
    frc_all = md.gather_owned(md.frc)
    torch.cuda.synchronize()
    if rank == 0:
        ref = ReplicatedMD(pot, t(pos), t(vel), t(mass), t(types, torch.int32), list(box), 0.25, 0, 1)
        for _ in range(steps):
            ref.step()
        ok = (torch.equal(ref.pos, md.pos) and torch.equal(ref.vel, vel_all) and torch.equal(ref.frc, frc_all))
        e_ref, k_ref = float(ref.potential_energy()), float(ref.kinetic_energy())
        print(f"world={world} atoms={n_atoms} steps={steps} bitwise_identical={ok} "
              f"dEpot={abs(e_pot - e_ref):.3e} dEkin={abs(e_kin - k_ref):.3e}")
        assert ok and abs(e_pot - e_ref) < 1e-9 * abs(e_ref) and abs(e_kin - k_ref) < 1e-12 * abs(k_ref)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
