"""Multi-GPU invariance check (run under torchrun on N GPUs; not collected by pytest):

    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tests/mgpu_check.py [atoms] [steps]
    torchrun ... tests/mgpu_check.py [atoms] [steps] halo [skin] [rebuild_every] [full]
    torchrun ... tests/mgpu_check.py [atoms] [steps] brick [oracle]

Runs the same water box with ReplicatedMD on N ranks and, on rank 0, on a single rank; requires identical
neighbour sets implicitly through bitwise identical positions / velocities / forces after `steps` MD steps
(per-atom results do not depend on the ownership split: same kernels, same per-atom arithmetic and order).
With `halo` the N-rank run is the brick-decomposed HaloMD (ghost-atom exchange over NCCL, migration); the neighbour
order inside a row then differs from the single-GPU one, so agreement is required to 1e-10 relative,
not bitwise.  `full` selects PANTEA_FORCE_FULL (reverse halo).
`brick` runs the brick-decomposed engine of csrc/mgpu.cu (ghost positions stored into the peers' mailboxes over NVLink by
the integration kernel, CUDA-graph steps) and requires, against the single-GPU device loop (pantea_md_run) on every
rank: exactly one owner per atom, bitwise identical positions / velocities / forces (the global cell grid keeps the
neighbour order of every row), and the sorted neighbour sets of the final state equal to the single-GPU ones; with
`oracle` also forces <= 1e-10 against the CPU oracle for the final positions.
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


def main_halo(n_atoms, steps, skin, every, full):
    from pantea_b200.halo import HaloMD
    rank, world, local = init_distributed()
    dev = torch.device("cuda", local)
    nnp = NeuralNetworkPotential.from_runner(ROOT / "tests" / "golden" / "h2o.json")
    nnp.load()
    pot = nnp.device_potential()
    pos, types, box = water_box(n_atoms)
    vel, mass = md_velocities(types), water_masses(types)
    t = lambda a, dt=torch.float64: torch.as_tensor(np.ascontiguousarray(a), dtype=dt, device=dev)  # noqa: E731
    args = (pot, t(pos), t(vel), t(mass), t(types, torch.int32), list(box), 0.25)
    md = HaloMD(*args, rank, world, skin=skin, rebuild_every=every, force_mode=1 if full else 0)
    for _ in range(steps):
        md.step()
    md.validate()
    md.check_capacity()
    e_pot, e_kin = float(md.potential_energy()), float(md.kinetic_energy())
    pos_all, vel_all, frc_all = md.gather_owned(md.pos), md.gather_owned(md.vel), md.gather_owned(md.frc)
    counts = torch.tensor([md.n_own, md.n_ghost], dtype=torch.float64, device=dev)
    gathered = [torch.zeros_like(counts) for _ in range(world)]
    if world > 1:
        dist.all_gather(gathered, counts)
    torch.cuda.synchronize()
    if rank == 0:
        ref = HaloMD(*args, 0, 1, force_mode=1 if full else 0)   # one brick, no ghosts: the single-GPU path
        for _ in range(steps):
            ref.step()
        ref.check_capacity()
        rp, rv, rf = ref.gather_owned(ref.pos), ref.gather_owned(ref.vel), ref.gather_owned(ref.frc)
        d = pos_all - rp
        bx = torch.tensor(list(box), dtype=torch.float64, device=dev)
        d -= bx * torch.round(d / bx)
        e_ref, k_ref = float(ref.potential_energy()), float(ref.kinetic_energy())
        err = (float(d.abs().max()), float((vel_all - rv).abs().max() / rv.abs().max()),
               float((frc_all - rf).abs().max() / rf.abs().max()))
        print(f"halo world={world} dims={md.domain.grid.dims} atoms={n_atoms} steps={steps} skin={skin} every={every} "
              f"full={full} owned/ghosts per rank={[tuple(int(x) for x in g.tolist()) for g in gathered]} "
              f"rebuilds={md.rebuilds} rollbacks={md.rollbacks} max|dx|={err[0]:.3e} rel dv={err[1]:.3e} rel dF={err[2]:.3e} "
              f"dEpot={abs(e_pot - e_ref):.3e} dEkin={abs(e_kin - k_ref):.3e}")
        assert err[0] < 1e-9 and err[1] < 1e-9 and err[2] < 1e-8
        assert abs(e_pot - e_ref) < 1e-9 * abs(e_ref) and abs(e_kin - k_ref) < 1e-9 * abs(k_ref)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main_brick(n_atoms, steps, with_oracle):
    import ctypes as C

    from pantea_b200 import _lib, engine
    from pantea_b200.brick import BrickMD
    rank, world, local = init_distributed()
    dev = torch.device("cuda", local)
    lib = _lib.load()
    nnp = NeuralNetworkPotential.from_runner(ROOT / "tests" / "golden" / "h2o.json")
    nnp.load()
    pot = nnp.device_potential()
    pos, types, box = water_box(n_atoms)
    vel, mass = md_velocities(types), water_masses(types)
    t = lambda a, dt=torch.float64: torch.as_tensor(np.ascontiguousarray(a), dtype=dt, device=dev)  # noqa: E731
    p0, v0, ty = t(pos), t(vel), t(types, torch.int32)
    md = BrickMD(pot, p0, v0, ty, list(box), 0.25, rank, world)
    owned0 = md.owned_count()
    md.run(steps)
    md.check_capacity()
    gp, gv, gf, count = md.gather()
    torch.cuda.synchronize()
    # single-GPU device loop on this rank (every rank does the same; small systems)
    n = len(pos)
    ws = engine.Workspace(pot, n, engine.estimate_max_neighbors(pot.r_cutoff, n / float(np.prod(box)), n), torch.float64)
    rp, rv = p0.clone(), v0.clone()
    ws.bind(rp, ty, list(box), pot.r_cutoff)
    _, _, rf = ws.energy_forces(False, True)
    params = _lib.MDParams(0.25, 0.0, 0.0, 3.166811563e-6, 0, 1, 0, 0)
    _lib.check(lib.pantea_md_run(ws.handle, _lib.ptr(rp), _lib.ptr(rv), _lib.ptr(rf), _lib.ptr(t(mass)), _lib.ptr(ty), n,
                                 _lib.box_arg(list(box)), steps, C.byref(params), None, _lib.stream_ptr()))
    _lib.check(lib.pantea_neighbor_status(ws.handle, None, _lib.stream_ptr()))
    one_owner = bool((count == 1).all())
    same = torch.equal(gp, rp) and torch.equal(gv, rv) and torch.equal(gf, rf)
    dmax = [float((a - b).abs().max()) for a, b in ((gp, rp), (gv, rv), (gf, rf))]
    msg = (f"brick world={world} dims={md.grid.dims} atoms={n} steps={steps} owned(rank {rank})={owned0}->{md.owned_count()} "
           f"one_owner={one_owner} bitwise_identical={same} max|d|(x,v,F)={dmax}")
    ok = one_owner and same
    if with_oracle and rank == 0:
        from oracle import c_oracle
        from oracle.spec import load_potential
        specs = load_potential(ROOT / "tests" / "golden" / "h2o.json")
        c_oracle.set_num_threads(16)
        _, _, f_o = c_oracle.energy_forces(specs, gp.cpu().numpy(), types, box)
        f_o = torch.as_tensor(f_o, device=dev)
        rms = float(f_o.pow(2).mean().sqrt())
        over = float(((gf - f_o).abs() / (1e-10 * (f_o.abs() + rms))).max())
        row_ptr, _ = c_oracle.neighbors(gp.cpu().numpy(), types, box, 12.0)
        ws.bind(gp.contiguous(), ty, list(box), pot.r_cutoff)
        rp_gpu, _ = ws.neighbor_lists()
        nb_equal = bool(np.array_equal(rp_gpu.cpu().numpy(), row_ptr))
        msg += f" forces_vs_oracle(err/tol)={over:.3e} neighbour_counts_equal={nb_equal}"
        ok = ok and over <= 1.0 and nb_equal
    print(msg, flush=True)
    flag = torch.tensor([0.0 if ok else 1.0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(flag, op=dist.ReduceOp.MAX)
    md.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if float(flag.item()) != 0.0:
        raise SystemExit(1)


def main():
    n_atoms = int(sys.argv[1]) if len(sys.argv) > 1 else 24000
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    if len(sys.argv) > 3 and sys.argv[3] == "brick":
        return main_brick(n_atoms, steps, len(sys.argv) > 4 and sys.argv[4] == "oracle")
    if len(sys.argv) > 3 and sys.argv[3] == "halo":
        skin = float(sys.argv[4]) if len(sys.argv) > 4 else 0.0
        every = int(sys.argv[5]) if len(sys.argv) > 5 else 1
        return main_halo(n_atoms, steps, skin, every, len(sys.argv) > 6 and sys.argv[6] == "full")
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
