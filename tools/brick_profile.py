"""Per-kernel device times of the brick engine's step (rank 0), plus step time from CUDA events.
usage: [torchrun ...] python tools/brick_profile.py [atoms] [steps]"""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from pantea_b200.brick import BrickMD  # noqa: E402
from pantea_b200.distributed import init_distributed  # noqa: E402
from pantea_b200.potentials import NeuralNetworkPotential  # noqa: E402
from pantea_b200.utils.synthetic import md_velocities, water_box  # noqa: E402

n_atoms = int(sys.argv[1]) if len(sys.argv) > 1 else 99999
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
rank, world, local = init_distributed()
dev = torch.device("cuda", local)
nnp = NeuralNetworkPotential.from_runner(ROOT / "tests" / "golden" / "h2o.json")
nnp.load()
pot = nnp.device_potential()
pos, types, box = water_box(n_atoms)
vel = md_velocities(types)
t = lambda a, dt=torch.float64: torch.as_tensor(np.ascontiguousarray(a), dtype=dt, device=dev)  # noqa: E731
md = BrickMD(pot, t(pos), t(vel), t(types, torch.int32), list(box), 0.25, rank, world)
md.run(25)
md.check_capacity()
md.reset(t(pos), t(vel))
md.run(3)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
md.run(steps)
b.record()
torch.cuda.synchronize()
if rank == 0:
    print(f"world={world} atoms={len(pos)} {a.elapsed_time(b) / steps:.4f} ms/step over {steps} graph steps (no L2 flush), owned {md.owned_count()}")
from torch.profiler import ProfilerActivity, profile
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    md.run(5)
    torch.cuda.synchronize()
if rank == 0:
    tot = 0.0
    for ev in sorted(prof.key_averages(), key=lambda e: -e.device_time_total)[:24]:
        tot += ev.device_time_total / 5 / 1e3
        print(f"{ev.key[:80]:80s} n={ev.count:3d} per-step={ev.device_time_total / 5 / 1e3:8.4f} ms")
    print(f"sum of kernel times per step: {tot:.4f} ms")
md.check_capacity()
md.close()
if world > 1:
    import torch.distributed as dist
    dist.barrier()
    dist.destroy_process_group()
