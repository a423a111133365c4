#!/usr/bin/env python
"""Headline benchmark: HDNNP water MD throughput (atom-steps/s) on N B200s of one node.

    python bench.py --gpus 1 --steps 20 --warmup 3                 # this implementation
    python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 # CPU arm (oracle port on host cores)
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[3]): 100 000-atom periodic water box, potential tests/golden/h2o.json
(7 symmetry functions, rc = 12 Bohr, MLPs 3-5-5-1 / 4-5-5-1), NVE velocity Verlet exactly as the
reference (no mass in the integrator), dt = 0.25 a.u.; strong scaling over the GPUs.
A "step" is one MD step: position update + wrap, neighbour build, fused symmetry-function / MLP /
force kernel, velocity update.  One JSON line is printed by rank 0 (contract in the task statement).

Timing: W >= 3 warm-up steps, then K steps each bracketed by CUDA events on the launching stream,
with an L2 flush (256 MB write) between the timed steps; sum of the K event intervals, max over ranks.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
GOLDEN = ROOT / "tests" / "golden"
METRIC = "HDNNP MD atom-steps/sec (force evals)"
UNIT = "atom-steps/s"
DT = 0.25
SEGMENT = 25  # timed steps replay 25-step pieces of the trajectory from the initial state (see run_b200)
# algorithmic flop weights per unit (SURVEY.md 8d / DESIGN.md): pair, radial-SF (G2), triplet-SF (G3), per-atom rest
FLOP_PAIR, FLOP_RAD, FLOP_TRIP, FLOP_INTEGRATE = 74.0, 39.0, 160.0, 40.0
FLOP_MLP = {1: 540.0, 2: 580.0}  # H, O networks of h2o.json (forward + input gradient)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--atoms", type=int, default=int(os.environ.get("PANTEA_BENCH_ATOMS", "100000")))
    ap.add_argument("--dtype", default="f64", choices=["f64", "f32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="diagnostics only: skip the oracle comparison before timing")
    ap.add_argument("--no-extra", action="store_true",
                    help="skip the secondary workloads (configs[1] preprocessing batch, configs[4] 10^6-atom box, FP32 mode)")
    ap.add_argument("--engine", default="brick", choices=["brick", "replicated"],
                    help="brick (default): spatial bricks, ghost positions pushed into the peers' mailboxes over NVLink by "
                         "the integration kernel, CUDA-graph steps (csrc/mgpu.cu); replicated: round-1 scheme, all "
                         "positions all-gathered with NCCL every step")
    ap.add_argument("--skin", type=float, default=0.0,
                    help="Verlet skin (Bohr) of the secondary measurement `verlet_skin`; 0 skips it")
    ap.add_argument("--kernel-times", action="store_true",
                    help="after the timed region, print a per-kernel breakdown (CUPTI, diagnostic only) to stderr")
    ap.add_argument("--no-flush", action="store_true", help="diagnostics only: skip the L2 flush between steps")
    ap.add_argument("--halo", default="off", choices=["auto", "on", "off"],
                    help="secondary measurement `halo_exchange`: the same K steps with the brick-decomposed HaloMD "
                         "(ghost-atom exchange); auto = when N > 1 (default off: opt-in)")
    ap.add_argument("--halo-skin", type=float, default=4.4, help="ghost-shell skin (Bohr) of the halo measurement")
    ap.add_argument("--halo-every", type=int, default=8, help="ghost lists rebuilt every this many steps")
    args = ap.parse_args()
    args.atoms = 3 * (args.atoms // 3)  # whole water molecules: "100 000 atoms" = 33 333 molecules = 99 999 atoms
    return args


def workload_name(n_atoms: int) -> str:
    return (f"{n_atoms}-atom periodic water box NVE MD, tests/golden/h2o.json HDNNP (rc=12 Bohr), dt=0.25 a.u. "
            "[BASELINE.json configs[3]]")


# ----------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index: int) -> None:
        self.proc = None
        self.index = device_index
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(device_index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        sm, mx, power, reasons = [], [], [], set()
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def host_cpus() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def pin_oracle_threads(c_oracle) -> int:
    """The CPU arm always uses every host core it may run on (or PANTEA_REF_THREADS): torchrun exports
    OMP_NUM_THREADS=1 to its workers, which would otherwise decide the CPU number at N > 1."""
    want = int(os.environ.get("PANTEA_REF_THREADS", "0")) or host_cpus()
    c_oracle.set_num_threads(want)
    return c_oracle.num_threads()


def parity_check(md, lib, _lib, pos_h, types_h, box_h, dtype, rank, world, dev):
    """Forces and neighbour counts of the initial state, all atoms, against the CPU oracle (rank 0 compares; the owned
    rows of every rank are gathered first).  Criterion per force component: |dF| <= rtol * (|F_oracle| + rms(F_oracle)),
    rtol = 1e-10 (FP64) / 1e-5 (FP32) -- element-wise with an absolute floor at rtol * rms, because single components
    cancel to ~1e-8 of the terms they are summed from.  The pure element-wise maximum is reported beside it."""
    import torch
    import torch.distributed as dist

    n = len(pos_h)
    own = torch.zeros(n, dtype=torch.int32, device=dev)
    _lib.check(lib.pantea_neighbor_counts(md.ws.handle, _lib.ptr(own), _lib.stream_ptr()))
    one_owner = True
    if hasattr(md, "gather"):  # brick engine: owned rows by role
        _, _, frc, owners = md.gather()
        frc = frc.double()
        one_owner = bool((owners == 1).all())
        counts = torch.where(md.read()[3] == 2, own, torch.zeros_like(own))
    else:
        frc = md.gather_owned(md.frc).double()
        counts = torch.zeros(n, dtype=torch.int32, device=dev)
        counts[md.lo:md.hi] = own[md.lo:md.hi]
    if world > 1:
        dist.all_reduce(counts, op=dist.ReduceOp.SUM)
    out = None
    if rank == 0:
        from oracle import c_oracle  # the checker
        from oracle.spec import load_potential

        threads = pin_oracle_threads(c_oracle)
        specs = load_potential(GOLDEN / "h2o.json")
        t0 = time.perf_counter()
        _, _, f_o = c_oracle.energy_forces(specs, pos_h, types_h, box_h)
        # (the oracle's neighbour export is serial: counts are compared up to 3 x 10^5 atoms, forces always)
        row_ptr = c_oracle.neighbors(pos_h, types_h, box_h, 12.0)[0] if n <= 300000 else None
        cpu_s = time.perf_counter() - t0
        f_o = torch.as_tensor(f_o, device=dev)
        rtol = 1e-10 if dtype == torch.float64 else 1e-5
        rms = float(f_o.pow(2).mean().sqrt())
        err = (frc - f_o).abs()
        over = float((err / (rtol * (f_o.abs() + rms))).max())
        rel = float((err / f_o.abs().clamp_min(1e-300)).max())
        n_equal = bool((counts.cpu().numpy() == np.diff(row_ptr).astype(np.int32)).all()) if row_ptr is not None else None
        out = {"n_atoms": n, "max_err_over_tol": over, "rtol": rtol, "criterion": "|dF| <= rtol*(|F|+rms(F)) per component",
               "max_rel_F": rel, "max_abs_dF": float(err.max()), "rms_F": rms, "neighbors_equal": n_equal,
               "one_owner_per_atom": one_owner,
               "oracle": f"all {n} atoms, oracle/hdnnp_oracle.c, {threads} threads, {cpu_s:.1f} s",
               "ok": bool(over <= 1.0 and n_equal is not False and one_owner)}
    flag = torch.tensor([0.0 if (out is None or out["ok"]) else 1.0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(flag, op=dist.ReduceOp.MAX)
    if float(flag.item()) != 0.0:
        if rank == 0:
            print(json.dumps({"parity": out}), file=sys.stderr)
        raise SystemExit("bench.py: GPU forces / neighbour counts differ from the oracle; no throughput number is reported")
    return out


# ----------------------------------------------------------------------------------------------- secondary workloads
def run_preprocess(pot, rank, world, dev, n_structs=10000, atoms=192, batch=1250):
    """BASELINE.json configs[1]: ACSF values + gradients of 10 000 synthetic 192-atom water structures, structure s on
    rank s mod world (reference acsf.py:47-120), then the scaler statistics of every element over the whole set
    (trainer.py:68-88, scaler.py:249-283): per-batch two-pass statistics kernel, merged across batches and ranks.
    Parity: eight random structures against the C oracle (values and gradients, 1e-10), statistics against torch."""
    import torch

    from pantea_b200 import engine
    from pantea_b200.descriptors.scaler import DescriptorScaler
    from pantea_b200.distributed import all_reduce_max, merge_scaler_params
    from pantea_b200.utils.synthetic import water_box

    mine = list(range(rank, n_structs, world))
    ws = engine.Workspace(pot, batch * atoms, atoms - 1, torch.float64)
    batches = []
    for b0 in range(0, len(mine), batch):
        ids = mine[b0:b0 + batch]
        structs = [water_box(atoms, seed=2024 + s) for s in ids]
        pos = torch.as_tensor(np.concatenate([x[0] for x in structs]), device=dev)
        types = torch.as_tensor(np.concatenate([x[1] for x in structs]), dtype=torch.int32, device=dev)
        boxes = torch.as_tensor(np.stack([x[2] for x in structs]), device=dev)
        ptr = torch.arange(len(ids) + 1, dtype=torch.int32, device=dev) * atoms
        idx = {el: torch.nonzero(types == pot.type_of[el]).flatten().to(torch.int32) for el in ("H", "O")}
        batches.append((ids, structs, pos, types, boxes, ptr, idx))

    def process(keep=False):
        params = {"H": None, "O": None}
        kept = []
        for ids, structs, pos, types, boxes, ptr, idx in batches:
            ws.bind_batch(pos, types, ptr, boxes, pot.r_cutoff, check=False)
            res = {}
            for el in ("H", "O"):
                G, dG = ws.acsf(pot.slot(el), pot.n_symfunc[el], idx[el], True, True)
                params[el] = DescriptorScaler.fit(G) if params[el] is None else DescriptorScaler.partial_fit(params[el], G)
                res[el] = (G, dG)
            if keep:
                kept.append(res)
        return params, kept

    process()            # warm-up: capacities, lazy allocations, GPU clocks back up after the CPU legs ...
    warm, _ = process()
    for el in ("H", "O"):
        merge_scaler_params(warm[el])  # ... and the collectives' first-use set-up
    torch.cuda.synchronize()
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    # three timed passes (barrier before each, max over ranks of each, best of the three: a pass whose outputs are all
    # retained grows the allocator pool inside the timed region); the outputs the parity check reads come from one more
    # untimed pass
    best, passes = None, []
    for _ in range(3):
        if world > 1:
            dist.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        params, _ = process()
        merged = {el: merge_scaler_params(params[el]) for el in ("H", "O")}
        b.record()
        torch.cuda.synchronize()
        ms = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device=dev)
        all_reduce_max(ms)
        passes.append(round(float(ms.item()), 3))
        best = ms if best is None else torch.minimum(best, ms)
    ms = best
    _, kept = process(keep=True)
    ok, checked = True, 0
    if rank == 0:
        from oracle import c_oracle
        from oracle.spec import load_potential
        specs = load_potential(GOLDEN / "h2o.json")
        rng = np.random.default_rng(11)
        for _ in range(8):
            bi = int(rng.integers(len(batches)))
            si = int(rng.integers(len(batches[bi][0])))
            p0, t0, b0 = batches[bi][1][si]
            for spec, el in zip(specs, ("H", "O")):
                centres = np.nonzero(t0 == spec.atom_type)[0]
                G_o, dG_o = c_oracle.acsf(spec, p0, t0, b0, centres)
                k = len(centres)
                G, dG = kept[bi][el]
                sl = slice(si * k, (si + 1) * k)
                ok &= bool(np.abs(G[sl].cpu().numpy() - G_o).max() < 1e-10 * np.abs(G_o).max())
                ok &= bool(np.abs(dG[sl].cpu().numpy() - dG_o).max() < 1e-10 * np.abs(dG_o).max())
            checked += 1
        if world == 1:  # the merged statistics against torch reductions over all descriptor rows
            for el in ("H", "O"):
                allG = torch.cat([r[el][0] for r in kept])
                ok &= bool(torch.allclose(merged[el].mean, allG.mean(0), rtol=1e-10, atol=1e-14))
                ok &= bool(torch.allclose(merged[el].sigma, allG.std(0, unbiased=False), rtol=1e-8, atol=1e-14))
                ok &= bool(torch.equal(merged[el].minval, allG.min(0).values) and torch.equal(merged[el].maxval, allG.max(0).values))
    t_s = float(ms.item()) * 1e-3
    return {"workload": f"{n_structs} x {atoms}-atom water structures, ACSF values + gradients + scaler statistics "
                        "[BASELINE.json configs[1]]", "value": n_structs / t_s, "unit": "structures/s",
            "atoms_per_s": n_structs * atoms / t_s, "ms_total": float(ms.item()), "timing": "best of 3 passes, each the max over ranks", "ms_passes": passes,
            "split": f"structure s on rank s mod {world}",
            "scaler_samples": {el: int(merged[el].nsamples) for el in ("H", "O")},
            "parity": {"structures_checked": checked, "ok": ok} if rank == 0 else None}


def run_million(pot, rank, world, dev, dtype, steps=20):
    """BASELINE.json configs[4]: 10^6-atom water box on `world` GPUs (strong scaling) and 125 000 atoms per GPU (weak
    scaling), brick engine; forces of the initial state checked against the oracle on a sample of 20 000 atoms."""
    import torch

    from pantea_b200.brick import BrickMD
    from pantea_b200.distributed import all_reduce_max
    from pantea_b200.utils.synthetic import md_velocities, water_box

    out = {}
    for name, n_atoms in (("strong_1e6", 999999), ("weak_125k_per_gpu", 3 * (125000 * world // 3))):
        pos_h, types_h, box_h = water_box(n_atoms)
        vel_h = md_velocities(types_h)
        t = lambda a, dt=dtype: torch.as_tensor(np.ascontiguousarray(a), dtype=dt, device=dev)  # noqa: E731
        p0, v0 = t(pos_h), t(vel_h)
        md = BrickMD(pot, p0, v0, t(types_h, torch.int32), [float(b) for b in box_h], DT, rank, world)
        _, _, frc, owners = md.gather()
        parity = None
        if rank == 0:
            from oracle import c_oracle
            from oracle.spec import load_potential
            specs = load_potential(GOLDEN / "h2o.json")
            pin_oracle_threads(c_oracle)
            m = min(n_atoms, 20000)
            _, _, f_o = c_oracle.energy_forces(specs, pos_h, types_h, box_h, begin=0, end=m)
            f_o = torch.as_tensor(f_o[:m], device=dev)
            rtol = 1e-10 if dtype == torch.float64 else 1e-5
            rms = float(f_o.pow(2).mean().sqrt())
            over = float(((frc[:m].double() - f_o).abs() / (rtol * (f_o.abs() + rms))).max())
            parity = {"atoms_checked": m, "max_err_over_tol": over, "rtol": rtol,
                      "one_owner_per_atom": bool((owners == 1).all()), "ok": bool(over <= 1.0 and (owners == 1).all())}
        for attempt in range(4):  # settle the capacities on one 25-step segment
            md.run(SEGMENT)
            try:
                md.check_capacity()
                ok = 0.0
            except Exception:
                ok = 1.0
            flag = torch.tensor([ok], dtype=torch.float64, device=dev)
            all_reduce_max(flag)
            md.reset(p0, v0)
            if float(flag.item()) == 0.0:
                break
        md.run(3)
        torch.cuda.synchronize()
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        md.run(steps)
        b.record()
        torch.cuda.synchronize()
        md.check_capacity()
        ms = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device=dev)
        all_reduce_max(ms)
        out[name] = {"atoms": n_atoms, "value": n_atoms * steps / (float(ms.item()) * 1e-3), "unit": UNIT,
                     "ms_per_step": float(ms.item()) / steps, "steps": steps, "bricks": list(md.grid.dims),
                     "l2": "inputs larger than L2 (pair lists > 1 GB per step), no flush", "parity": parity}
        md.close()
        del md
        torch.cuda.empty_cache()
    return out


# ----------------------------------------------------------------------------------------------- reference arm
def run_reference(args) -> None:
    """CPU arm.  The reference's own JAX implementation cannot run here (jax/flax/ase are not installed and there
    is no network; its dense O(N^3) algorithm could not hold a 100k-atom box anyway), so the oracle port
    (oracle/hdnnp_oracle.c: same semantics, neighbour lists + analytic gradients, OpenMP) is timed on all host
    cores on a bounded sample of the same workload: forces for the first M atoms of the same box per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # the real thing first: the unmodified reference installed (--no-deps) under baseline/_ref
    ref_status = "baseline/_ref not present"
    ref_dir = ROOT / "baseline" / "_ref"
    if ref_dir.exists():
        sys.path.insert(0, str(ref_dir))
        try:
            from pantea.potentials import NeuralNetworkPotential as _RefNNP  # noqa: F401
            ref_status = "importable"
        except Exception as exc:  # ModuleNotFoundError: jax (not installable: no network, no wheel)
            ref_status = f"pantea.potentials not importable: {type(exc).__name__}: {exc}"
        finally:
            sys.path.remove(str(ref_dir))
            for mod in [m for m in sys.modules if m == "pantea" or m.startswith("pantea.")]:
                del sys.modules[mod]
    from oracle import c_oracle
    from oracle.spec import load_potential
    from pantea_b200.utils.synthetic import water_box

    specs = load_potential(GOLDEN / "h2o.json")
    pos, types, box = water_box(args.atoms)
    n = len(pos)
    threads = pin_oracle_threads(c_oracle)
    # size the per-step sample for ~3 s of CPU work
    t0 = time.perf_counter()
    c_oracle.energy_forces(specs, pos, types, box, begin=0, end=min(n, 512))
    rate = min(n, 512) / (time.perf_counter() - t0)
    m = int(max(256, min(n, rate * 3.0)))
    for _ in range(args.warmup):
        c_oracle.energy_forces(specs, pos, types, box, begin=0, end=min(m, 2048))
    t0 = time.perf_counter()
    for k in range(args.steps):
        lo = (k * m) % max(n - m, 1)
        c_oracle.energy_forces(specs, pos, types, box, begin=lo, end=lo + m)
    elapsed = time.perf_counter() - t0
    value = m * args.steps / elapsed
    sample = f"forces+energies of {m} of the {n} atoms per step (cell-list neighbour gather over the full box)"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * elapsed / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(n), "note": "reference JAX path not runnable here; oracle port timed",
                   "reference_install": ref_status},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "host_cpus": host_cpus(), "kind": "port",
                         "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------- GPU arm
def run_b200(args) -> None:
    import torch

    from pantea_b200 import _lib, engine
    from pantea_b200.distributed import ReplicatedMD, all_reduce_max, init_distributed
    from pantea_b200.potentials import NeuralNetworkPotential
    from pantea_b200.utils.synthetic import md_velocities, water_box, water_masses
    import torch.distributed as dist

    rank, world, local = init_distributed()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a GPU (no CPU fallback on the product path)")
    dev = torch.device("cuda", local)
    lib = _lib.load()
    dtype = torch.float64 if args.dtype == "f64" else torch.float32
    code = _lib.dtype_code(dtype)

    nnp = NeuralNetworkPotential.from_runner(GOLDEN / "h2o.json")
    nnp.load()
    pot = nnp.device_potential()
    pos_h, types_h, box_h = water_box(args.atoms)
    vel_h, mass_h = md_velocities(types_h), water_masses(types_h)
    n = len(pos_h)
    box = [float(b) for b in box_h]
    t = lambda a, dt=dtype: torch.as_tensor(np.ascontiguousarray(a), dtype=dt, device=dev)  # noqa: E731
    brick = args.engine == "brick"
    if brick:
        from pantea_b200.brick import BrickMD
        md = BrickMD(pot, t(pos_h), t(vel_h), t(types_h, torch.int32), box, DT, rank, world)
    else:
        md = ReplicatedMD(pot, t(pos_h), t(vel_h), t(mass_h), t(types_h, torch.int32), box, DT, rank, world)
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def flush():
        if not args.no_flush:
            _lib.check(lib.pantea_l2_flush(_lib.ptr(flush_buf), flush_buf.numel(), _lib.stream_ptr()))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up, then K timed steps ---------------------------------------------------------------
    # The reference integrator carries no mass, so the synthetic box densifies within ~100 steps
    # and the per-atom work drifts.  To keep the workload stationary for any K, the timed steps replay SEGMENT-step
    # pieces of the trajectory from the initial state (the reset is outside the timed events).  The untimed warm-up
    # runs one full segment, which also sizes the pair-list / shared-memory capacities for everything that follows.
    pos0, vel0 = t(pos_h), t(vel_h)
    max_seen = 0

    def capacity_ok(m=None) -> bool:
        nonlocal max_seen
        bad = 0.0
        try:
            max_seen = max(max_seen, (m or md).check_capacity())
        except _lib.CapacityError:
            bad = 1.0
        flag = torch.tensor([bad], dtype=torch.float64, device=dev)
        all_reduce_max(flag)
        return float(flag.item()) == 0.0

    warmup = max(args.warmup, 3)  # timing rule: at least 3 warm-up steps
    warmup = max(warmup, SEGMENT)
    launches_per_reset = 0

    def settle_capacities(m=None) -> None:
        """Untimed: run one full segment until no capacity flag is raised (the library grows its buffers), then reset."""
        nonlocal launches_per_reset
        m = m or md
        for attempt in range(6):
            for _ in range(warmup):
                m.step()
            ok = capacity_ok(m)  # on overflow the library has raised the capacity: run the segment again
            l0 = lib.pantea_launch_count()
            m.reset(pos0, vel0)
            launches_per_reset = lib.pantea_launch_count() - l0
            if ok:
                return
        raise SystemExit("bench.py: capacities did not settle")

    def timed_steps(k_steps: int, m=None):
        """K steps, each bracketed by CUDA events on the launching stream (L2 flushed before each); returns the summed
        step time (ms, max over ranks), the wall time and the kernels launched by the steps themselves."""
        m = m or md
        barrier()
        launches0 = lib.pantea_launch_count()
        starts = [torch.cuda.Event(enable_timing=True) for _ in range(k_steps)]
        stops = [torch.cuda.Event(enable_timing=True) for _ in range(k_steps)]
        extra_launches = 0
        wall0 = time.perf_counter()
        for k in range(k_steps):
            if k > 0 and k % SEGMENT == 0:
                m.reset(pos0, vel0)
                extra_launches += launches_per_reset
            flush()
            extra_launches += 0 if args.no_flush else 1
            starts[k].record()
            m.step()
            stops[k].record()
        barrier()
        wall_s = time.perf_counter() - wall0
        n_launch = lib.pantea_launch_count() - launches0 - extra_launches
        ms = sum(s.elapsed_time(e) for s, e in zip(starts, stops))
        ms_t = torch.tensor([ms], dtype=torch.float64, device=dev)
        all_reduce_max(ms_t)
        return float(ms_t.item()), wall_s, n_launch

    parity = None if args.no_parity else parity_check(md, lib, _lib, pos_h, types_h, box_h, dtype, rank, world, dev)
    sampler = ClockSampler(torch.cuda.current_device()) if rank == 0 else None  # before the warm-up: the timed region
    settle_capacities()                                                        # alone is shorter than a sampling period
    barrier()
    ms_total, wall, launches = timed_steps(args.steps)
    clocks = sampler.stop() if sampler else None
    value = n * args.steps / (ms_total * 1e-3)
    if not capacity_ok():
        raise SystemExit("bench.py: a capacity flag was raised inside the timed region; the measurement is void")
    max_seen_main = max_seen
    fp32_info = None
    if brick and dtype == torch.float64 and not args.no_extra:
        # secondary measurement: the same K steps in the mixed mode (symmetry functions in single precision on the same
        # double-precision state; pantea_workspace_set_compute_precision) -- the north star's "FP32 mode"
        md.ws.set_compute_precision(32)
        md.reset(pos0, vel0)
        par32 = None if args.no_parity else parity_check(md, lib, _lib, pos_h, types_h, box_h, torch.float32, rank, world, dev)
        settle_capacities()
        ms32, _, _ = timed_steps(args.steps)
        if not capacity_ok():
            raise SystemExit("bench.py: a capacity flag was raised inside the mixed-precision timed region")
        fp32_info = {"value": n * args.steps / (ms32 * 1e-3), "unit": UNIT, "ms_per_step": ms32 / args.steps,
                     "speedup_over_f64": ms_total / ms32,
                     "what": "same K steps, symmetry functions and their gradients in FP32 (difference vectors formed in "
                             "FP64, state / scaler / networks / integrator FP64); tolerance of the mode 1e-5",
                     "parity": par32}
        md.ws.set_compute_precision(64)
        md.reset(pos0, vel0)
    engine_info = None
    if brick:
        own = torch.tensor([float(md.owned_count())], dtype=torch.float64, device=dev)
        own_max = own.clone()
        all_reduce_max(own_max)
        engine_info = {"bricks": list(md.grid.dims), "owned_atoms_max_over_ranks": int(own_max.item()),
                       "own_cap": int(md.own_cap), "halo": "ghost positions stored by the integration kernel into the "
                       "peers' mailboxes (CUDA IPC over NVLink), step-number flags, no collective call in the step"}
        brick_md = md
        # the diagnostics below (roofline counters, e2e through pantea_neighbor_build / pantea_energy_forces with host
        # buffers) use a plain index-range-owned workspace of the same size
        md = ReplicatedMD(pot, t(pos_h), t(vel_h), t(mass_h), t(types_h, torch.int32), box, DT, rank, world)
        for _ in range(2):
            md.step()
        capacity_ok(md)

    # ---- secondary measurement: the same K steps with Verlet-skin reuse of rows and pair lists ------------------
    # (the headline above rebuilds the neighbour rows every step, like the reference; this one is reported beside it)
    skin_info = None
    if args.skin > 0.0:
        md.ws.set_skin(args.skin)
        md.reset(pos0, vel0)
        settle_capacities()
        b0, r0 = md.ws.rebuild_counts()
        ms_skin, _, _ = timed_steps(args.steps)
        b1, r1 = md.ws.rebuild_counts()
        if not capacity_ok():
            raise SystemExit("bench.py: a capacity flag was raised inside the Verlet-skin timed region")
        skin_info = {"skin_bohr": args.skin, "value": n * args.steps / (ms_skin * 1e-3), "unit": UNIT,
                     "ms_per_step": ms_skin / args.steps, "neighbor_builds": b1 - b0, "row_rebuilds": r1 - r0,
                     "what": "same K steps; rows gathered with rc + skin, rebuilt (device-side decision) when an atom "
                             "moved > skin/2; identical energies/forces up to summation order"}
        md.ws.set_skin(0.0)
        md.reset(pos0, vel0)
        settle_capacities()

    # ---- secondary measurement: brick decomposition with ghost-atom halo exchange -----------------
    # same K steps, same segments; the headline stays the replicated-coordinates scheme until the halo path wins
    halo_info = None
    if args.halo == "on" or (args.halo == "auto" and world > 1):
        from pantea_b200.halo import HaloMD
        pos_full = t(pos_h)  # HaloMD takes the full initial arrays on every rank and keeps its brick
        vel_full = t(vel_h)
        hmd = HaloMD(pot, pos_full, vel_full, t(mass_h), t(types_h, torch.int32), box, DT, rank, world,
                     skin=args.halo_skin if args.halo_every > 1 else 0.0, rebuild_every=args.halo_every)
        pos0_keep, vel0_keep = pos0, vel0
        pos0, vel0 = pos_full, vel_full
        settle_capacities(hmd)
        hmd.rebuild_events = []
        rb0, rl0 = hmd.rebuilds, hmd.rollbacks
        ms_halo, _, halo_launches = timed_steps(args.steps, hmd)
        hmd.validate()
        torch.cuda.synchronize()
        rebuild_ms = [a.elapsed_time(b) for a, b in hmd.rebuild_events]
        if not capacity_ok(hmd):
            raise SystemExit("bench.py: a capacity flag was raised inside the halo-exchange timed region")
        shares = torch.tensor([float(hmd.n_own), float(hmd.n_ghost)], dtype=torch.float64, device=dev)
        all_reduce_max(shares)
        halo_info = {"value": n * args.steps / (ms_halo * 1e-3), "unit": UNIT, "ms_per_step": ms_halo / args.steps,
                     "bricks": list(hmd.domain.grid.dims), "skin_bohr": hmd.skin, "rebuild_every": hmd.rebuild_every,
                     "rebuilds": hmd.rebuilds - rb0, "rollbacks": hmd.rollbacks - rl0,
                     "rebuild_ms_mean": statistics.mean(rebuild_ms) if rebuild_ms else None,
                     "max_owned_per_rank": int(shares[0].item()),
                     "max_ghosts_per_rank": int(shares[1].item()), "gpu_launches": int(halo_launches),
                     "what": "same K steps with brick decomposition: ghost positions by one all_to_all_single per step "
                             "(fixed lists between rebuilds), migration + list rebuild every `rebuild_every` steps"}
        pos0, vel0 = pos0_keep, vel0_keep
        del hmd

    if args.kernel_times and rank == 0:  # diagnostic: never feeds a reported number
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for _ in range(5):
                flush()
                md.step()
            torch.cuda.synchronize()
        rows = sorted(prof.key_averages(), key=lambda e: -e.device_time_total)[:12]
        for e in rows:
            print(f"[kernel-times] {e.key[:70]:70s} n={e.count:4d} avg={e.device_time_total / e.count / 1e3:8.4f} ms",
                  file=sys.stderr)

    mx = C.c_int32(max_seen_main)

    # ---- roofline of the dominant kernel (fused atom kernel), rank 0 ---------------------------------
    roofline = None
    extra = {}
    if rank == 0:
        # work units of one force evaluation: the generic kernels count the live units of the reference algorithm
        # (pairs, radial-SF and triplet-SF evaluations); the fast path reports the pair-list entries it actually walks
        # (Gaussian screening drops triplets whose weight is below 4e-18 of the group's largest)
        counters = torch.zeros(4, dtype=torch.int64, device=dev)
        _lib.check(lib.pantea_workspace_set_counters(md.ws.handle, _lib.ptr(counters)))
        lib.pantea_set_fast_path(0)
        md._forces(md.frc_new)
        lib.pantea_set_fast_path(1)
        md._forces(md.frc_new)
        torch.cuda.synchronize()
        _lib.check(lib.pantea_workspace_set_counters(md.ws.handle, None))
        n_pair, n_rad, n_trip, n_walked = (int(x) for x in counters[:4].tolist())
        fast_path = n_walked > 0
        n_own = md.hi - md.lo
        own_types = types_h[md.lo:md.hi]
        flops_mlp = sum(FLOP_MLP[int(x)] for x in own_types)
        flops_reference = FLOP_PAIR * n_pair + FLOP_RAD * n_rad + FLOP_TRIP * n_trip            # SURVEY 8(d) convention
        flops_executed = FLOP_PAIR * n_pair + FLOP_RAD * n_rad + FLOP_TRIP * (n_walked if fast_path else n_trip)
        reps = 5
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
        eval_ms = []
        if fast_path:
            _lib.check(lib.pantea_eval_timing(1, None))
        for a, b in ev:
            flush()
            a.record()
            md._forces(md.frc_new)
            b.record()
            if fast_path:
                ms_k = C.c_float(0.0)
                _lib.check(lib.pantea_eval_timing(1, C.byref(ms_k)))
                eval_ms.append(ms_k.value)
        torch.cuda.synchronize()
        if fast_path:
            _lib.check(lib.pantea_eval_timing(0, None))
        k_ms = statistics.mean(a.elapsed_time(b) for a, b in ev)
        dom_ms = statistics.mean(eval_ms) if eval_ms else k_ms
        # measured FMA-pipe peak (the bound of this kernel; MEASURED_PEAKS.json only holds HBM and bf16 tensor peaks)
        scratch = torch.zeros(16, dtype=torch.float64, device=dev)
        fl = C.c_double(0.0)
        peak = 0.0
        for _ in range(3):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            _lib.check(lib.pantea_bench_fma(code, 4096, 148 * 8, 256, _lib.ptr(scratch), C.byref(fl), _lib.stream_ptr()))
            b.record()
            torch.cuda.synchronize()
            peak = max(peak, fl.value / (a.elapsed_time(b) * 1e-3) / 1e12)
        traffic = None
        tfile = ROOT / "profiles" / "traffic.json"
        if tfile.exists():
            try:
                traffic = json.loads(tfile.read_text()).get("eval_kernel_dram_bytes_per_launch")
            except (ValueError, OSError):
                traffic = None
        achieved = flops_reference / (dom_ms * 1e-3) / 1e12
        roofline = {"bound": "fp64_pipe" if dtype == torch.float64 else "fp32_pipe",
                    "kernel": ("hdnnp_eval2_kernel (dominant kernel: neighbour records, radial and angular symmetry functions "
                               "with central gradients)" if fast_path else
                               "force evaluation = pair_filter_kernel + hdnnp_eval_kernel + mlp_force_kernel"),
                    "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                    "frac": achieved / peak if peak else None, "traffic": traffic,
                    "what": "achieved = ALGORITHMIC flops of the reference algorithm (SURVEY 8(d): 74/pair + 39/radial SF + "
                            "160/live triplet SF) / CUDA-event duration of the kernel; frac_executed counts only the triplets "
                            "the kernel walks after Gaussian screening",
                    "frac_executed": flops_executed / (dom_ms * 1e-3) / 1e12 / peak if peak else None,
                    "peak_source": "measured live: FMA microbenchmark (pantea_bench_fma), CUDA-core pipe; "
                                   "MEASURED_PEAKS.json holds no FP64 figure",
                    "kernel_ms": dom_ms, "algorithmic_flops_per_launch": flops_reference,
                    "executed_flops_per_launch": flops_executed,
                    "units_per_launch": {"atoms": n_own, "pairs": n_pair, "radial_sf": n_rad, "triplet_sf": n_trip,
                                         "triplet_sf_walked": n_walked if fast_path else n_trip},
                    "force_evaluation": {"what": "pair filter + evaluation + network kernels (everything but the neighbour rows)",
                                         "ms": k_ms, "algorithmic_flops": flops_reference + flops_mlp,
                                         "frac": (flops_reference + flops_mlp) / (k_ms * 1e-3) / 1e12 / peak if peak else None}}
        extra["kernel_share_of_step"] = k_ms / (ms_total / args.steps)
        peaks_file = ROOT / "MEASURED_PEAKS.json"
        if peaks_file.exists():
            hbm = json.loads(peaks_file.read_text()).get("hbm_gbs")
            extra["hbm_gbs_measured_peak"] = hbm

    # ---- end-to-end through the public C-ABI path with HOST buffers ---------------------------------
    # per step: pinned host positions -> device, neighbour build + fused energy/force kernel for the rank's
    # atoms, forces + energy back to pinned host memory, host waits for the result.
    e2e_steps = max(3, min(args.steps, 10))
    host_pos = torch.as_tensor(pos_h, dtype=dtype).pin_memory()
    host_frc = torch.empty((md.hi - md.lo, 3), dtype=dtype).pin_memory()
    host_e = torch.empty(1, dtype=dtype).pin_memory()
    d_pos = torch.empty((n, 3), dtype=dtype, device=dev)
    d_frc = torch.zeros((n, 3), dtype=dtype, device=dev)
    d_e = torch.zeros(1, dtype=dtype, device=dev)

    def e2e_step():
        d_pos.copy_(host_pos, non_blocking=True)
        md.ws.bind(d_pos, md.types, box, pot.r_cutoff, check=False, owned=(md.lo, md.hi))
        _lib.check(lib.pantea_energy_forces(md.ws.handle, None, _lib.ptr(d_frc), _lib.ptr(d_e), 0, _lib.stream_ptr()))
        host_frc.copy_(d_frc[md.lo:md.hi], non_blocking=True)
        host_e.copy_(d_e, non_blocking=True)
        torch.cuda.synchronize()

    for _ in range(2):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    barrier()
    e2e_t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    all_reduce_max(e2e_t)
    esz = 8 if dtype == torch.float64 else 4
    e2e = {"value": n * e2e_steps / float(e2e_t.item()), "unit": UNIT, "h2d_bytes_per_step": n * 3 * esz,
           "d2h_bytes_per_step": (md.hi - md.lo) * 3 * esz + esz, "steps": e2e_steps,
           "what": "host positions (pinned) -> pantea_neighbor_build + pantea_energy_forces -> host forces+energy, per rank"}

    # ---- CPU baseline beside it (rank 0, N = 1 only) -------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import c_oracle  # the checker, timed as the CPU baseline only
        from oracle.spec import load_potential

        specs = load_potential(GOLDEN / "h2o.json")
        pin_oracle_threads(c_oracle)
        t0 = time.perf_counter()
        c_oracle.energy_forces(specs, pos_h, types_h, box_h, begin=0, end=min(n, 512))
        rate = min(n, 512) / (time.perf_counter() - t0)
        m = int(max(256, min(n, rate * 12.0)))
        t0 = time.perf_counter()
        c_oracle.energy_forces(specs, pos_h, types_h, box_h, begin=0, end=m)
        el = time.perf_counter() - t0
        cpu = {"value": m / el, "unit": UNIT, "cores": c_oracle.num_threads(), "host_cpus": host_cpus(), "kind": "port",
               "sample": f"one force+energy evaluation of {m} of the {n} atoms ({el:.1f} s); oracle/hdnnp_oracle.c, "
                         "OpenMP, cell-list gather; the reference's JAX path cannot run here (no jax)"}

    line = None
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": args.dtype, "data": "synthetic",
            "config": {"workload": workload_name(n), "atoms": n,
                       "parallelism": (f"brick-decomposed x{world} {tuple(engine_info['bricks'])}, ghost halo over NVLink peer memory"
                                       if brick else f"replicated-coords block-owned x{world}"),
                       "engine": engine_info,
                       "l2": "flushed between timed steps (256 MB write)" if not args.no_flush else "not flushed",
                       "max_neighbors_seen": int(mx.value), "force_mode": "reference (central-role gradient)",
                       "trajectory": f"timed steps replay {SEGMENT}-step segments from the initial state (untimed reset)"},
            "parity": parity,
            "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
            "fp32": fp32_info, "verlet_skin": skin_info, "halo_exchange": halo_info,
            "wall_s_timed_region": wall,
        }
        line.update(extra)
    # ---- secondary workloads (all ranks take part).  The headline line above is complete before they start; a watchdog
    #      prints it without them if they fail to finish (a rank that died inside a collective must not cost the line)
    import threading

    def emit(extras_done, note=None):
        if rank == 0:
            line["preprocess"], line["million"] = extras_done.get("preprocess"), extras_done.get("million")
            if note:
                line["extras_note"] = note
            print(json.dumps(line), flush=True)

    extras = {}
    run_extras = not args.no_extra and os.environ.get("PANTEA_BENCH_EXTRA", "1") != "0" and args.atoms <= 200000
    if run_extras:
        def bail():
            emit(extras, "secondary workloads did not finish within the time limit")
            os._exit(0)
        watchdog = threading.Timer(float(os.environ.get("PANTEA_BENCH_EXTRA_LIMIT_S", "240")), bail)
        watchdog.daemon = True
        watchdog.start()
        del md
        if brick:
            brick_md.close()
        torch.cuda.empty_cache()
        for name, fn in (("preprocess", lambda: run_preprocess(pot, rank, world, dev)),
                         ("million", lambda: run_million(pot, rank, world, dev, dtype))):
            try:
                extras[name] = fn()
            except Exception as exc:  # noqa: BLE001
                extras[name] = {"error": f"{type(exc).__name__}: {exc}"}
                if world > 1:
                    break  # the ranks may be out of step: no further collectives
        watchdog.cancel()
    emit(extras)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main() -> None:
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
