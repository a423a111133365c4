"""Brick-decomposed molecular dynamics with ghost-atom halo exchange (north star
"large-system MD: the box is spatially domain-decomposed across the GPUs of one box, with ghost-atom halo
exchange").  One process per GPU, `torch.distributed` over NCCL (NVLink / NVSwitch).

Scheme.  The periodic box is cut into `px x py x pz` bricks (8 ranks -> 2x2x2, 4 -> 2x2x1, 2 -> 2x1x1).  A rank
*owns* the atoms whose wrapped position lies in its brick and additionally holds *ghost* copies of every foreign
atom within `r_cutoff + skin` (periodic distance) of the brick.  Local arrays are `[owned | ghosts]`; all
coordinates stay the true wrapped coordinates of the global box, so the library's periodic cell list and its
single-shift minimum image (reference `atoms/box.py:112-117`) apply unchanged -- the neighbour set of an owned
atom, as a set of global ids, is exactly the single-GPU one -- and the energy/force kernels run over the owned
range `[0, n_own)` of the local arrays (`pantea_workspace_set_owned_range`).

Per step, between two rebuilds of the ghost lists: ONE launch of `pantea_halo_pack` (gathers the fixed send list and
checks the Verlet criterion of the ghost shell on the device) and ONE `all_to_all_single` of ghost positions
(grouped NCCL send/recv with fixed message sizes, written straight into the ghost rows) -- no host
synchronisation.  The reference force (`force.py:16-43`) is the central-role derivative, so no reverse
communication is needed; `force_mode = FORCE_FULL` adds the reverse halo (ghost rows sent back
with the splits swapped, `pantea_halo_unpack_add`).

Every `rebuild_every` steps (after the position update, before the force evaluation) the lists are rebuilt:
atoms migrate to the brick they now sit in (position, velocity, force, mass, type, global id: `all_to_all_single`
with exchanged counts), and the send lists are re-selected.  With `rebuild_every = 1` and `skin = 0` this is
exact by construction.  With longer segments the sticky device flag of `pantea_halo_pack` is read at the next
rebuild (max over ranks); if an atom moved more than `skin / 2` the segment is rolled back to the snapshot taken
after the previous rebuild and repeated with half the interval, so results never depend on a violated skin.

`BrickGrid` and `HaloDomain` are device-agnostic host logic (torch ops + collectives; covered on CPU with gloo,
world_size 2); `HaloMD` drives the CUDA library and raises without it (no CPU fallback).
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def brick_dims(world: int, box: Sequence[float]) -> Tuple[int, int, int]:
    """Factorisation px * py * pz = world with the smallest brick surface (ties: more cuts along x, then y)."""
    best, best_key = (world, 1, 1), None
    for px in range(1, world + 1):
        if world % px:
            continue
        for py in range(1, world // px + 1):
            if (world // px) % py:
                continue
            pz = world // (px * py)
            a, b, c = box[0] / px, box[1] / py, box[2] / pz
            key = (round(a * b + b * c + a * c, 9), -px, -py)
            if best_key is None or key < best_key:
                best, best_key = (px, py, pz), key
    return best


class BrickGrid:
    """`px x py x pz` bricks of a periodic orthorhombic box; rank = (ix * py + iy) * pz + iz.  `cuts[d]` are the
    `p_d - 1` interior brick boundaries along d (ascending; default: equal widths) -- `balanced()` places them at the
    atom-count quantiles of a configuration, which evens out the owned atoms when the density varies along an axis."""

    def __init__(self, box: Sequence[float], world: int, dims: Optional[Sequence[int]] = None,
                 cuts: Optional[Sequence[Sequence[float]]] = None) -> None:
        self.box = [float(b) for b in box]
        self.world = int(world)
        self.dims = tuple(int(d) for d in dims) if dims is not None else brick_dims(self.world, self.box)
        if self.dims[0] * self.dims[1] * self.dims[2] != self.world:
            raise ValueError(f"brick grid {self.dims} does not match {self.world} ranks")
        if cuts is None:
            cuts = [[k * self.box[d] / self.dims[d] for k in range(1, self.dims[d])] for d in range(3)]
        self.cuts = [[float(c) for c in cuts[d]] for d in range(3)]
        for d in range(3):
            if len(self.cuts[d]) != self.dims[d] - 1 or sorted(self.cuts[d]) != self.cuts[d]:
                raise ValueError(f"cuts[{d}] must hold {self.dims[d] - 1} ascending boundaries")
        self.edges = [[0.0, *self.cuts[d], self.box[d]] for d in range(3)]
        self._cache = {}

    @classmethod
    def balanced(cls, box: Sequence[float], world: int, positions: torch.Tensor,
                 dims: Optional[Sequence[int]] = None) -> "BrickGrid":
        """Brick boundaries at the k / p_d quantiles of the atoms' (wrapped) coordinates along every cut axis."""
        grid = cls(box, world, dims)
        cuts = []
        for d in range(3):
            p = grid.dims[d]
            if p == 1:
                cuts.append([])
                continue
            x = torch.sort(positions[:, d].double()).values
            n = x.numel()
            cuts.append([float(x[min(n - 1, (k * n) // p)]) for k in range(1, p)])
        return cls(box, world, grid.dims, cuts)

    def coords(self, rank: int) -> Tuple[int, int, int]:
        px, py, pz = self.dims
        return rank // (py * pz), (rank // pz) % py, rank % pz

    def bounds(self, rank: int) -> Tuple[List[float], List[float]]:
        c = self.coords(rank)
        return [self.edges[d][c[d]] for d in range(3)], [self.edges[d][c[d] + 1] for d in range(3)]

    def _tensors(self, device):
        key = str(device)
        if key not in self._cache:
            lo = torch.tensor([self.bounds(r)[0] for r in range(self.world)], dtype=torch.float64, device=device)
            hi = torch.tensor([self.bounds(r)[1] for r in range(self.world)], dtype=torch.float64, device=device)
            cuts = [torch.tensor(self.cuts[d], dtype=torch.float64, device=device) for d in range(3)]
            self._cache[key] = (lo, hi, cuts, torch.arange(self.world, device=device))
        return self._cache[key]

    def owner(self, pos: torch.Tensor) -> torch.Tensor:
        """Owning rank of every (wrapped) position: single-valued, so every atom has exactly one owner
        (a coordinate equal to a boundary belongs to the upper brick)."""
        _, _, cuts, _ = self._tensors(pos.device)
        idx = []
        for d in range(3):
            if self.dims[d] == 1:
                idx.append(torch.zeros(pos.shape[0], dtype=torch.int64, device=pos.device))
            else:
                idx.append(torch.bucketize(torch.remainder(pos[:, d].double(), self.box[d]).contiguous(), cuts[d], right=True))
        return (idx[0] * self.dims[1] + idx[1]) * self.dims[2] + idx[2]

    def distance2_all(self, pos: torch.Tensor) -> torch.Tensor:
        """[world, n] squared periodic distances of every position to every brick (0 inside)."""
        lo, hi, _, _ = self._tensors(pos.device)
        d2 = torch.zeros((self.world, pos.shape[0]), dtype=torch.float64, device=pos.device)
        for d in range(3):
            if self.dims[d] == 1:
                continue  # the bricks span the whole (periodic) box along d
            x, L = pos[:, d].double()[None, :], self.box[d]
            l, h = lo[:, d, None], hi[:, d, None]
            gap = torch.minimum(torch.remainder(l - x, L), torch.remainder(x - h, L))
            gap = torch.where((x >= l) & (x < h), torch.zeros_like(gap), gap)
            d2 += gap * gap
        return d2

    def distance2(self, pos: torch.Tensor, rank: int) -> torch.Tensor:
        return self.distance2_all(pos)[rank]

    def ghost_masks(self, pos: torch.Tensor, owner: torch.Tensor, r_halo: float) -> torch.Tensor:
        """[world, n]: atom needed as a ghost by brick r = not owned by r and within r_halo of it (slightly inclusive)."""
        lim = r_halo * (1.0 + 1e-12) + 1e-12
        ranks = self._tensors(pos.device)[3]
        return (owner[None, :] != ranks[:, None]) & (self.distance2_all(pos) <= lim * lim)

    def ghost_mask(self, pos: torch.Tensor, owner: torch.Tensor, rank: int, r_halo: float) -> torch.Tensor:
        return self.ghost_masks(pos, owner, r_halo)[rank]


class HaloDomain:
    """Ownership, migration and ghost lists of one rank (host logic; tensors may live on any device)."""

    def __init__(self, box: Sequence[float], r_halo: float, rank: int, world: int,
                 dims: Optional[Sequence[int]] = None, group=None, grid: Optional[BrickGrid] = None) -> None:
        self.grid = grid if grid is not None else BrickGrid(box, world, dims)
        self.r_halo, self.rank, self.world, self.group = float(r_halo), int(rank), int(world), group
        self.send_idx: Optional[torch.Tensor] = None
        self.send_splits: List[int] = [0] * world
        self.recv_splits: List[int] = [0] * world
        self.rev_order: Optional[torch.Tensor] = None
        self.rev_first: Optional[torch.Tensor] = None

    # ------------------------------------------------------------------------------ collectives
    def exchange_counts(self, send_counts: Sequence[int], device) -> List[int]:
        if self.world == 1:
            return list(send_counts)
        s = torch.tensor(list(send_counts), dtype=torch.int64, device=device)
        r = torch.empty_like(s)
        dist.all_to_all_single(r, s, group=self.group)
        return [int(x) for x in r.tolist()]

    def all_to_all_rows(self, rows: torch.Tensor, send_splits: Sequence[int], recv_splits: Sequence[int],
                        out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Rows [sum(send_splits), ...] grouped by destination -> rows [sum(recv_splits), ...] grouped by source."""
        n_recv = int(sum(recv_splits))
        if out is None:
            out = torch.empty((n_recv, *rows.shape[1:]), dtype=rows.dtype, device=rows.device)
        if self.world == 1:
            out.copy_(rows)
            return out
        dist.all_to_all_single(out, rows.contiguous(), list(recv_splits), list(send_splits), group=self.group)
        return out

    # ------------------------------------------------------------------------------ migration
    def migrate(self, pos_owned: torch.Tensor, arrays: Sequence[torch.Tensor], gid: torch.Tensor):
        """Send every owned atom to the rank whose brick holds its (wrapped) position.  `arrays` are per-atom arrays
        (first axis = owned atoms); returns the new arrays and global ids (arrival order: by source rank, stable).
        Two payload messages per peer: all floating-point columns in one, all integer columns (and the ids) in one."""
        if self.world == 1:
            return list(arrays), gid
        n = int(pos_owned.shape[0])
        dest = self.grid.owner(pos_owned)
        order = torch.argsort(dest, stable=True)
        send = [int(x) for x in torch.bincount(dest, minlength=self.world).tolist()]
        recv = self.exchange_counts(send, pos_owned.device)
        fl = [i for i, a in enumerate(arrays) if a.is_floating_point()]
        it = [i for i, a in enumerate(arrays) if not a.is_floating_point()]
        out: List[Optional[torch.Tensor]] = [None] * len(arrays)
        if fl:
            cols = torch.cat([arrays[i].reshape(n, -1) for i in fl], dim=1)[order]
            got = self.all_to_all_rows(cols, send, recv)
            c0 = 0
            for i in fl:
                w = arrays[i].reshape(n, -1).shape[1]
                out[i] = got[:, c0:c0 + w].reshape(-1, *arrays[i].shape[1:]).contiguous()
                c0 += w
        cols = torch.cat([gid.reshape(n, 1)] + [arrays[i].reshape(n, -1).to(torch.int64) for i in it], dim=1)[order]
        got = self.all_to_all_rows(cols, send, recv)
        c0 = 1
        for i in it:
            w = arrays[i].reshape(n, -1).shape[1]
            out[i] = got[:, c0:c0 + w].reshape(-1, *arrays[i].shape[1:]).to(arrays[i].dtype).contiguous()
            c0 += w
        return out, got[:, 0].contiguous()

    # ------------------------------------------------------------------------------ ghost lists
    def build_lists(self, pos_owned: torch.Tensor) -> int:
        """Select, per destination rank, the owned atoms its brick needs as ghosts (one batched distance evaluation and
        one compaction for all destinations); exchange the counts.  Returns the number of ghosts this rank receives."""
        n_own = int(pos_owned.shape[0])
        dev = pos_owned.device
        if self.world == 1:
            self.send_idx = torch.zeros(0, dtype=torch.int64, device=dev)
            self.send_splits, self.recv_splits = [0], [0]
        else:
            mine = torch.full((n_own,), self.rank, dtype=torch.int64, device=dev)
            dst, idx = torch.nonzero(self.grid.ghost_masks(pos_owned, mine, self.r_halo), as_tuple=True)  # sorted by dst
            self.send_idx = idx.contiguous()
            self.send_splits = [int(x) for x in torch.bincount(dst, minlength=self.world).tolist()]
            self.recv_splits = self.exchange_counts(self.send_splits, dev)
        self.rev_order = self.rev_first = None
        self._n_own = n_own
        return int(sum(self.recv_splits))

    def reverse_index(self) -> Tuple[torch.Tensor, torch.Tensor]:
        """(order, first) of the reverse halo: send-list entries sorted by owned atom and each atom's range in them."""
        if self.rev_order is None:
            dev = self.send_idx.device
            self.rev_order = torch.argsort(self.send_idx, stable=True)
            first = torch.zeros(self._n_own + 1, dtype=torch.int64, device=dev)
            if self.send_idx.numel():
                first[1:] = torch.cumsum(torch.bincount(self.send_idx, minlength=self._n_own), 0)
            self.rev_first = first
        return self.rev_order, self.rev_first

    @property
    def n_send(self) -> int:
        return int(sum(self.send_splits))

    @property
    def n_ghost(self) -> int:
        return int(sum(self.recv_splits))

    def forward(self, owned_rows: torch.Tensor, out: Optional[torch.Tensor] = None, packed: bool = False) -> torch.Tensor:
        """Ghost rows of a per-atom array: owned rows (or the already packed send buffer) -> [n_ghost, ...]."""
        rows = owned_rows if packed else owned_rows.index_select(0, self.send_idx)
        return self.all_to_all_rows(rows, self.send_splits, self.recv_splits, out=out)

    def reverse(self, ghost_rows: torch.Tensor) -> torch.Tensor:
        """Ghost rows back to their owners: [n_ghost, ...] -> [n_send, ...] in send-list order."""
        return self.all_to_all_rows(ghost_rows, self.recv_splits, self.send_splits)

    def gather_global(self, owned_rows: torch.Tensor, gid: torch.Tensor, n_total: int) -> torch.Tensor:
        """Assemble the full [n_total, ...] array from every rank's owned rows (diagnostics / tests)."""
        out = torch.zeros((n_total, *owned_rows.shape[1:]), dtype=owned_rows.dtype, device=owned_rows.device)
        if self.world == 1:
            out[gid] = owned_rows
            return out
        counts = self.exchange_counts([int(gid.numel())] * self.world, gid.device)  # everybody's owned count
        send = [int(gid.numel())] * self.world
        gids = self.all_to_all_rows(gid.repeat(self.world), send, counts)
        rows = self.all_to_all_rows(owned_rows.repeat(self.world, *([1] * (owned_rows.dim() - 1))), send, counts)
        out[gids] = rows
        return out


class HaloMD:
    """Velocity-Verlet MD (reference integrator, no mass: `molecular_dynamics.py:16-30`) of one periodic box on
    `world` GPUs with brick decomposition and ghost-atom halo exchange.  Same interface as `ReplicatedMD`; the full
    initial arrays are passed to every rank, which keeps only the atoms of its brick.  `balance=True` places the brick
    boundaries at the atom-count quantiles of the initial configuration."""

    def __init__(self, device_potential, positions: torch.Tensor, velocities: torch.Tensor, masses: torch.Tensor,
                 types: torch.Tensor, box: Sequence[float], time_step: float, rank: int = 0, world: int = 1,
                 thermostat=None, kb: float = 3.166811563e-6, skin: float = 0.0, rebuild_every: int = 1,
                 dims: Optional[Sequence[int]] = None, force_mode: int = 0, balance: bool = True) -> None:
        from pantea_b200 import _lib, engine

        self._lib, self.lib, self._engine = _lib, _lib.load(), engine
        _lib.require_cuda()
        if box is None:
            raise ValueError("HaloMD needs a periodic box")
        if skin < 0.0 or rebuild_every < 1:
            raise ValueError("skin must be >= 0 and rebuild_every >= 1")
        if rebuild_every > 1 and skin <= 0.0:
            raise ValueError("rebuild_every > 1 needs a skin > 0 (ghost shell of r_cutoff + skin)")
        self.pot = device_potential
        self.n, self.dt, self.box = int(positions.shape[0]), float(time_step), [float(b) for b in box]
        self.rank, self.world = int(rank), int(world)
        self.dtype, self.code = positions.dtype, _lib.dtype_code(positions.dtype)
        self.dev = positions.device
        self.thermostat, self.kb = thermostat, kb
        self.skin, self.force_mode = float(skin), int(force_mode)
        self.rebuild_every = self._every0 = int(rebuild_every)
        grid = (BrickGrid.balanced(self.box, world, positions, dims) if balance and world > 1
                else BrickGrid(self.box, world, dims))
        self.domain = HaloDomain(self.box, device_potential.r_cutoff + self.skin, rank, world, grid=grid)
        self.ke = torch.zeros(1, dtype=torch.float64, device=self.dev)
        self.violated = torch.zeros(1, dtype=torch.int32, device=self.dev)
        self.ws = None
        self.rebuilds = 0
        self.rollbacks = 0
        self.rebuild_events = None  # set to a list to collect (start, stop) CUDA events around every rebuild (diagnostic)
        self._global0 = (masses.reshape(-1).to(self.dtype), types.to(torch.int32))
        self.reset(positions, velocities)

    # ------------------------------------------------------------------------------ state
    def reset(self, positions: torch.Tensor, velocities: torch.Tensor) -> None:
        """(Re)start from full global arrays (same atoms, box, potential)."""
        masses, types = self._global0
        pos = positions.to(self.dtype)
        mine = torch.nonzero(self.domain.grid.owner(pos) == self.rank, as_tuple=True)[0]
        self.gid = mine
        n_own = int(mine.numel())
        self.steps = 0
        self._since = 0
        self.rebuild_every = self._every0
        self._install(pos[mine].contiguous(), velocities[mine].to(self.dtype).contiguous(),
                      torch.zeros((n_own, 3), dtype=self.dtype, device=self.dev), masses[mine].contiguous(),
                      types[mine].contiguous())
        self._evaluate(first=True)
        self.frc[:n_own].copy_(self.frc_new[:n_own])
        self._snapshot()

    def _install(self, pos: torch.Tensor, vel: torch.Tensor, frc: torch.Tensor, mass: torch.Tensor,
                 types: torch.Tensor) -> None:
        """Build the ghost lists for the given owned arrays and set up the local [owned | ghost] arrays."""
        dom, n_own = self.domain, int(pos.shape[0])
        n_ghost = dom.build_lists(pos)
        n_loc = n_own + n_ghost
        self.n_own, self.n_ghost, self.n_local = n_own, n_ghost, n_loc
        self.pos = torch.empty((n_loc, 3), dtype=self.dtype, device=self.dev)
        self.pos[:n_own] = pos
        self.types = torch.empty(n_loc, dtype=torch.int32, device=self.dev)
        self.types[:n_own] = types
        self.vel, self.mass = vel, mass
        rows = n_loc if self.force_mode == self._lib.FORCE_FULL else n_own  # full mode also writes the ghost rows
        self.frc = frc
        self.frc_new = torch.empty((rows, 3), dtype=self.dtype, device=self.dev)
        self.send_buf = torch.empty((dom.n_send, 3), dtype=self.dtype, device=self.dev)
        if self.world > 1:
            dom.forward(pos, out=self.pos[n_own:])
            dom.forward(types, out=self.types[n_own:])
        self.pos_ref = pos.clone() if self.rebuild_every > 1 else None
        self.violated.zero_()
        self.lo, self.hi = 0, n_own  # owned range of the local arrays (bench / diagnostics)
        if self.ws is None or self.ws.max_atoms < n_loc:
            if self.ws is not None:
                self.ws.close()
            density = self.n / (self.box[0] * self.box[1] * self.box[2])
            cap_atoms = int(1.25 * n_loc) + 1024
            self.ws = self._engine.Workspace(self.pot, cap_atoms, self._engine.estimate_max_neighbors(
                self.pot.r_cutoff, density, self.n), self.dtype)
        self.rebuilds += 1

    def _snapshot(self) -> None:
        n = self.n_own
        self._snap = (self.pos[:n].clone(), self.vel.clone(), self.frc.clone(), self.steps)

    # ------------------------------------------------------------------------------ pieces of a step
    def _exchange_positions(self) -> None:
        _lib, dom = self._lib, self.domain
        check = self.pos_ref is not None
        _lib.check(self.lib.pantea_halo_pack(
            _lib.ptr(self.pos), _lib.ptr(dom.send_idx), dom.n_send, _lib.ptr(self.send_buf),
            _lib.ptr(self.pos_ref) if check else None, self.n_own if check else 0, _lib.box_arg(self.box),
            0.5 * self.skin, _lib.ptr(self.violated) if check else None, self.code, _lib.stream_ptr()))
        if self.world > 1:
            dom.forward(self.send_buf, out=self.pos[self.n_own:], packed=True)

    def _evaluate(self, first: bool = False, e_atom: Optional[torch.Tensor] = None) -> None:
        """Neighbour rows over [owned | ghosts] in the global box, energies / forces of the owned atoms -> frc_new."""
        _lib = self._lib
        self.ws.bind(self.pos, self.types, self.box, self.pot.r_cutoff, check=first, owned=(0, self.n_own))
        _lib.check(self.lib.pantea_energy_forces(self.ws.handle, _lib.ptr(e_atom), _lib.ptr(self.frc_new), None,
                                                 self.force_mode, _lib.stream_ptr()))
        if self.force_mode == _lib.FORCE_FULL and self.world > 1:
            dom = self.domain
            back = dom.reverse(self.frc_new[self.n_own:])
            order, first_ = dom.reverse_index()
            _lib.check(self.lib.pantea_halo_unpack_add(_lib.ptr(self.frc_new), _lib.ptr(back), _lib.ptr(order),
                                                       _lib.ptr(first_), self.n_own, self.code, _lib.stream_ptr()))

    def _segment_violated(self) -> bool:
        if self.pos_ref is None:
            return False
        flag = self.violated.clone()
        if self.world > 1:
            dist.all_reduce(flag, op=dist.ReduceOp.MAX)
        return bool(int(flag.item()))

    def _rebuild(self) -> None:
        """Migration + new ghost lists from the current owned positions (host-synchronous)."""
        n = self.n_own
        if self.rebuild_events is not None:
            ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            ev[0].record()
        arrays = [self.pos[:n], self.vel, self.frc, self.mass, self.types[:n]]
        (pos, vel, frc, mass, types), self.gid = self.domain.migrate(self.pos[:n], arrays, self.gid)
        self._install(pos, vel, frc, mass, types)
        self._since = 0
        if self.rebuild_events is not None:
            ev[1].record()
            self.rebuild_events.append(ev)

    def _advance(self) -> int:
        """One step; returns 0, or the number of steps that were rolled back (a violated ghost-shell skin)."""
        _lib, lib, st = self._lib, self.lib, self._lib.stream_ptr()
        box_c = _lib.box_arg(self.box)
        _lib.check(lib.pantea_md_update_positions(_lib.ptr(self.pos), _lib.ptr(self.vel), _lib.ptr(self.frc), 0,
                                                  self.n_own, box_c, self.dt, self.code, st))
        rebuilt = False
        if self._since + 1 >= self.rebuild_every:
            if self._segment_violated():
                return self._rollback()
            self._rebuild()
            rebuilt = True
        else:
            self._exchange_positions()
            self._since += 1
        self._evaluate()
        _lib.check(lib.pantea_md_update_velocities(_lib.ptr(self.vel), _lib.ptr(self.frc), _lib.ptr(self.frc_new), 0,
                                                   self.n_own, self.dt, self.code, st))
        if self.thermostat is not None:
            self.kinetic_energy()
            _lib.check(lib.pantea_md_rescale_velocities(_lib.ptr(self.vel), 0, self.n_own, _lib.ptr(self.ke), self.n,
                                                        self.dt, self.thermostat.time_constant,
                                                        self.thermostat.target_temperature, self.kb, self.code, st))
        self.steps += 1
        if rebuilt and self.rebuild_every > 1:
            self._snapshot()
        return 0

    def _rollback(self, in_flight: bool = True) -> int:
        pos, vel, frc, step0 = self._snap
        undone = self.steps - step0 + (1 if in_flight else 0)  # completed steps of the segment (+ the one in flight)
        if self.rebuild_every <= 1:
            raise RuntimeError("HaloMD: an atom moved more than skin/2 within one step; increase the skin")
        n = self.n_own
        self.pos[:n].copy_(pos)
        self.vel.copy_(vel)
        self.frc.copy_(frc)
        self.steps = step0
        self.rebuild_every = max(1, self.rebuild_every // 2)
        self._since = self.rebuild_every  # the repeated segment starts with a rebuild
        self.violated.zero_()
        if self.rebuild_every == 1:
            self.pos_ref = None
        self.rollbacks += 1
        return undone

    def _run(self, todo: int) -> None:
        while todo > 0:
            undone = self._advance()
            todo += undone - 1 if undone else -1

    def step(self) -> None:
        self._run(1)

    def validate(self) -> None:
        """Host-synchronous check of the segment in flight (call before reading results when rebuild_every > 1): if an
        atom has left the ghost-shell skin since the last rebuild, the segment is rolled back and repeated."""
        while self._segment_violated():
            self._run(self._rollback(in_flight=False))

    # ------------------------------------------------------------------------------ observables
    def check_capacity(self) -> int:
        import ctypes as C
        mx = C.c_int32(0)
        self._lib.check(self.lib.pantea_neighbor_status(self.ws.handle, C.byref(mx), self._lib.stream_ptr()))
        return int(mx.value)

    def kinetic_energy(self) -> torch.Tensor:
        _lib = self._lib
        self.ke.zero_()
        if self.n_own:
            _lib.check(self.lib.pantea_md_kinetic_energy(_lib.ptr(self.vel), _lib.ptr(self.mass), 0, self.n_own,
                                                         _lib.ptr(self.ke), self.code, _lib.stream_ptr()))
        if self.world > 1:
            dist.all_reduce(self.ke, op=dist.ReduceOp.SUM)
        return self.ke

    def potential_energy(self) -> torch.Tensor:
        keep = self.frc_new.clone()
        e_atom = torch.zeros(self.n_local, dtype=self.dtype, device=self.dev)
        self._evaluate(e_atom=e_atom)
        self.frc_new.copy_(keep)
        e = e_atom[: self.n_own].double().sum().reshape(1)
        if self.world > 1:
            dist.all_reduce(e, op=dist.ReduceOp.SUM)
        return e

    def gather_owned(self, t: torch.Tensor) -> torch.Tensor:
        """Full [n, ...] array (global atom order) from every rank's owned rows."""
        return self.domain.gather_global(t[: self.n_own].contiguous(), self.gid, self.n)
