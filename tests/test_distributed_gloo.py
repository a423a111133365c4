"""world_size-2 gloo test of the multi-GPU host logic: block ownership, the in-place position all-gather and the
scalar reductions used by ReplicatedMD (the CUDA kernels themselves are covered by the gpu tests)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pantea_b200.distributed import BlockLayout, all_reduce_max, all_reduce_sum


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_atoms, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        layout = BlockLayout(n_atoms, rank, world)
        ref = torch.arange(n_atoms * 3, dtype=torch.float64).reshape(n_atoms, 3)
        buf = layout.allocate(torch.zeros_like(ref))
        lo, hi = layout.owned()
        buf[lo:hi] = ref[lo:hi] + 1.0          # every rank "integrates" only its own block ...
        layout.exchange(buf)                   # ... and one in-place all-gather makes everybody consistent
        ok_gather = torch.equal(buf[:n_atoms], ref + 1.0)
        s = all_reduce_sum(torch.tensor([float(hi - lo)], dtype=torch.float64))
        m = all_reduce_max(torch.tensor([float(rank)], dtype=torch.float64))
        out[rank] = (ok_gather, float(s), float(m), lo, hi)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_atoms", [10, 11, 96])
def test_block_layout_exchange_world2(n_atoms):
    world = 2
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, _free_port(), n_atoms, out), nprocs=world, join=True)
        res = dict(out)
    assert all(res[r][0] for r in range(world))
    assert res[0][1] == n_atoms and res[0][2] == world - 1
    assert res[0][3] == 0 and res[0][4] == res[1][3] and res[1][4] == n_atoms     # contiguous, complete cover


def test_block_layout_properties():
    for n, w in ((10, 4), (99999, 8), (7, 8), (1, 1)):
        covered = []
        for r in range(w):
            lo, hi = BlockLayout(n, r, w).owned()
            assert 0 <= lo <= hi <= n and hi - lo <= BlockLayout(n, r, w).block
            covered += list(range(lo, hi))
        assert covered == list(range(n))
        assert BlockLayout(n, 0, w).padded % w == 0 and BlockLayout(n, 0, w).padded >= n


def _scaler_worker(rank, world, port, out):
    from pantea_b200.descriptors.scaler import DescriptorScaler
    from pantea_b200.distributed import merge_scaler_params
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        data = torch.from_numpy(np.random.default_rng(42).normal(2.0, 3.0, size=(37, 5)))
        mine = data[rank::world]                       # structures are split index mod world
        params = DescriptorScaler.fit(mine[:3])
        params = DescriptorScaler.partial_fit(params, mine[3:])
        merged = merge_scaler_params(params)
        empty = merge_scaler_params(params if rank == 0 else None)   # a rank without samples of the element
        out[rank] = (merged.nsamples.item(), merged.mean.numpy(), merged.sigma.numpy(), merged.minval.numpy(),
                     merged.maxval.numpy(), empty.nsamples.item(), empty.mean.numpy())
    finally:
        dist.destroy_process_group()


def test_scaler_statistics_merge_across_ranks_world2():
    """SURVEY 8(e) dataset preprocessing: per-rank scaler statistics merged into the whole-dataset ones."""
    world = 2
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_scaler_worker, args=(world, _free_port(), out), nprocs=world, join=True)
        res = dict(out)
    data = np.random.default_rng(42).normal(2.0, 3.0, size=(37, 5))
    for r in range(world):
        n, mean, sigma, mn, mx, n_e, mean_e = res[r]
        assert n == 37
        np.testing.assert_allclose(mean, data.mean(0), rtol=1e-13)
        np.testing.assert_allclose(sigma, data.std(0), rtol=1e-12)
        np.testing.assert_array_equal(mn, data.min(0))
        np.testing.assert_array_equal(mx, data.max(0))
        assert n_e == len(data[0::2])
        np.testing.assert_allclose(mean_e, data[0::2].mean(0), rtol=1e-13)
    for k in range(1, 5):
        np.testing.assert_array_equal(res[0][k], res[1][k])        # every rank holds bitwise the same statistics


# ---------------------------------------------------------------------------------------------- halo exchange
def _halo_worker(rank, world, port, dims, out):
    from pantea_b200.halo import HaloDomain
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(5)
        box, rc, n = [30.0, 24.0, 27.0], 6.0, 900
        pos = torch.from_numpy(rng.uniform(0.0, 1.0, (n, 3)) * np.asarray(box))
        vel = torch.from_numpy(rng.normal(size=(n, 3)))
        types = torch.from_numpy(rng.integers(1, 3, n).astype(np.int32))
        dom = HaloDomain(box, rc, rank, world, dims)
        gid = torch.nonzero(dom.grid.owner(pos) == rank, as_tuple=True)[0]
        own_pos, own_vel, own_types = pos[gid].clone(), vel[gid].clone(), types[gid].clone()
        res = {}
        for phase in range(2):
            n_ghost = dom.build_lists(own_pos)
            ghost_pos, ghost_gid = dom.forward(own_pos), dom.forward(gid)
            assert ghost_pos.shape[0] == n_ghost == dom.n_ghost
            # reverse halo: every ghost row returns 1 -> an owned atom collects its number of ghost copies
            back = dom.reverse(torch.ones((n_ghost, 1), dtype=torch.float64))
            copies = torch.zeros(len(gid), dtype=torch.float64).index_add_(0, dom.send_idx, back[:, 0])
            rev_first = dom.reverse_index()[1]
            occ = (rev_first[1:] - rev_first[:-1]).double()
            res[phase] = (gid.numpy().copy(), own_pos.numpy().copy(), own_vel.numpy().copy(), own_types.numpy().copy(),
                          ghost_gid.numpy().copy(), ghost_pos.numpy().copy(), copies.numpy().copy(),
                          bool(torch.equal(copies, occ)),
                          dom.gather_global(own_vel, gid, n).numpy().copy(), pos.numpy().copy())
            # every atom moves (same draw on every rank), is wrapped, and migrates to its new brick
            step = torch.from_numpy(np.random.default_rng(77).normal(scale=2.0, size=(n, 3)))
            pos = torch.remainder(pos + step, torch.tensor(box, dtype=torch.float64))
            own_pos = pos[gid].clone()
            (own_pos, own_vel, own_types), gid = dom.migrate(own_pos, [own_pos, own_vel, own_types], gid)
        out[rank] = res
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("dims", [(2, 1, 1), (1, 1, 2), (2, 2, 1)])
def test_halo_domain_migration_and_ghosts_world2(dims):
    """brick ownership, ghost selection, forward / reverse halo and migration on 2 ranks (and on 4 ranks in a
    2 x 2 x 1 grid: several peers per rank, atoms needed by more than one brick) -- gloo, CPU."""
    world = dims[0] * dims[1] * dims[2]
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_halo_worker, args=(world, _free_port(), dims, out), nprocs=world, join=True)
        res = dict(out)
    rng = np.random.default_rng(5)
    box, rc, n = np.asarray([30.0, 24.0, 27.0]), 6.0, 900
    pos = rng.uniform(0.0, 1.0, (n, 3)) * box
    vel = rng.normal(size=(n, 3))
    types = rng.integers(1, 3, n).astype(np.int32)
    for phase in range(2):
        if phase == 1:
            pos = res[0][phase][9]
            assert not np.array_equal(res[0][0][0], res[0][1][0])                 # atoms did change owner
        gids = [res[r][phase][0] for r in range(world)]
        assert sorted(np.concatenate(gids).tolist()) == list(range(n))          # every atom has exactly one owner
        d = pos[:, None, :] - pos[None, :, :]
        d -= box * np.rint(d / box)
        within = (d ** 2).sum(-1) <= rc * rc
        for r in range(world):
            gid, own_pos, own_vel, own_types, ghost_gid, ghost_pos, copies, rev_ok, vel_all, _ = res[r][phase]
            np.testing.assert_array_equal(own_pos, pos[gid])                      # payloads follow their atoms
            np.testing.assert_array_equal(own_vel, vel[gid])
            np.testing.assert_array_equal(own_types, types[gid])
            np.testing.assert_array_equal(ghost_pos, pos[ghost_gid])
            np.testing.assert_array_equal(vel_all, vel)
            assert rev_ok and len(set(ghost_gid.tolist())) == len(ghost_gid) and not set(ghost_gid) & set(gid)
            local = set(gid.tolist()) | set(ghost_gid.tolist())
            needed = set(np.nonzero(within[gid].any(0))[0].tolist())             # anything within rc of an owned atom
            assert needed <= local
            expect = sum(np.isin(gid, res[o][phase][4]).astype(float) for o in range(world) if o != r)
            np.testing.assert_array_equal(copies, expect)      # one returned row per brick that holds a ghost copy


@pytest.mark.parametrize("world,expect", [(8, (2, 2, 2)), (4, (2, 2, 1)), (2, (2, 1, 1)), (6, (3, 2, 1)), (1, (1, 1, 1))])
def test_brick_grid_covers_every_neighbour(world, expect):
    """Ghost criterion of the brick decomposition: owned + ghost atoms of a brick contain every atom within rc (minimum
    image) of an owned atom; ownership is a partition.  Pure host logic, no collective."""
    from pantea_b200.halo import BrickGrid
    rng = np.random.default_rng(11)
    box, rc, n = np.asarray([26.0, 26.0, 26.0]), 6.0, 1500
    grid = BrickGrid(box.tolist(), world)
    assert grid.dims == expect
    pos = torch.from_numpy(rng.uniform(0.0, 1.0, (n, 3)) * box)
    pos[:8] = torch.tensor([[0.0, 0.0, 0.0], [13.0, 13.0, 13.0], [25.999999999, 0.0, 13.0], [12.999999999, 13.0, 0.0],
                            [13.0, 0.0, 25.999999999], [6.0, 19.0, 13.0], [19.0, 6.0, 7.0], [7.0, 6.0, 19.0]],
                           dtype=torch.float64)
    owner = grid.owner(pos)
    assert owner.min() >= 0 and owner.max() < world
    d = pos[:, None, :] - pos[None, :, :]
    d -= torch.from_numpy(box) * torch.round(d / torch.from_numpy(box))
    within = (d ** 2).sum(-1) <= rc * rc
    n_ghost = 0
    for r in range(world):
        own = owner == r
        lo, hi = grid.bounds(r)
        for k in range(3):
            assert (pos[own, k] >= lo[k] - 1e-9).all() and (pos[own, k] <= hi[k] + 1e-9).all()
        ghost = grid.ghost_mask(pos, owner, r, rc)
        assert not (ghost & own).any()
        needed = within[own].any(0)
        assert not (needed & ~(own | ghost)).any()
        n_ghost += int(ghost.sum())
    if world == 1:
        assert n_ghost == 0
