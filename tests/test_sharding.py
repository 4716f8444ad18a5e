"""world_size-2 gloo tests (CPU) of the atom-sharding host logic: partition, halo index lists
and the all-to-all-v exchange ordering (metatrain_b200/sharded.py)."""
import os
import tempfile

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from metatrain_b200.neighbors import neighbor_list
from metatrain_b200.sharded import Halo, brick_owner, build_shard, slab_owner
from metatrain_b200.systems import replicate, water_384
from oracle.pet_oracle import reverse_edge_map_sorted


def _worker(rank, world, init_file, result_file):
    dist.init_process_group("gloo", init_method=f"file://{init_file}", rank=rank, world_size=world)
    box = replicate(water_384(), (2, 1, 1))
    nl = neighbor_list(box["positions"], box["cell"], True, 4.5)
    gi, gj, gs = nl
    n_edges = len(gi)
    rev_global = reverse_edge_map_sorted(torch.from_numpy(gi), torch.from_numpy(gj),
                                         torch.from_numpy(gs)).numpy()
    shard = build_shard(box["positions"], box["cell"], nl, rank, world, partition="slabs")
    owner = slab_owner(box["positions"], box["cell"], world)
    ok = True
    # every atom has exactly one owner, equal counts
    counts = np.bincount(owner, minlength=world)
    ok &= counts.max() - counts.min() <= 1
    ok &= np.array_equal(shard.own_ids, np.nonzero(owner == rank)[0])
    # local edges are exactly the rows of owned atoms
    ok &= np.array_equal(shard.edge_global, np.nonzero(owner[gi] == rank)[0])
    ok &= np.array_equal(shard.local_ids[shard.centers], gi[shard.edge_global])
    ok &= np.array_equal(shard.local_ids[shard.neighbors], gj[shard.edge_global])
    ok &= bool((shard.centers < len(shard.own_ids)).all())
    # exchange a tagged payload: row e carries its global edge id
    x_global = torch.arange(n_edges, dtype=torch.float32)[:, None] * torch.ones(1, 4)
    x_loc = x_global[shard.edge_global]
    send_in = np.concatenate(shard.halo_send)
    recv_in = np.concatenate(shard.halo_recv)
    halo = Halo(len(recv_in), torch.from_numpy(send_in), [len(a) for a in shard.halo_send],
                [len(a) for a in shard.halo_recv], torch.from_numpy(recv_in))
    got = halo.exchange(x_loc[halo.send_idx])
    # ghost slot k must now hold the row of the global reverse of our k-th halo edge
    expect = rev_global[shard.edge_global[recv_in]]
    ok &= np.array_equal(got[:, 0].numpy().astype(np.int64), expect)
    # halo edges are exactly the local edges whose neighbour is not ours
    is_halo = owner[gj[shard.edge_global]] != rank
    ok &= np.array_equal(np.sort(recv_in), np.nonzero(is_halo)[0])
    ok &= np.array_equal(np.sort(send_in), np.nonzero(is_halo)[0])
    # and the reverse of every non-halo local edge is local
    local_set = set(shard.edge_global.tolist())
    ok &= all(int(rev_global[e]) in local_set for e in shard.edge_global[~is_halo][:2000])
    with open(f"{result_file}.{rank}", "w") as fh:
        fh.write(f"{int(ok)} {len(recv_in)} {len(shard.edge_global)}")
    dist.destroy_process_group()


def test_halo_lists_and_exchange_two_ranks():
    world = 2
    with tempfile.TemporaryDirectory() as tmp:
        init_file, result_file = os.path.join(tmp, "init"), os.path.join(tmp, "res")
        mp.spawn(_worker, args=(world, init_file, result_file), nprocs=world, join=True)
        res = [open(f"{result_file}.{r}").read().split() for r in range(world)]
    assert all(r[0] == "1" for r in res), res
    assert int(res[0][1]) == int(res[1][1]) > 0          # mirror-image halo sets
    assert int(res[0][2]) + int(res[1][2]) == 2 * 14520   # all edges owned exactly once


def test_slab_owner_balanced_for_eight_ranks():
    box = replicate(water_384(), (2, 2, 2))
    owner = slab_owner(box["positions"], box["cell"], 8)
    counts = np.bincount(owner, minlength=8)
    assert counts.sum() == len(owner) and counts.max() - counts.min() <= 1


def test_brick_owner_balanced_and_compact():
    box = replicate(water_384(), (2, 2, 2))
    nl = neighbor_list(box["positions"], box["cell"], True, 4.5)
    for world in (2, 4, 8):
        owner = brick_owner(box["positions"], box["cell"], world)
        counts = np.bincount(owner, minlength=world)
        assert counts.sum() == len(owner) and counts.max() - counts.min() <= world
    # 8 bricks cut fewer edges than 8 slabs of the same cubic box
    cut = lambda own: int((own[nl[0]] != own[nl[1]]).sum())  # noqa: E731
    assert cut(brick_owner(box["positions"], box["cell"], 8)) < cut(slab_owner(box["positions"], box["cell"], 8))
    # every shard sees mirror-image halo sets with every peer
    shards = [build_shard(box["positions"], box["cell"], nl, r, 4) for r in range(4)]
    for a in range(4):
        for b in range(4):
            assert len(shards[a].halo_send[b]) == len(shards[b].halo_recv[a])
