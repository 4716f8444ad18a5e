"""Atom-sharded evaluation of ONE large structure over the GPUs of a box.

The reference has no multi-GPU inference (it points users at LAMMPS domain decomposition,
``src/metatrain/pet/modules/transformer.py:122-125``).  PET's per-atom transformer blocks are
independent; the only cross-atom dataflow is the reversed-message gather
``out.flat[reverse_neighbor_index]`` once per GNN layer (``src/metatrain/pet/modules/backend.py:559-566``)
and, in the force path, the scatter of edge gradients onto the neighbour atom
(autograd of ``structures.py:220``).  So:

* atoms are split into ``world`` spatial bricks (recursive equal-count bisection along the cell
  axes, ``brick_owner``; or slabs along the longest axis, ``slab_owner``);
  a rank owns its atoms and **all edges centred on them** (full CSR rows -> attention stays local);
* a *halo edge* is an owned edge ``(i -> j, S)`` whose neighbour ``j`` lives on a peer ``p``.  Its
  reversed edge ``(j -> i, -S)`` is one of ``p``'s halo edges towards us — the two sets are
  mirror images, so one index list per peer serves both directions;
* per exchange each rank sends the rows of its halo edges towards ``p`` (ordered by the key of the
  *reversed* edge) and receives ``p``'s rows into ghost slots (ordered by the key of its own halo
  edges), i.e. an all-to-all-v over NVLink (``torch.distributed.all_to_all_single`` = grouped NCCL
  send/recv).  Exchanges per step: one per GNN layer forward (128 floats per halo edge), one per
  GNN layer backward (128 floats), one for the edge gradients (3 floats); plus one all-reduce of
  the per-structure energy and one of the [N, 3] position gradient;
* everything else (geometry, GEMMs, attention, readout) runs unchanged on the local rows.
"""
import contextlib
from dataclasses import dataclass
from typing import Dict, List, Optional

import numpy as np
import torch
import torch.distributed as dist

Tensor = torch.Tensor


#: optional profiler: a list that receives ``(label, start_event, end_event, bytes_sent)`` per collective
#: (bench.py's instrumented steps); None = no events recorded
comm_timer: Optional[list] = None


@contextlib.contextmanager
def _timed(label: str, nbytes: int):
    """Brackets one collective with CUDA events when ``comm_timer`` is set."""
    if comm_timer is None:
        yield
        return
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    yield
    b.record()
    comm_timer.append((label, a, b, nbytes))


@dataclass
class Halo:
    """Index lists of one rank's halo edges (all tensors on the compute device)."""
    n_ghost: int                 # ghost rows appended to edge arrays (= number of halo edges)
    send_idx: Tensor             # [n_ghost] local CSR edge ids, grouped by peer, send order
    send_splits: List[int]       # rows sent to each rank
    recv_splits: List[int]       # rows received from each rank (same numbers, by symmetry)
    halo_edges: Tensor           # [n_ghost] local CSR edge ids in ghost-slot (receive) order
    group: Optional[object] = None

    def exchange(self, rows: Tensor, out: Optional[Tensor] = None) -> Tensor:
        """rows = x[send_idx] ([n_ghost, D]) -> the peers' rows for our ghost slots."""
        rows = rows.contiguous()
        if out is None:
            out = torch.empty_like(rows)
        if self.n_ghost == 0 and sum(self.send_splits) == 0:
            # still a collective: every rank must take part
            pass
        with _timed("halo_all_to_all", rows.numel() * rows.element_size()):
            dist.all_to_all_single(out, rows, self.recv_splits, self.send_splits, group=self.group)
        return out


def slab_owner(positions: np.ndarray, cell: np.ndarray, world: int) -> np.ndarray:
    """Owner rank of every atom: equal-count slabs along the longest cell vector."""
    frac = positions @ np.linalg.inv(cell)
    frac -= np.floor(frac)
    axis = int(np.argmax(np.linalg.norm(cell, axis=1)))
    order = np.argsort(frac[:, axis], kind="stable")
    owner = np.empty(len(positions), dtype=np.int64)
    owner[order] = (np.arange(len(positions)) * world) // len(positions)
    return owner


def _grid_factors(world: int, lengths: np.ndarray):
    """(pa, pb, pc) with pa pb pc = world minimising the cut surface of bricks in a cell with the
    given edge lengths."""
    best, best_cost = (1, 1, world), None
    for pa in range(1, world + 1):
        if world % pa:
            continue
        for pb in range(1, world // pa + 1):
            if (world // pa) % pb:
                continue
            pc = world // (pa * pb)
            # cut planes per axis x area of one plane
            cost = ((pa > 1) * pa * lengths[1] * lengths[2] + (pb > 1) * pb * lengths[0] * lengths[2]
                    + (pc > 1) * pc * lengths[0] * lengths[1])
            if best_cost is None or cost < best_cost - 1e-9:
                best, best_cost = (pa, pb, pc), cost
    return best


def brick_owner(positions: np.ndarray, cell: np.ndarray, world: int) -> np.ndarray:
    """Owner rank of every atom: ``pa x pb x pc`` bricks (recursive equal-count bisection along
    the three cell axes), which cut ~(pa + pb + pc) planes instead of the ``world`` planes of
    slabs: at 8 ranks a cubic box loses 3 x 2 half-planes of halo instead of 8."""
    frac = positions @ np.linalg.inv(cell)
    frac -= np.floor(frac)
    grid = _grid_factors(world, np.linalg.norm(cell, axis=1))
    owner = np.zeros(len(positions), dtype=np.int64)
    groups = [np.arange(len(positions))]
    for axis, parts in enumerate(grid):
        nxt = []
        for ids in groups:
            order = ids[np.argsort(frac[ids, axis], kind="stable")]
            split = (np.arange(len(order)) * parts) // max(len(order), 1)
            for q in range(parts):
                nxt.append(order[split == q])
        groups = nxt
    for r, ids in enumerate(groups):
        owner[ids] = r
    return owner


@dataclass
class Shard:
    """One rank's part of a structure, host side (numpy) + halo description."""
    rank: int
    world: int
    n_atoms_global: int
    own_ids: np.ndarray          # global ids of owned atoms (ascending)
    local_ids: np.ndarray        # global ids of [owned | ghost] atoms = rows of the local arrays
    centers: np.ndarray          # local edge list (local atom ids), neighbor-list order
    neighbors: np.ndarray
    cell_shifts: np.ndarray
    edge_global: np.ndarray      # index of each local edge in the global neighbor list
    # halo description in *input edge order* (converted to CSR order on the device)
    halo_recv: List[np.ndarray]  # per peer: local input-edge ids in ghost-slot order
    halo_send: List[np.ndarray]  # per peer: local input-edge ids in send order


def build_shard(positions: np.ndarray, cell: np.ndarray, nl, rank: int, world: int,
                partition: str = "bricks") -> Shard:
    """Partition a periodic structure.  ``nl = (i, j, S)`` is its full neighbor list sorted by
    centre (``metatrain_b200.neighbors.neighbor_list``); ``partition`` = "bricks" or "slabs"."""
    gi, gj, gs = nl
    n = len(positions)
    split = {"bricks": brick_owner, "slabs": slab_owner}[partition]
    owner = split(np.asarray(positions, dtype=np.float64), np.asarray(cell, dtype=np.float64), world)
    own_ids = np.nonzero(owner == rank)[0]
    mine = np.nonzero(owner[gi] == rank)[0]          # local edges, still sorted by centre
    li, lj, ls = gi[mine], gj[mine], gs[mine]
    ghost_ids = np.unique(lj[owner[lj] != rank])
    local_ids = np.concatenate([own_ids, ghost_ids])
    lookup = np.full(n, -1, dtype=np.int64)
    lookup[local_ids] = np.arange(len(local_ids))
    halo_recv, halo_send = [], []
    peer_of_edge = owner[lj]
    for p in range(world):
        sel = np.nonzero(peer_of_edge == p)[0] if p != rank else np.zeros(0, dtype=np.int64)
        if len(sel):
            # ghost-slot order: by (centre, neighbour, shift) of our own halo edge
            k_recv = np.lexsort((ls[sel, 2], ls[sel, 1], ls[sel, 0], lj[sel], li[sel]))
            # send order: by the key of the reversed edge (neighbour, centre, -shift)
            k_send = np.lexsort((-ls[sel, 2], -ls[sel, 1], -ls[sel, 0], li[sel], lj[sel]))
            halo_recv.append(sel[k_recv])
            halo_send.append(sel[k_send])
        else:
            halo_recv.append(np.zeros(0, dtype=np.int64))
            halo_send.append(np.zeros(0, dtype=np.int64))
    return Shard(rank, world, n, own_ids, local_ids, lookup[li], lookup[lj], ls, mine,
                 halo_recv, halo_send)


def attach_halo(topo, shard: Shard, device, group=None, device_lists=None) -> None:
    """Translate the shard's halo lists to CSR edge ids and patch ``topo.rev`` so that the
    reverse of a halo edge points at its ghost slot.  Two ghost bases are used: the forward
    token buffer is laid out [E edge rows | N centre rows | H ghost rows], backward edge
    arrays are [E | H]."""
    E, N = topo.n_edges, topo.n_atoms
    # CSR edge k came from input edge perm[k]  ->  inverse map input edge -> CSR edge
    inv = torch.empty(len(shard.centers), dtype=torch.int64, device=device)
    inv[topo.perm.long()] = torch.arange(E, device=device)
    if device_lists is not None and "halo_recv" in device_lists:
        # (the halo lists only depend on the shard: uploaded once, or with the step's other inputs)
        recv_d, send_d = device_lists["halo_recv"], device_lists["halo_send"]
    else:
        recv_in = np.concatenate(shard.halo_recv) if shard.world > 1 else np.zeros(0, dtype=np.int64)
        send_in = np.concatenate(shard.halo_send) if shard.world > 1 else np.zeros(0, dtype=np.int64)
        recv_d, send_d = torch.from_numpy(recv_in).to(device), torch.from_numpy(send_in).to(device)
    halo_edges = inv[recv_d]
    send_idx = inv[send_d]
    H = int(halo_edges.numel())
    slots = torch.arange(H, device=device, dtype=torch.int32)
    rev_fwd = topo.rev.clone()
    rev_bwd = topo.rev.clone()
    rev_fwd[halo_edges] = slots + (E + N)
    rev_bwd[halo_edges] = slots + E
    topo.rev = rev_fwd
    topo.rev_bwd = rev_bwd
    topo.halo = Halo(H, send_idx, [len(a) for a in shard.halo_send],
                     [len(a) for a in shard.halo_recv], halo_edges, group)


def shard_to_host_tensors(shard: Shard, pin_memory: bool = False) -> Dict[str, Tensor]:
    """The rank's index lists as (optionally pinned) host tensors."""
    cat = lambda parts: (np.concatenate(parts) if shard.world > 1 else np.zeros(0, dtype=np.int64))  # noqa: E731
    out = dict(ids=torch.from_numpy(np.ascontiguousarray(shard.local_ids)),
               centers=torch.from_numpy(shard.centers.astype(np.int32)),
               neighbors=torch.from_numpy(shard.neighbors.astype(np.int32)),
               shifts=torch.from_numpy(np.ascontiguousarray(shard.cell_shifts.astype(np.int32))),
               halo_recv=torch.from_numpy(np.ascontiguousarray(cat(shard.halo_recv)).astype(np.int64)),
               halo_send=torch.from_numpy(np.ascontiguousarray(cat(shard.halo_send)).astype(np.int64)))
    return {k: v.pin_memory() for k, v in out.items()} if pin_memory else out


def shard_to_device(shard: Shard, device, host: Optional[Dict[str, Tensor]] = None) -> Dict[str, Tensor]:
    host = host if host is not None else shard_to_host_tensors(shard)
    return {k: v.to(device, non_blocking=True) for k, v in host.items()}


def evaluate_sharded(backend, shard: Shard, positions: Tensor, species: Tensor, cell: Tensor,
                     target: str = "energy", gradients: bool = True, group=None,
                     device_lists: Optional[Dict[str, Tensor]] = None) -> Dict[str, Tensor]:
    """Energy (+ dE/dr for ALL atoms, summed over ranks) of one structure sharded by atoms.

    ``positions`` / ``species`` are the full (replicated) arrays on this rank's device; every rank
    returns the same ``energies [1, P]`` and ``dE_dpos [N, 3]``."""
    from . import engine
    from .backend import _EdgeGeometry, _Features, _Predict

    dev = positions.device
    pos = positions.detach().clone().requires_grad_(gradients)
    lists = device_lists if device_lists is not None else shard_to_device(shard, dev)
    ids, centers, neighbors, shifts = (lists[k] for k in ("ids", "centers", "neighbors", "shifts"))
    pos_local = pos.index_select(0, ids)
    z_nodes = backend.species_to_species_index[species.long()].index_select(0, ids)
    cells = cell.reshape(1, 3, 3)
    sysidx = torch.zeros(len(shard.local_ids), dtype=torch.int64, device=dev)
    backend._check_inference()
    if backend.num_neighbors_adaptive is not None or backend.system_conditioning is not None:
        raise NotImplementedError("evaluate_sharded: adaptive cutoff / system conditioning are not built "
                                  "for atom-sharded runs")
    topo = engine.build_topology(pos_local, centers, neighbors, shifts, cells, sysidx, z_nodes,
                                 backend.cutoff, check_symmetric=False,
                                 n_rows=len(shard.own_ids))
    attach_halo(topo, shard, dev, group, lists)
    vec, dist_, fc = _EdgeGeometry.apply(pos_local, cells, topo, backend.cutoff,
                                         backend.cutoff_width, backend._cutoff_id)
    h, m = _Features.apply(vec, dist_, fc, backend, topo, None, None)
    atomic = _Predict.apply(h, m, fc, backend, topo, target, 0, None)    # [n_own, P]
    energy = atomic.sum(dim=0, keepdim=True)
    out = {"atomic_local": atomic.detach()}
    if gradients:
        # one all-reduce for the position gradient (every row is non-zero on exactly one rank: the
        # force scatter already pulled the halo contributions in) and the per-structure energy
        (grad,) = torch.autograd.grad(energy.sum(), pos)
        packed = torch.cat([grad.reshape(-1), energy.detach().reshape(-1)])
        with _timed("all_reduce", packed.numel() * packed.element_size()):
            dist.all_reduce(packed, group=group)
        n3 = grad.numel()
        out["dE_dpos"] = packed[:n3].view_as(grad)
        out["energies"] = packed[n3:].view_as(energy)
    else:
        total = energy.detach().clone()
        dist.all_reduce(total, group=group)
        out["energies"] = total
    return out
