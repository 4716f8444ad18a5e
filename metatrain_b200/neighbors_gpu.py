"""Device-side neighbor list: the GPU counterpart of
``metatrain_b200.neighbors.neighbor_list`` / vesin (``src/metatrain/utils/neighbor_lists.py:131``).

Returns ``(centers, neighbors, cell_shifts)`` as int32 CUDA tensors, grouped by centre, ready for
``B200PETBackend.preprocess``: positions never leave the GPU and the ~8 MB neighbor list of a
10k-atom box is never copied over PCIe.  Pairs are accepted with a 2e-6 relative margin on the
cutoff; the model's own symmetric filter (``petb200_nl_filter_count``) takes the final decision.
"""
import ctypes
from typing import Tuple

import torch

from . import lib
from .lib import call, ptr

Tensor = torch.Tensor


def neighbor_list_gpu(positions: Tensor, cell: Tensor, periodic: bool, cutoff: float
                      ) -> Tuple[Tensor, Tensor, Tensor]:
    """Full neighbor list of one structure on the device of ``positions``."""
    if not positions.is_cuda:
        raise RuntimeError("neighbor_list_gpu: expected CUDA positions (use neighbors.neighbor_list on the host)")
    dev = positions.device
    pos = positions.detach().to(torch.float32).contiguous()
    n = pos.shape[0]
    i32 = torch.int32
    empty = (torch.empty(0, dtype=i32, device=dev), torch.empty(0, dtype=i32, device=dev),
             torch.empty((0, 3), dtype=i32, device=dev))
    if n == 0:
        return empty
    if periodic:
        cell_h = [float(v) for v in cell.detach().reshape(-1).tolist()]
        origin_h = [0.0, 0.0, 0.0]
    else:
        lo = pos.min(0).values - 1e-3
        ext = torch.clamp(pos.max(0).values - lo + 1e-3, min=float(cutoff))
        origin_h = [float(v) for v in lo.tolist()]
        e = [float(v) for v in ext.tolist()]
        cell_h = [e[0], 0.0, 0.0, 0.0, e[1], 0.0, 0.0, 0.0, e[2]]
    cell_c = (ctypes.c_float * 9)(*cell_h)
    origin_c = (ctypes.c_float * 3)(*origin_h)
    handle = lib.load()
    n_bins = handle.petb200_nl_num_bins(cell_c, int(periodic), float(cutoff), n)
    if n_bins <= 0:
        raise ValueError("neighbor_list_gpu: singular cell")
    ws_bytes = handle.petb200_nl_workspace(n, n_bins)
    workspace = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    offsets = torch.empty(n + 1, dtype=i32, device=dev)
    call("nl_count", ptr(pos), n, cell_c, origin_c, int(periodic), float(cutoff), ptr(workspace),
         ws_bytes, ptr(offsets))
    n_pairs = int(offsets[n].item())  # sizes the output: one 4-byte device->host read
    if n_pairs == 0:
        return empty
    centers = torch.empty(n_pairs, dtype=i32, device=dev)
    neighbors = torch.empty(n_pairs, dtype=i32, device=dev)
    shifts = torch.empty((n_pairs, 3), dtype=i32, device=dev)
    call("nl_fill", n, cell_c, origin_c, int(periodic), float(cutoff), ptr(workspace), ws_bytes,
         ptr(offsets), ptr(centers), ptr(neighbors), ptr(shifts))
    return centers, neighbors, shifts


class VerletNeighborList:
    """Skin-based reuse of the device neighbor list across MD steps (a Verlet list).

    The list is built for ``cutoff + skin`` and handed out unchanged while no atom has moved more
    than ``skin / 2`` since the build (and the cell is unchanged) — then no pair inside the
    model cutoff can be missing.  The backend treats such a list like the reference treats a
    non-strict neighbor list: pairs beyond the model cutoff are dropped in ``preprocess``
    (``src/metatrain/pet/modules/structures.py:265-272``; here ``petb200_nl_filter_count``).
    One 4-byte device->host read per step decides whether to rebuild.
    """

    def __init__(self, cutoff: float, skin: float = 0.5, periodic: bool = True):
        if skin < 0:
            raise ValueError("VerletNeighborList: skin must be non-negative")
        self.cutoff, self.skin, self.periodic = float(cutoff), float(skin), bool(periodic)
        self._ref_pos = None
        self._ref_cell = None
        self._lists = None
        self.n_builds = 0
        self.n_reuses = 0

    def update(self, positions: Tensor, cell: Tensor) -> Tuple[Tensor, Tensor, Tensor]:
        """``(centers, neighbors, cell_shifts)`` valid for ``positions`` (non-strict: may contain
        pairs up to ``cutoff + skin``)."""
        pos = positions.detach()
        stale = (self._lists is None or self._ref_pos.shape != pos.shape
                 or not torch.equal(self._ref_cell, cell.detach()))
        if not stale:
            moved = float((pos - self._ref_pos).square().sum(dim=1).max().sqrt())
            stale = moved > 0.5 * self.skin
        if stale:
            self._lists = neighbor_list_gpu(pos, cell, self.periodic, self.cutoff + self.skin)
            self._ref_pos = pos.clone()
            self._ref_cell = cell.detach().clone()
            self.n_builds += 1
        else:
            self.n_reuses += 1
        return self._lists
