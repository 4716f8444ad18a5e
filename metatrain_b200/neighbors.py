"""Host-side neighbor list (the role ``vesin`` plays for the reference).

The reference builds a full neighbor list per ``System`` on the CPU with the C library
vesin (``src/metatrain/utils/neighbor_lists.py:125-201``, call ``:131``) inside the
DataLoader, outside ``mtt eval``'s timer.  This is the equivalent host component: a
linked-cell search in numpy that returns every ordered pair ``(i, j, S)`` with
``|r_j + S.cell - r_i| <= cutoff`` (no ``(i, i, 0)``), sorted by centre, for orthorhombic
and triclinic cells of any size (several periodic images per pair when the cell is
smaller than the cutoff).  A CUDA cell-list builder is the next row of SURVEY.md 8(f).
"""
from typing import Tuple

import numpy as np


def _ragged_pairs(count_per_row: np.ndarray, start_per_row: np.ndarray):
    """For row r emit (r, start[r] + k) for k < count[r], vectorised."""
    total = int(count_per_row.sum())
    rows = np.repeat(np.arange(len(count_per_row)), count_per_row)
    first = np.cumsum(count_per_row) - count_per_row
    within = np.arange(total) - np.repeat(first, count_per_row)
    return rows, np.repeat(start_per_row, count_per_row) + within


def neighbor_list(positions, cell, periodic: bool, cutoff: float
                  ) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Full neighbor list ``(centers, neighbors, cell_shifts)`` (int64), centre-sorted."""
    pos = np.asarray(positions, dtype=np.float64)
    n = len(pos)
    empty = (np.zeros(0, np.int64), np.zeros(0, np.int64), np.zeros((0, 3), np.int64))
    if n == 0:
        return empty
    if not periodic:
        # open boundaries: a bounding box plays the role of the cell, no images
        lo = pos.min(0) - 1e-6
        extent = np.maximum(pos.max(0) - lo + 1e-6, cutoff)
        cell_m = np.diag(extent)
        frac = (pos - lo) / extent
        wrap = np.zeros((n, 3), np.int64)
        base = frac
    else:
        cell_m = np.asarray(cell, dtype=np.float64)
        frac = pos @ np.linalg.inv(cell_m)
        wrap = np.floor(frac).astype(np.int64)
        base = frac - wrap
    vol = abs(np.linalg.det(cell_m))
    heights = np.array([vol / np.linalg.norm(np.cross(cell_m[(k + 1) % 3], cell_m[(k + 2) % 3]))
                        for k in range(3)])
    nbins = np.maximum(1, np.floor(heights / cutoff).astype(np.int64))
    reach = np.ceil(cutoff / (heights / nbins) - 1e-12).astype(np.int64)
    reach = np.maximum(reach, 1)
    bins = np.minimum((base * nbins).astype(np.int64), nbins - 1)
    flat = (bins[:, 0] * nbins[1] + bins[:, 1]) * nbins[2] + bins[:, 2]
    order = np.argsort(flat, kind="stable")
    n_cells = int(nbins.prod())
    bin_count = np.bincount(flat, minlength=n_cells)
    bin_start = np.cumsum(bin_count) - bin_count
    wrapped = base @ cell_m if periodic else pos
    out_i, out_j, out_s = [], [], []
    rng = [range(-int(r), int(r) + 1) for r in reach]
    for da in rng[0]:
        for db in rng[1]:
            for dc in rng[2]:
                tgt = bins + np.array([da, db, dc])
                img = np.floor_divide(tgt, nbins)
                if not periodic:
                    ok = (img == 0).all(axis=1)
                    if not ok.any():
                        continue
                else:
                    ok = np.ones(n, bool)
                tb = tgt - img * nbins
                tflat = (tb[:, 0] * nbins[1] + tb[:, 1]) * nbins[2] + tb[:, 2]
                cnt = np.where(ok, bin_count[np.where(ok, tflat, 0)], 0)
                ii, jpos = _ragged_pairs(cnt, bin_start[np.where(ok, tflat, 0)])
                if len(ii) == 0:
                    continue
                jj = order[jpos]
                shift = img[ii]
                d = wrapped[jj] + shift @ cell_m - wrapped[ii]
                keep = (np.einsum("ij,ij->i", d, d) <= cutoff * cutoff)
                keep &= ~((ii == jj) & (shift == 0).all(axis=1))
                if keep.any():
                    out_i.append(ii[keep])
                    out_j.append(jj[keep])
                    out_s.append(shift[keep])
    if not out_i:
        return empty
    i = np.concatenate(out_i)
    j = np.concatenate(out_j)
    s = np.concatenate(out_s)
    # shifts relative to the caller's (unwrapped) positions
    s = s - wrap[j] + wrap[i]
    # exact distance test on the caller's coordinates (what the model will recompute)
    r = pos[j] - pos[i] + (s @ cell_m if periodic else 0.0)
    keep = np.linalg.norm(r, axis=1) <= cutoff
    i, j, s = i[keep], j[keep], s[keep]
    key = np.lexsort((s[:, 2], s[:, 1], s[:, 0], j, i))
    return i[key].astype(np.int64), j[key].astype(np.int64), s[key].astype(np.int64)
