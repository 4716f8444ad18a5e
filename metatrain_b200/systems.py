"""Synthetic periodic boxes and batching helpers for the benchmark / evaluator loop.

Inputs named by BASELINE.json: the 64-atom diamond-Si box (synthesised — no Si file exists
in the reference's tests/resources), and the water boxes obtained by tiling
the reference's 384-atom periodic water fixture (``tests/resources/periodic_water.data``;
positions stored in ``metatrain_b200/data/water_384.npz`` by ``tests/golden/make_golden.py``)
3x3x3 -> 10 368 atoms, 6x6x7 -> 96 768 atoms.
"""
import os
from typing import Dict, List, Sequence

import numpy as np
import torch

from .neighbors import neighbor_list

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")


def water_384() -> Dict[str, np.ndarray]:
    z = np.load(os.path.join(_DATA, "water_384.npz"))
    return dict(Z=z["Z"], positions=z["positions"], cell=z["cell"], pbc=True)


def replicate(frame: Dict[str, np.ndarray], reps: Sequence[int], jitter: float = 0.0,
              seed: int = 1) -> Dict[str, np.ndarray]:
    na, nb, nc = reps
    cell = np.asarray(frame["cell"], dtype=np.float64)
    shifts = np.array([[a, b, c] for a in range(na) for b in range(nb) for c in range(nc)])
    pos = (frame["positions"][None, :, :] + (shifts @ cell)[:, None, :]).reshape(-1, 3)
    if jitter > 0:
        pos = pos + np.random.default_rng(seed).normal(0.0, jitter, pos.shape)
    return dict(Z=np.tile(frame["Z"], len(shifts)), positions=pos,
                cell=cell * np.array([[na], [nb], [nc]]), pbc=True)


def silicon_box(reps: int = 2, a: float = 5.431, sigma: float = 0.05, seed: int = 0):
    fcc = np.array([[0, 0, 0], [0, 0.5, 0.5], [0.5, 0, 0.5], [0.5, 0.5, 0]])
    basis = np.concatenate([fcc, fcc + 0.25]) * a
    unit = dict(Z=np.full(8, 14, dtype=np.int64), positions=basis, cell=np.eye(3) * a, pbc=True)
    box = replicate(unit, (reps, reps, reps))
    rng = np.random.default_rng(seed)
    box["positions"] = box["positions"] + rng.normal(0.0, sigma, box["positions"].shape)
    return box


def make_batch(frames: List[Dict[str, np.ndarray]], cutoff: float, pin_memory: bool = False
               ) -> Dict[str, torch.Tensor]:
    """Host tensors in the layout of ``concatenate_structures``
    (src/metatrain/pet/modules/structures.py:17-112): fp32 positions/cells, int32 indices."""
    pos, cen, nei, sh, Z, sysi, cells = [], [], [], [], [], [], []
    off = 0
    for k, f in enumerate(frames):
        i, j, S = neighbor_list(f["positions"], f["cell"], bool(f["pbc"]), cutoff)
        pos.append(np.asarray(f["positions"], dtype=np.float32))
        cen.append(i + off)
        nei.append(j + off)
        sh.append(S)
        Z.append(f["Z"])
        sysi.append(np.full(len(f["Z"]), k))
        cells.append(np.asarray(f["cell"], dtype=np.float32))
        off += len(f["Z"])
    batch = dict(
        positions=torch.from_numpy(np.concatenate(pos)),
        centers=torch.from_numpy(np.concatenate(cen).astype(np.int32)),
        neighbors=torch.from_numpy(np.concatenate(nei).astype(np.int32)),
        species=torch.from_numpy(np.concatenate(Z).astype(np.int32)),
        cells=torch.from_numpy(np.stack(cells)),
        cell_shifts=torch.from_numpy(np.concatenate(sh).astype(np.int32).reshape(-1, 3)),
        system_indices=torch.from_numpy(np.concatenate(sysi).astype(np.int64)),
    )
    if pin_memory:
        batch = {k: v.pin_memory() for k, v in batch.items()}
    return batch
