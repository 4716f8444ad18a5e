"""Batched evaluator loop — the plain-tensor mirror of ``metatrain.cli.eval._eval_targets``
(``src/metatrain/cli/eval.py:148-310``).

Same semantics as the reference loop: batches of ``batch_size`` structures in dataset order
(the last batch may be smaller); the neighbor list is built on the host per batch *outside*
the timer (the reference does it in the DataLoader collate, ``eval.py:199-214``); 10 warm-up
batches (``:219-232``); each batch is moved to the device (``batch_to``, ``:240-244``), then
``evaluate_model`` is timed with a ``cuda.synchronize`` on the far side (``:246-256``); RMSE /
MAE of per-atom-normalised energies and of raw position gradients are accumulated
(``average_by_num_atoms``, ``RMSEAccumulator`` / ``MAEAccumulator``, ``:259-279``); the log line
reports total time and mean +- std "ms per atom" (``:302-310``).  Dataset readers, writers and
the TensorMap plumbing are out of scope: structures are dicts (``Z``, ``positions``, ``cell``,
``pbc``) and optional targets dicts (``energy``, ``forces``).
"""
import itertools
import logging
import time
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from .evaluate import evaluate
from .systems import make_batch

logger = logging.getLogger(__name__)


class _Accumulator:
    """Sum of squared / absolute errors and element counts per key (metrics.py:51,244)."""

    def __init__(self) -> None:
        self.sse: Dict[str, float] = {}
        self.sae: Dict[str, float] = {}
        self.count: Dict[str, int] = {}

    def update(self, key: str, pred: torch.Tensor, target: torch.Tensor) -> None:
        diff = (pred.double() - target.double()).reshape(-1)
        self.sse[key] = self.sse.get(key, 0.0) + float((diff * diff).sum())
        self.sae[key] = self.sae.get(key, 0.0) + float(diff.abs().sum())
        self.count[key] = self.count.get(key, 0) + diff.numel()

    def finalize(self) -> Dict[str, float]:
        out = {}
        for key, n in self.count.items():
            out[f"{key} RMSE"] = (self.sse[key] / n) ** 0.5
            out[f"{key} MAE"] = self.sae[key] / n
        return out


def eval_targets(
    backend,
    structures: Sequence[dict],
    targets: Optional[Sequence[dict]] = None,
    target: str = "energy",
    batch_size: int = 1,
    gradients: bool = True,
    warm_up: bool = True,
    device: str = "cuda:0",
) -> Dict[str, object]:
    """Evaluate ``backend`` on ``structures``; returns predictions, metrics and timings."""
    if len(structures) == 0:
        logger.info("This dataset is empty. No evaluation will be performed.")
        return {"energies": [], "forces": [], "metrics": {}, "ms_per_atom": (float("nan"), float("nan"))}
    cutoff = backend.cutoff
    batches = [list(range(i, min(i + batch_size, len(structures))))
               for i in range(0, len(structures), batch_size)]

    def host_batch(idx: List[int]):
        return make_batch([structures[i] for i in idx], cutoff, pin_memory=True)

    if warm_up:
        logger.info("Warming up the model with 10 batches...")
        for idx in itertools.islice(itertools.cycle(batches), 10):
            evaluate(backend, **{k: v.to(device) for k, v in host_batch(idx).items()},
                     target=target, gradients=gradients)
    acc = _Accumulator()
    energies, forces, per_atom = [], [], []
    total = 0.0
    for idx in batches:
        host = host_batch(idx)                                  # neighbor list: outside the timer
        dev = {k: v.to(device, non_blocking=True) for k, v in host.items()}   # batch_to
        torch.cuda.synchronize()
        start = time.time()
        out = evaluate(backend, **dev, target=target, gradients=gradients)
        torch.cuda.synchronize()
        taken = time.time() - start
        n_atoms = [len(structures[i]["Z"]) for i in idx]
        total += taken
        per_atom.append(taken / sum(n_atoms))
        e = out["energies"].cpu()
        energies.extend(e[k] for k in range(len(idx)))
        if gradients:
            f = -out["dE_dpos"].cpu()
            forces.extend(torch.split(f, n_atoms))
        if targets is not None:
            for k, i in enumerate(idx):
                if "energy" in targets[i]:
                    acc.update(f"{target} (per atom)", e[k] / n_atoms[k],
                               torch.as_tensor(targets[i]["energy"]).reshape(-1) / n_atoms[k])
                if gradients and "forces" in targets[i]:
                    acc.update(f"{target} forces", forces[len(forces) - len(idx) + k],
                               torch.as_tensor(targets[i]["forces"]))
    per_atom = np.array(per_atom)
    mean, std = float(per_atom.mean()), float(per_atom.std())
    logger.info(f"Evaluation time: {total:.2f} s [{1000.0 * mean:.4f} ± {1000.0 * std:.4f} ms per atom]")
    return {"energies": energies, "forces": forces, "metrics": acc.finalize(),
            "ms_per_atom": (1000.0 * mean, 1000.0 * std), "total_time_s": total}
