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


class PipelinedEvaluator:
    """Throughput form of the loop: the host->device copy of batch ``i + 1`` runs on a copy stream
    while batch ``i`` computes, and results come back asynchronously into pinned buffers — the job the
    reference leaves to its ``DataLoader`` (pinned batches prepared ahead by worker processes,
    ``src/metatrain/cli/eval.py:199-214``) plus ``batch_to`` (``:240-244``).

        ev = PipelinedEvaluator(backend, target)
        ticket = ev.submit(host_batch)          # enqueue H2D (pinned host tensors)
        for nxt in more_batches:
            upcoming = ev.submit(nxt)           # next copy overlaps ...
            out = ev.run(ticket)                # ... this evaluation; out = pinned host tensors
            ticket = upcoming
    """

    def __init__(self, backend, target: str = "energy", gradients: bool = True, device: str = "cuda:0"):
        self.backend, self.target, self.gradients = backend, target, gradients
        self.device = torch.device(device)
        self.copy_stream = torch.cuda.Stream(self.device)
        self._out: Dict[str, torch.Tensor] = {}

    def submit(self, host: Dict[str, torch.Tensor]):
        """Start the host->device copy of one batch on the copy stream; returns a ticket."""
        with torch.cuda.stream(self.copy_stream):
            dev = {k: v.to(self.device, non_blocking=True) for k, v in host.items()}
            done = torch.cuda.Event()
            done.record(self.copy_stream)
        return dev, done

    def run(self, ticket) -> Dict[str, torch.Tensor]:
        """Evaluate a submitted batch; returns energies (and dE_dpos) as pinned HOST tensors that are
        valid when the call returns (one event wait, no device-wide synchronisation)."""
        dev, done = ticket
        compute = torch.cuda.current_stream(self.device)
        compute.wait_event(done)
        for v in dev.values():
            v.record_stream(compute)
        out = evaluate(self.backend, **dev, target=self.target, gradients=self.gradients)
        keys = ["energies"] + (["dE_dpos"] if self.gradients else [])
        for k in keys:
            buf = self._out.get(k)
            if buf is None or buf.shape != out[k].shape:
                buf = self._out[k] = torch.empty(out[k].shape, dtype=out[k].dtype).pin_memory()
            buf.copy_(out[k], non_blocking=True)
        back = torch.cuda.Event()
        back.record(compute)
        back.synchronize()
        return {k: self._out[k] for k in keys}


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
