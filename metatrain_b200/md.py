"""MD-loop evaluator: energy + forces of one structure, step after step.

The reference's MD path runs, every step, the host neighbor list, ``preprocess`` with its
host synchronisation, and ~600 eager launches (``src/metatrain/utils/neighbor_lists.py:125-201``,
``src/metatrain/pet/modules/structures.py:292-294``).  Here the neighbor list is a device-side
Verlet list (``neighbors_gpu.VerletNeighborList``) and, while it is reused, the whole step —
edge geometry, features, readout, per-structure sum and the backward to the positions — is
replayed as ONE CUDA graph: no host synchronisation, no per-kernel launch cost.  Small systems
(hundreds of atoms), where the eager step is bound by host launch overhead, gain the most.

The captured topology holds every pair of the skin list, including those currently beyond
the model cutoff: their cutoff factor is exactly 0, so they carry zero weight in the readout
and 1e-15 (the reference's own padding weight, ``transformer.py:109-110``) as attention keys.
"""
from typing import Dict, Optional

import torch

from .backend import B200PETBackend
from .evaluate import sum_over_atoms
from .neighbors_gpu import VerletNeighborList

Tensor = torch.Tensor


class GraphedEvaluator:
    """Energy and +dE/dr of a fixed set of atoms in a fixed cell, for successive positions.

    :param backend: a ``B200PETBackend`` on a CUDA device, in eval mode (fixed cutoff).
    :param species: atomic numbers ``[N]``.
    :param cell: ``[3, 3]`` cell (ignored when not periodic).
    :param skin: Verlet skin in Angstrom; the graph is re-captured whenever an atom has moved
        more than ``skin / 2`` since the list was built.
    """

    def __init__(self, backend: B200PETBackend, species: Tensor, cell: Tensor, periodic: bool = True,
                 skin: float = 0.3, target: str = "energy", use_graph: bool = True,
                 charge: int = 0, spin_multiplicity: int = 1):
        if backend.num_neighbors_adaptive is not None:
            raise NotImplementedError("GraphedEvaluator: the adaptive cutoff re-derives the topology "
                                      "from the positions every step; use evaluate()")
        dev = next(backend.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("GraphedEvaluator: the backend must live on a CUDA device")
        self.backend, self.target, self.use_graph = backend, target, bool(use_graph)
        self.species = species.to(dev)
        self.cell = cell.to(dev, torch.float32).contiguous()
        self.cells = self.cell.reshape(1, 3, 3)
        self.system_indices = torch.zeros(self.species.shape[0], dtype=torch.long, device=dev)
        self._charge = torch.tensor([charge], dtype=torch.long, device=dev)
        self._spin = torch.tensor([spin_multiplicity], dtype=torch.long, device=dev)
        if backend.system_conditioning is not None:
            backend.system_conditioning.validate(self._charge, self._spin)
        self.verlet = VerletNeighborList(backend.cutoff, skin, periodic)
        self._pos = torch.zeros((self.species.shape[0], 3), device=dev, requires_grad=True)
        self._graph: Optional[torch.cuda.CUDAGraph] = None
        self._out: Optional[Dict[str, Tensor]] = None
        self._topo = None
        self._z_nodes = None
        self._list_id = -1
        self.n_captures = 0
        self.n_replays = 0

    # one step on the current topology, reading self._pos
    def _step(self) -> Dict[str, Tensor]:
        be = self.backend
        emit, be.emit_nef = be.emit_nef, False
        try:
            batch = be.preprocess_on_topology(self._pos, self.cells, self._topo, self._z_nodes)
            if be.system_conditioning is not None:
                batch["charge"], batch["spin_multiplicity"] = self._charge, self._spin
            nodes, edges = be.calculate_features(batch)
            pred, _, _ = be.predict(nodes, edges, batch, self.cells, self.system_indices, [self.target])
        finally:
            be.emit_nef = emit
        atomic = pred[self.target][0]
        energies = sum_over_atoms(atomic, self.system_indices, 1)
        (grad,) = torch.autograd.grad([energies], [self._pos], grad_outputs=[torch.ones_like(energies)])
        return {"energies": energies.detach(), "atomic": atomic.detach(), "dE_dpos": grad}

    def _capture(self) -> None:
        # warm up on a side stream (lazy initialisation: packed weights, kernel attributes)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(2):
                self._step()
        torch.cuda.current_stream().wait_stream(side)
        self._graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._graph):
            self._out = self._step()
        self.n_captures += 1

    def __call__(self, positions: Tensor) -> Dict[str, Tensor]:
        """``energies [1, P]``, ``atomic [N, P]``, ``dE_dpos [N, 3]`` at ``positions``.  The
        returned tensors are the graph's output buffers: they are overwritten by the next call."""
        pos = positions.detach().to(self._pos.device, torch.float32)
        centers, neighbors, shifts = self.verlet.update(pos, self.cell)
        rebuilt = self.verlet.n_builds != self._list_id
        with torch.no_grad():
            self._pos.copy_(pos)
        if rebuilt:
            self._list_id = self.verlet.n_builds
            # the skin list is filtered by the model's exactly symmetric pair test at the list
            # radius (the device list's own acceptance test is not bitwise symmetric for pairs that
            # cross a periodic boundary; its 2e-6 margin makes it a superset of this selection)
            self._topo, self._z_nodes = self.backend.build_topology(
                pos, centers, neighbors, self.species, self.cells, shifts, self.system_indices,
                list_cutoff=self.backend.cutoff + self.verlet.skin)
            self._graph = None
            if self.use_graph:
                self._capture()
        if self._graph is None:
            return self._step()
        self._graph.replay()
        self.n_replays += 1
        return self._out
