"""Energy + force evaluation through the backend: the plain-tensor mirror of
``metatrain.utils.evaluate_model.evaluate_model`` (``src/metatrain/utils/evaluate_model.py:21-160``).

Same recipe as the reference: re-leaf the positions (``_prepare_system``, ``:294-348``;
optionally the strain trick ``:310-321``), run the three backend stages, sum per structure
(``src/metatrain/utils/sum_over_atoms.py:31``), then ``torch.autograd.grad`` of the summed
energies w.r.t. positions/strain (``src/metatrain/utils/output_gradient.py:34-40``).  The
returned position gradient is **+dE/dr** like the reference's ``"positions"`` gradient
block (``evaluate_model.py:128-141``); forces are its negative.
"""
from typing import Dict, Optional

import torch

from .lib import call, ptr

Tensor = torch.Tensor


class _SumOverAtoms(torch.autograd.Function):
    @staticmethod
    def forward(ctx, atomic, struct_ptr, system_indices):
        n_struct = struct_ptr.shape[0] - 1
        out = torch.empty((n_struct, atomic.shape[1]), device=atomic.device, dtype=torch.float32)
        call("sum_over_atoms", ptr(atomic.contiguous()), ptr(struct_ptr), n_struct,
             atomic.shape[1], ptr(out))
        ctx.save_for_backward(system_indices)
        return out

    @staticmethod
    def backward(ctx, g):
        (system_indices,) = ctx.saved_tensors
        return g[system_indices], None, None


def sum_over_atoms(atomic: Tensor, system_indices: Tensor, n_structures: int) -> Tensor:
    """Per-structure sums of per-atom predictions; atoms of a structure are contiguous
    (they are, after ``concatenate_structures``, structures.py:17-112)."""
    idx = system_indices.long()
    # structure b owns the contiguous, ascending block idx == b: its bounds come from a binary
    # search (no device->host synchronisation, unlike torch.bincount)
    bounds = torch.arange(n_structures + 1, device=atomic.device, dtype=idx.dtype)
    struct_ptr = torch.searchsorted(idx, bounds).to(torch.int32)
    return _SumOverAtoms.apply(atomic, struct_ptr, idx)


def selection_mask(selected_atoms: Tensor, system_indices: Tensor) -> Tensor:
    """Boolean mask ``[N]`` of the selected atoms of a batch.  ``selected_atoms``: a boolean mask
    (returned as is) or ``[n, 2]`` integer ``(system, atom-within-system)`` pairs."""
    if selected_atoms.dtype == torch.bool:
        if selected_atoms.shape != system_indices.shape:
            raise ValueError("selected_atoms mask must have one entry per atom")
        return selected_atoms.to(system_indices.device)
    sel = selected_atoms.to(system_indices.device).long()
    if sel.dim() != 2 or sel.shape[1] != 2:
        raise ValueError("selected_atoms must be a boolean mask [N] or (system, atom) pairs [n, 2]")
    idx = system_indices.long()
    n = idx.shape[0]
    # atoms of a system are contiguous: first atom of every system by binary search
    n_sys = int(idx[-1].item()) + 1 if n else 0
    first = torch.searchsorted(idx, torch.arange(n_sys + 1, device=idx.device))
    mask = torch.zeros(n, dtype=torch.bool, device=idx.device)
    if sel.shape[0]:
        sys_ok = (sel[:, 0] >= 0) & (sel[:, 0] < n_sys)
        sys_c = sel[:, 0].clamp(0, max(n_sys - 1, 0))
        glob = first[sys_c] + sel[:, 1]
        ok = sys_ok & (sel[:, 1] >= 0) & (glob < first[sys_c + 1])
        mask[glob[ok]] = True   # pairs that are not in the batch select nothing (mts.slice semantics)
    return mask


def evaluate(
    backend,
    positions: Tensor,
    centers: Tensor,
    neighbors: Tensor,
    species: Tensor,
    cells: Tensor,
    cell_shifts: Tensor,
    system_indices: Tensor,
    target: str = "energy",
    gradients: bool = True,
    strain: bool = False,
    charge: Optional[Tensor] = None,
    spin_multiplicity: Optional[Tensor] = None,
    selected_atoms: Optional[Tensor] = None,
) -> Dict[str, Tensor]:
    """One energy(+forces, +strain gradient) evaluation of a batch of structures.

    Returns ``energies [B, P]``, ``atomic [N, P]`` and, if requested, ``dE_dpos [N, 3]``
    and ``dE_dstrain [3, 3]`` (a single strain shared by the batch).

    ``selected_atoms`` mirrors the argument of ``PET.forward``
    (``src/metatrain/pet/model.py:282``; applied at ``:724`` / ``:921-925`` by slicing the per-atom
    predictions before the sum over atoms): either a boolean mask ``[N]`` or ``[n, 2]`` integer
    ``(system, atom)`` pairs (the values of the reference's ``Labels``).  Energies are then summed
    over the selected atoms only, ``atomic`` holds their rows (in batch order), and the position
    gradient is that of the selected atoms' energies w.r.t. ALL positions — what an MD engine with
    domain decomposition needs: ghost atoms out to ``interaction_range = num_gnn_layers * cutoff``
    (``model.py:1004``) take part in the message passing but are not summed.
    """
    pos = positions.detach().clone().requires_grad_(gradients)
    pos_in, cells_in = pos, cells
    eps: Optional[Tensor] = None
    if strain:
        eps = torch.eye(3, device=pos.device, dtype=pos.dtype, requires_grad=True)
        pos_in = pos @ eps
        cells_in = cells @ eps
    # PET.forward hands its cutoff_width_adaptive hyper to preprocess (model.py:417-426)
    batch = backend.preprocess(pos_in, centers, neighbors, species, cells_in, cell_shifts,
                               system_indices,
                               float(backend.hypers.get("cutoff_width_adaptive", 1.0)))
    if backend.system_conditioning is not None:
        # what PET.forward does between the stages (pet/model.py:464-471): systems without the data
        # default to charge 0 / spin multiplicity 1 (:1173-1174)
        n_sys = cells.shape[0]
        charge = torch.zeros(n_sys, dtype=torch.long, device=pos.device) if charge is None else charge
        spin_multiplicity = (torch.ones(n_sys, dtype=torch.long, device=pos.device)
                             if spin_multiplicity is None else spin_multiplicity)
        backend.system_conditioning.validate(charge, spin_multiplicity)
        batch["charge"], batch["spin_multiplicity"] = charge, spin_multiplicity
        batch["system_indices"] = system_indices
    nodes, edges = backend.calculate_features(batch)
    pred, _, _ = backend.predict(nodes, edges, batch, cells_in, system_indices, [target])
    atomic = torch.cat(pred[target], dim=1) if len(pred[target]) > 1 else pred[target][0]
    mask: Optional[Tensor] = None
    if selected_atoms is not None:
        mask = selection_mask(selected_atoms, system_indices)
        atomic = atomic * mask.reshape((-1,) + (1,) * (atomic.dim() - 1)).to(atomic.dtype)
    if atomic.dim() > 2:  # tensorial per-atom outputs (non_conservative_stress: [N, 3, 3, P])
        energies = sum_over_atoms(atomic.reshape(atomic.shape[0], -1), system_indices,
                                  cells.shape[0]).reshape((cells.shape[0],) + tuple(atomic.shape[1:]))
    else:
        energies = sum_over_atoms(atomic, system_indices, cells.shape[0])
    out = {"energies": energies.detach(), "atomic": atomic.detach() if mask is None else atomic.detach()[mask]}
    if gradients:
        wrt = [pos] + ([eps] if strain else [])
        grads = torch.autograd.grad(
            [energies], wrt, grad_outputs=[torch.ones_like(energies)])
        out["dE_dpos"] = grads[0]
        if strain:
            out["dE_dstrain"] = grads[1]
    return out
