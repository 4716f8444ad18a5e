"""TorchScript packaging of the hot path (SURVEY.md 8(b) level B3, 8(f) rank 2).

The reference exports a model as TorchScript — ``AtomisticModel(self.eval(), metadata,
capabilities).save(path, collect_extensions=...)`` (``src/metatrain/pet/model.py:990-1021``,
``src/metatrain/cli/export.py:243-266``) — and MD engines load it back together with the shared
libraries of its custom operators (``src/metatrain/utils/io.py:183-184``).  A scripted module can
call neither ``ctypes`` nor Python ``autograd.Function``s, so the engine is exposed a second time as
TorchScript operators (``csrc/torch_ops.cpp`` -> ``csrc/libpetb200_torch.so``):

    torch.ops.petb200.topology(...)    CSR topology of a batch (one device->host read)
    torch.ops.petb200.pet_atomic(...)  per-atom predictions, differentiable w.r.t. positions / cells

``ExportedPET`` is the scriptable module around them: same tensor inputs as
``PETBackend.preprocess`` (``backend.py:238``), packed weights frozen into the module at export
time.  ``extension_libraries()`` lists what ``collect_extensions`` has to ship.

Supported here: the default PET configuration (PreLN + RMSNorm + SwiGLU, feedforward featurizer,
fixed cutoff, no system conditioning, one readout layer) on a tensor-core precision — the
configuration the C++ stage schedule (``csrc/schedule.cu``) is built for.
"""
import os
from typing import List, Optional, Tuple

import torch

from . import engine, lib
from .lib import PREC_FP32

Tensor = torch.Tensor

_CSRC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc")
_OPS_PATH = os.path.join(_CSRC, "libpetb200_torch.so")
_loaded = False


def torch_ops_path() -> str:
    return _OPS_PATH


def build_torch_ops(verbose: bool = False) -> str:
    """Compile ``csrc/torch_ops.cpp`` against this interpreter's libtorch, in-tree, linked to
    ``libpetb200.so`` (found next to it through ``$ORIGIN``)."""
    import shutil
    import torch.utils.cpp_extension as ext

    lib.build()
    src = os.path.join(_CSRC, "torch_ops.cpp")
    if os.path.isfile(_OPS_PATH) and os.path.getmtime(_OPS_PATH) >= max(
            os.path.getmtime(src), os.path.getmtime(os.path.join(_CSRC, "..", "..", "include", "petb200.h"))):
        return _OPS_PATH
    build_dir = os.path.join(_CSRC, "build_torch_ops")
    os.makedirs(build_dir, exist_ok=True)
    ext.load(name="petb200_torch", sources=[src], with_cuda=True, is_python_module=False,
             build_directory=build_dir, verbose=verbose,
             extra_cflags=["-O2", "-std=c++17"],
             extra_ldflags=[f"-L{_CSRC}", "-lpetb200", "-Wl,-rpath,\\$$ORIGIN:\\$$ORIGIN/.."])
    shutil.copyfile(os.path.join(build_dir, "petb200_torch.so"), _OPS_PATH)
    global _loaded
    _loaded = True   # ext.load registered the operators with this process already
    return _OPS_PATH


def load_torch_ops() -> None:
    """Register the ``petb200::`` operators with this process (once).  No fallback: raises when the
    library has not been built."""
    global _loaded
    if _loaded:
        return
    if not os.path.isfile(_OPS_PATH):
        raise RuntimeError(f"{_OPS_PATH} is missing: run metatrain_b200.export.build_torch_ops() "
                           "(or __graft_entry__.build()). There is no fallback path.")
    lib.load()   # libpetb200.so first (the operator library links against it)
    torch.ops.load_library(_OPS_PATH)
    _loaded = True


def extension_libraries() -> List[str]:
    """Shared libraries a scripted ``ExportedPET`` needs at load time (what
    ``collect_extensions`` copies, src/metatrain/cli/export.py:252-266)."""
    return [lib.library_path(), _OPS_PATH]


def export_weights(backend, target: str) -> Tuple[List[Tensor], List[int], List[float]]:
    """Flat weight list + metadata of ``torch.ops.petb200.pet_atomic`` (order mirrored by
    ``Weights`` in csrc/torch_ops.cpp): per GNN layer 10 token-builder tensors, 22 tensors per
    attention layer, 5 combine tensors; then the node / edge embeddings and 18 head tensors.
    Matrices are in the bf16 hi/lo split format of ``petb200_split_bf16``."""
    hyp = backend.hypers
    if backend._precision == PREC_FP32:
        raise NotImplementedError("export: choose a tensor-core precision (bf16x3)")
    if (engine._is_generic(hyp) or backend.featurizer_type != "feedforward"
            or backend.num_neighbors_adaptive is not None or backend.system_conditioning is not None):
        raise NotImplementedError("export: only the default PET configuration (PreLN + RMSNorm + SwiGLU, "
                                  "feedforward featurizer, fixed cutoff, no conditioning) is packaged")
    pw = backend._packed()
    sp = lambda w: engine.split_weight(w.contiguous() if not w.is_contiguous() else w, pw).clone()  # noqa: E731
    empty = torch.empty(0, device=pw.edge_emb.device)
    out: List[Tensor] = []
    for L, C in zip(pw.gnn, pw.combine):
        if engine._stage_structs(pw, L) is None or C["img"] is None:
            raise NotImplementedError("export: layer shape outside the fused kernels (d_pet = 128 needed)")
        d = L["w_geo"].shape[0]
        width = L["w1_t"].shape[0]
        c_img = L["c_img"] if (L["c_img"] is not None and engine.USE_FUSED_CHAINS) else [empty.to(torch.uint8)] * 2
        out += [sp(L["w1m"]), sp(L["w1_t"][width - d:].contiguous()), L["b_fold"], L["geo_fold"],
                L["nbr_fold"] if L["nbr_fold"] is not None else empty, sp(L["w2"]), sp(L["w2_t"]), L["b2"],
                c_img[0], c_img[1]]
        for T in L["tl"]:
            out += [T["qkv_img"], T["b_qkv"], sp(T["w_qkv_t"]), sp(T["w_o"]), sp(T["w_o_t"]), T["b_o"],
                    T["mlp_img"][0], T["mlp_img"][1], T["b_in"], T["b_out"],
                    sp(T["w_con"]), sp(T["w_con_t"]), T["b_con"], sp(T["w_exp"]), sp(T["w_exp_t"]), T["b_exp"],
                    sp(T["wc_in"]), sp(T["wc_in_t"]), T["bc_in"], sp(T["wc_out"]), sp(T["wc_out_t"]), T["bc_out"]]
        out += [C["img"][0], C["img"][1], C["s_vec"], C["b_fold"], C["b_b"]]
    if len(pw.heads[target]) != 1:
        raise NotImplementedError("export: one readout layer expected")
    H = pw.heads[target][0]
    out += [pw.node_emb[0], pw.edge_emb]
    out += [sp(H["n1"]), H["n1_b"], sp(H["n2"]), H["n2_b"], sp(H["e1"]), H["e1_b"], sp(H["e2"]), H["e2_b"],
            sp(H["n1_t"]), sp(H["n2_t"]), sp(H["e1_t"]), sp(H["e2_t"]), H["wn"], H["bn"], H["we"], H["be"]]
    fused_head = H["e_img"] is not None and engine.USE_FUSED_CHAINS and H["wn"].shape[0] == 1
    out += [H["e_img"][0], H["e_img"][1]] if fused_head else [empty.to(torch.uint8)] * 2
    out = [t.detach().contiguous() for t in out]
    dff = pw.gnn[0]["tl"][0]["w_out"].shape[1]
    meta = [len(pw.gnn), len(pw.gnn[0]["tl"]), hyp["d_pet"], hyp["d_node"], hyp["num_heads"], dff,
            H["n2"].shape[0], H["wn"].shape[0], backend._precision, backend._cutoff_id]
    fmeta = [float(backend.cutoff), float(backend.cutoff_width), float(hyp["attention_temperature"]),
             float(H["be"][0])]
    return out, [int(v) for v in meta], fmeta


class ExportedPET(torch.nn.Module):
    """Scriptable energy model over the ``petb200::`` operators.

    ``forward(positions, centers, neighbors, species, cells, cell_shifts, system_indices,
    selected_atoms=None)`` returns ``(energies [B, P], atomic [N, P])`` (``selected_atoms``: optional
    boolean mask of the atoms to sum, the rest — e.g. ghost atoms within ``interaction_range`` of a
    domain — only take part in the message passing); forces are ``-autograd.grad(energies.sum(), positions)``
    exactly as an MD engine computes them from the reference's exported model
    (``src/metatrain/utils/evaluate_model.py:21-160``)."""

    def __init__(self, backend, target: str = "energy"):
        super().__init__()
        load_torch_ops()
        weights, meta, fmeta = export_weights(backend, target)
        self.weights: List[Tensor] = weights
        self.meta: List[int] = meta
        self.fmeta: List[float] = fmeta
        self.cutoff: float = float(backend.cutoff)
        self.interaction_range: float = float(backend.cutoff) * len(backend.gnn_layers)  # model.py:1004
        # what an atom's energy really depends on with the feedforward featurizer: the last message
        # update (backend.py:559-575) still mixes in the reversed token of the last GNN layer
        self.receptive_range: float = float(backend.cutoff) * (len(backend.gnn_layers) + 1)
        self.register_buffer("species_to_species_index", backend.species_to_species_index.detach().clone())

    def forward(self, positions: Tensor, centers: Tensor, neighbors: Tensor, species: Tensor, cells: Tensor,
                cell_shifts: Tensor, system_indices: Tensor,
                selected_atoms: Optional[Tensor] = None) -> Tuple[Tensor, Tensor]:
        z = species.to(torch.int64)
        table = self.species_to_species_index
        outside = (z < 0) | (z >= table.shape[0])
        z_nodes = torch.where(outside, torch.full_like(z, -1), table[z.clamp(0, table.shape[0] - 1)])
        topo = torch.ops.petb200.topology(positions, cells, centers, neighbors, cell_shifts, system_indices,
                                          z_nodes, self.cutoff)
        atomic = torch.ops.petb200.pet_atomic(positions, cells, topo, self.weights, self.meta, self.fmeta)
        if selected_atoms is not None:   # boolean mask [N] (pet/model.py:282,921-925): ghosts are not summed
            atomic = atomic * selected_atoms.to(atomic.dtype).unsqueeze(1)
        energies = torch.zeros((cells.shape[0], atomic.shape[1]), dtype=atomic.dtype, device=atomic.device)
        energies = energies.index_add(0, system_indices.to(torch.int64), atomic)
        return energies, atomic
