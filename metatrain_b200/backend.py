"""``B200PETBackend`` — drop-in replacement of the reference's tensor backend.

Mirrors ``metatrain.pet.modules.backend.PETBackend``
(``src/metatrain/pet/modules/backend.py:12``): same constructor, ``add_output`` /
``remove_output``, and the three plain-tensor stages ``preprocess`` (``:238``),
``calculate_features`` (``:344``) and ``predict`` (``:420``) that ``PET.forward`` calls
(``src/metatrain/pet/model.py:417,500,530``); same ``state_dict`` keys/order/initial values
(see ``parameters.py``), same exception types for invalid hyper-parameters.

Internally nothing is padded: edges live in a CSR layout and every stage is one
``torch.autograd.Function`` whose forward/backward are hand-scheduled sequences of CUDA
kernels from ``libpetb200.so`` (``engine.py``).  The padded NEF tensors that ``PET.forward``
may read from ``batch_data`` (``padding_mask``, ``cutoff_factors``, ``edge_distances``, ...,
``model.py:436-459,512-513``) are still produced, from the CSR data, so that the wrapper
keeps working unchanged.

Featurizers: ``feedforward`` (default) and ``residual`` (backend.py:589-649); transformer types
``PreLN`` (default) and ``PostLN`` (transformer.py:236-262); normalisation ``RMSNorm`` (default) or
``LayerNorm``; activation ``SwiGLU`` (default) or ``SiLU``; fixed cutoff (default) or the adaptive
cutoff with the ``solver`` (adaptive_cutoff.py:110-229) or ``grid`` (:232-395) method.  Not built
yet (raise ``NotImplementedError``): weight gradients (training).  Diagnostic capture fires the
``node_backbone`` / ``edge_backbone`` hook points only.
System conditioning (charge / spin embeddings, conditioning.py) is built.
"""
from typing import Dict, List, Optional, Tuple

import torch

from . import engine
from .lib import CUTOFF_BUMP, CUTOFF_COSINE, PREC_BF16, PREC_BF16X3, PREC_FP32, call, ptr
from .parameters import PETParameters

Tensor = torch.Tensor

_AVAILABLE_NORMALIZATIONS = ["LayerNorm", "RMSNorm"]
_AVAILABLE_TRANSFORMER_TYPES = ["PostLN", "PreLN"]
_AVAILABLE_ACTIVATIONS = ["SiLU", "SwiGLU"]
_PRECISIONS = {"fp32": PREC_FP32, "bf16x3": PREC_BF16X3, "bf16": PREC_BF16}


def _require_cuda(t: Tensor, what: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(
            f"B200PETBackend.{what}: expected CUDA tensors; this backend has no CPU path")
    # every kernel is launched on the CURRENT device's current stream (lib.call): tensors living
    # on another device would be touched from the wrong context / an unordered stream
    if t.device.index != torch.cuda.current_device():
        raise RuntimeError(
            f"B200PETBackend.{what}: tensors live on {t.device} but the current CUDA device is "
            f"cuda:{torch.cuda.current_device()}; wrap the call in `with torch.cuda.device(...)` "
            "or call torch.cuda.set_device first")


# ------------------------------------------------------------------ autograd stages
class _EdgeGeometry(torch.autograd.Function):
    """positions, cells -> (edge vectors, embedder distances, cutoff factors), CSR order."""

    @staticmethod
    def forward(ctx, positions, cells, topo, cutoff, width, func):
        pos = positions.detach().to(torch.float32).contiguous()
        cel = cells.detach().to(torch.float32).contiguous()
        vec, dist, fc = engine.edges_forward(topo, pos, cel, cutoff, width, func)
        ctx.topo, ctx.params = topo, (cutoff, width, func)
        ctx.in_dtypes = (positions.dtype, cells.dtype)
        ctx.save_for_backward(vec, dist)
        return vec, dist, fc

    @staticmethod
    def backward(ctx, d_vec, d_dist, d_fc):
        vec, dist = ctx.saved_tensors
        cutoff, width, func = ctx.params
        need_cells = ctx.needs_input_grad[1]
        c = lambda g: None if g is None else g.contiguous()  # noqa: E731
        d_pos, d_cells = engine.edges_backward(
            ctx.topo, vec, dist, c(d_vec), c(d_dist), c(d_fc), cutoff, width, func, need_cells)
        d_pos = d_pos.to(ctx.in_dtypes[0])
        if d_cells is not None:
            d_cells = d_cells.to(ctx.in_dtypes[1])
        return d_pos, d_cells, None, None, None, None


class _AdaptiveEdgeGeometry(torch.autograd.Function):
    """``_EdgeGeometry`` with per-pair cutoffs from the adaptive-cutoff solver: the backward adds
    the gradient that reaches the positions through the per-atom cutoffs (the implicit-function
    step, adaptive_cutoff.py:203-227)."""

    @staticmethod
    def forward(ctx, positions, cells, topo, adaptive, cutoff, width, func):
        pos = positions.detach().to(torch.float32).contiguous()
        cel = cells.detach().to(torch.float32).contiguous()
        vec, dist, fc = engine.adaptive_edges_forward(topo, adaptive, pos, cel, width, func)
        ctx.topo, ctx.adaptive, ctx.params = topo, adaptive, (cutoff, width, func)
        ctx.in_dtypes = (positions.dtype, cells.dtype)
        ctx.save_for_backward(vec, dist)
        return vec, dist, fc

    @staticmethod
    def backward(ctx, d_vec, d_dist, d_fc):
        vec, dist = ctx.saved_tensors
        cutoff, width, func = ctx.params
        need_cells = ctx.needs_input_grad[1]
        c = lambda g: None if g is None else g.contiguous()  # noqa: E731
        d_pos, d_cells = engine.adaptive_edges_backward(
            ctx.topo, ctx.adaptive, vec, dist, c(d_vec), c(d_dist), c(d_fc), cutoff, width, func,
            need_cells)
        d_pos = d_pos.to(ctx.in_dtypes[0])
        if d_cells is not None:
            d_cells = d_cells.to(ctx.in_dtypes[1])
        return d_pos, d_cells, None, None, None, None, None


class _Features(torch.autograd.Function):
    """(edge vectors, distances, cutoff factors) -> (node features, edge messages).  With the
    residual featurizer the outputs are the node features of every GNN layer followed by the edge
    features of every GNN layer (2 L tensors)."""

    @staticmethod
    def forward(ctx, vec, dist, fc, backend, topo, charge=None, spin_multiplicity=None):
        pw = backend._packed()
        prec = backend._precision
        cond = (engine.conditioning_table(pw, charge, spin_multiplicity)
                if backend.system_conditioning is not None else None)
        ctx.backend, ctx.topo, ctx.pw, ctx.fc, ctx.prec = backend, topo, pw, fc, prec
        ctx.residual = backend.featurizer_type == "residual"
        args = (pw, backend.hypers, topo, vec.contiguous(), dist.contiguous(), fc.contiguous(), prec, cond)
        if ctx.residual:
            nodes, edges, ctx.saved = engine.features_forward_residual(*args)
            return tuple(nodes) + tuple(edges)
        h, m, ctx.saved = engine.features_forward(*args)
        return h, m

    @staticmethod
    def backward(ctx, *grads):
        topo = ctx.topo
        if ctx.residual:
            n_layers = len(grads) // 2
            d_vec, d_dist, d_fc = engine.features_backward_residual(
                ctx.pw, ctx.backend.hypers, topo, ctx.fc.contiguous(), ctx.saved,
                list(grads[:n_layers]), list(grads[n_layers:]), ctx.prec)
            return d_vec, d_dist, d_fc, None, None, None, None
        d_h, d_m = grads
        if d_h is None:
            d_h = torch.zeros((topo.n_atoms, ctx.backend.d_node), device=ctx.fc.device, dtype=torch.float32)
        if d_m is None:
            d_m = torch.zeros((topo.n_edges, ctx.backend.d_pet), device=ctx.fc.device, dtype=torch.float32)
        d_vec, d_dist, d_fc = engine.features_backward(
            ctx.pw, ctx.backend.hypers, topo, ctx.fc.contiguous(), ctx.saved, d_h, d_m, ctx.prec)
        return d_vec, d_dist, d_fc, None, None, None, None


class _Predict(torch.autograd.Function):
    """(node features, edge messages, cutoff factors) -> per-atom predictions [N, P] of one
    readout layer."""

    @staticmethod
    def forward(ctx, h, m, fc, backend, topo, name, layer, sink=None):
        pw = backend._packed()
        prec = backend._precision
        atomic, saved = engine.predict_forward(pw, topo, name, h.contiguous(), m.contiguous(),
                                               fc.contiguous(), prec, layer)
        if sink is not None:  # last-layer features (outputs of the node / edge heads), no gradient
            # (the fused edge head keeps only the pre-activation of its second Linear: the features are
            # silu of it, formed only when the padded NEF view is produced anyway)
            e2 = saved["e2"]
            if e2 is None and backend.emit_nef:
                e2 = torch.nn.functional.silu(saved["e2p"])
            sink.append((saved["n2"], e2))
        ctx.topo, ctx.pw, ctx.saved, ctx.fc, ctx.name, ctx.prec, ctx.layer = topo, pw, saved, fc, name, prec, layer
        return atomic

    @staticmethod
    def backward(ctx, d_atomic):
        d_h, d_m, d_fc = engine.predict_backward(ctx.pw, ctx.topo, ctx.name, ctx.fc.contiguous(),
                                                 ctx.saved, d_atomic, ctx.prec, ctx.layer)
        return d_h, d_m, d_fc, None, None, None, None, None


class _CsrToNef(torch.autograd.Function):
    """[E, ...] CSR edge array -> zero-padded [N, M, ...] NEF array (nef.py:169-201)."""

    @staticmethod
    def forward(ctx, x, topo):
        d = 1
        for s in x.shape[1:]:
            d *= s
        out = torch.empty((topo.n_atoms, topo.max_row) + tuple(x.shape[1:]), device=x.device, dtype=torch.float32)
        call("csr_to_nef", ptr(x.contiguous()), ptr(topo.row_ptr), topo.n_atoms, topo.n_edges,
             topo.max_row, d, ptr(out))
        ctx.topo, ctx.d, ctx.tail = topo, d, tuple(x.shape[1:])
        return out

    @staticmethod
    def backward(ctx, g):
        topo = ctx.topo
        out = torch.empty((topo.n_edges,) + ctx.tail, device=g.device, dtype=torch.float32)
        call("nef_to_csr", ptr(g.contiguous()), ptr(topo.row_ptr), ptr(topo.ctr), topo.n_atoms,
             topo.n_edges, topo.max_row, ctx.d, ptr(out))
        return out, None


class _NefToCsr(torch.autograd.Function):
    """padded NEF array -> CSR edge array (nef.py:204-218)."""

    @staticmethod
    def forward(ctx, x, topo):
        tail = tuple(x.shape[2:])
        d = 1
        for s in tail:
            d *= s
        out = torch.empty((topo.n_edges,) + tail, device=x.device, dtype=torch.float32)
        call("nef_to_csr", ptr(x.contiguous()), ptr(topo.row_ptr), ptr(topo.ctr), topo.n_atoms,
             topo.n_edges, topo.max_row, d, ptr(out))
        ctx.topo, ctx.d, ctx.tail = topo, d, tail
        return out

    @staticmethod
    def backward(ctx, g):
        topo = ctx.topo
        out = torch.empty((topo.n_atoms, topo.max_row) + ctx.tail, device=g.device, dtype=torch.float32)
        call("csr_to_nef", ptr(g.contiguous()), ptr(topo.row_ptr), topo.n_atoms, topo.n_edges,
             topo.max_row, ctx.d, ptr(out))
        return out, None


def _process_non_conservative_stress(tensor: Tensor, cells: Tensor, system_indices: Tensor,
                                     num_properties: int) -> Tensor:
    """[N, 9 P] direct stress predictions -> [N, 3, 3, P], divided by the cell volume (infinite
    for the zero cells of non-periodic systems) and symmetrised (backend.py:780-813).  O(N)
    bookkeeping on the readout's output, done with tensor ops."""
    t = tensor.reshape(-1, 3, 3, num_properties)
    volumes = torch.abs(torch.det(cells.to(tensor.dtype)))
    volumes = torch.where(volumes == 0.0, torch.full_like(volumes, float("inf")), volumes)
    t = t / volumes[system_indices.long()].reshape(-1, 1, 1, 1)
    return 0.5 * (t + t.transpose(1, 2))


# ------------------------------------------------------------------------ the module
class B200PETBackend(PETParameters):
    """CUDA (sm_100a) implementation of the PET tensor backend.

    :param hypers: PET ``ModelHypers`` (``src/metatrain/pet/documentation.py:156-259``).
    :param atomic_types: sorted list of supported atomic numbers.
    :param precision: GEMM arithmetic — ``"bf16x3"`` (default: tcgen05, 2-term bf16 split,
        fp32 accumulate; forces within ~1e-5 eV/A of the fp32 reference), ``"fp32"`` (FFMA,
        the exact-fp32 parity path) or ``"bf16"`` (single-pass tcgen05; ~1e-2 eV/A).
    """

    NUM_FEATURE_TYPES: int = 2

    def __init__(self, hypers: dict, atomic_types: List[int], precision: str = "bf16x3") -> None:
        # same validation and exception types as the reference
        if hypers["normalization"] not in _AVAILABLE_NORMALIZATIONS:  # transformer.py:329-333
            raise ValueError(f"Unknown normalization flag: {hypers['normalization']}. "
                             f"Please choose from: {_AVAILABLE_NORMALIZATIONS}")
        if hypers["transformer_type"] not in _AVAILABLE_TRANSFORMER_TYPES:  # :335-339
            raise ValueError(f"Unknown transformer flag: {hypers['transformer_type']}. "
                             f"Please choose from: {_AVAILABLE_TRANSFORMER_TYPES}")
        if hypers["activation"] not in _AVAILABLE_ACTIVATIONS:  # :342-346
            raise ValueError(f"Unknown activation flag: {hypers['activation']}. "
                             f"Please choose from: {_AVAILABLE_ACTIVATIONS}")
        if hypers["d_pet"] % hypers["num_heads"] != 0:  # transformer.py:80-81
            raise ValueError("total dimension is not divisible by the number of heads")
        if hypers["cutoff_function"].lower() not in ("bump", "cosine"):  # structures.py:312-316
            raise ValueError(f"Unknown cutoff function type: {hypers['cutoff_function']}. "
                             f"Supported types are 'Cosine' and 'Bump'.")
        unsupported = []
        if hypers["featurizer_type"] not in ("feedforward", "residual"):
            raise ValueError(f"Unknown featurizer type: {hypers['featurizer_type']}")
        if (hypers.get("num_neighbors_adaptive") is not None
                and hypers.get("adaptive_cutoff_method", "solver").lower() not in ("solver", "grid")):
            raise ValueError("adaptive_cutoff_method must be 'grid' or 'solver', got "
                             + hypers["adaptive_cutoff_method"])  # structures.py:244-248
        # widths: the token width / head geometry is what the attention and row-wise kernels are built
        # for; the node width and the feed-forward width only enter dense contractions
        if (hypers["d_pet"], hypers["d_head"], hypers["num_heads"]) != (128, 128, 8):
            unsupported.append("d_pet/d_head/num_heads other than 128/128/8")
        if hypers["d_node"] != 256:
            # (d_node == d_pet is a different layer in the reference: no centre contraction / expansion
            # / centre feed-forward at all, transformer.py:189-201)
            unsupported.append("d_node other than 256")
        dff = hypers["d_feedforward"]
        if dff % 128 != 0 and not (precision != "fp32" and dff % 64 == 0 and dff <= 512):
            # multiples of 128 run on every path; other multiples of 64 only exist inside the fused
            # feed-forward kernels (tensor-core precisions, width <= 512)
            unsupported.append("d_feedforward that is neither a multiple of 128 nor a multiple of 64 <= 512 "
                               "on a tensor-core precision")
        if unsupported:
            raise NotImplementedError(
                "B200PETBackend: not built yet (SURVEY.md 8(f).4): " + ", ".join(unsupported))
        if precision not in _PRECISIONS:
            raise ValueError(f"unknown precision {precision!r}; choose from {list(_PRECISIONS)}")
        super().__init__(hypers, atomic_types)
        self.hypers = dict(hypers)
        self.featurizer_type = hypers["featurizer_type"]
        self.nl_is_strict = bool(hypers["long_range"]["enable"])
        self.cutoff = float(hypers["cutoff"])
        self.cutoff_function = hypers["cutoff_function"]
        self.cutoff_width = float(hypers["cutoff_width"])
        self.num_heads = int(hypers["num_heads"])
        self.num_neighbors_adaptive = (float(hypers["num_neighbors_adaptive"])
                                       if hypers.get("num_neighbors_adaptive") is not None else None)
        self.adaptive_cutoff_method = hypers.get("adaptive_cutoff_method", "solver")
        self._precision = _PRECISIONS[precision]
        self._cutoff_id = CUTOFF_BUMP if self.cutoff_function.lower() == "bump" else CUTOFF_COSINE
        self._pw: Optional[engine.PackedWeights] = None
        #: also emit the reference's padded NEF tensors from preprocess()/calculate_features()
        #: (needed by the PET wrapper for feature / diagnostic outputs; energies+forces
        #: only need the CSR handles)
        self.emit_nef = True

    # ------------------------------------------------------------------ internals
    def set_precision(self, precision: str) -> None:
        if precision == "fp32" and self.hypers["d_feedforward"] % 128 != 0:
            raise NotImplementedError("B200PETBackend: the fp32 path needs d_feedforward to be a multiple of 128")
        self._precision = _PRECISIONS[precision]

    def _invalidate_packed(self) -> None:
        self._pw = None
        self.__dict__.pop("_petb200_param_list", None)

    def add_output(self, *args, **kwargs):
        self._invalidate_packed()
        return super().add_output(*args, **kwargs)

    def remove_output(self, *args, **kwargs):
        self._invalidate_packed()
        return super().remove_output(*args, **kwargs)

    def _apply(self, *args, **kwargs):
        self._invalidate_packed()
        return super()._apply(*args, **kwargs)

    def load_state_dict(self, *args, **kwargs):
        self._invalidate_packed()
        return super().load_state_dict(*args, **kwargs)

    def _packed(self) -> engine.PackedWeights:
        if self._pw is None or not self._pw.is_current(self):
            self._pw = engine.PackedWeights(self)
        return self._pw

    def _check_inference(self) -> None:
        if self.training and torch.is_grad_enabled() and any(
                p.requires_grad for p in self.parameters()):
            raise NotImplementedError(
                "B200PETBackend: weight gradients / double backward (training) are not built "
                "yet (SURVEY.md 8(f).3); call .eval() or freeze the parameters")

    # ------------------------------------------------------------------ stage 1
    def preprocess(
        self,
        positions: Tensor,
        centers: Tensor,
        neighbors: Tensor,
        species: Tensor,
        cells: Tensor,
        cell_shifts: Tensor,
        system_indices: Tensor,
        cutoff_width_adaptive: float = 1.0,
    ) -> Dict[str, Tensor]:
        """Same signature and returned keys as ``PETBackend.preprocess`` (backend.py:238)."""
        _require_cuda(positions, "preprocess")
        topo, z_nodes = self.build_topology(positions, centers, neighbors, species, cells,
                                            cell_shifts, system_indices)
        return self.preprocess_on_topology(positions, cells, topo, z_nodes, cutoff_width_adaptive)

    def build_topology(self, positions, centers, neighbors, species, cells, cell_shifts,
                       system_indices, list_cutoff: Optional[float] = None):
        """The integer half of ``preprocess`` (a4-a6): CSR rows, reverse-edge map.  The one
        device->host read of the stage (edge count) happens here.  ``list_cutoff`` (> the model
        cutoff) keeps the pairs of a skin list out to that radius (beyond the model cutoff their
        cutoff factor is 0); the topology then stays valid while the list does — what
        ``md.GraphedEvaluator`` replays.  The pairs are always selected by the symmetric filter
        (``petb200_nl_filter_count``), never taken from the caller's list unfiltered: a device
        neighbor list may accept a pair at its radius in one direction only."""
        _require_cuda(positions, "preprocess")
        table = self.species_to_species_index
        if table.device != positions.device:
            raise RuntimeError(f"B200PETBackend.preprocess: inputs on {positions.device}, parameters on "
                               f"{table.device}")
        z = species.long()
        # atomic numbers outside the table or not in atomic_types: the reference fails in
        # nn.Embedding (index out of range); here the flag travels with the stage's one
        # device->host read and raises a ValueError
        outside = (z < 0) | (z >= table.shape[0])
        z_nodes = torch.where(outside, torch.full_like(z, -1), table[z.clamp(0, table.shape[0] - 1)])
        topo = engine.build_topology(positions, centers, neighbors, cell_shifts, cells,
                                     system_indices, z_nodes,
                                     self.cutoff if list_cutoff is None else float(list_cutoff),
                                     species_for_error=species)
        return topo, z_nodes

    def preprocess_on_topology(self, positions, cells, topo, z_nodes,
                               cutoff_width_adaptive: float = 1.0) -> Dict[str, Tensor]:
        """The floating-point half of ``preprocess`` on a given topology (no host
        synchronisation with a fixed cutoff)."""
        atomic_cutoffs = None
        if self.num_neighbors_adaptive is None:
            vec, dist, fc = _EdgeGeometry.apply(positions, cells, topo, self.cutoff,
                                                self.cutoff_width, self._cutoff_id)
        else:
            # per-atom cutoffs, then the topology of the pairs inside the pair cutoffs
            topo, adaptive = engine.adaptive_topology(
                topo, positions.detach().to(torch.float32).contiguous(),
                cells.detach().to(torch.float32).contiguous(), self.cutoff,
                self.num_neighbors_adaptive, float(cutoff_width_adaptive),
                self.adaptive_cutoff_method.lower())
            vec, dist, fc = _AdaptiveEdgeGeometry.apply(positions, cells, topo, adaptive, self.cutoff,
                                                        self.cutoff_width, self._cutoff_id)
            atomic_cutoffs = adaptive.atomic_cutoffs.to(positions.dtype)
        if not self.emit_nef:
            batch_data = {"element_indices_nodes": z_nodes}
        else:
            batch_data = self._nef_batch_data(topo, positions, z_nodes, vec, dist, fc)
            if atomic_cutoffs is not None:
                batch_data["atomic_cutoffs_stats"] = atomic_cutoffs
        # CSR handles consumed by calculate_features / predict (not part of the reference's
        # dictionary; plain tensors + one opaque python attribute)
        vec._petb200_topology = topo
        batch_data["_petb200_vec"] = vec
        batch_data["_petb200_dist"] = dist
        batch_data["_petb200_fc"] = fc
        return batch_data

    def _nef_batch_data(self, topo, positions, z_nodes, vec, dist, fc) -> Dict[str, Tensor]:
        """The padded NEF view of the batch that ``PET.forward`` may read
        (model.py:436-459,512-513,545,560); derived from the CSR data."""
        n, m = topo.n_atoms, topo.max_row
        counts = (topo.row_ptr[1:] - topo.row_ptr[:-1]).long()
        mask = torch.arange(m, device=positions.device)[None, :] < counts[:, None]
        slot = torch.arange(topo.n_edges, device=positions.device) - topo.row_ptr[:-1].long()[topo.ctr.long()]
        z_nbr_nef = torch.zeros((n, m), device=positions.device, dtype=torch.long)
        rev_flat = torch.arange(n * m, device=positions.device).reshape(n, m).clone()
        if topo.n_edges > 0:
            ctr_l, rev_l = topo.ctr.long(), topo.rev.long()
            z_nbr_nef[ctr_l, slot] = topo.z_neighbors.long()
            rev_flat[ctr_l, slot] = ctr_l[rev_l] * m + slot[rev_l]
        vec_nef = _CsrToNef.apply(vec, topo)
        dist_nef = _CsrToNef.apply(dist, topo)
        # the reference's padded distances are sqrt(0 + 1e-15) (structures.py:330)
        dist_nef = torch.where(mask, dist_nef, torch.full_like(dist_nef, 1e-15 ** 0.5))
        batch_data: Dict[str, Tensor] = {
            "element_indices_nodes": z_nodes,
            "element_indices_neighbors": z_nbr_nef,
            "edge_vectors": vec_nef,
            "edge_distances": dist_nef,
            "padding_mask": mask,
            "reverse_neighbor_index": rev_flat,
            "cutoff_factors": _CsrToNef.apply(fc, topo),
            "atomic_cutoffs_stats": torch.full((n,), self.cutoff, device=positions.device,
                                               dtype=positions.dtype),
            "centers": topo.ctr.long(),
            "neighbors": topo.col.long(),
            "nef_to_edges_neighbor": slot,
            "cell_shifts": topo.shift,
        }
        return batch_data

    @staticmethod
    def _topology_of(batch_data: Dict[str, Tensor]) -> engine.Topology:
        try:
            return batch_data["_petb200_vec"]._petb200_topology
        except (KeyError, AttributeError):
            raise RuntimeError(
                "B200PETBackend: batch_data was not produced by this backend's preprocess() "
                "(CSR handles missing)") from None

    # ------------------------------------------------------------------ stage 2
    def calculate_features(
        self, batch_data: Dict[str, Tensor], capture_diagnostics: bool = False
    ) -> Tuple[List[Tensor], List[Tensor]]:
        """Same contract as ``PETBackend.calculate_features`` (backend.py:344): returns
        ``([node features [N, d_node]], [edge features [N, M, d_pet]])``.  With
        ``capture_diagnostics`` the outputs pass through the ``node_backbone`` / ``edge_backbone``
        identity modules so that forward hooks registered on them fire (the hook points inside the
        featurizer do not exist in the fused implementation)."""
        self._check_inference()
        topo = self._topology_of(batch_data)
        charge = spin = None
        if self.system_conditioning is not None:
            # set by the wrapper after preprocess (pet/model.py:464-471; backend.py:375-378)
            charge, spin = batch_data["charge"], batch_data["spin_multiplicity"]
            if charge.shape[0] != topo.n_structures or spin.shape[0] != topo.n_structures:
                raise ValueError("calculate_features: one charge and one spin multiplicity per system expected")
        outs = _Features.apply(batch_data["_petb200_vec"], batch_data["_petb200_dist"],
                               batch_data["_petb200_fc"], self, topo, charge, spin)
        n_layers = len(outs) // 2   # 1 (feedforward) or num_gnn_layers (residual featurizer)
        nodes, edges = list(outs[:n_layers]), list(outs[n_layers:])
        if not self.emit_nef:
            return nodes, edges  # CSR [E, d_pet]; predict() recognises it by its rank
        nef = []
        for m in edges:
            m_nef = _CsrToNef.apply(m, topo)
            m_nef._petb200_csr = m  # lets predict() skip the NEF round trip
            nef.append(m_nef)
        if capture_diagnostics and not torch.jit.is_scripting() and not torch.jit.is_tracing():
            # backend.py:396-415: hooks on node_backbone[i] / edge_backbone[i] see the raw backbone
            # features.  Hooks on modules INSIDE the featurizer never fire here: the GNN layers run
            # as fused kernels, their sub-modules only hold parameters.
            nodes = [self.node_backbone[i](h) for i, h in enumerate(nodes)]
            # (a hook that replaces the tensor drops the CSR shortcut attribute: predict() then
            # gathers the NEF tensor back to CSR)
            nef = [self.edge_backbone[i](m_nef) for i, m_nef in enumerate(nef)]
        return nodes, nef

    # ------------------------------------------------------------------ stage 3
    def predict(
        self,
        node_features_list: List[Tensor],
        edge_features_list: List[Tensor],
        batch_data: Dict[str, Tensor],
        cells: Tensor,
        system_indices: Tensor,
        requested_output_names: List[str],
    ) -> Tuple[Dict[str, List[Tensor]], Dict[str, List[Tensor]], Dict[str, List[Tensor]]]:
        """Same contract as ``PETBackend.predict`` (backend.py:420).  The last-layer feature
        dictionaries (outputs of the node / edge heads, backend.py:651-687; read by the wrapper for
        the ``mtt::aux::{target}_last_layer_features`` outputs) hold, for every REQUESTED output,
        one ``[N, d_head]`` node tensor and one edge tensor (``[N, M, d_head]`` zero-padded NEF, or
        CSR ``[E, d_head]`` when ``emit_nef`` is off) per readout layer; they are detached."""
        self._check_inference()
        topo = self._topology_of(batch_data)
        fc = batch_data["_petb200_fc"]
        if len(node_features_list) != self.num_readout_layers or len(edge_features_list) != self.num_readout_layers:
            raise ValueError(f"predict: expected {self.num_readout_layers} node / edge feature tensors")
        edges_csr = []
        for m_in in edge_features_list:
            m = getattr(m_in, "_petb200_csr", None)
            if m is None:
                m = m_in if m_in.dim() == 2 else _NefToCsr.apply(m_in, topo)
            edges_csr.append(m)
        atomic_predictions: Dict[str, List[Tensor]] = {}
        node_ll: Dict[str, List[Tensor]] = {}
        edge_ll: Dict[str, List[Tensor]] = {}
        for name in self.node_last_layers.keys():
            if name not in requested_output_names:
                continue
            # node + edge contributions summed over the readout layers (backend.py:469-481)
            atomic = None
            sink: List[Tuple[Tensor, Tensor]] = []
            for layer, (h, m) in enumerate(zip(node_features_list, edges_csr)):
                part = _Predict.apply(h, m, fc, self, topo, name, layer, sink)
                atomic = part if atomic is None else atomic + part
            sizes = self._packed().heads[name][0]["block_sizes"]
            blocks = list(torch.split(atomic, sizes, dim=1))
            if name == "non_conservative_stress":  # backend.py:483-490
                blocks[0] = _process_non_conservative_stress(blocks[0], cells, system_indices,
                                                             blocks[0].shape[1] // 9)
            atomic_predictions[name] = blocks
            node_ll[name] = [n for n, _ in sink]
            # (CSR-only mode with the fused edge head: the edge features are not materialised)
            edge_ll[name] = [_CsrToNef.apply(e, topo) if self.emit_nef else e for _, e in sink if e is not None]
        return atomic_predictions, node_ll, edge_ll
