"""Parameter containers of the B200 PET backend.

The drop-in contract (level B1) requires the ``state_dict`` of our backend
to have exactly the reference's keys, shapes and *order* (``species_to_species_index``
first — ``src/metatrain/pet/modules/backend.py:63-71`` — because
``PET.load_checkpoint`` probes the dtype from the first float entry,
``src/metatrain/pet/model.py:978-981``), and the seed-0 initial values must equal the
reference's so that its known-answer test (``pet/tests/test_regression.py:66-74``)
transfers.  Both follow from creating the same leaf modules in the same order; the leaf
modules here are *only* parameter holders — their ``forward`` is never called, all
arithmetic happens in the CUDA engine (``metatrain_b200/csrc``).

Creation order mirrored (no code shared): ``backend.py:73-136`` (GNN layers, combination
norms/MLPs, embedders, empty head dicts), ``transformer.py:181-196`` (attention, norms,
MLP, centre contraction/expansion/MLP), ``transformer.py:446-461`` (geometry embedder,
compress, neighbour embedder), ``backend.py:171-217`` (heads + last layers per target).
"""
from math import prod
from typing import Dict, List, Optional

import torch
from torch import nn


class _Holder(nn.Module):
    """A module that only owns sub-modules/parameters; calling it is a bug."""

    def forward(self, *args, **kwargs):  # pragma: no cover
        raise RuntimeError("parameter holder: arithmetic lives in the CUDA engine")


def _ff_holder(d_model: int, d_ff: int, activation: str) -> _Holder:
    """FeedForward (transformer.py:21-39): SwiGLU projects to value and gate, SiLU to d_ff."""
    h = _Holder()
    h.w_in = nn.Linear(d_model, 2 * d_ff if activation.lower() == "swiglu" else d_ff)
    h.w_out = nn.Linear(d_ff, d_model)
    return h


def _transformer_layer_holder(d_pet: int, d_node: int, d_ff: int, norm: str = "RMSNorm",
                              activation: str = "SwiGLU") -> _Holder:
    norm_class = getattr(nn, norm)  # transformer.py:181
    layer = _Holder()
    layer.attention = _Holder()
    layer.attention.input_linear = nn.Linear(d_pet, 3 * d_pet)
    layer.attention.output_linear = nn.Linear(d_pet, d_pet)
    layer.norm_attention = norm_class(d_pet)
    layer.norm_mlp = norm_class(d_pet)
    layer.mlp = _ff_holder(d_pet, d_ff, activation)
    layer.center_contraction = nn.Linear(d_node, d_pet)
    layer.center_expansion = nn.Linear(d_pet, d_node)
    layer.norm_center_features = norm_class(d_node)
    layer.center_mlp = _ff_holder(d_node, 2 * d_node, activation)
    return layer


def _gnn_layer_holder(d_pet, d_node, d_ff, n_attention, n_species, is_first, norm="RMSNorm",
                      activation="SwiGLU") -> _Holder:
    g = _Holder()
    g.trans = _Holder()
    g.trans.layers = nn.ModuleList(
        [_transformer_layer_holder(d_pet, d_node, d_ff, norm, activation) for _ in range(n_attention)]
    )
    g.edge_embedder = nn.Linear(4, d_pet)
    n_merge = 2 if is_first else 3
    g.compress = nn.Sequential(
        nn.Linear(n_merge * d_pet, d_pet), nn.SiLU(), nn.Linear(d_pet, d_pet)
    )
    if not is_first:
        g.neighbor_embedder = nn.Embedding(n_species, d_pet)
    return g


def _head_holder(d_in: int, d_head: int) -> nn.Sequential:
    return nn.Sequential(
        nn.Linear(d_in, d_head), nn.SiLU(), nn.Linear(d_head, d_head), nn.SiLU()
    )


class SystemConditioning(_Holder):
    """Parameters of ``SystemConditioningEmbedding`` (``src/metatrain/pet/modules/conditioning.py:8-100``)
    under the reference's names: per-system charge and spin-multiplicity embeddings, concatenated
    and projected to ``d_out`` by Linear -> SiLU -> Linear (the last one zero-initialised, :40-48).
    ``validate`` mirrors the reference's range check (:53-79); the wrapper calls it
    (``pet/model.py:468``)."""

    required_data_keys: List[str] = ["charge", "spin_multiplicity"]

    def __init__(self, d_out: int, max_charge: int = 10, max_spin_multiplicity: int = 10) -> None:
        super().__init__()
        self.max_charge = max_charge
        self.max_spin_multiplicity = max_spin_multiplicity
        self.charge_embedding = nn.Embedding(2 * max_charge + 1, d_out)
        self.spin_multiplicity_embedding = nn.Embedding(max_spin_multiplicity, d_out)
        gate = nn.Linear(d_out, d_out)
        nn.init.zeros_(gate.weight)
        nn.init.zeros_(gate.bias)
        self.project = nn.Sequential(nn.Linear(2 * d_out, d_out), nn.SiLU(), gate)

    def validate(self, charge: torch.Tensor, spin_multiplicity: torch.Tensor) -> None:
        if (charge < -self.max_charge).any() or (charge > self.max_charge).any():
            raise ValueError(
                f"charge values must be in [{-self.max_charge}, {self.max_charge}], got "
                f"min={charge.min().item()}, max={charge.max().item()}. Increase max_charge in "
                f"model hypers to support wider charge ranges.")
        if (spin_multiplicity < 1).any() or (spin_multiplicity > self.max_spin_multiplicity).any():
            raise ValueError(
                f"spin_multiplicity values must be in [1, {self.max_spin_multiplicity}], got "
                f"min={spin_multiplicity.min().item()}, max={spin_multiplicity.max().item()}. Increase "
                f"max_spin_multiplicity in model hypers to support higher spin multiplicities.")


class PETParameters(nn.Module):
    """Owns every learnable tensor of the PET backend under the reference's names."""

    def __init__(self, hypers: dict, atomic_types: List[int]) -> None:
        super().__init__()
        self.d_pet = int(hypers["d_pet"])
        self.d_node = int(hypers["d_node"])
        self.d_head = int(hypers["d_head"])
        self.d_feedforward = int(hypers["d_feedforward"])
        self.num_gnn_layers = int(hypers["num_gnn_layers"])
        self.num_attention_layers = int(hypers["num_attention_layers"])
        n_species = len(atomic_types)

        self.register_buffer(
            "species_to_species_index", torch.full((max(atomic_types) + 1,), -1)
        )
        for i, z in enumerate(atomic_types):
            self.species_to_species_index[z] = i

        self.gnn_layers = nn.ModuleList(
            [
                _gnn_layer_holder(self.d_pet, self.d_node, self.d_feedforward,
                                  self.num_attention_layers, n_species, l == 0,
                                  hypers.get("normalization", "RMSNorm"), hypers.get("activation", "SwiGLU"))
                for l in range(self.num_gnn_layers)
            ]
        )
        if hypers.get("featurizer_type", "feedforward") == "feedforward":
            self.num_readout_layers = 1  # backend.py:93-107
            self.combination_norms = nn.ModuleList(
                [nn.LayerNorm(2 * self.d_pet) for _ in range(self.num_gnn_layers)]
            )
            self.combination_mlps = nn.ModuleList(
                [
                    nn.Sequential(nn.Linear(2 * self.d_pet, 2 * self.d_pet), nn.SiLU(),
                                  nn.Linear(2 * self.d_pet, self.d_pet))
                    for _ in range(self.num_gnn_layers)
                ]
            )
        else:  # residual featurizer: one readout per GNN layer, no combination MLPs (:108-111)
            self.num_readout_layers = self.num_gnn_layers
            self.combination_norms = nn.ModuleList()
            self.combination_mlps = nn.ModuleList()
        self.node_embedders = nn.ModuleList(
            [nn.Embedding(n_species, self.d_node) for _ in range(self.num_readout_layers)]
        )
        self.edge_embedder = nn.Embedding(n_species, self.d_pet)
        if hypers.get("system_conditioning"):  # backend.py:121-130
            self.system_conditioning: Optional[SystemConditioning] = SystemConditioning(
                self.d_node, int(hypers["max_charge"]), int(hypers["max_spin_multiplicity"]))
        else:
            self.system_conditioning = None
        self.node_heads = nn.ModuleDict()
        self.edge_heads = nn.ModuleDict()
        self.node_last_layers = nn.ModuleDict()
        self.edge_last_layers = nn.ModuleDict()
        # parameter-free hook points for diagnostic outputs (backend.py:138-155)
        ident = lambda n: nn.ModuleList([nn.Identity() for _ in range(n)])  # noqa: E731
        self.gnn_layers_post_mp_node = ident(self.num_gnn_layers)
        self.gnn_layers_post_mp_edge = ident(self.num_gnn_layers)
        self.node_backbone = ident(self.num_readout_layers)
        self.edge_backbone = ident(self.num_readout_layers)

    def add_output(self, target_name: str, output_shapes: Dict[str, List[int]]) -> None:
        r = range(self.num_readout_layers)
        self.node_heads[target_name] = nn.ModuleList(
            [_head_holder(self.d_node, self.d_head) for _ in r])
        self.edge_heads[target_name] = nn.ModuleList(
            [_head_holder(self.d_pet, self.d_head) for _ in r])
        self.node_last_layers[target_name] = nn.ModuleList(
            [nn.ModuleDict({k: nn.Linear(self.d_head, prod(s)) for k, s in output_shapes.items()})
             for _ in r])
        self.edge_last_layers[target_name] = nn.ModuleList(
            [nn.ModuleDict({k: nn.Linear(self.d_head, prod(s)) for k, s in output_shapes.items()})
             for _ in r])

    def remove_output(self, target_name: str) -> None:
        for table in (self.node_heads, self.edge_heads, self.node_last_layers,
                      self.edge_last_layers):
            if target_name in table:
                del table[target_name]
