"""Shared helpers for the parity tests."""
import ast
import os
import random

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GOLDEN_CASES = ["qm9_5", "water_384", "water_384_nonstrict", "carbon_5", "si_64",
                "si_64_cosine", "si_64_cutoff7", "co_periodic", "ragged_mix",
                # residual featurizer (backend.py:589-649): one readout per GNN layer
                "qm9_5_residual", "water_384_residual",
                # PostLN transformer layers (transformer.py:236-262)
                "qm9_5_postln", "water_384_postln",
                # the original PET layer (PostLN + LayerNorm + SiLU + residual featurizer) and the
                # PreLN layer with LayerNorm + SiLU
                "water_384_classic", "qm9_5_classic", "water_384_preln_ln_silu",
                # adaptive cutoff, solver method (adaptive_cutoff.py:110-229)
                "water_384_adaptive", "qm9_5_adaptive", "carbon_5_adaptive", "ragged_mix_adaptive",
                "water_384_adaptive_grid", "carbon_5_adaptive_grid", "ragged_mix_adaptive_grid",
                # LoRA adapters (finetuning.py:322-378), merged into the packed weights
                "water_384_lora", "qm9_5_lora_wide",
                # system conditioning (conditioning.py:8-100)
                "qm9_5_conditioned", "qm9_5_conditioned_residual",
                # direct stress head (backend.py:780-813)
                "stress_head_mix",
                # all variants combined in one model / one ragged batch
                "kitchen_sink",
                # feed-forward width other than the default (a multiple of 128: every path)
                "water_384_dff640"]

# pet/documentation.py:159-259 defaults
DEFAULT_HYPERS = dict(
    cutoff=4.5, num_neighbors_adaptive=None, adaptive_cutoff_method="solver",
    cutoff_function="Bump", cutoff_width=0.5, cutoff_width_adaptive=1.0, d_pet=128,
    d_head=128, d_node=256, d_feedforward=256, num_heads=8, num_attention_layers=2,
    num_gnn_layers=2, normalization="RMSNorm", activation="SwiGLU",
    attention_temperature=1.0, transformer_type="PreLN", featurizer_type="feedforward",
    zbl=False, long_range=dict(enable=False), system_conditioning=False, max_charge=10,
    max_spin_multiplicity=10,
)


def load_golden(name):
    g = dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz")))
    g["target"] = str(g["target"])
    g["hypers"] = dict(DEFAULT_HYPERS)
    g["hypers"].update(ast.literal_eval(str(g["hypers_override"])))
    g["lora"] = g["hypers"].pop("_lora", None)  # test-only: LoRA adapters injected after construction
    g["gate_seed"] = g["hypers"].pop("_gate_seed", None)  # test-only: conditioning gate re-drawn
    g["atomic_types"] = [int(z) for z in g["atomic_types"]]
    g["out_shape"] = [int(v) for v in g["out_shape"]] if "out_shape" in g else [1]
    return g


def apply_lora(module, g):
    """Inject the golden case's LoRA adapters (same seed and order as make_golden.py did on the
    reference).  Returns the extra state-dict entry the oracle needs (the adapter scale)."""
    if g.get("gate_seed") is not None:  # same draw as oracle/ref_loader.build_reference_backend
        torch.manual_seed(g["gate_seed"])
        gate = module.system_conditioning.project[2]
        with torch.no_grad():
            gate.weight.normal_(0.0, 0.05)
            gate.bias.normal_(0.0, 0.05)
    if not g.get("lora"):
        return {}
    from metatrain_b200.finetuning import inject_lora_layers
    lora = g["lora"]
    torch.manual_seed(lora["seed"])
    inject_lora_layers(module, tuple(lora["target_modules"]), rank=lora["rank"], alpha=lora["alpha"])
    return {"lora.scaling": torch.tensor(lora["alpha"] / lora["rank"])}


def seed_all(seed=0):
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)


def golden_inputs(g, device="cpu", dtype=torch.float32):
    t = lambda k: torch.tensor(g[k]).to(device)  # noqa: E731
    out = dict(
        positions=t("positions").to(dtype), centers=t("centers"), neighbors=t("neighbors"),
        species=t("species"), cells=t("cells").to(dtype), cell_shifts=t("cell_shifts"),
        system_indices=t("system_indices"),
    )
    if "charge" in g:  # system conditioning inputs
        out.update(charge=t("charge"), spin_multiplicity=t("spin_multiplicity"))
    return out


def weight_fingerprint(state_dict):
    rows = []
    for v in state_dict.values():
        v64 = v.detach().to(torch.float64).cpu()
        rows.append([float(v64.sum()), float((v64 * v64).sum())])
    return np.array(rows, dtype=np.float64)


LONG_BOX_CASES = ["water_long_1x1x8", "water_long_1x1x24"]


def load_long_box(name):
    """Compact goldens of the elongated water boxes (make_golden.make_long_box_case): fp32 positions,
    cell and species plus the unmodified reference's fp32 / fp64 outputs on exactly those inputs.
    The (deterministic) neighbor list is rebuilt here."""
    from metatrain_b200.neighbors import neighbor_list
    g = load_golden(name)
    pos = g["positions"].astype(np.float64)
    cell = g["cells"][0].astype(np.float64)
    i, j, S = neighbor_list(pos, cell, True, g["hypers"]["cutoff"])
    assert len(i) == int(g["n_edges"]), "neighbor list differs from the one the golden was made with"
    g["nl"] = (i, j, S)
    g["batch"] = dict(
        positions=torch.from_numpy(g["positions"]), centers=torch.from_numpy(i.astype(np.int32)),
        neighbors=torch.from_numpy(j.astype(np.int32)), species=torch.from_numpy(g["species"].astype(np.int32)),
        cells=torch.from_numpy(g["cells"]), cell_shifts=torch.from_numpy(S.astype(np.int32).reshape(-1, 3)),
        system_indices=torch.zeros(len(pos), dtype=torch.int64))
    return g
