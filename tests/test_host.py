"""CPU tests of the host-side logic: C-ABI surface, parameter/state-dict contract,
neighbor list, hyper-parameter validation (no GPU compute)."""
import ctypes
import os
import sys
import re

import numpy as np
import pytest
import torch

from helpers import DEFAULT_HYPERS, load_golden, seed_all, weight_fingerprint
from metatrain_b200 import B200PETBackend, lib
from metatrain_b200.neighbors import neighbor_list
from metatrain_b200.systems import make_batch, replicate, silicon_box, water_384
from oracle.structures import neighbor_list as oracle_neighbor_list

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "petb200.h")).read()
    return sorted(set(re.findall(r"PETB200_API\s+[\w\s\*]+?\b(petb200_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    if not os.path.isfile(lib.library_path()):
        lib.build()
    handle = ctypes.CDLL(lib.library_path())
    names = declared_symbols()
    assert len(names) >= 25
    for name in names:
        assert hasattr(handle, name), f"{name} declared in include/petb200.h but not exported"
    assert set(names) == set(lib._SIGNATURES), "ctypes table out of sync with the header"
    assert handle.petb200_version() >= 1


def declared_prototypes():
    """name -> list of C parameter declarations, parsed from include/petb200.h."""
    text = open(os.path.join(ROOT, "include", "petb200.h")).read()
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    protos = {}
    for m in re.finditer(r"PETB200_API\s+[\w\s\*]+?\b(petb200_\w+)\s*\(([^;]*?)\)\s*;", text, flags=re.S):
        params = [p.strip() for p in m.group(2).replace("\n", " ").split(",")]
        protos[m.group(1)] = [] if params in ([""], ["void"]) else params
    return protos


def test_ctypes_signatures_match_the_header():
    """Every hand-written ctypes argument list in lib._SIGNATURES must agree, position by position,
    with the C prototype (a drifted list corrupts the call silently)."""
    def ctype_of(param):
        if "*" in param or "petb200_stream_t" in param:
            return ctypes.c_void_p
        base = param.replace("const", "").split()[0]
        return {"int64_t": ctypes.c_int64, "int": ctypes.c_int, "float": ctypes.c_float,
                "size_t": ctypes.c_size_t}[base]
    protos = declared_prototypes()
    assert set(protos) == set(lib._SIGNATURES)
    for name, params in protos.items():
        expected = [ctype_of(p) for p in params]
        assert lib._SIGNATURES[name] == expected, f"{name}: ctypes {lib._SIGNATURES[name]} vs header {params}"


def test_torchscript_operator_library_loads_and_registers_its_schemas():
    """B3: csrc/libpetb200_torch.so (TORCH_LIBRARY(petb200, ...)) loads without a GPU and registers the
    operators a scripted model calls; no compute here."""
    import torch
    from metatrain_b200 import export
    export.build_torch_ops()
    export.load_torch_ops()
    topo, atomic = torch.ops.petb200.topology.default, torch.ops.petb200.pet_atomic.default
    assert "Tensor[]" in str(topo._schema) and "float cutoff" in str(topo._schema)
    assert str(atomic._schema).endswith("-> Tensor")
    assert all(os.path.isfile(p) for p in export.extension_libraries())
    with pytest.raises((RuntimeError, NotImplementedError)):   # no CPU path
        torch.ops.petb200.topology(torch.zeros(2, 3), torch.zeros(1, 3, 3), torch.zeros(0, dtype=torch.int64),
                                   torch.zeros(0, dtype=torch.int64), torch.zeros(0, 3, dtype=torch.int64),
                                   torch.zeros(2, dtype=torch.int64), torch.zeros(2, dtype=torch.int64), 4.5)


def test_state_dict_contract():
    g = load_golden("qm9_5")
    seed_all(0)
    be = B200PETBackend(g["hypers"], g["atomic_types"])
    be.add_output(g["target"], {g["target"] + "___0": [1]})
    sd = be.state_dict()
    first = next(iter(sd))
    assert first == "species_to_species_index" and not sd[first].is_floating_point()
    np.testing.assert_array_equal(weight_fingerprint(sd), g["weight_fingerprint"])
    assert sum(p.numel() for p in be.parameters()) == 2903298  # 4 species + 1 head
    be.remove_output(g["target"])
    assert not any("heads" in k or "last_layers" in k for k in be.state_dict())


def test_hyper_validation_matches_reference_error_types():
    bad = dict(DEFAULT_HYPERS, normalization="BatchNorm")
    with pytest.raises(ValueError, match="Unknown normalization flag"):
        B200PETBackend(bad, [1])
    with pytest.raises(ValueError, match="Unknown transformer flag"):
        B200PETBackend(dict(DEFAULT_HYPERS, transformer_type="x"), [1])
    with pytest.raises(ValueError, match="Unknown activation flag"):
        B200PETBackend(dict(DEFAULT_HYPERS, activation="relu"), [1])
    with pytest.raises(ValueError, match="not divisible"):
        B200PETBackend(dict(DEFAULT_HYPERS, num_heads=7), [1])
    with pytest.raises(ValueError, match="Unknown cutoff function"):
        B200PETBackend(dict(DEFAULT_HYPERS, cutoff_function="step"), [1])
    res = B200PETBackend(dict(DEFAULT_HYPERS, featurizer_type="residual"), [1])  # built: backend.py:589-649
    assert res.num_readout_layers == DEFAULT_HYPERS["num_gnn_layers"] and len(res.combination_mlps) == 0
    B200PETBackend(dict(DEFAULT_HYPERS, transformer_type="PostLN"), [1])  # built: transformer.py:236-262
    B200PETBackend(dict(DEFAULT_HYPERS, normalization="LayerNorm", activation="SiLU"), [1])  # built
    B200PETBackend(dict(DEFAULT_HYPERS, num_neighbors_adaptive=16), [1])  # built: adaptive_cutoff.py:110-229
    B200PETBackend(dict(DEFAULT_HYPERS, num_neighbors_adaptive=16, adaptive_cutoff_method="grid"), [1])  # :232-395
    with pytest.raises(ValueError, match="must be 'grid' or 'solver'"):
        B200PETBackend(dict(DEFAULT_HYPERS, num_neighbors_adaptive=16, adaptive_cutoff_method="x"), [1])
    cond = B200PETBackend(dict(DEFAULT_HYPERS, system_conditioning=True), [1])  # built: conditioning.py
    assert "system_conditioning.project.2.weight" in cond.state_dict()
    with pytest.raises(ValueError, match="charge values must be in"):
        cond.system_conditioning.validate(torch.tensor([11]), torch.tensor([1]))
    with pytest.raises(ValueError, match="spin_multiplicity values must be in"):
        cond.system_conditioning.validate(torch.tensor([0]), torch.tensor([0]))


def test_lora_injection_mirrors_reference_state_dict_keys():
    """finetuning.py:322-378: wrapped Linears keep the reference's sub-module names, so the keys
    of a LoRA-finetuned checkpoint (`...input_linear.linear.weight`, `...lora_A.weight`) load."""
    from metatrain_b200.finetuning import LoRALinear, inject_lora_layers
    be = B200PETBackend(DEFAULT_HYPERS, [1, 8])
    be.add_output("energy", {"energy___0": [1]})
    before = set(be.state_dict())
    inject_lora_layers(be, ("input_linear", "output_linear"), rank=4, alpha=8.0)
    after = set(be.state_dict())
    stem = "gnn_layers.0.trans.layers.0.attention.input_linear"
    assert stem + ".weight" in before and stem + ".weight" not in after
    assert {stem + ".linear.weight", stem + ".linear.bias", stem + ".lora_A.weight", stem + ".lora_B.weight"} <= after
    n_wrapped = sum(isinstance(m, LoRALinear) for m in be.modules())
    assert n_wrapped == 2 * DEFAULT_HYPERS["num_gnn_layers"] * DEFAULT_HYPERS["num_attention_layers"]
    assert len(after) == len(before) + 3 * n_wrapped - n_wrapped  # +A, +B per wrapped layer; weight/bias renamed
    inject_lora_layers(be, ("input_linear",))  # already wrapped: not an nn.Linear any more, left alone
    assert sum(isinstance(m, LoRALinear) for m in be.modules()) == n_wrapped


def test_graphed_evaluator_needs_a_cuda_backend():
    from metatrain_b200 import GraphedEvaluator
    be = B200PETBackend(DEFAULT_HYPERS, [1, 8])
    be.add_output("energy", {"energy___0": [1]})
    with pytest.raises(RuntimeError, match="CUDA device"):
        GraphedEvaluator(be, torch.tensor([1, 8]), torch.eye(3))
    ad = B200PETBackend(dict(DEFAULT_HYPERS, num_neighbors_adaptive=8), [1, 8])
    with pytest.raises(NotImplementedError, match="adaptive cutoff"):
        GraphedEvaluator(ad, torch.tensor([1, 8]), torch.eye(3))


def test_bench_reads_measured_peaks_tolerantly(tmp_path, monkeypatch):
    """bench.py's roofline denominators come from the driver-written MEASURED_PEAKS.json; whatever
    its key names, a malformed or missing file must fall back instead of killing the bench."""
    import json
    sys.path.insert(0, ROOT)
    import bench
    monkeypatch.setattr(bench, "ROOT", str(tmp_path))
    assert bench.peaks() == (6650.0, 1590.0, 1400.0, "fallback")
    for content in ({"hbm_gbs": 6556.8, "bf16_tflops": 1651, "bf16_tflops_sustained": 1403.7},
                    {"hbm": {"copy_GBs": 6556.8}, "bf16": {"burst_tflops": 1651.0, "sustained_tflops": 1403.7}}):
        (tmp_path / "MEASURED_PEAKS.json").write_text(json.dumps(content))
        assert bench.peaks() == (6556.8, 1651.0, 1403.7, "measured")
    (tmp_path / "MEASURED_PEAKS.json").write_text("not json")
    assert bench.peaks()[3] == "fallback"


def test_no_cpu_fallback():
    be = B200PETBackend(DEFAULT_HYPERS, [1, 8])
    be.add_output("energy", {"energy___0": [1]})
    b = make_batch([water_384()], 4.5)
    with pytest.raises(RuntimeError, match="no CPU path"):
        be.preprocess(b["positions"], b["centers"], b["neighbors"], b["species"], b["cells"],
                      b["cell_shifts"], b["system_indices"], 1.0)


@pytest.mark.parametrize("case", ["water_384", "carbon_5", "si_64", "ragged_mix", "co_periodic"])
def test_host_neighbor_list_reproduces_golden_lists(case):
    g = load_golden(case)
    for b in range(g["cells"].shape[0]):
        sel = g["system_indices"] == b
        pos, cell = g["positions"][sel].astype(np.float64), g["cells"][b].astype(np.float64)
        periodic = bool(np.abs(cell).sum() > 0)
        i, j, s = neighbor_list(pos, cell, periodic, 4.5)
        off = int(np.nonzero(sel)[0][0]) if sel.any() else 0
        e = np.isin(g["centers"], np.nonzero(sel)[0])
        ref = set(zip((g["centers"][e] - off).tolist(), (g["neighbors"][e] - off).tolist(),
                      map(tuple, g["cell_shifts"][e].tolist())))
        assert set(zip(i.tolist(), j.tolist(), map(tuple, s.tolist()))) == ref


def test_host_neighbor_list_wrapped_positions_and_big_box():
    box = silicon_box()
    a = neighbor_list(box["positions"], box["cell"], True, 4.5)
    b = oracle_neighbor_list(box["positions"], box["cell"], True, 4.5)
    assert all(np.array_equal(x, y) for x, y in zip(a, b))
    moved = box["positions"] + np.array([31.0, -17.0, 5.5])  # far outside the home cell
    i, j, s = neighbor_list(moved, box["cell"], True, 4.5)
    r = moved[j] - moved[i] + s @ box["cell"]
    assert len(i) == len(a[0]) and np.linalg.norm(r, axis=1).max() <= 4.5
    big = replicate(water_384(), (2, 2, 2))
    i, j, s = neighbor_list(big["positions"], big["cell"], True, 4.5)
    assert len(i) == 8 * 14520
