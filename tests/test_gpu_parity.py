"""GPU parity tests proper: the CUDA path (through the C ABI) against the committed golden
vectors of the unmodified reference, against the oracle on the same seeded inputs, and —
at BASELINE.json's full size — through size-independent properties."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from helpers import (GOLDEN_CASES, LONG_BOX_CASES, apply_lora, golden_inputs, load_golden,  # noqa: E402
                     load_long_box, seed_all)
from metatrain_b200 import B200PETBackend, evaluate  # noqa: E402
from metatrain_b200.systems import make_batch, replicate, water_384  # noqa: E402
from oracle import pet_oracle  # noqa: E402

DEV = "cuda:0"
FORCE_TOL = 1e-4  # eV/A, BASELINE.json north_star ("forces within 1e-4 eV/A of reference")


def make_backend(g, precision="fp32"):
    seed_all(0)
    be = B200PETBackend(g["hypers"], g["atomic_types"], precision=precision)
    be.add_output(g["target"], {g["target"] + "___0": g["out_shape"]})
    apply_lora(be, g)
    return be.to(DEV).eval()


@pytest.mark.parametrize("case", GOLDEN_CASES)
def test_matches_reference_golden(case):
    g = load_golden(case)
    be = make_backend(g)
    strain = "ref32_dE_dstrain" in g
    out = evaluate(be, **golden_inputs(g, DEV), target=g["target"], strain=strain)
    e = out["energies"].cpu().numpy()
    scale = max(1.0, float(np.abs(g["ref32_energies"]).max()))
    assert np.abs(e - g["ref32_energies"]).max() <= 1e-5 * scale
    assert np.abs(out["atomic"].cpu().numpy() - g["ref32_atomic"]).max() <= 2e-5
    f_err = np.abs(out["dE_dpos"].cpu().numpy() - g["ref32_dE_dpos"]).max()
    assert f_err <= FORCE_TOL, f"force max-abs-err {f_err:.2e}"
    assert f_err <= 2e-5, f"fp32 path should sit at the fp32 noise floor, got {f_err:.2e}"
    if strain:
        assert np.abs(out["dE_dstrain"].cpu().numpy() - g["ref32_dE_dstrain"]).max() <= 1e-4
    if "ref64_dE_dpos" in g:
        assert np.abs(out["dE_dpos"].cpu().numpy() - g["ref64_dE_dpos"]).max() <= FORCE_TOL


@pytest.mark.parametrize("case", GOLDEN_CASES)
def test_tensor_core_split_meets_force_tolerance(case):
    """bf16x3 (tcgen05, 2-term split): forces within the north-star 1e-4 eV/A of the
    reference; single-pass bf16 is reported but only sanity-bounded."""
    g = load_golden(case)
    be = make_backend(g, precision="bf16x3")
    out = evaluate(be, **golden_inputs(g, DEV), target=g["target"])
    f_err = np.abs(out["dE_dpos"].cpu().numpy() - g["ref32_dE_dpos"]).max()
    e_err = np.abs(out["energies"].cpu().numpy() - g["ref32_energies"]).max()
    n = len(g["species"])
    print(f"{case}: bf16x3 force max-abs-err {f_err:.2e} eV/A, energy err/atom {e_err / n:.2e}")
    assert f_err <= FORCE_TOL
    assert e_err / n <= 1e-5 * max(1.0, float(np.abs(g["ref32_energies"]).max()) / n)
    be.set_precision("bf16")
    out1 = evaluate(be, **golden_inputs(g, DEV), target=g["target"])
    f1 = np.abs(out1["dE_dpos"].cpu().numpy() - g["ref32_dE_dpos"]).max()
    print(f"{case}: bf16 single-pass force max-abs-err {f1:.2e} eV/A")
    assert f1 <= 0.2


@pytest.mark.parametrize("precision", ["fp32", "bf16x3"])
@pytest.mark.parametrize("case", LONG_BOX_CASES)
def test_elongated_boxes_match_reference(case, precision):
    """The boxes of the atom-sharded runs are up to 24x longer than the seed box (z to 376 A): fp32
    coordinates there carry ~3e-5 A of rounding.  The unmodified reference was run on exactly these
    fp32-rounded coordinates (fp32 and fp64); its own fp32-vs-fp64 force difference is the floor
    (2.5e-5 at 1x1x8, 7.7e-5 at 1x1x24)."""
    g = load_long_box(case)
    be = make_backend(g, precision=precision)
    be.emit_nef = False
    out = evaluate(be, **{k: v.to(DEV) for k, v in g["batch"].items()}, target=g["target"])
    f = out["dE_dpos"].cpu().numpy()
    err32 = np.abs(f - g["ref32_dE_dpos"]).max()
    err64 = np.abs(f - g["ref64_dE_dpos"]).max()
    floor = np.abs(g["ref32_dE_dpos"] - g["ref64_dE_dpos"]).max()
    print(f"{case} {precision}: force err vs ref fp32 {err32:.2e}, vs ref fp64 {err64:.2e} "
          f"(reference fp32 vs fp64: {floor:.2e})")
    assert err64 <= FORCE_TOL, f"force max-abs-err vs the fp64 reference {err64:.2e}"
    assert err32 <= FORCE_TOL + floor
    e_ref = float(g["ref64_energies"].ravel()[0])
    assert abs(float(out["energies"]) - e_ref) <= 2e-6 * abs(e_ref)


def test_feed_forward_width_inside_the_fused_kernel_only():
    """d_feedforward = 192 (a multiple of 64, not of 128) exists only inside the fused feed-forward
    kernels: served on the tensor-core precisions, rejected loudly on the fp32 path."""
    g = load_golden("qm9_5_dff192")
    be = make_backend(g, precision="bf16x3")
    out = evaluate(be, **golden_inputs(g, DEV), target=g["target"])
    assert np.abs(out["energies"].cpu().numpy() - g["ref32_energies"]).max() <= 1e-4
    assert np.abs(out["dE_dpos"].cpu().numpy() - g["ref32_dE_dpos"]).max() <= FORCE_TOL
    with pytest.raises(NotImplementedError):
        make_backend(g, precision="fp32")
    with pytest.raises(NotImplementedError):
        be.set_precision("fp32")


def test_lora_adapters_are_merged_and_repacked_on_update():
    """LoRA adapters (finetuning.py:357-378) are merged into the packed weights; zeroing every
    lora_B in place must fall back to the base model's golden (the packed weights follow the
    parameters' versions)."""
    from metatrain_b200.finetuning import LoRALinear
    g, base = load_golden("water_384_lora"), load_golden("water_384")
    be = make_backend(g)
    adapted = [m for m in be.modules() if isinstance(m, LoRALinear)]
    assert len(adapted) == 2 * g["hypers"]["num_gnn_layers"] * g["hypers"]["num_attention_layers"]
    out = evaluate(be, **golden_inputs(g, DEV), target=g["target"])
    assert np.abs(out["dE_dpos"].cpu().numpy() - g["ref32_dE_dpos"]).max() <= 2e-5
    with torch.no_grad():
        for m in adapted:
            m.lora_B.weight.zero_()
    out = evaluate(be, **golden_inputs(g, DEV), target=g["target"])
    assert np.abs(out["energies"].cpu().numpy() - base["ref32_energies"]).max() <= 1e-5 * abs(base["ref32_energies"]).max()
    assert np.abs(out["dE_dpos"].cpu().numpy() - base["ref32_dE_dpos"]).max() <= 2e-5


def test_reference_hard_coded_energies():
    """src/metatrain/pet/tests/test_regression.py:66-74, same assert_close tolerance."""
    g = load_golden("qm9_5")
    out = evaluate(make_backend(g), **golden_inputs(g, DEV), target=g["target"], gradients=False)
    expected = torch.tensor([1.146098375320, 0.171331465244, 0.539504408836, 0.861489117146,
                             0.177449733019])
    torch.testing.assert_close(out["energies"].cpu().ravel(), expected)


@pytest.mark.parametrize("case", ["water_384_adaptive", "qm9_5_adaptive", "carbon_5_adaptive",
                                  "ragged_mix_adaptive", "water_384_adaptive_grid",
                                  "carbon_5_adaptive_grid", "ragged_mix_adaptive_grid"])
def test_adaptive_cutoffs_match_reference(case):
    """Per-atom cutoffs of the solver (adaptive_cutoff.py:110-229) and the pairs kept by the
    symmetrised pair cutoffs (structures.py:253-262), against the unmodified reference."""
    g = load_golden(case)
    be = make_backend(g)
    inp = golden_inputs(g, DEV)
    bd = be.preprocess(inp["positions"], inp["centers"], inp["neighbors"], inp["species"],
                       inp["cells"], inp["cell_shifts"], inp["system_indices"], 1.0)
    rc = bd["atomic_cutoffs_stats"].cpu().numpy()
    assert rc.shape == g["ref32_atomic_cutoffs"].shape
    assert np.abs(rc - g["ref32_atomic_cutoffs"]).max() <= 2e-5
    assert int(bd["padding_mask"].sum()) == int(g["ref32_n_edges_kept"])
    assert rc.min() >= g["hypers"]["cutoff"] / 16 and rc.max() <= g["hypers"]["cutoff"]


@pytest.mark.parametrize("case", ["si_64", "ragged_mix", "carbon_5", "carbon_5_adaptive",
                                  "water_384_adaptive", "carbon_5_adaptive_grid"])
def test_stages_match_oracle(case):
    """Stage-by-stage comparison with the oracle on identical weights and inputs."""
    g = load_golden(case)
    be = make_backend(g)
    sd = {k: v.detach().cpu() for k, v in be.state_dict().items()}
    ref = pet_oracle.energy_and_gradients(sd, g["hypers"], **golden_inputs(g), target=g["target"])
    inp = golden_inputs(g, DEV)
    pos = inp["positions"].clone().requires_grad_(True)
    bd = be.preprocess(pos, inp["centers"], inp["neighbors"], inp["species"], inp["cells"],
                       inp["cell_shifts"], inp["system_indices"], 1.0)
    rb = ref["batch"]
    # integer work is bit exact
    assert torch.equal(bd["padding_mask"].cpu(), rb["mask"])
    assert torch.equal(bd["element_indices_nodes"].cpu(), rb["z_nodes"])
    assert torch.equal(bd["element_indices_neighbors"].cpu()[rb["mask"]], rb["z_neighbors"][rb["mask"]])
    assert torch.equal(bd["reverse_neighbor_index"].cpu()[rb["mask"]], rb["reverse_flat"][rb["mask"]])
    # padded slots differ by construction (the reference aliases edge 0 there, we write zeros)
    torch.testing.assert_close(bd["edge_vectors"].detach().cpu()[rb["mask"]],
                               rb["edge_vectors"].detach()[rb["mask"]], atol=1e-6, rtol=1e-6)
    torch.testing.assert_close(bd["edge_distances"].detach().cpu()[rb["mask"]],
                               rb["edge_distances"].detach()[rb["mask"]], atol=1e-6, rtol=1e-6)
    torch.testing.assert_close(bd["cutoff_factors"].detach().cpu(), rb["cutoff_factors"].detach(),
                               atol=2e-6, rtol=1e-5)
    nodes, edges = be.calculate_features(bd)
    assert nodes[0].shape == ref["node_features"].shape and edges[0].shape == ref["edge_features"].shape
    torch.testing.assert_close(nodes[0].detach().cpu(), ref["node_features"], atol=2e-4, rtol=1e-4)
    m = rb["mask"][..., None]
    torch.testing.assert_close(edges[0].detach().cpu() * m, ref["edge_features"] * m, atol=2e-4, rtol=1e-4)
    pred, node_ll, edge_ll = be.predict(nodes, edges, bd, inp["cells"], inp["system_indices"], [g["target"]])
    torch.testing.assert_close(pred[g["target"]][0].detach().cpu(), ref["atomic"], atol=2e-5, rtol=1e-5)
    # last-layer features (backend.py:651-687)
    assert len(node_ll[g["target"]]) == len(ref["node_last_layer_features"]) == be.num_readout_layers
    for ours, theirs in zip(node_ll[g["target"]], ref["node_last_layer_features"]):
        torch.testing.assert_close(ours.cpu(), theirs, atol=2e-4, rtol=1e-4)
    for ours, theirs in zip(edge_ll[g["target"]], ref["edge_last_layer_features"]):
        torch.testing.assert_close(ours.cpu() * m, theirs * m, atol=2e-4, rtol=1e-4)
    (grad,) = torch.autograd.grad(pred[g["target"]][0].sum(), pos)
    assert (grad.cpu() - ref["dE_dpos"]).abs().max() <= 2e-5


def test_predict_accepts_plain_nef_tensor():
    """The B1 contract: predict() consumes an [N, M, d] edge tensor (e.g. one modified by
    the caller); gradients flow through the NEF<->CSR conversion."""
    g = load_golden("si_64")
    be = make_backend(g)
    inp = golden_inputs(g, DEV)
    pos = inp["positions"].clone().requires_grad_(True)
    bd = be.preprocess(pos, inp["centers"], inp["neighbors"], inp["species"], inp["cells"],
                       inp["cell_shifts"], inp["system_indices"], 1.0)
    nodes, edges = be.calculate_features(bd)
    plain = edges[0] * 1.0  # a new tensor without the CSR shortcut attribute
    a, _, _ = be.predict(nodes, [plain], bd, inp["cells"], inp["system_indices"], [g["target"]])
    b, _, _ = be.predict(nodes, edges, bd, inp["cells"], inp["system_indices"], [g["target"]])
    torch.testing.assert_close(a[g["target"]][0], b[g["target"]][0], atol=0, rtol=0)
    (ga,) = torch.autograd.grad(a[g["target"]][0].sum(), pos, retain_graph=True)
    (gb,) = torch.autograd.grad(b[g["target"]][0].sum(), pos)
    torch.testing.assert_close(ga, gb, atol=1e-6, rtol=1e-6)


def test_capture_diagnostics_fires_backbone_hooks():
    """backend.py:396-415: with capture_diagnostics the raw backbone features pass through the
    node_backbone / edge_backbone identity modules, where the wrapper hangs its forward hooks."""
    g = load_golden("si_64")
    be = make_backend(g)
    inp = golden_inputs(g, DEV)
    bd = be.preprocess(inp["positions"], inp["centers"], inp["neighbors"], inp["species"],
                       inp["cells"], inp["cell_shifts"], inp["system_indices"], 1.0)
    seen = {}
    hooks = [be.node_backbone[0].register_forward_hook(lambda m, i, o: seen.__setitem__("node", o)),
             be.edge_backbone[0].register_forward_hook(lambda m, i, o: seen.__setitem__("edge", o))]
    nodes, edges = be.calculate_features(bd, capture_diagnostics=True)
    for h in hooks:
        h.remove()
    assert seen["node"] is nodes[0] and seen["edge"] is edges[0]
    assert seen["edge"].shape == (bd["padding_mask"].shape[0], bd["padding_mask"].shape[1], be.d_pet)
    plain_nodes, plain_edges = be.calculate_features(bd)
    assert torch.equal(plain_nodes[0], nodes[0]) and torch.equal(plain_edges[0], edges[0])


@pytest.mark.parametrize("case", ["water_384", "ragged_mix", "qm9_5_residual", "carbon_5"])
def test_cxx_stage_schedule_equals_python_schedule(case):
    """petb200_gnn_fwd / _bwd (the C++ stage-level schedule, csrc/schedule.cu) enqueue the same kernels
    in the same order as the per-op Python schedule: results must be bit-identical."""
    from metatrain_b200 import engine
    g = load_golden(case)
    be = make_backend(g, "bf16x3")
    outs = {}
    for flag in (True, False):
        engine.USE_STAGE_SCHEDULE = flag
        try:
            outs[flag] = evaluate(be, **golden_inputs(g, DEV), target=g["target"], strain=True)
        finally:
            engine.USE_STAGE_SCHEDULE = True
    for k in ("energies", "atomic", "dE_dpos"):
        assert torch.equal(outs[True][k], outs[False][k]), k
    assert np.abs(outs[True]["dE_dpos"].cpu().numpy() - g["ref32_dE_dpos"]).max() <= FORCE_TOL


def test_selected_atoms_domain_decomposition_reproduces_the_periodic_box():
    """``selected_atoms`` (pet/model.py:282,724): cut the periodic water box into two open clusters
    (own atoms + ghost images), evaluate each with only its own atoms selected: per-atom energies
    and the summed position gradients must reproduce the periodic evaluation — what LAMMPS domain
    decomposition relies on.  Ghosts are taken out to (num_gnn_layers + 1) * cutoff: the feedforward
    featurizer's last message update (backend.py:559-575) mixes in the reversed token of the LAST
    GNN layer, so an atom's energy depends on atoms up to 3 cutoffs away, one more than the
    ``interaction_range = num_gnn_layers * cutoff`` the reference declares (model.py:1004)."""
    from metatrain_b200.neighbors import neighbor_list
    g = load_golden("water_384")
    be = make_backend(g, "fp32")
    full = evaluate(be, **golden_inputs(g, DEV), target=g["target"])
    box = water_384()
    pos, cell, Z = box["positions"], box["cell"], box["Z"]
    n = len(Z)
    rng_cut = be.cutoff * (len(be.gnn_layers) + 1)
    frac = (pos @ np.linalg.inv(cell)) % 1.0
    images = np.array([[a, b, c] for a in (-1, 0, 1) for b in (-1, 0, 1) for c in (-1, 0, 1)])
    all_pos = (pos[None] + (images @ cell)[:, None]).reshape(-1, 3)
    all_id = np.tile(np.arange(n), len(images))
    home = np.repeat((images == 0).all(1), n)
    e_atomic = np.zeros(n)
    grad = np.zeros((n, 3))
    for side in (0, 1):
        own = np.nonzero((frac[:, 0] < 0.5) == (side == 0))[0]
        d2 = ((all_pos[:, None, :] - pos[own][None, :, :]) ** 2).sum(-1).min(1)
        keep = np.nonzero(d2 <= rng_cut ** 2)[0]
        is_own = home[keep] & np.isin(all_id[keep], own)
        cp, cz, cid = all_pos[keep], Z[all_id[keep]], all_id[keep]
        i, j, S = neighbor_list(cp, np.eye(3), False, be.cutoff)
        batch = dict(positions=torch.tensor(cp, dtype=torch.float32), centers=torch.tensor(i), neighbors=torch.tensor(j),
                     species=torch.tensor(cz), cells=torch.zeros(1, 3, 3), cell_shifts=torch.tensor(S).reshape(-1, 3),
                     system_indices=torch.zeros(len(cp), dtype=torch.long))
        out = evaluate(be, **{k: v.to(DEV) for k, v in batch.items()}, target=g["target"],
                       selected_atoms=torch.tensor(is_own))
        assert out["atomic"].shape[0] == int(is_own.sum())
        e_atomic[cid[is_own]] = out["atomic"].cpu().numpy()[:, 0]
        np.add.at(grad, cid, out["dE_dpos"].cpu().numpy())
        assert abs(float(out["energies"]) - float(out["atomic"].sum())) <= 1e-3
    assert np.abs(e_atomic - full["atomic"].cpu().numpy()[:, 0]).max() <= 2e-5
    assert np.abs(grad - full["dE_dpos"].cpu().numpy()).max() <= 5e-5
    # (system, atom) pairs select the same atoms as the mask
    pairs = torch.tensor([[0, 0], [0, 5], [0, 383], [3, 1]])
    sel = evaluate(be, **golden_inputs(g, DEV), target=g["target"], selected_atoms=pairs)
    assert torch.equal(sel["atomic"], full["atomic"][[0, 5, 383]])
    assert abs(float(sel["energies"]) - float(full["atomic"][[0, 5, 383]].sum())) <= 1e-4


def test_csr_only_mode_equals_default():
    g = load_golden("water_384")
    be = make_backend(g)
    a = evaluate(be, **golden_inputs(g, DEV), target=g["target"])
    be.emit_nef = False
    b = evaluate(be, **golden_inputs(g, DEV), target=g["target"])
    assert torch.equal(a["energies"], b["energies"]) and torch.equal(a["dE_dpos"], b["dE_dpos"])


def test_deterministic():
    g = load_golden("water_384")
    be = make_backend(g)
    a = evaluate(be, **golden_inputs(g, DEV), target=g["target"])
    b = evaluate(be, **golden_inputs(g, DEV), target=g["target"])
    assert torch.equal(a["dE_dpos"], b["dE_dpos"]) and torch.equal(a["atomic"], b["atomic"])


def test_empty_and_errors():
    g = load_golden("qm9_5")
    be = make_backend(g)
    z = lambda *s, dt=torch.float32: torch.zeros(*s, dtype=dt, device=DEV)  # noqa: E731
    # empty system, src/metatrain/pet/tests/test_functionality.py:79-103
    out = evaluate(be, z(0, 3), z(0, dt=torch.int32), z(0, dt=torch.int32), z(0, dt=torch.int32),
                   z(1, 3, 3), z(0, 3, dt=torch.int32), z(0, dt=torch.int64), target=g["target"],
                   gradients=False)
    assert out["atomic"].numel() == 0 and float(out["energies"].abs().sum()) == 0.0
    # a one-directional neighbor list must be rejected loudly (the reference silently
    # produces garbage)
    inp = golden_inputs(g, DEV)
    half = inp["centers"] < inp["neighbors"]
    with pytest.raises(ValueError, match="not symmetric"):
        be.preprocess(inp["positions"], inp["centers"][half], inp["neighbors"][half], inp["species"],
                      inp["cells"], inp["cell_shifts"][half], inp["system_indices"], 1.0)
    with pytest.raises(RuntimeError, match="no CPU path"):
        be.preprocess(*(t.cpu() for t in (inp["positions"], inp["centers"], inp["neighbors"],
                                          inp["species"], inp["cells"], inp["cell_shifts"],
                                          inp["system_indices"])), 1.0)
    be.train()
    with pytest.raises(NotImplementedError, match="training"):
        evaluate(be, **inp, target=g["target"])


@pytest.mark.parametrize("precision", ["fp32", "bf16x3"])
@pytest.mark.parametrize("adaptive", [None, 4])
def test_atoms_without_any_neighbour(adaptive, precision):
    """A batch whose neighbor list is empty (every atom isolated): zero-length rows everywhere,
    with the fixed and with the adaptive cutoff; compared with the oracle on the same weights."""
    g = load_golden("qm9_5")
    g["hypers"] = dict(g["hypers"], num_neighbors_adaptive=adaptive)
    be = make_backend(g, precision=precision)
    sd = {k: v.detach().cpu() for k, v in be.state_dict().items()}
    pos = torch.tensor([[0.0, 0, 0], [20.0, 0, 0], [0, 30.0, 0], [5.0, 5.0, 50.0]])
    inp = dict(positions=pos, centers=torch.zeros(0, dtype=torch.long), neighbors=torch.zeros(0, dtype=torch.long),
               species=torch.tensor([1, 6, 8, 7]), cells=torch.zeros(2, 3, 3),
               cell_shifts=torch.zeros(0, 3, dtype=torch.long), system_indices=torch.tensor([0, 0, 1, 1]))
    ref = pet_oracle.energy_and_gradients(sd, g["hypers"], **inp, target=g["target"])
    out = evaluate(be, **{k: v.to(DEV) for k, v in inp.items()}, target=g["target"])
    torch.testing.assert_close(out["energies"].cpu(), ref["energies"], atol=1e-5, rtol=1e-5)
    torch.testing.assert_close(out["atomic"].cpu(), ref["atomic"], atol=1e-5, rtol=1e-5)
    assert float(out["dE_dpos"].abs().max()) == 0.0 and float(ref["dE_dpos"].abs().max()) == 0.0


@pytest.mark.parametrize("precision", ["fp32", "bf16x3"])
@pytest.mark.parametrize("variant", ["fixed", "adaptive_solver", "adaptive_grid"])
def test_reference_autograd_system_against_fp64_oracle(variant, precision):
    """The system of the reference's autograd test (utils/testing/autograd.py:24-43): two carbon atoms
    in a 2 A cubic cell — ~95 periodic images per atom inside the cutoff, i.e. rows longer than the
    tensor-core attention tile and shifts up to +-3.  Forces against the oracle evaluated in fp64
    (whose own gradient is gradcheck-ed in tests/test_oracle.py)."""
    from helpers import DEFAULT_HYPERS
    from oracle.structures import neighbor_list
    hyp = dict(DEFAULT_HYPERS)
    if variant != "fixed":
        hyp.update(num_neighbors_adaptive=20.0, adaptive_cutoff_method=variant.split("_")[1])
    seed_all(0)
    be = B200PETBackend(hyp, [6], precision=precision)
    be.add_output("energy", {"energy___0": [1]})
    be = be.to(DEV).eval()
    sd64 = {k: (v.detach().cpu().double() if v.is_floating_point() else v.detach().cpu())
            for k, v in be.state_dict().items()}
    pos0 = np.array([[0.0, 0.0, 0.0], [0.9, 0.9, 0.9]])
    cell = 2.0 * np.eye(3)
    i, j, S = neighbor_list(pos0, cell, True, 4.5)
    inp = dict(positions=torch.tensor(pos0), centers=torch.tensor(i), neighbors=torch.tensor(j),
               species=torch.tensor([6, 6]), cells=torch.tensor(cell)[None], cell_shifts=torch.tensor(S),
               system_indices=torch.zeros(2, dtype=torch.long))
    ref = pet_oracle.energy_and_gradients(sd64, hyp, **inp, target="energy", with_strain=True)
    dev_inp = {k: (v.float() if v.is_floating_point() else v).to(DEV) for k, v in inp.items()}
    out = evaluate(be, **dev_inp, target="energy", strain=True)
    assert abs(float(out["energies"]) - float(ref["energies"])) <= 2e-5 * max(1.0, abs(float(ref["energies"])))
    assert (out["dE_dpos"].cpu().double() - ref["dE_dpos"]).abs().max() <= FORCE_TOL
    assert (out["dE_dstrain"].cpu().double() - ref["dE_dstrain"]).abs().max() <= 5e-4


def test_float64_inputs_and_empty_structure_in_a_batch():
    """fp64 positions / cells are accepted (computed in fp32, gradients returned in the input dtype,
    like a float32 model fed by an fp64 MD engine), and a structure without atoms in the middle of a
    batch gets a zero energy row (sum_over_atoms.py:31 semantics)."""
    g = load_golden("carbon_5")
    be = make_backend(g)
    inp = golden_inputs(g, DEV)
    ref = evaluate(be, **inp, target=g["target"], strain=True)
    inp64 = dict(inp, positions=inp["positions"].double(), cells=inp["cells"].double())
    out = evaluate(be, **inp64, target=g["target"], strain=True)
    assert out["dE_dpos"].dtype == torch.float64 and out["dE_dstrain"].dtype == torch.float64
    assert (out["dE_dpos"].float() - ref["dE_dpos"]).abs().max() <= 1e-6
    assert (out["dE_dstrain"].float() - ref["dE_dstrain"]).abs().max() <= 1e-5
    assert (out["energies"] - ref["energies"]).abs().max() <= 1e-6
    # insert an empty structure as system 2: later systems shift up by one
    sysi = inp["system_indices"].clone()
    sysi[sysi >= 2] += 1
    cells = torch.cat([inp["cells"][:2], torch.eye(3, device=DEV)[None] * 7.0, inp["cells"][2:]])
    out = evaluate(be, **dict(inp, system_indices=sysi, cells=cells), target=g["target"])
    assert out["energies"].shape[0] == 6 and float(out["energies"][2].abs().sum()) == 0.0
    keep = [0, 1, 3, 4, 5]
    assert (out["energies"][keep] - ref["energies"]).abs().max() <= 1e-6
    assert (out["dE_dpos"] - ref["dE_dpos"]).abs().max() <= 1e-6


def test_default_dtype_float64_does_not_leak_into_work_buffers():
    """Atomistic test suites often run under torch.set_default_dtype(torch.float64): every buffer
    handed to the C ABI must still be fp32."""
    g = load_golden("si_64")
    be = make_backend(g, "bf16x3")    # fp32 parameters (seeded init under the fp32 default)
    ref = evaluate(be, **golden_inputs(g, DEV), target=g["target"], strain=True)
    torch.set_default_dtype(torch.float64)
    try:
        be._invalidate_packed()       # weights are re-packed under the fp64 default as well
        out = evaluate(be, **golden_inputs(g, DEV), target=g["target"], strain=True)
    finally:
        torch.set_default_dtype(torch.float32)
    for k in ("energies", "dE_dpos"):
        assert torch.equal(out[k].float(), ref[k]), k
    # the strain gradient is assembled by torch matmuls of the autograd graph (pos^T d_pos): under
    # the fp64 default those run in the inputs' fp32 all the same, but not bit-reproducibly
    torch.testing.assert_close(out["dE_dstrain"].float(), ref["dE_dstrain"], rtol=1e-5, atol=1e-4)


def test_unsupported_atomic_type_and_foreign_device_raise():
    g = load_golden("qm9_5")            # atomic_types [1, 6, 7, 8]
    be = make_backend(g)
    inp = golden_inputs(g, DEV)
    inp["species"] = inp["species"].clone()
    inp["species"][3] = 5               # inside the lookup table, not a model type
    with pytest.raises(ValueError, match="atomic types .*5"):
        evaluate(be, **inp, target=g["target"])
    inp["species"][3] = 117             # beyond the lookup table
    with pytest.raises(ValueError, match="atomic types"):
        evaluate(be, **inp, target=g["target"])
    if torch.cuda.device_count() > 1:
        with pytest.raises(RuntimeError, match="current CUDA device"):
            evaluate(be, **golden_inputs(g, "cuda:1"), target=g["target"])
        torch.cuda.synchronize()


def test_evaluator_loop_matches_reference_semantics():
    """eval_targets mirrors cli/eval.py:_eval_targets: batching (last batch smaller), per-atom
    energy metrics, force metrics, predictions equal to direct evaluation."""
    from metatrain_b200 import eval_targets
    g = load_golden("carbon_5")
    be = make_backend(g)
    structures, targets = [], []
    for b in range(5):
        sel = g["system_indices"] == b
        structures.append(dict(Z=g["species"][sel], positions=g["positions"][sel].astype(np.float64),
                               cell=g["cells"][b].astype(np.float64), pbc=True))
        targets.append(dict(energy=g["ref32_energies"][b], forces=-g["ref32_dE_dpos"][sel]))
    res = eval_targets(be, structures, targets, target=g["target"], batch_size=2, device=DEV)
    assert len(res["energies"]) == 5 and len(res["forces"]) == 5
    e = np.array([float(x) for x in res["energies"]])
    assert np.abs(e - g["ref32_energies"].ravel()).max() <= 1e-4
    assert res["metrics"][g["target"] + " forces RMSE"] <= 2e-5
    assert res["metrics"][g["target"] + " (per atom) MAE"] <= 1e-5
    assert res["ms_per_atom"][0] > 0


def test_pipelined_evaluator_overlaps_copies_without_changing_results():
    """PipelinedEvaluator: the H2D copy of the next batch runs on a copy stream while this one is
    evaluated; results (pinned host tensors) equal a plain evaluate() of the same batch."""
    from metatrain_b200.eval_loop import PipelinedEvaluator
    cases = [load_golden(c) for c in ("water_384", "si_64", "water_384")]
    be = {c["target"]: None for c in cases}
    g0 = cases[0]
    backend = make_backend(g0, "bf16x3")
    hosts = [{k: v.pin_memory() for k, v in golden_inputs(g0).items()} for _ in range(3)]
    hosts[1]["positions"] = (hosts[1]["positions"] + 0.01).pin_memory()
    ev = PipelinedEvaluator(backend, g0["target"], device=DEV)
    ticket = ev.submit(hosts[0])
    for k in range(3):
        upcoming = ev.submit(hosts[(k + 1) % 3])
        got = {n: t.clone() for n, t in ev.run(ticket).items()}
        ticket = upcoming
        ref = evaluate(backend, **{n: t.to(DEV) for n, t in hosts[k].items()}, target=g0["target"])
        assert torch.equal(got["energies"], ref["energies"].cpu())
        assert torch.equal(got["dE_dpos"], ref["dE_dpos"].cpu())
    assert be is not None


@pytest.mark.parametrize("case", ["water_384", "carbon_5", "si_64", "qm9_5", "co_periodic", "ragged_mix"])
def test_gpu_neighbor_list_equals_host_list(case):
    """petb200_nl_count/fill (GPU cell list) produce the same pair set as the host list that
    the golden files were generated with (after the model's own cutoff filter)."""
    from metatrain_b200.neighbors_gpu import neighbor_list_gpu
    g = load_golden(case)
    for b in range(g["cells"].shape[0]):
        sel = g["system_indices"] == b
        pos = torch.tensor(g["positions"][sel], device=DEV)
        cell = torch.tensor(g["cells"][b], device=DEV)
        periodic = bool(np.abs(g["cells"][b]).sum() > 0)
        i, j, s = neighbor_list_gpu(pos, cell, periodic, 4.5)
        # exact cutoff decision in fp64 on the same fp32 inputs
        p64, c64 = g["positions"][sel].astype(np.float64), g["cells"][b].astype(np.float64)
        i, j, s = i.cpu().numpy(), j.cpu().numpy(), s.cpu().numpy()
        d = np.linalg.norm(p64[j] - p64[i] + s @ c64, axis=1) if len(i) else np.zeros(0)
        assert (d <= 4.5 * (1 + 1e-5)).all()
        keep = d <= 4.5
        got = set(zip(i[keep].tolist(), j[keep].tolist(), map(tuple, s[keep].tolist())))
        off = int(np.nonzero(sel)[0][0]) if sel.any() else 0
        e = np.isin(g["centers"], np.nonzero(sel)[0])
        ref = set(zip((g["centers"][e] - off).tolist(), (g["neighbors"][e] - off).tolist(),
                      map(tuple, g["cell_shifts"][e].tolist())))
        assert len(got) == keep.sum() and got == ref
        assert (np.diff(i) >= 0).all()  # grouped by centre


def test_gpu_neighbor_list_end_to_end_10k(water_10k):
    """Positions in, energies + forces out, nothing but positions crosses PCIe."""
    from metatrain_b200.neighbors_gpu import neighbor_list_gpu
    g, be, batch, out = water_10k
    i, j, s = neighbor_list_gpu(batch["positions"], batch["cells"][0], True, 4.5)
    assert i.shape[0] >= 392040
    out2 = evaluate(be, batch["positions"], i, j, batch["species"], batch["cells"], s,
                    batch["system_indices"], target=g["target"])
    assert abs(float(out2["energies"]) - float(out["energies"])) <= 2e-6 * abs(float(out["energies"]))
    assert (out2["dE_dpos"] - out["dE_dpos"]).abs().max() <= 2e-5


def test_verlet_list_reuse_matches_fresh_lists():
    """MD-style loop: the skin list is reused while atoms move < skin/2 and gives the energies and
    forces of a freshly built exact list at every step (the backend drops the extra pairs)."""
    from metatrain_b200.neighbors_gpu import VerletNeighborList, neighbor_list_gpu
    g = load_golden("water_384")
    be = make_backend(g, precision="bf16x3")
    be.emit_nef = False
    box = water_384()
    cell = torch.tensor(box["cell"], dtype=torch.float32, device=DEV)
    pos = torch.tensor(box["positions"], dtype=torch.float32, device=DEV)
    z = torch.tensor(box["Z"], device=DEV)
    sysidx = torch.zeros(len(z), dtype=torch.int64, device=DEV)
    cutoff = float(g["hypers"]["cutoff"])
    vl = VerletNeighborList(cutoff, skin=0.6, periodic=True)
    gen = torch.Generator(device="cpu").manual_seed(3)
    for step in range(6):
        i, j, s = vl.update(pos, cell)
        out = evaluate(be, pos, i, j, z, cell[None], s, sysidx, target=g["target"])
        fi, fj, fs = neighbor_list_gpu(pos, cell, True, cutoff)
        ref = evaluate(be, pos, fi, fj, z, cell[None], fs, sysidx, target=g["target"])
        assert abs(float(out["energies"]) - float(ref["energies"])) <= 2e-5 * abs(float(ref["energies"]))
        assert (out["dE_dpos"] - ref["dE_dpos"]).abs().max() <= 2e-5
        pos = pos + 0.04 * torch.randn(pos.shape, generator=gen).to(DEV)
    assert vl.n_builds >= 1 and vl.n_reuses >= 2 and vl.n_builds + vl.n_reuses == 6


def test_graphed_md_evaluator_on_a_molecule():
    """Non-periodic system (zero cell): the bounding-box neighbor list path of the MD evaluator."""
    from metatrain_b200 import GraphedEvaluator
    from metatrain_b200.neighbors_gpu import neighbor_list_gpu
    g = load_golden("qm9_5")
    be = make_backend(g, precision="bf16x3")
    inp = golden_inputs(g, DEV)
    sel = inp["system_indices"] == 3
    pos, species, cell = inp["positions"][sel].clone(), inp["species"][sel], torch.zeros(3, 3, device=DEV)
    md = GraphedEvaluator(be, species, cell, periodic=False, skin=0.4, target=g["target"])
    gen = torch.Generator(device="cpu").manual_seed(5)
    for _ in range(6):
        out = md(pos)
        e, f = out["energies"].clone(), out["dE_dpos"].clone()
        c, n, s = neighbor_list_gpu(pos, cell, False, be.cutoff)
        ref = evaluate(be, pos, c.long(), n.long(), species, cell[None], s,
                       torch.zeros(len(species), dtype=torch.long, device=DEV), target=g["target"])
        assert (e - ref["energies"]).abs().max() <= 2e-5 and (f - ref["dE_dpos"]).abs().max() <= 2e-5
        pos = pos + 0.05 * torch.randn(pos.shape, generator=gen).to(DEV)


@pytest.mark.parametrize("use_graph", [True, False])
def test_graphed_md_evaluator_matches_eager_steps(use_graph):
    """md.GraphedEvaluator (device Verlet list + one CUDA graph per list) against evaluate() with
    a fresh exact neighbor list, along a random walk that forces list rebuilds / re-captures."""
    from metatrain_b200 import GraphedEvaluator
    from metatrain_b200.neighbors_gpu import neighbor_list_gpu
    g = load_golden("si_64")
    be = make_backend(g, precision="bf16x3")
    inp = golden_inputs(g, DEV)
    cell = inp["cells"][0]
    md = GraphedEvaluator(be, inp["species"], cell, periodic=True, skin=0.3, target=g["target"],
                          use_graph=use_graph)
    gen = torch.Generator(device="cpu").manual_seed(3)
    pos = inp["positions"].clone()
    for step in range(12):
        out = md(pos)
        e, f = out["energies"].clone(), out["dE_dpos"].clone()
        c, n, s = neighbor_list_gpu(pos, cell, True, be.cutoff)
        ref = evaluate(be, pos, c.long(), n.long(), inp["species"], inp["cells"], s, inp["system_indices"],
                       target=g["target"])
        assert (e - ref["energies"]).abs().max() <= 2e-5 * max(1.0, float(ref["energies"].abs().max()))
        assert (f - ref["dE_dpos"]).abs().max() <= 2e-5, f"step {step}"
        pos = pos + 0.04 * torch.randn(pos.shape, generator=gen).to(DEV)
    assert md.verlet.n_builds >= 2 and md.verlet.n_reuses >= 4
    if use_graph:
        assert md.n_captures == md.verlet.n_builds and md.n_replays == 12


def test_graphed_md_evaluator_survives_many_list_rebuilds_on_a_small_periodic_box():
    """Small periodic box = most pairs cross a boundary: the device list's acceptance test is not
    bitwise symmetric there, the topology must come from the symmetric filter (no sporadic
    'neighbor list is not symmetric')."""
    from metatrain_b200.md import GraphedEvaluator
    g = load_golden("si_64")
    be = make_backend(g, "bf16x3")
    inp = golden_inputs(g, DEV)
    ev = GraphedEvaluator(be, inp["species"], inp["cells"][0], skin=0.05, target=g["target"], use_graph=False)
    gen = torch.Generator(device=DEV).manual_seed(3)
    pos = inp["positions"].clone()
    for _ in range(60):
        pos = pos + 0.03 * torch.randn(pos.shape, generator=gen, device=DEV)
        out = ev(pos)
        assert torch.isfinite(out["dE_dpos"]).all()
    assert ev.verlet.n_builds >= 20


def test_neighbor_order_invariance():
    """Shuffling the neighbor list changes nothing but fp summation order."""
    g = load_golden("si_64")
    be = make_backend(g)
    inp = golden_inputs(g, DEV)
    a = evaluate(be, **inp, target=g["target"])
    perm = torch.randperm(inp["centers"].shape[0], generator=torch.Generator().manual_seed(0)).to(DEV)
    inp2 = dict(inp, centers=inp["centers"][perm], neighbors=inp["neighbors"][perm],
                cell_shifts=inp["cell_shifts"][perm])
    b = evaluate(be, **inp2, target=g["target"])
    torch.testing.assert_close(a["energies"], b["energies"], atol=1e-4, rtol=1e-6)
    assert (a["dE_dpos"] - b["dE_dpos"]).abs().max() <= 2e-5


# ------------------------------------------------ BASELINE.json config 2: 10k-atom water
@pytest.fixture(scope="module", params=["fp32", "bf16x3"])
def water_10k(request):
    """The bench configuration, on the fp32 parity path and on the tensor-core product path
    (tcgen05 GEMMs, mma.sync attention, fused feed-forward kernels) that bench.py measures."""
    g = load_golden("water_384")
    be = make_backend(g, precision=request.param)
    be.emit_nef = False
    box = replicate(water_384(), (3, 3, 3))
    batch = {k: v.to(DEV) for k, v in make_batch([box], 4.5).items()}
    return g, be, batch, evaluate(be, **batch, target=g["target"])


def test_10k_replication_property(water_10k):
    """A 3x3x3 tiling of the 384-atom box has 27x its energy and tiled forces: pins the
    full-size configuration to the reference's golden for the seed box."""
    g, be, batch, out = water_10k
    n = 384
    assert batch["positions"].shape[0] == 27 * n and batch["centers"].shape[0] == 392040
    e_ref = float(g["ref64_energies"].ravel()[0])
    assert abs(float(out["energies"]) / 27.0 - e_ref) <= 2e-5 * abs(e_ref)
    f = out["dE_dpos"].cpu().numpy().reshape(27, n, 3)
    assert np.abs(f - g["ref64_dE_dpos"][None]).max() <= FORCE_TOL
    assert np.abs(out["atomic"].cpu().numpy().reshape(27, n) - g["ref64_atomic"].reshape(1, n)).max() <= 5e-5


def test_10k_physical_invariants(water_10k):
    g, be, batch, out = water_10k
    # Newton's third law: the net force vanishes
    assert out["dE_dpos"].sum(0).abs().max() <= 5e-3
    # rigid translation (atoms leave the home cell; shifts stay valid) leaves E unchanged
    moved = dict(batch, positions=batch["positions"] + torch.tensor([3.3, -1.7, 0.9], device=DEV))
    out2 = evaluate(be, **moved, target=g["target"])
    assert abs(float(out2["energies"]) - float(out["energies"])) <= 2e-6 * abs(float(out["energies"]))
    assert (out2["dE_dpos"] - out["dE_dpos"]).abs().max() <= FORCE_TOL
