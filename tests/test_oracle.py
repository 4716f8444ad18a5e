"""CPU tests: pin the oracle (oracle/pet_oracle.py) against the reference's goldens.

* ``qm9_5`` energies are hard-coded in the reference itself
  (src/metatrain/pet/tests/test_regression.py:66-74);
* every other fixture under tests/golden/ was produced by running the unmodified
  reference backend (tests/golden/make_golden.py);
* when /root/reference is mounted the oracle is also compared live with it.
"""
import numpy as np
import pytest
import torch

from helpers import (DEFAULT_HYPERS, GOLDEN_CASES, apply_lora, golden_inputs, load_golden, load_long_box, seed_all,
                     weight_fingerprint)
from metatrain_b200.parameters import PETParameters
from oracle import pet_oracle, ref_loader
from oracle.structures import neighbor_list

REFERENCE_HARD_CODED = np.array(  # test_regression.py:66-74
    [1.146098375320, 0.171331465244, 0.539504408836, 0.861489117146, 0.177449733019])


def seeded_state_dict(g):
    seed_all(0)
    p = PETParameters(g["hypers"], g["atomic_types"])
    p.add_output(g["target"], {g["target"] + "___0": g["out_shape"]})
    extra = apply_lora(p, g)
    sd = p.state_dict()
    sd.update(extra)
    return sd


@pytest.mark.parametrize("case", GOLDEN_CASES + ["qm9_5_dff192"])
def test_seeded_weights_equal_reference(case):
    g = load_golden(case)
    sd = seeded_state_dict(g)
    sd.pop("lora.scaling", None)
    fp = weight_fingerprint(sd)
    assert fp.shape == g["weight_fingerprint"].shape
    np.testing.assert_array_equal(fp, g["weight_fingerprint"])


def test_oracle_reproduces_reference_hard_coded_energies():
    g = load_golden("qm9_5")
    out = pet_oracle.energy_and_gradients(
        seeded_state_dict(g), g["hypers"], **golden_inputs(g), target=g["target"],
        with_gradients=False)
    # the reference asserts with torch.testing.assert_close fp32 defaults
    torch.testing.assert_close(out["energies"].ravel(),
                               torch.tensor(REFERENCE_HARD_CODED, dtype=torch.float32))


@pytest.mark.parametrize("case", GOLDEN_CASES + ["qm9_5_dff192"])
def test_oracle_matches_golden(case):
    g = load_golden(case)
    strain = "ref32_dE_dstrain" in g
    out = pet_oracle.energy_and_gradients(
        seeded_state_dict(g), g["hypers"], **golden_inputs(g), target=g["target"],
        with_strain=strain)
    scale = max(1.0, float(np.abs(g["ref32_energies"]).max()))
    assert np.abs(out["energies"].numpy() - g["ref32_energies"]).max() <= 2e-6 * scale
    assert np.abs(out["atomic"].numpy() - g["ref32_atomic"]).max() <= 5e-6
    assert np.abs(out["dE_dpos"].numpy() - g["ref32_dE_dpos"]).max() <= 5e-6
    if strain:
        assert np.abs(out["dE_dstrain"].numpy() - g["ref32_dE_dstrain"]).max() <= 2e-5
    assert int(out["batch"]["mask"].sum()) == int(g["ref32_n_edges_kept"])


def test_oracle_matches_reference_on_elongated_box():
    """Elongated box of the atom-sharded runs (1x1x8 tiling, z to 125 A), fp32 coordinates: the oracle
    reproduces the unmodified reference's fp32 outputs, and the reference's own fp32-vs-fp64 force
    difference there (the accuracy floor of fp32 coordinates this far from the origin) is recorded."""
    g = load_long_box("water_long_1x1x8")
    np.testing.assert_array_equal(weight_fingerprint(seeded_state_dict(g)), g["weight_fingerprint"])
    batch = {k: (v.long() if not v.is_floating_point() else v) for k, v in g["batch"].items()}
    out = pet_oracle.energy_and_gradients(seeded_state_dict(g), g["hypers"], **batch, target=g["target"])
    assert abs(float(out["energies"]) - float(g["ref32_energies"].ravel()[0])) <= 2e-6 * abs(float(g["ref32_energies"].ravel()[0]))
    assert np.abs(out["dE_dpos"].numpy() - g["ref32_dE_dpos"]).max() <= 2e-5
    floor = np.abs(g["ref32_dE_dpos"] - g["ref64_dE_dpos"]).max()
    assert floor <= 1e-4


def test_oracle_fp64_matches_reference_fp64():
    g = load_golden("si_64")
    sd = {k: (v.double() if v.is_floating_point() else v) for k, v in seeded_state_dict(g).items()}
    out = pet_oracle.energy_and_gradients(
        sd, g["hypers"], **golden_inputs(g, dtype=torch.float64), target=g["target"],
        with_strain=True)
    assert np.abs(out["dE_dpos"].numpy() - g["ref64_dE_dpos"]).max() <= 1e-10
    assert np.abs(out["dE_dstrain"].numpy() - g["ref64_dE_dstrain"]).max() <= 1e-9


def test_nonstrict_equals_strict():
    # src/metatrain/pet/tests/test_non_strict_nl.py:175-215
    a, b = load_golden("water_384"), load_golden("water_384_nonstrict")
    assert len(b["centers"]) > len(a["centers"])
    np.testing.assert_allclose(a["ref32_energies"], b["ref32_energies"], rtol=2e-6)
    np.testing.assert_allclose(a["ref32_dE_dpos"], b["ref32_dE_dpos"], atol=5e-6)


def test_manual_attention_equals_sdpa():
    # src/metatrain/pet/tests/test_functionality.py:162-181
    g = load_golden("si_64")
    sd = seeded_state_dict(g)
    a = pet_oracle.energy_and_gradients(sd, g["hypers"], **golden_inputs(g), manual_attention=False)
    b = pet_oracle.energy_and_gradients(sd, g["hypers"], **golden_inputs(g), manual_attention=True)
    torch.testing.assert_close(a["atomic"], b["atomic"], atol=1e-5, rtol=1e-5)
    torch.testing.assert_close(a["dE_dpos"], b["dE_dpos"], atol=1e-5, rtol=1e-5)


@pytest.mark.parametrize("variant", ["fixed", "adaptive_solver", "adaptive_grid"])
def test_oracle_gradcheck_on_the_reference_autograd_system(variant):
    """The reference's autograd test (src/metatrain/utils/testing/autograd.py:24-94): two carbon
    atoms at (0,0,0) and (0.9,0.9,0.9) in a 2 A cubic cell, fp64, torch.autograd.gradcheck of the
    summed energy w.r.t. the positions — restated on the oracle, also through both adaptive-cutoff
    methods (whose gradient comes from the implicit-function step / the probe weights)."""
    hyp = dict(DEFAULT_HYPERS)
    if variant != "fixed":
        hyp.update(num_neighbors_adaptive=20.0, adaptive_cutoff_method=variant.split("_")[1])
    seed_all(0)
    p = PETParameters(hyp, [6])
    p.add_output("energy", {"energy___0": [1]})
    sd = {k: (v.double() if v.is_floating_point() else v) for k, v in p.state_dict().items()}
    pos0 = np.array([[0.0, 0.0, 0.0], [0.9, 0.9, 0.9]])
    cell = 2.0 * np.eye(3)
    i, j, S = neighbor_list(pos0, cell, True, 4.5)
    fixed = dict(centers=torch.tensor(i), neighbors=torch.tensor(j), species=torch.tensor([6, 6]),
                 cells=torch.tensor(cell)[None], cell_shifts=torch.tensor(S),
                 system_indices=torch.zeros(2, dtype=torch.long))

    def energy(positions):
        batch = pet_oracle.build_batch(positions, fixed["centers"], fixed["neighbors"], fixed["species"],
                                       fixed["cells"], fixed["cell_shifts"], fixed["system_indices"],
                                       sd["species_to_species_index"], hyp["cutoff"], hyp["cutoff_function"],
                                       hyp["cutoff_width"], False, hyp["num_neighbors_adaptive"],
                                       hyp["cutoff_width_adaptive"], hyp["adaptive_cutoff_method"])
        node, msg = pet_oracle.features(sd, hyp, batch)
        return pet_oracle.predict(sd, hyp, node, msg, batch, "energy").sum()

    positions = torch.tensor(pos0, dtype=torch.float64, requires_grad=True)
    # the solver's gradient is exact at the converged root only: 10 Newton steps leave ~1e-9
    tol = dict(atol=1e-6, rtol=1e-4) if variant == "adaptive_solver" else {}
    assert torch.autograd.gradcheck(energy, positions, fast_mode=True, **tol)


@pytest.mark.parametrize("case", ["train_qm9_2", "train_carbon_1"])
def test_oracle_parameter_gradients_match_reference_training_goldens(case):
    """What a training step differentiates (the CUDA path does not build
    weight gradients yet): loss = sum(E) + 0.1 sum |dE/dr|^2 with the force term from
    create_graph=True (output_gradient.py:34-40), manual attention as the reference uses when
    training (backend.py:380-384), fp64.  Pins the oracle's gradient w.r.t. EVERY parameter to the
    unmodified reference, so that the wgrad / double-backward kernels have a checker."""
    g = load_golden(case)
    seed_all(0)
    p = PETParameters(g["hypers"], g["atomic_types"])
    p.add_output(g["target"], {g["target"] + "___0": [1]})
    names = [n for n, _ in p.named_parameters()]
    assert names == [str(n) for n in g["train_param_names"]]
    w = {k: (v.double() if v.is_floating_point() else v) for k, v in p.state_dict().items()}
    for n in names:
        w[n] = w[n].clone().requires_grad_(True)
    inp = golden_inputs(g, dtype=torch.float64)
    pos = inp["positions"].clone().requires_grad_(True)
    batch = pet_oracle.build_batch(pos, inp["centers"], inp["neighbors"], inp["species"], inp["cells"],
                                   inp["cell_shifts"], inp["system_indices"], w["species_to_species_index"],
                                   g["hypers"]["cutoff"], g["hypers"]["cutoff_function"],
                                   g["hypers"]["cutoff_width"])
    node, msg = pet_oracle.features(w, g["hypers"], batch, manual_attention=True)
    atomic = pet_oracle.predict(w, g["hypers"], node, msg, batch, g["target"])
    energies = torch.zeros(inp["cells"].shape[0], atomic.shape[1], dtype=torch.float64).index_add_(
        0, inp["system_indices"], atomic)
    (de_dr,) = torch.autograd.grad(energies.sum(), pos, create_graph=True)
    loss = energies.sum() + 0.1 * (de_dr ** 2).sum()
    grads = torch.autograd.grad(loss, [w[n] for n in names], allow_unused=True)
    np.testing.assert_allclose(float(loss.detach()), float(g["train_loss"]), rtol=1e-10)
    rows = np.array([[0.0, 0.0] if gr is None else [float(gr.sum()), float((gr * gr).sum())] for gr in grads])
    np.testing.assert_allclose(rows, g["train_grad_fingerprint"], rtol=1e-7, atol=1e-12)


def test_oracle_neighbor_list_definition():
    """All ordered (i, j, S) with |r_j + S.cell - r_i| <= rc, no (i, i, 0) — checked
    against an O(N^2 * images) enumeration on the triclinic multi-image carbon cell."""
    g = load_golden("carbon_5")
    sel = g["system_indices"] == 0
    pos = g["positions"][sel].astype(np.float64)
    cell = g["cells"][0].astype(np.float64)
    i, j, S = neighbor_list(pos, cell, True, 4.5)
    found = set(zip(i.tolist(), j.tolist(), map(tuple, S.tolist())))
    expect = set()
    R = 4
    for a in range(-R, R + 1):
        for b in range(-R, R + 1):
            for c in range(-R, R + 1):
                s = np.array([a, b, c])
                d = np.linalg.norm(pos[None] + s @ cell - pos[:, None], axis=2)
                for ii, jj in zip(*np.nonzero(d <= 4.5)):
                    if ii == jj and a == b == c == 0:
                        continue
                    expect.add((int(ii), int(jj), (a, b, c)))
    assert found == expect
    # symmetric: the reverse of every edge exists (nef.py:88-166 relies on it)
    assert all((jj, ii, (-s[0], -s[1], -s[2])) in found for ii, jj, s in found)


@pytest.mark.skipif(not ref_loader.reference_available(), reason="reference not mounted")
def test_oracle_matches_live_reference():
    g = load_golden("ragged_mix")
    be = ref_loader.build_reference_backend(g["atomic_types"], g["target"], None).eval()
    inp = golden_inputs(g)
    pos = inp["positions"].clone().requires_grad_(True)
    bd = be.preprocess(pos, inp["centers"], inp["neighbors"], inp["species"], inp["cells"],
                       inp["cell_shifts"], inp["system_indices"], 1.0)
    nodes, edges = be.calculate_features(bd)
    pred, _, _ = be.predict(nodes, edges, bd, inp["cells"], inp["system_indices"], [g["target"]])
    (grad,) = torch.autograd.grad(pred[g["target"]][0].sum(), pos)
    out = pet_oracle.energy_and_gradients(be.state_dict(), g["hypers"], **inp, target=g["target"])
    torch.testing.assert_close(out["atomic"], pred[g["target"]][0].detach(), atol=2e-6, rtol=1e-5)
    torch.testing.assert_close(out["dE_dpos"], grad, atol=2e-6, rtol=1e-5)
    torch.testing.assert_close(out["node_features"], nodes[0].detach(), atol=1e-5, rtol=1e-5)
