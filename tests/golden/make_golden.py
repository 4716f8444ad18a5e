"""Generate the committed golden vectors by running the UNMODIFIED reference backend.

Run in the build container only (needs ``/root/reference``):

    python tests/golden/make_golden.py

For every case it stores the inputs (species, fp32 positions, cell, the neighbor list fed
to both sides) and the outputs of the reference's ``PETBackend.preprocess ->
calculate_features -> predict -> torch.autograd.grad`` (energies, per-atom energies,
+dE/dr, strain gradient, feature check-sums), fp32 and (for the accuracy floor) fp64,
plus a fingerprint of the seed-0 weights so the tests can prove that the product module
initialises identically to the reference.

Reference entry points exercised: ``src/metatrain/pet/modules/backend.py:238,344,420``;
the golden energies of case ``qm9_5`` are additionally hard-coded in the reference at
``src/metatrain/pet/tests/test_regression.py:66-74``.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import ref_loader  # noqa: E402
from oracle.structures import (  # noqa: E402
    neighbor_list,
    read_lammps_atomic,
    read_xyz_frames,
    silicon_box,
)

RES = os.path.join(ref_loader.REFERENCE_ROOT, "tests", "resources")


def weight_fingerprint(state_dict):
    """[n_tensors, 2] (sum, sum of squares) in float64, in state-dict order."""
    rows = []
    for v in state_dict.values():
        v64 = v.detach().to(torch.float64)
        rows.append([float(v64.sum()), float((v64 * v64).sum())])
    return np.array(rows, dtype=np.float64)


def batch_frames(frames, cutoff):
    pos, cen, nei, sh, Z, sysi, cells = [], [], [], [], [], [], []
    off = 0
    for k, f in enumerate(frames):
        i, j, S = neighbor_list(f["positions"], f["cell"], f["pbc"], cutoff)
        pos.append(f["positions"])
        cen.append(i + off)
        nei.append(j + off)
        sh.append(S)
        Z.append(f["Z"])
        sysi.append(np.full(len(f["Z"]), k, dtype=np.int64))
        cells.append(f["cell"])
        off += len(f["Z"])
    return dict(
        positions=np.concatenate(pos).astype(np.float32),
        centers=np.concatenate(cen).astype(np.int64),
        neighbors=np.concatenate(nei).astype(np.int64),
        cell_shifts=np.concatenate(sh).astype(np.int64).reshape(-1, 3),
        species=np.concatenate(Z).astype(np.int64),
        system_indices=np.concatenate(sysi).astype(np.int64),
        cells=np.stack(cells).astype(np.float32),
    )


def run_reference(backend, inp, target, dtype, with_strain):
    t = lambda a, dt=None: torch.tensor(a) if dt is None else torch.tensor(a).to(dt)  # noqa: E731
    pos = t(inp["positions"], dtype).requires_grad_(True)
    cells = t(inp["cells"], dtype)
    strain = torch.eye(3, dtype=dtype, requires_grad=True)
    pos_in, cells_in = (pos @ strain, cells @ strain) if with_strain else (pos, cells)
    sysi = t(inp["system_indices"])
    bd = backend.preprocess(pos_in, t(inp["centers"]), t(inp["neighbors"]), t(inp["species"]),
                            cells_in, t(inp["cell_shifts"]), sysi, 1.0)
    if "charge" in inp:  # what PET.forward adds for system conditioning (pet/model.py:464-471)
        bd["charge"], bd["spin_multiplicity"] = t(inp["charge"]), t(inp["spin_multiplicity"])
        bd["system_indices"] = sysi
    nodes, edges = backend.calculate_features(bd)
    pred, _, _ = backend.predict(nodes, edges, bd, cells_in, sysi, [target])
    atomic = pred[target][0]
    energies = torch.zeros((cells.shape[0],) + tuple(atomic.shape[1:]), dtype=dtype).index_add_(0, sysi, atomic)
    wrt = [pos] + ([strain] if with_strain else [])
    grads = torch.autograd.grad(energies.sum(), wrt)
    out = dict(
        energies=energies.detach().numpy(),
        atomic=atomic.detach().numpy(),
        dE_dpos=grads[0].numpy(),
        node_features_sum=np.array([float(nodes[0].double().sum()), float(nodes[0].double().abs().sum())]),
        edge_features_sum=np.array([float((edges[0] * bd["padding_mask"][..., None]).double().sum()),
                                    float((edges[0] * bd["padding_mask"][..., None]).double().abs().sum())]),
        n_edges_kept=np.array(int(bd["padding_mask"].sum())),
        atomic_cutoffs=bd["atomic_cutoffs_stats"].detach().numpy(),
    )
    if with_strain:
        out["dE_dstrain"] = grads[1].numpy()
    return out


ONLY = set(sys.argv[1:])  # optional: regenerate just these cases


def make_case(name, frames, atomic_types, target="energy", hypers=None, nl_cutoff=4.5,
              with_strain=False, fp64=True, charge=None, spin_multiplicity=None, out_shape=(1,)):
    if ONLY and name not in ONLY:
        path = os.path.join(HERE, name + ".npz")
        return dict(np.load(path)) if os.path.exists(path) else None
    inp = batch_frames(frames, nl_cutoff)
    if charge is not None:
        inp["charge"] = np.asarray(charge, dtype=np.int64)
        inp["spin_multiplicity"] = np.asarray(spin_multiplicity, dtype=np.int64)
    be32 = ref_loader.build_reference_backend(atomic_types, target, hypers, out_shape=out_shape).eval()
    fp = weight_fingerprint(be32.state_dict())
    ref32 = run_reference(be32, inp, target, torch.float32, with_strain)
    payload = dict(inp)
    payload["atomic_types"] = np.array(atomic_types, dtype=np.int64)
    payload["target"] = np.array(target)
    payload["hypers_override"] = np.array(repr(hypers or {}))
    payload["out_shape"] = np.array(out_shape, dtype=np.int64)
    payload["weight_fingerprint"] = fp
    for k, v in ref32.items():
        payload["ref32_" + k] = v
    if fp64:
        be64 = ref_loader.build_reference_backend(atomic_types, target, hypers, dtype=torch.float64,
                                                  out_shape=out_shape).eval()
        ref64 = run_reference(be64, inp, target, torch.float64, with_strain)
        for k in ("energies", "atomic", "dE_dpos", "dE_dstrain"):
            if k in ref64:
                payload["ref64_" + k] = ref64[k]
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **payload)
    n_e = int(ref32["n_edges_kept"])
    print(f"{name}: N={len(inp['species'])} E_in={len(inp['centers'])} E_kept={n_e} "
          f"E={ref32['energies'].ravel()[:5]} -> {os.path.getsize(path) / 1024:.0f} KiB")
    return payload


def make_training_case(name, frames, atomic_types, target="energy", hypers=None, nl_cutoff=4.5):
    """Pins what a training step differentiates (not built on the GPU yet):
    the reference backend in train mode, fp64, loss = sum(E) + 0.1 * sum(|dE/dr|^2) with the force
    term built by ``create_graph=True`` (src/metatrain/utils/output_gradient.py:34-40), gradient
    w.r.t. every parameter.  Stored: the loss and (sum, sum of squares) of every parameter gradient,
    in state-dict order."""
    if ONLY and name not in ONLY:
        return
    inp = batch_frames(frames, nl_cutoff)
    be = ref_loader.build_reference_backend(atomic_types, target, hypers, dtype=torch.float64).train()
    t = lambda a: torch.tensor(a)  # noqa: E731
    pos = t(inp["positions"]).double().requires_grad_(True)
    cells, sysi = t(inp["cells"]).double(), t(inp["system_indices"])
    bd = be.preprocess(pos, t(inp["centers"]), t(inp["neighbors"]), t(inp["species"]), cells,
                       t(inp["cell_shifts"]), sysi, 1.0)
    nodes, edges = be.calculate_features(bd)
    pred, _, _ = be.predict(nodes, edges, bd, cells, sysi, [target])
    atomic = pred[target][0]
    energies = torch.zeros(cells.shape[0], atomic.shape[1], dtype=torch.float64).index_add_(0, sysi, atomic)
    (de_dr,) = torch.autograd.grad(energies.sum(), pos, create_graph=True)
    loss = energies.sum() + 0.1 * (de_dr ** 2).sum()
    names = [n for n, _ in be.named_parameters()]
    grads = torch.autograd.grad(loss, list(be.parameters()), allow_unused=True)
    rows = [[0.0, 0.0] if gr is None else [float(gr.sum()), float((gr * gr).sum())] for gr in grads]
    payload = dict(inp)
    payload.update(atomic_types=np.array(atomic_types, dtype=np.int64), target=np.array(target),
                   hypers_override=np.array(repr(hypers or {})), out_shape=np.array([1], dtype=np.int64),
                   train_loss=np.array(float(loss)), train_grad_fingerprint=np.array(rows),
                   train_param_names=np.array(names))
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **payload)
    print(f"{name}: loss {float(loss):.9f}, {len(names)} parameter gradients -> {os.path.getsize(path) / 1024:.0f} KiB")


def make_long_box_case(name, frame, reps, atomic_types, fp64=True):
    """Elongated periodic boxes for the atom-sharded path (BASELINE.json configs[3] is a slab-sharded
    box ~8x longer than the seed box): the reference is run on the SAME fp32-rounded coordinates the
    GPU path sees (up to ~375 A at 1x1x24), fp32 and fp64.  Compact file: positions + outputs only;
    the tests rebuild the (deterministic) neighbor list themselves."""
    if ONLY and name not in ONLY:
        return
    from oracle.structures import replicate
    box = replicate(frame, reps)
    box["positions"] = box["positions"].astype(np.float32).astype(np.float64)  # what both sides are fed
    inp = batch_frames([box], 4.5)
    payload = dict(positions=inp["positions"], cells=inp["cells"], species=inp["species"],
                   reps=np.array(reps, dtype=np.int64), n_edges=np.array(len(inp["centers"])),
                   atomic_types=np.array(atomic_types, dtype=np.int64), target=np.array("energy"),
                   hypers_override=np.array(repr({})), out_shape=np.array([1], dtype=np.int64))
    be32 = ref_loader.build_reference_backend(atomic_types, "energy", None).eval()
    payload["weight_fingerprint"] = weight_fingerprint(be32.state_dict())
    ref32 = run_reference(be32, inp, "energy", torch.float32, False)
    for k in ("energies", "dE_dpos"):
        payload["ref32_" + k] = ref32[k]
    if fp64:
        be64 = ref_loader.build_reference_backend(atomic_types, "energy", None, dtype=torch.float64).eval()
        ref64 = run_reference(be64, inp, "energy", torch.float64, False)
        for k in ("energies", "dE_dpos"):
            payload["ref64_" + k] = ref64[k]
        print(f"{name}: reference fp32 vs fp64 force max-abs diff "
              f"{np.abs(ref32['dE_dpos'] - ref64['dE_dpos']).max():.3e}")
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **payload)
    print(f"{name}: N={len(inp['species'])} E={len(inp['centers'])} E={ref32['energies'].ravel()} "
          f"-> {os.path.getsize(path) / 1024:.0f} KiB")


def main():
    torch.set_num_threads(8)
    qm9 = read_xyz_frames(os.path.join(RES, "qm9_reduced_100.xyz"), 5)
    p = make_case("qm9_5", qm9, [1, 6, 7, 8], target="mtt::U0")
    hard_coded = np.array([1.146098375320, 0.171331465244, 0.539504408836,
                           0.861489117146, 0.177449733019])  # test_regression.py:66-74
    err = np.abs(p["ref32_energies"].ravel() - hard_coded).max()
    print("qm9_5 vs reference's hard-coded goldens: max abs diff", err)
    assert err < 1e-5

    carbon = read_xyz_frames(os.path.join(RES, "carbon_reduced_100.xyz"), 5)
    make_training_case("train_qm9_2", qm9[:2], [1, 6, 7, 8], target="mtt::U0")
    make_training_case("train_carbon_1", carbon[:1], [6])
    water = read_lammps_atomic(os.path.join(RES, "periodic_water.data"), {1: 1, 2: 8})
    make_case("water_384", [water], [1, 8])
    # elongated boxes of the atom-sharded runs: 1x1x8 (3 072 atoms, z to 125 A) and 1x1x24
    # (9 216 atoms, z to 376 A), same fp32-rounded coordinates on both sides
    make_long_box_case("water_long_1x1x8", water, (1, 1, 8), [1, 8])
    make_long_box_case("water_long_1x1x24", water, (1, 1, 24), [1, 8])
    # seed box of the 10k / 100k water benchmarks (positions only)
    np.savez_compressed(os.path.join(os.path.dirname(os.path.dirname(HERE)), "metatrain_b200",
                                     "data", "water_384.npz"),
                        Z=water["Z"], positions=water["positions"], cell=water["cell"])
    # non-strict neighbor list: pairs out to 5.5 A are handed in and must be dropped
    make_case("water_384_nonstrict", [water], [1, 8], nl_cutoff=5.5, fp64=False)

    # residual featurizer (backend.py:589-649)
    make_case("qm9_5_residual", qm9, [1, 6, 7, 8], target="mtt::U0", hypers=dict(featurizer_type="residual"))
    make_case("water_384_residual", [water], [1, 8], hypers=dict(featurizer_type="residual"), fp64=False)
    # PostLN transformer layers (transformer.py:236-262)
    make_case("qm9_5_postln", qm9, [1, 6, 7, 8], target="mtt::U0", hypers=dict(transformer_type="PostLN"))
    make_case("water_384_postln", [water], [1, 8], hypers=dict(transformer_type="PostLN"), fp64=False)
    # the original PET layer: PostLN + LayerNorm + SiLU + residual featurizer; PreLN with LayerNorm + SiLU
    classic = dict(transformer_type="PostLN", normalization="LayerNorm", activation="SiLU",
                   featurizer_type="residual")
    make_case("water_384_classic", [water], [1, 8], hypers=classic, fp64=False)
    make_case("qm9_5_classic", qm9, [1, 6, 7, 8], target="mtt::U0", hypers=classic)
    make_case("water_384_preln_ln_silu", [water], [1, 8],
              hypers=dict(normalization="LayerNorm", activation="SiLU"), fp64=False)

    # system conditioning (conditioning.py): per-system charge / spin embeddings added to the node
    # features after every GNN layer; the zero-initialised gate is re-drawn (seed 2) so that the
    # branch is active ("_gate_seed" is a test-only pseudo hyper, see oracle/ref_loader.py)
    make_case("qm9_5_conditioned", qm9, [1, 6, 7, 8], target="mtt::U0",
              hypers=dict(system_conditioning=True, _gate_seed=2),
              charge=[-1, 0, 1, 2, 0], spin_multiplicity=[1, 2, 3, 1, 2])
    make_case("qm9_5_conditioned_residual", qm9, [1, 6, 7, 8], target="mtt::U0", fp64=False,
              hypers=dict(system_conditioning=True, featurizer_type="residual", _gate_seed=2),
              charge=[3, -2, 0, 1, -10], spin_multiplicity=[10, 1, 2, 4, 3])

    # direct (non-conservative) stress head: [N, 9] -> [N, 3, 3, 1], / volume, symmetrised
    # (backend.py:483-490, 780-813); periodic boxes and a zero-cell molecule (volume -> inf)
    make_case("stress_head_mix", [carbon[0], carbon[1], qm9[0]], [1, 6, 7, 8], target="non_conservative_stress",
              out_shape=(3, 3), fp64=False)

    # LoRA adapters (finetuning.py:322-378) on the attention projections (the reference default
    # target modules) and on every feed-forward / compress Linear; adapters seeded with 1
    make_case("water_384_lora", [water], [1, 8], fp64=False,
              hypers=dict(_lora=dict(rank=4, alpha=8.0, seed=1, target_modules=["input_linear", "output_linear"])))
    make_case("qm9_5_lora_wide", qm9, [1, 6, 7, 8], target="mtt::U0",
              hypers=dict(_lora=dict(rank=8, alpha=4.0, seed=1,
                                     target_modules=["input_linear", "output_linear", "w_in", "w_out",
                                                     "center_contraction", "center_expansion"])))

    make_case("carbon_5", carbon, [6], with_strain=False)
    # other layer widths: feed-forward width a multiple of 64 inside (192) and outside (640) the fused
    # kernel's range
    make_case("qm9_5_dff192", qm9, [1, 6, 7, 8], target="mtt::U0", hypers=dict(d_feedforward=192), fp64=False)
    make_case("water_384_dff640", [water], [1, 8], hypers=dict(d_feedforward=640), fp64=False)

    # adaptive cutoff, solver method (adaptive_cutoff.py:110-229, structures.py:222-262)
    make_case("water_384_adaptive", [water], [1, 8], hypers=dict(num_neighbors_adaptive=16))
    make_case("qm9_5_adaptive", qm9, [1, 6, 7, 8], target="mtt::U0", hypers=dict(num_neighbors_adaptive=6))
    make_case("carbon_5_adaptive", carbon, [6], hypers=dict(num_neighbors_adaptive=10), with_strain=True)
    # the legacy grid method (adaptive_cutoff.py:232-395)
    grid = dict(adaptive_cutoff_method="grid")
    make_case("water_384_adaptive_grid", [water], [1, 8], hypers=dict(num_neighbors_adaptive=16, **grid))
    make_case("carbon_5_adaptive_grid", carbon, [6], hypers=dict(num_neighbors_adaptive=10, **grid),
              with_strain=True)

    si = silicon_box()
    make_case("si_64", [si], [14], with_strain=True)
    make_case("si_64_cosine", [si], [14], hypers=dict(cutoff_function="Cosine"), fp64=False)
    # 70 neighbours per atom (cutoff 7 A): rows longer than the 64 tokens the tensor-core attention
    # kernels hold, i.e. the long-row fallback of the bf16x3 path
    make_case("si_64_cutoff7", [si], [14], hypers=dict(cutoff=7.0), nl_cutoff=7.0, fp64=False)

    # the periodic 2-atom system of pet/tests/test_backend.py:68-80 (strain gradient)
    co = dict(Z=np.array([6, 8]), positions=np.array([[0.0, 0.0, 0.0], [1.5, 1.5, 1.5]]),
              cell=3.5 * np.eye(3), pbc=True)
    make_case("co_periodic", [co], [1, 6, 7, 8], with_strain=True)

    # H2O of test_backend.py:57-65 + isolated C (test_functionality.py:106-159) + a
    # dissociated pair, batched -> ragged rows, rows with zero neighbours
    h2o = dict(Z=np.array([8, 1, 1]),
               positions=np.array([[0.0, 0.0, 0.119], [0.0, 0.757, -0.477], [0.0, -0.757, -0.477]]),
               cell=np.zeros((3, 3)), pbc=False)
    lone = dict(Z=np.array([6]), positions=np.zeros((1, 3)), cell=np.zeros((3, 3)), pbc=False)
    pair = dict(Z=np.array([6, 6]), positions=np.array([[0.0, 0, 0], [0, 0, 100.0]]),
                cell=np.zeros((3, 3)), pbc=False)
    make_case("ragged_mix", [h2o, lone, pair, qm9[0]], [1, 6, 7, 8])
    # the same ragged batch through the adaptive cutoff: atoms without neighbours solve to the
    # maximum cutoff, rows shorter than the target keep every pair
    make_case("ragged_mix_adaptive", [h2o, lone, pair, qm9[0], carbon[0]], [1, 6, 7, 8],
              hypers=dict(num_neighbors_adaptive=8))
    # everything at once: original PET layer (PostLN + LayerNorm + SiLU), residual featurizer, adaptive
    # cutoff, system conditioning, LoRA adapters on every Linear of the transformer layers, Cosine
    # cutoff function, periodic and non-periodic structures in one batch, strain gradient
    make_case("kitchen_sink", [h2o, lone, pair, qm9[0], carbon[0], carbon[3]], [1, 6, 7, 8],
              hypers=dict(classic, num_neighbors_adaptive=8, system_conditioning=True, _gate_seed=2,
                          cutoff_function="Cosine",
                          _lora=dict(rank=4, alpha=8.0, seed=1,
                                     target_modules=["input_linear", "output_linear", "w_in", "w_out"])),
              charge=[0, 1, -1, 0, 2, 0], spin_multiplicity=[1, 2, 1, 3, 1, 2], with_strain=True, fp64=False)
    make_case("ragged_mix_adaptive_grid", [h2o, lone, pair, qm9[0], carbon[0]], [1, 6, 7, 8],
              hypers=dict(num_neighbors_adaptive=8, adaptive_cutoff_method="grid"))


if __name__ == "__main__":
    main()
