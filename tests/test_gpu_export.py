"""B3 packaging: the hot path as TorchScript operators.  A scripted module over
``torch.ops.petb200.*`` must survive ``torch.jit.save`` / ``torch.jit.load`` (what
``AtomisticModel.save`` and ``load_atomistic_model`` do, src/metatrain/pet/model.py:990-1021,
src/metatrain/utils/io.py:183-184) and reproduce the reference goldens, forces included."""
import io

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from helpers import golden_inputs, load_golden, seed_all  # noqa: E402
from metatrain_b200 import B200PETBackend, evaluate  # noqa: E402
from metatrain_b200.export import ExportedPET, extension_libraries  # noqa: E402

DEV = "cuda:0"


def _backend(g):
    seed_all(0)
    be = B200PETBackend(g["hypers"], g["atomic_types"], precision="bf16x3")
    be.add_output(g["target"], {g["target"] + "___0": g["out_shape"]})
    return be.to(DEV).eval()


def _run(module, inp, strain=False):
    pos = inp["positions"].clone().requires_grad_(True)
    eps = torch.eye(3, device=DEV, requires_grad=True)
    p_in, c_in = (pos @ eps, inp["cells"] @ eps) if strain else (pos, inp["cells"])
    energies, atomic = module(p_in, inp["centers"], inp["neighbors"], inp["species"], c_in, inp["cell_shifts"],
                              inp["system_indices"])
    grads = torch.autograd.grad(energies.sum(), [pos] + ([eps] if strain else []))
    return energies.detach(), atomic.detach(), grads


@pytest.mark.parametrize("case", ["water_384", "qm9_5", "si_64", "ragged_mix", "carbon_5"])
def test_scripted_module_round_trips_and_matches_reference(case):
    g = load_golden(case)
    be = _backend(g)
    inp = golden_inputs(g, DEV)
    strain = "ref32_dE_dstrain" in g
    eager = evaluate(be, **inp, target=g["target"], strain=strain)
    scripted = torch.jit.script(ExportedPET(be, g["target"]))
    buf = io.BytesIO()
    torch.jit.save(scripted, buf)
    buf.seek(0)
    loaded = torch.jit.load(buf, map_location=DEV)
    for module in (scripted, loaded):
        energies, atomic, grads = _run(module, inp, strain)
        # same kernels in the same order as the eager backend: per-atom values are bit-identical
        # (the per-structure sum is an index_add here, a segmented sum there; the cutoff-factor
        # gradients of readout and attention are accumulated in a different order)
        assert torch.equal(atomic, eager["atomic"])
        torch.testing.assert_close(grads[0], eager["dE_dpos"], rtol=0, atol=2e-6)
        torch.testing.assert_close(energies, eager["energies"], rtol=2e-6, atol=1e-5)
        scale = max(1.0, float(np.abs(g["ref32_energies"]).max()))
        assert np.abs(energies.cpu().numpy() - g["ref32_energies"]).max() <= 1e-5 * scale
        assert np.abs(grads[0].cpu().numpy() - g["ref32_dE_dpos"]).max() <= 1e-4
        if strain:
            assert np.abs(grads[1].cpu().numpy() - g["ref32_dE_dstrain"]).max() <= 1e-4
    assert loaded.interaction_range == pytest.approx(be.cutoff * len(be.gnn_layers))
    # selected_atoms: only the selected atoms are summed, all of them take part in the message passing
    mask = torch.zeros(inp["positions"].shape[0], dtype=torch.bool, device=DEV)
    mask[::3] = True
    e_sel, a_sel = loaded(inp["positions"], inp["centers"], inp["neighbors"], inp["species"], inp["cells"],
                          inp["cell_shifts"], inp["system_indices"], mask)
    ref_sel = evaluate(be, **inp, target=g["target"], selected_atoms=mask, gradients=False)
    torch.testing.assert_close(e_sel, ref_sel["energies"], rtol=2e-6, atol=1e-5)
    assert torch.equal(a_sel[mask], ref_sel["atomic"])


def test_export_rejects_what_is_not_packaged_and_lists_its_extensions():
    g = load_golden("water_384_adaptive")
    with pytest.raises(NotImplementedError):
        ExportedPET(_backend(g), g["target"])
    libs = extension_libraries()
    assert [p.rsplit("/", 1)[1] for p in libs] == ["libpetb200.so", "libpetb200_torch.so"]


def test_scripted_module_rejects_bad_inputs():
    g = load_golden("qm9_5")
    module = torch.jit.script(ExportedPET(_backend(g), g["target"]))
    inp = golden_inputs(g, DEV)
    half = inp["centers"].shape[0] // 2
    with pytest.raises(RuntimeError, match="not symmetric"):
        module(inp["positions"], inp["centers"][:half], inp["neighbors"][:half], inp["species"], inp["cells"],
               inp["cell_shifts"][:half], inp["system_indices"])
    bad = inp["species"].clone()
    bad[0] = 5
    with pytest.raises(RuntimeError, match="atomic types"):
        module(inp["positions"], inp["centers"], inp["neighbors"], bad, inp["cells"], inp["cell_shifts"],
               inp["system_indices"])
