"""Multi-GPU parity: the atom-sharded evaluation (halo all-to-all over NCCL) must reproduce
the single-GPU result — which itself is pinned to the reference goldens."""
import os
import tempfile

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu

from helpers import LONG_BOX_CASES, load_golden, load_long_box, seed_all  # noqa: E402


def _make_backend(g, dev, precision):
    from metatrain_b200 import B200PETBackend
    seed_all(0)
    be = B200PETBackend(g["hypers"], g["atomic_types"], precision=precision)
    be.add_output(g["target"], {g["target"] + "___0": [1]})
    be = be.to(dev).eval()
    be.emit_nef = False
    return be


def _worker(rank, world, init_file, out_file, reps, precision, long_case=None):
    from metatrain_b200.neighbors import neighbor_list
    from metatrain_b200.sharded import build_shard, evaluate_sharded
    from metatrain_b200.systems import replicate, water_384
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", init_method=f"file://{init_file}", rank=rank, world_size=world,
                            device_id=dev)
    g = load_golden("water_384")
    be = _make_backend(g, dev, precision)
    if long_case is not None:
        lg = load_long_box(long_case)
        box = dict(positions=lg["positions"].astype(np.float64), cell=lg["cells"][0].astype(np.float64),
                   Z=lg["species"])
        nl = lg["nl"]
    else:
        box = replicate(water_384(), reps)
        nl = neighbor_list(box["positions"], box["cell"], True, 4.5)
    shard = build_shard(box["positions"], box["cell"], nl, rank, world)
    pos = torch.tensor(box["positions"], dtype=torch.float32, device=dev)
    species = torch.tensor(box["Z"], dtype=torch.int32, device=dev)
    cell = torch.tensor(box["cell"], dtype=torch.float32, device=dev)
    out = evaluate_sharded(be, shard, pos, species, cell, target=g["target"])
    if rank == 0:
        np.savez(out_file, energies=out["energies"].cpu().numpy(), dE_dpos=out["dE_dpos"].cpu().numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("precision", ["fp32", "bf16x3"])
@pytest.mark.parametrize("world,reps", [(2, (2, 1, 1)), (2, (1, 1, 1)), (1, (1, 1, 1))])
def test_sharded_matches_single_gpu(world, reps, precision):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    from metatrain_b200 import evaluate
    from metatrain_b200.systems import make_batch, replicate, water_384
    g = load_golden("water_384")
    with tempfile.TemporaryDirectory() as tmp:
        out_file = os.path.join(tmp, "out.npz")
        mp.spawn(_worker, args=(world, os.path.join(tmp, "init"), out_file, reps, precision),
                 nprocs=world, join=True)
        got = dict(np.load(out_file))
    be = _make_backend(g, "cuda:0", precision)
    batch = {k: v.to("cuda:0") for k, v in make_batch([replicate(water_384(), reps)], 4.5).items()}
    ref = evaluate(be, **batch, target=g["target"])
    e_ref = float(ref["energies"])
    assert abs(float(got["energies"].ravel()[0]) - e_ref) <= 2e-6 * abs(e_ref)
    assert np.abs(got["dE_dpos"] - ref["dE_dpos"].cpu().numpy()).max() <= 2e-5
    tiles = reps[0] * reps[1] * reps[2]
    f = got["dE_dpos"].reshape(tiles, 384, 3)
    assert np.abs(f - g["ref32_dE_dpos"][None]).max() <= 1e-4


@pytest.mark.parametrize("precision", ["fp32", "bf16x3"])
@pytest.mark.parametrize("world", [1, 2])
@pytest.mark.parametrize("case", LONG_BOX_CASES)
def test_sharded_elongated_box_matches_reference(case, world, precision):
    """BASELINE.json configs[3]-style boxes (slabs along a long axis) against the UNMODIFIED
    reference run on the same fp32-rounded coordinates (tests/golden/make_golden.py,
    make_long_box_case): forces within 1e-4 eV/A of the fp64 reference."""
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    g = load_long_box(case)
    with tempfile.TemporaryDirectory() as tmp:
        out_file = os.path.join(tmp, "out.npz")
        mp.spawn(_worker, args=(world, os.path.join(tmp, "init"), out_file, None, precision, case),
                 nprocs=world, join=True)
        got = dict(np.load(out_file))
    err64 = np.abs(got["dE_dpos"] - g["ref64_dE_dpos"]).max()
    err32 = np.abs(got["dE_dpos"] - g["ref32_dE_dpos"]).max()
    floor = np.abs(g["ref32_dE_dpos"] - g["ref64_dE_dpos"]).max()
    print(f"{case} world {world} {precision}: vs ref fp64 {err64:.2e}, vs ref fp32 {err32:.2e} (floor {floor:.2e})")
    assert err64 <= 1e-4
    assert err32 <= 1e-4 + floor
    e_ref = float(g["ref64_energies"].ravel()[0])
    assert abs(float(got["energies"].ravel()[0]) - e_ref) <= 2e-6 * abs(e_ref)
