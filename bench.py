#!/usr/bin/env python
"""Benchmark of the PET energy+forces hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl petb200|reference]

One "step" = one energy+forces evaluation (preprocess -> features -> predict ->
autograd.grad to positions) of the 10 368-atom periodic water box (BASELINE.json
configs[1]; the reference's 384-atom water fixture tiled 3x3x3).  Metric: atom-steps/s.

* ``value``       inputs (positions + neighbor list) resident in HBM, CUDA-event timing.
* ``e2e``         the same metric through the public API (`metatrain_b200.evaluate`) with
                  pinned HOST buffers: H2D of positions + neighbor list and D2H of
                  energies + forces inside the timed region of every step.
* ``roofline``    dominant kernel (the GEMM), timed per launch with CUDA events on the
                  launching stream in extra instrumented steps; algorithmic FLOPs =
                  sum 2*M*N*K of the launches.  ``edge_scatter`` = the HBM-bound message
                  reversal + LayerNorm kernel (2060 B of algorithmic traffic per edge).
* ``cpu_baseline``/``--impl reference``  the oracle port (oracle/pet_oracle.py: the
                  reference's algorithm with the reference's torch CPU primitives) on the
                  host cores, bounded sample of the same workload.

N > 1: one process per GPU (torchrun).  Default `--multi sharded`: ONE box N times larger
(3x3x3N tiling, 10 368 atoms per GPU) is sharded by atoms (metatrain_b200/sharded.py): slabs,
halo-edge all-to-all-v over NCCL per GNN layer forward and backward, energy / force all-reduce
-> weak scaling; value = atoms of the whole box x steps / max-over-ranks time.
`--multi independent`: each rank evaluates its own 10k box (no data-path collective).
"""
import argparse
import contextlib
import json
import os
import subprocess
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

CUTOFF = 4.5
REPS = (3, 3, 3)
TARGET = "energy"
METRIC = "atom-steps/sec (energy+forces), PET 10k-atom water box"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="petb200", choices=["petb200", "reference"])
    ap.add_argument("--precision", default=os.environ.get("PETB200_PRECISION", "bf16x3"),
                    choices=["bf16x3", "fp32", "bf16"])
    ap.add_argument("--reps", type=int, nargs=3, default=list(REPS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--multi", default="sharded", choices=["sharded", "independent"])
    ap.add_argument("--config4", action="store_true",
                    help="also time the 96 768-atom box of BASELINE.json configs[3] (always on at 8 GPUs)")
    return ap.parse_args()


def workload_config(reps, world, multi):
    """The `config` object, identical in the petb200 and the reference arm."""
    n = 384 * reps[0] * reps[1] * reps[2]
    return {"workload": f"PET default hypers, periodic water box {reps[0]}x{reps[1]}x{reps[2]} tiling "
                        f"of the 384-atom fixture ({n} atoms per GPU), cutoff 4.5 A, energy + forces "
                        "(BASELINE.json configs[1])",
            "atoms_per_gpu": n, "n_gpus": world,
            "parallelism": ("1 GPU" if world == 1 else
                            f"one {n * world}-atom box sharded by atoms over {world} GPUs, halo "
                            "all-to-all-v + all-reduce over NCCL" if multi == "sharded" else
                            f"{world} independent boxes (1 per GPU)"),
            "cache": "per-step working set (GBs of activations) >> 126 MB L2; no explicit flush"}


def hypers():
    from helpers import DEFAULT_HYPERS
    return dict(DEFAULT_HYPERS)


def seeded_state_dict():
    from helpers import seed_all
    from metatrain_b200.parameters import PETParameters
    seed_all(0)
    p = PETParameters(hypers(), [1, 8])
    p.add_output(TARGET, {TARGET + "___0": [1]})
    return p.state_dict()


# ------------------------------------------------------------------ CPU reference arm
def oracle_step(sd, hyp, batch):
    from oracle import pet_oracle
    return pet_oracle.energy_and_gradients(sd, hyp, **batch, target=TARGET)


def time_oracle(reps, steps, warmup):
    """atom-steps/s of the oracle port on the host cores for a `reps` tiling."""
    from metatrain_b200.systems import make_batch, replicate, water_384
    torch.set_num_threads(os.cpu_count())
    sd, hyp = seeded_state_dict(), hypers()
    batch = make_batch([replicate(water_384(), reps)], CUTOFF)
    batch = {k: (v.long() if not v.is_floating_point() else v) for k, v in batch.items()}
    n_atoms = batch["positions"].shape[0]
    for _ in range(warmup):
        oracle_step(sd, hyp, batch)
    t0 = time.perf_counter()
    for _ in range(steps):
        oracle_step(sd, hyp, batch)
    dt = (time.perf_counter() - t0) / max(steps, 1)
    return n_atoms / dt, dt, n_atoms


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # always the full BASELINE configs[1] box (same config as the GPU arm): ~6 s per step on the
    # GPU box's host cores
    reps = tuple(args.reps)
    value, dt, n_atoms = time_oracle(reps, args.steps, max(args.warmup, 1))
    sample = (f"the full per-GPU box: water {reps[0]}x{reps[1]}x{reps[2]} tiling ({n_atoms} atoms), "
              f"{args.steps} steps after {max(args.warmup, 1)} warm-up, oracle port on all host cores")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "atom-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(reps, int(os.environ.get("WORLD_SIZE", "1")), args.multi),
        "cpu_baseline": {"value": value, "unit": "atom-steps/s", "cores": os.cpu_count(),
                         "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "atom-steps/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# -------------------------------------------------------------------------- GPU arm
class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index):
        self.idx = device_index
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                 "-i", str(self.idx), "-lms", "100"], stdout=subprocess.PIPE,
                stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for row in out.strip().splitlines():
            f = [x.strip() for x in row.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, flag in zip(names, f[5:9]):
                if flag.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


class KernelTimer:
    """CUDA-event timing of individual C-ABI calls on the launching stream."""

    def __init__(self, names=None):
        self.names = set(names) if names is not None else None   # None: every entry point
        self.records = []
        self.shapes = []

    @contextlib.contextmanager
    def __call__(self, name, args):
        if self.names is not None and name not in self.names:
            yield
            return
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        yield
        b.record()
        work, byt = 0.0, 0.0
        if name == "gemm":
            m, n, k = args[6], args[7], args[8]
            work = 2.0 * m * n * k
            epi = args[16]
            out_cols = n // 2 if epi == 2 else (2 * n if epi == 4 else n)
            aux_cols = {1: n, 2: n, 3: n, 4: 2 * n, 5: n}.get(epi, 0) if (args[13] or args[14]) else 0
            extra = (n if args[11] else 0) + (n if args[17] else 0)
            byt = 4.0 * m * (k + out_cols + aux_cols + extra)  # algorithmic fp32 bytes
        elif name in ("combine_ln_fwd",):
            work = float(args[4])  # edges
        elif name == "combine_scatter_bwd":
            # the edge scatter proper: out[e] = base[e] + d_cat[e, :d] + d_cat[rev[e], d:]
            byt = float(args[3]) * (4.0 * 4 * args[4] + 4.0)
        elif name in ("combine_fwd", "combine_bwd"):
            # fused combine block: 2 contractions (256x256 + 256x128), gather of t / t[rev]
            edges = float(args[7] if name == "combine_fwd" else args[10])
            work = 2.0 * edges * (256 * 256 + 256 * 128)
            byt = edges * (3080.0 if name == "combine_fwd" else 3592.0)
        elif name in ("mlp_fwd", "mlp_bwd"):
            # fused feed-forward block: rows, d, d_ff -> algorithmic flops (2 GEMMs forward, 3
            # backward incl. the recomputation) and MINIMUM bytes (x in, y out / x, dy in, dx out)
            fwd = name == "mlp_fwd"
            rows, d, dff = (args[5], args[6], args[7]) if fwd else (args[6], args[7], args[8])
            work = 2.0 * rows * d * 3 * dff * (1.0 if fwd else 5.0 / 3.0)
            byt = 4.0 * rows * d * (2 if fwd else 3)
        elif name in ("attention_fwd", "attention_bwd"):
            # tokens = E + N rows; fwd reads qkv (3d) writes out (d) + lse; bwd reads qkv, out, d_out
            # and writes d_qkv
            fwd = name == "attention_fwd"
            n_atoms, n_edges, heads, hd = (args[3], args[4], args[5], args[6]) if fwd else (args[6], args[7], args[8], args[9])
            d = heads * hd
            byt = 4.0 * (n_atoms + n_edges) * ((4 * d + heads) if fwd else (8 * d + 2 * heads))
        self.records.append((name, a, b, work, byt))
        if name == "gemm":
            self.shapes.append((len(self.records) - 1,
                                (int(args[6]), int(args[7]), int(args[8]), int(args[16]),
                                 int(args[17]), int(bool(args[11])))))

    def shape_totals(self):
        torch.cuda.synchronize()
        out = {}
        for idx, key in self.shapes:
            _, a, b, _, byt = self.records[idx]
            t, n, by = out.get(key, (0.0, 0, 0.0))
            out[key] = (t + a.elapsed_time(b) * 1e-3, n + 1, by + byt)
        return out

    def totals(self):
        torch.cuda.synchronize()
        out = {}
        for name, a, b, work, byt in self.records:
            t, w, n, by = out.get(name, (0.0, 0.0, 0, 0.0))
            out[name] = (t + a.elapsed_time(b) * 1e-3, w + work, n + 1, by + byt)
        return out


def peaks():
    """(HBM GB/s, bf16 TFLOP/s burst, bf16 TFLOP/s sustained, source).  The driver-written
    MEASURED_PEAKS.json is read tolerantly (its key names are not part of this repo): numbers are
    picked by what their (possibly nested) key says and by plausibility; anything not found falls
    back to the figure of /opt/skills/guides/B200_PROFILING.md."""
    hbm, burst, sust, src = 6650.0, 1590.0, 1400.0, "fallback"
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        flat = {}

        def walk(prefix, node):
            if isinstance(node, dict):
                for k, v in node.items():
                    walk(prefix + "." + str(k).lower(), v)
            elif isinstance(node, (int, float)) and not isinstance(node, bool):
                flat[prefix] = float(node)

        walk("", json.load(open(path)))
        bw = [v for k, v in flat.items() if any(t in k for t in ("hbm", "copy", "bandwidth", "gbs", "gb_s", "gbps"))
              and 1000.0 <= v <= 12000.0]
        tf = {k: v for k, v in flat.items() if any(t in k for t in ("bf16", "tflop", "tf_s", "tfs", "gemm", "cublas"))
              and 100.0 <= v <= 5000.0}
        got = False
        if bw:
            hbm, got = max(bw), True
        if tf:
            sus = [v for k, v in tf.items() if "sustain" in k]
            bur = [v for k, v in tf.items() if "burst" in k or "peak" in k]
            burst = max(bur) if bur else max(tf.values())
            sust = min(sus) if sus else min(tf.values())
            got = True
        if got:
            src = "measured"
    except (OSError, ValueError):
        pass
    return hbm, burst, sust, src


def run_petb200(args):
    import torch.distributed as dist

    from helpers import seed_all
    from metatrain_b200 import B200PETBackend, evaluate, lib
    from metatrain_b200.systems import make_batch, replicate, water_384

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # NCCL's own logging is left exactly as the caller configured it (the driver reads the
        # communicator sizes from NCCL_DEBUG=INFO output); only when nothing is configured is the
        # version banner kept off stdout (= 1 JSON line)
        if "NCCL_DEBUG" not in os.environ:
            os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)

    seed_all(0)
    be = B200PETBackend(hypers(), [1, 8], precision=args.precision)
    be.add_output(TARGET, {TARGET + "___0": [1]})
    be = be.to(dev).eval()
    be.emit_nef = False  # energies + forces only need the CSR handles

    sharded = world > 1 and args.multi == "sharded"

    def sharded_case(box):
        """One box sharded by atoms over all ranks: (resident step, e2e step, H2D tensors, edges)."""
        from metatrain_b200.neighbors import neighbor_list
        from metatrain_b200.sharded import (build_shard, evaluate_sharded, shard_to_device,
                                            shard_to_host_tensors)
        nl = box.get("nl") or neighbor_list(box["positions"], box["cell"], True, CUTOFF)
        shard = build_shard(box["positions"], box["cell"], nl, rank, world)
        host_l = shard_to_host_tensors(shard, pin_memory=True)
        host_pos = torch.tensor(box["positions"], dtype=torch.float32).pin_memory()
        host_z = torch.tensor(box["Z"], dtype=torch.int32).pin_memory()
        host_cell = torch.tensor(box["cell"], dtype=torch.float32).pin_memory()
        lists = shard_to_device(shard, dev, host_l)
        pos_d, z_d, cell_d = host_pos.to(dev), host_z.to(dev), host_cell.to(dev)
        n_all = len(box["Z"])
        e_h = torch.empty((1, 1), dtype=torch.float32).pin_memory()
        f_h = torch.empty((n_all, 3), dtype=torch.float32).pin_memory()

        def resident():
            return evaluate_sharded(be, shard, pos_d, z_d, cell_d, target=TARGET, device_lists=lists)

        def e2e():
            lists_in = shard_to_device(shard, dev, host_l)
            o = evaluate_sharded(be, shard, host_pos.to(dev, non_blocking=True),
                                 host_z.to(dev, non_blocking=True),
                                 host_cell.to(dev, non_blocking=True), target=TARGET,
                                 device_lists=lists_in)
            e_h.copy_(o["energies"], non_blocking=True)
            f_h.copy_(o["dE_dpos"], non_blocking=True)
            torch.cuda.synchronize()

        return dict(resident=resident, e2e=e2e, h2d=list(host_l.values()) + [host_pos, host_z, host_cell],
                    d2h=e_h.numel() * 4 + f_h.numel() * 4, n_edges=len(shard.centers), n_total=n_all,
                    halo_edges=int(sum(len(a) for a in shard.halo_recv)))

    def timed(fn, steps, warm=2):
        """ms per step of `fn`, CUDA events, max over ranks."""
        for _ in range(warm):
            fn()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(steps):
            fn()
        b.record()
        torch.cuda.synchronize()
        t = torch.tensor([a.elapsed_time(b) / steps], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    if sharded:
        reps = (args.reps[0], args.reps[1], args.reps[2] * world)
        box = replicate(water_384(), reps)
        case = sharded_case(box)
        step_resident, step_e2e, h2d_tensors = case["resident"], case["e2e"], case["h2d"]
        n_total, n_edges, d2h = case["n_total"], case["n_edges"], case["d2h"]
        n_atoms = n_total // world          # per-GPU share of the one big box
    else:
        box = replicate(water_384(), tuple(args.reps))
        host = make_batch([box], CUTOFF, pin_memory=True)
        resident = {k: v.to(dev) for k, v in host.items()}
        n_atoms = host["positions"].shape[0]
        n_edges = host["centers"].shape[0]
        n_total = n_atoms

        def step_resident():
            return evaluate(be, **resident, target=TARGET)

        e_host = torch.empty((1, 1), dtype=torch.float32).pin_memory()
        f_host = torch.empty((n_atoms, 3), dtype=torch.float32).pin_memory()
        d2h = e_host.numel() * 4 + f_host.numel() * 4

        # the batched evaluator loop in its throughput form: the H2D copy of the NEXT step's inputs is
        # enqueued on a copy stream before this step is evaluated (metatrain_b200.eval_loop.
        # PipelinedEvaluator); every step still pays its own H2D and D2H inside the timed region
        from metatrain_b200.eval_loop import PipelinedEvaluator
        pipe = PipelinedEvaluator(be, TARGET, device=str(dev))
        ticket = [pipe.submit(host)]

        def step_e2e():
            upcoming = pipe.submit(host)
            res = pipe.run(ticket[0])
            ticket[0] = upcoming
            return res

        h2d_tensors = list(host.values())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        out = step_resident()

    # ---- accuracy next to the throughput (any failure makes the run exit non-zero)
    from helpers import load_golden, load_long_box
    failures = []
    g = load_golden("water_384")
    tiles = n_total // 384
    f = out["dE_dpos"].cpu().numpy().reshape(tiles, 384, 3)
    accuracy = {}
    if not sharded:
        # 27 periodic copies of the seed box: every tile must reproduce the unmodified
        # reference's forces for the seed box (tests/golden/water_384.npz)
        force_err = float(np.abs(f - g["ref32_dE_dpos"][None]).max())
        energy_err = float(abs(float(out["energies"]) / tiles - float(g["ref32_energies"].ravel()[0])) / 384)
    else:
        # (1) the sharded evaluation against a single-GPU evaluation of the SAME box
        force_err_single = None
        if rank == 0:
            whole = {k: v.to(dev) for k, v in make_batch([box], CUTOFF).items()}
            ref_out = evaluate(be, **whole, target=TARGET)
            force_err_single = float((ref_out["dE_dpos"] - out["dE_dpos"]).abs().max())
            del whole, ref_out
            torch.cuda.empty_cache()
            if not force_err_single <= 1e-6:
                failures.append(f"sharded vs single-GPU forces differ by {force_err_single:.3e}")
        accuracy["force_max_abs_err_vs_single_gpu_eV_per_A"] = force_err_single
        # (2) the sharded path against the UNMODIFIED reference on an elongated box fed with the
        # same fp32-rounded coordinates (tests/golden/water_long_1x1x24.npz: 9 216 atoms, z to
        # 376 A; reference fp32 and fp64)
        lg = load_long_box("water_long_1x1x24")
        long_box = dict(positions=lg["positions"].astype(np.float64), cell=lg["cells"][0].astype(np.float64),
                        Z=lg["species"], nl=lg["nl"])
        lo = sharded_case(long_box)["resident"]()
        fl = lo["dE_dpos"].cpu().numpy()
        force_err = float(np.abs(fl - lg["ref64_dE_dpos"]).max())
        accuracy["long_box_1x1x24_force_err_vs_reference_fp64"] = force_err
        accuracy["long_box_1x1x24_force_err_vs_reference_fp32"] = float(np.abs(fl - lg["ref32_dE_dpos"]).max())
        accuracy["long_box_1x1x24_reference_fp32_vs_fp64"] = float(np.abs(lg["ref32_dE_dpos"] - lg["ref64_dE_dpos"]).max())
        energy_err = float(abs(float(lo["energies"]) - float(lg["ref64_energies"].ravel()[0])) / len(fl))
        # (3) the timed box itself: all tiles against the seed golden; coordinates up to
        # ~47 * world A from the origin are rounded differently from the golden's inputs, which
        # alone moves forces by ~1e-4 (see (2) for the like-for-like check)
        accuracy["timed_box_all_tiles_vs_seed_golden"] = float(np.abs(f - g["ref32_dE_dpos"][None]).max())
        del lo
        torch.cuda.empty_cache()
        for _ in range(2):   # the checks released the allocator cache: warm it up again
            step_resident()
    if not force_err <= 1e-4:
        failures.append(f"force max-abs-err {force_err:.3e} eV/A exceeds 1e-4")

    sampler = ClockSampler(local)
    barrier()
    if rank == 0:
        sampler.start()
    lib.launch_count = 0
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_start.record()
    for _ in range(args.steps):
        step_resident()
    t_end.record()
    barrier()
    launches = lib.launch_count
    clocks = sampler.stop() if rank == 0 else None
    sec = t_start.elapsed_time(t_end) * 1e-3
    if world > 1:
        t = torch.tensor([sec], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        sec = float(t)
    value = world * n_atoms * args.steps / sec

    # end to end through the public API with host buffers
    for _ in range(2):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    barrier()
    e2e_sec = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_sec], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_sec = float(t)
    e2e_value = world * n_atoms * args.steps / e2e_sec
    h2d = sum(v.numel() * v.element_size() for v in h2d_tensors)

    # ---- BASELINE.json's other multi-GPU readings of the metric (extra keys, fewer steps)
    strong, config4 = None, None
    if sharded:
        # strong scaling: the SAME 10 368-atom box over all GPUs ("10k-atom box @1/2/4/8")
        sbox = replicate(water_384(), tuple(args.reps))
        sc = sharded_case(sbox)
        ms = timed(sc["resident"], args.steps)
        strong = {"atoms": sc["n_total"], "n_gpus": world, "ms_per_step": ms,
                  "value": sc["n_total"] / ms * 1e3, "unit": "atom-steps/s", "scaling": "strong",
                  "edges_per_gpu": sc["n_edges"], "halo_edges_per_gpu": sc["halo_edges"]}
        del sc
        torch.cuda.empty_cache()
        if world == 8 or args.config4:
            # BASELINE.json configs[3]: the 6x6x7 tiling = 96 768 atoms over the GPUs of the box
            cbox = replicate(water_384(), (6, 6, 7))
            cc = sharded_case(cbox)
            ms = timed(cc["resident"], max(args.steps // 2, 3))
            config4 = {"workload": "water 6x6x7 tiling, 96 768 atoms (BASELINE.json configs[3])",
                       "atoms": cc["n_total"], "n_gpus": world, "ms_per_step": ms,
                       "value": cc["n_total"] / ms * 1e3, "unit": "atom-steps/s",
                       "edges_per_gpu": cc["n_edges"], "halo_edges_per_gpu": cc["halo_edges"]}
            del cc
            torch.cuda.empty_cache()

    # MD-engine style step: only positions go host->device, the neighbor list is rebuilt on
    # the GPU every step (petb200_nl_count / nl_fill), energies + forces come back
    md = None
    if not sharded:
        from metatrain_b200.neighbors_gpu import neighbor_list_gpu

        def step_md():
            pos_d = host["positions"].to(dev, non_blocking=True)
            i, j, sft = neighbor_list_gpu(pos_d, resident["cells"][0], True, CUTOFF)
            o = evaluate(be, pos_d, i, j, resident["species"], resident["cells"], sft,
                         resident["system_indices"], target=TARGET)
            e_host.copy_(o["energies"], non_blocking=True)
            f_host.copy_(o["dE_dpos"], non_blocking=True)
            torch.cuda.synchronize()

        for _ in range(2):
            step_md()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step_md()
        barrier()
        md_sec = time.perf_counter() - t0
        md = {"value": world * n_atoms * args.steps / md_sec, "unit": "atom-steps/s",
              "ms_per_step": md_sec / args.steps * 1e3,
              "h2d_bytes_per_step": host["positions"].numel() * 4, "d2h_bytes_per_step": d2h,
              "what": "positions H2D -> GPU cell-list neighbor list -> energy+forces -> D2H"}

        # the same with a Verlet (skin) list: thermal-size displacements every step, the list is
        # rebuilt only when an atom has moved more than skin / 2
        from metatrain_b200.neighbors_gpu import VerletNeighborList
        vl = VerletNeighborList(CUTOFF, skin=0.5, periodic=True)
        gen = torch.Generator().manual_seed(0)
        base_pos = host["positions"].clone()
        drift = torch.zeros_like(base_pos).pin_memory()

        def step_md_verlet():
            drift.add_(0.01 * torch.randn(base_pos.shape, generator=gen))
            pos_d = (base_pos + drift).pin_memory().to(dev, non_blocking=True)
            i, j, sft = vl.update(pos_d, resident["cells"][0])
            o = evaluate(be, pos_d, i, j, resident["species"], resident["cells"], sft,
                         resident["system_indices"], target=TARGET)
            e_host.copy_(o["energies"], non_blocking=True)
            f_host.copy_(o["dE_dpos"], non_blocking=True)
            torch.cuda.synchronize()

        for _ in range(2):
            step_md_verlet()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step_md_verlet()
        barrier()
        mdv_sec = time.perf_counter() - t0
        md["verlet"] = {"value": world * n_atoms * args.steps / mdv_sec, "unit": "atom-steps/s",
                        "ms_per_step": mdv_sec / args.steps * 1e3, "skin_A": 0.5,
                        "list_builds": vl.n_builds, "list_reuses": vl.n_reuses,
                        "what": "random-walk positions (0.01 A per step) H2D -> skin list reused while "
                                "max displacement < skin/2 -> energy+forces -> D2H"}

    # per-kernel roofline: instrumented extra steps (not part of the timed regions above)
    # (the per-op Python schedule issues the same kernels as the C++ stage schedule, one entry point
    # per kernel, so that each can be bracketed by events)
    from metatrain_b200 import engine as _engine
    timer = KernelTimer()
    lib.profile_hook = timer
    _engine.USE_STAGE_SCHEDULE = False
    for _ in range(3):
        step_resident()
    _engine.USE_STAGE_SCHEDULE = True
    lib.profile_hook = None
    tot = timer.totals()
    # communication share of the sharded step: every collective of 3 more steps (stage schedule on, as in
    # the timed region) bracketed by events on the compute stream, so a collective's time includes the wait
    # for the slowest peer.  Max over ranks.
    comm = None
    if sharded:
        from metatrain_b200 import sharded as _sh
        _sh.comm_timer = []
        for _ in range(3):
            step_resident()
        torch.cuda.synchronize()
        rec, _sh.comm_timer = _sh.comm_timer, None
        agg = {}
        for label, a, b, nb in rec:
            t, n, by = agg.get(label, (0.0, 0, 0))
            agg[label] = (t + a.elapsed_time(b), n + 1, by + nb)
        local = torch.tensor([agg.get(k, (0.0, 0, 0))[0] / 3 for k in ("halo_all_to_all", "all_reduce")],
                             device="cuda", dtype=torch.float64)
        hi, lo = local.clone(), local.clone()
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        comm = {"what": "collectives of the sharded step bracketed by CUDA events on the compute stream.  A rank "
                        "that reaches a collective early waits there for its peers, so max over ranks = transfer "
                        "+ load skew between ranks, min over ranks (the rank the others wait for) = the transfer "
                        "and launch cost itself",
                "halo_all_to_all": {"ms_per_step_min_over_ranks": float(lo[0]), "ms_per_step_max_over_ranks": float(hi[0]),
                                    "calls_per_step": agg.get("halo_all_to_all", (0, 0, 0))[1] // 3,
                                    "bytes_sent_per_step_this_rank": agg.get("halo_all_to_all", (0, 0, 0))[2] // 3},
                "all_reduce": {"ms_per_step_min_over_ranks": float(lo[1]), "ms_per_step_max_over_ranks": float(hi[1]),
                               "calls_per_step": agg.get("all_reduce", (0, 0, 0))[1] // 3,
                               "bytes_per_step": agg.get("all_reduce", (0, 0, 0))[2] // 3}}
    if os.environ.get("PETB200_GEMM_SHAPES") and rank == 0:
        shapes = {}
        for name, a, b, work, byt in timer.records:
            if name == "gemm":
                pass
        # re-walk with the raw args kept by the timer
        for key, (t_s, n_l, by) in sorted(timer.shape_totals().items(), key=lambda kv: -kv[1][0]):
            print(f"# gemm M={key[0]:7d} N={key[1]:5d} K={key[2]:5d} epi={key[3]} acc={key[4]} res={key[5]}: "
                  f"{n_l // 3:3d} launches/step {t_s / 3 * 1e3:8.3f} ms/step  {by / t_s * 1e-9:7.0f} GB/s",
                  file=sys.stderr)
    hbm, tf_burst, tf_sust, which = peaks()
    step_ms = sec / args.steps * 1e3
    # DRAM traffic per launch (dram__bytes_read.sum + dram__bytes_write.sum) of every kernel family,
    # written by tools/summarize_launches.py --json from the committed ncu launch list of this
    # command; only quoted for the workload it was captured on
    traffic = {}
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        if args.precision == "bf16x3" and tuple(args.reps) == tuple(REPS) and not sharded:
            traffic = tj.get("bytes_per_launch", {})
            traffic_src = tj.get("source")
    except (OSError, ValueError):
        pass

    def traffic_of(prefix):
        hits = [(k, v) for k, v in traffic.items() if k.startswith(prefix)]
        if not hits:
            return None
        n = sum(v["launches"] for _, v in hits)
        return sum(v["bytes_per_launch"] * v["launches"] for _, v in hits) / n

    # per C-ABI entry point: time, algorithmic flops / bytes, both roofline views
    kernels = {}
    for k, (t_s, work, n_l, byt) in tot.items():
        entry = {"ms_per_step": t_s / 3 * 1e3, "launches_per_step": n_l // 3,
                 "share_of_step": (t_s / 3 * 1e3) / step_ms}
        if byt:
            entry.update(achieved_gbs=byt / t_s * 1e-9, hbm_frac=byt / t_s * 1e-9 / hbm,
                         algorithmic_bytes_per_launch=byt / n_l)
        if work and k != "combine_ln_fwd":
            entry.update(achieved_tflops=work / t_s * 1e-12, tensor_frac=work / t_s * 1e-12 / tf_sust,
                         flops_per_launch=work / n_l)
        kernels[k] = entry
    # the dominant kernel family of the step: the dense contractions (gemm_tc_kernel).  SURVEY 8(d)
    # classifies them as tensor work: headline = algorithmic flops vs the measured sustained bf16
    # peak; the HBM view (they are tall-skinny, 32..128 flop/B) is the sub-key.
    g_t, g_flops, g_n, g_bytes = tot["gemm"]
    gemm_tflops = g_flops / g_t * 1e-12
    roofline = {
        "kernel": "gemm (all dense contractions of the step that are not inside a fused kernel: "
                  "gemm_tc_kernel)",
        "bound": "tensor", "achieved": gemm_tflops, "peak": tf_sust, "unit": "TFLOP/s",
        "frac": gemm_tflops / tf_sust, "traffic": traffic_of("gemm_tc_kernel"),
        "traffic_unit": "bytes per launch (ncu dram__bytes_read.sum + dram__bytes_write.sum, profiles/traffic.json)",
        "flops_per_launch": g_flops / g_n, "launches_per_step": g_n // 3,
        "ms_per_step": g_t / 3 * 1e3, "share_of_step": (g_t / 3 * 1e3) / step_ms,
        "precision": args.precision + " (3 MMAs per product: the issued-MMA ceiling is peak / 3)",
        "peak_source": which + " (MEASURED_PEAKS.json bf16_tflops_sustained)" if which == "measured" else which,
        "hbm_view": {"achieved": g_bytes / g_t * 1e-9, "peak": hbm, "unit": "GB/s",
                     "frac": g_bytes / g_t * 1e-9 / hbm, "algorithmic_bytes_per_launch": g_bytes / g_n},
    }
    # whole step: SURVEY 8(d) algorithmic flops (2.046 x forward) over the measured step time
    deg = np.bincount(resident["centers"].cpu().numpy(), minlength=n_atoms) if not sharded else None
    whole = None
    if deg is not None:
        f_fwd = 2001152.0 * n_edges + 4292864.0 * n_atoms + 2048.0 * float(((deg + 1.0) ** 2).sum())
        f_step = 2.046 * f_fwd
        whole = {"flops_per_step": f_step, "achieved_tflops": f_step / (step_ms * 1e-3) * 1e-12,
                 "peak_tflops": tf_sust, "frac": f_step / (step_ms * 1e-3) * 1e-12 / tf_sust}
    edge_scatter = None
    if "combine_ln_fwd" in tot:
        c_t, c_edges, c_n, _ = tot["combine_ln_fwd"]
        scatter_bytes = 2060.0 * c_edges  # 2x512 B read + 4 B rev + 1024 B write + 8 B stats per edge
        edge_scatter = {
            "kernel": "combine_ln_fwd (message reversal + LayerNorm)", "bound": "hbm",
            "achieved": scatter_bytes / c_t * 1e-9, "peak": hbm, "unit": "GB/s",
            "frac": scatter_bytes / c_t * 1e-9 / hbm, "traffic": traffic_of("combine_ln_fwd_kernel"),
            "algorithmic_bytes_per_launch": scatter_bytes / c_n, "peak_source": which,
            "avg_launch_us": c_t / c_n * 1e6}
    if "combine_scatter_bwd" in tot:
        c_t, _, c_n, byt = tot["combine_scatter_bwd"]
        edge_scatter = {
            "kernel": "combine_scatter_bwd (the edge scatter proper: gradient of the message reversal, "
                      "out[e] = base[e] + d_cat[e, :128] + d_cat[rev[e], 128:]; the forward gather is "
                      "fused into the tensor-bound combine_fwd kernel)", "bound": "hbm",
            "achieved": byt / c_t * 1e-9, "peak": hbm, "unit": "GB/s", "frac": byt / c_t * 1e-9 / hbm,
            "traffic": traffic_of("combine_scatter_bwd_kernel"),
            "algorithmic_bytes_per_launch": byt / c_n, "peak_source": which, "avg_launch_us": c_t / c_n * 1e6,
            "note": "algorithmic bytes = 2052 B/edge; the gathered half-rows are partly L2 hits, so the "
                    "fraction of the DRAM copy peak can exceed 1"}

    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    line = {
        "metric": METRIC, "value": value, "unit": "atom-steps/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": step_ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": {"fp32": "f32", "bf16x3": "bf16x3 (2-term split, fp32 accumulate)",
                  "bf16": "bf16"}[args.precision],
        "data": "synthetic",
        "config": workload_config(tuple(args.reps), world, args.multi),
        "details": {"atoms_total": n_total, "edges_per_gpu": n_edges},
        "force_max_abs_err_eV_per_A": force_err, "energy_abs_err_eV_per_atom": energy_err,
        "accuracy": accuracy, "parity_failures": failures,
        "e2e": {"value": e2e_value, "unit": "atom-steps/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "ms_per_step": e2e_sec / args.steps * 1e3},
        "e2e_device_neighbor_list": md,
        "strong_scaling": strong, "config4": config4, "comm": comm,
        "gpu_launches": launches, "clocks": clocks,
        "roofline": roofline, "whole_step": whole, "edge_scatter": edge_scatter, "kernels": kernels,
    }
    if not args.no_cpu_baseline and world == 1:
        cpu_steps = 6
        v, dt, n = time_oracle((2, 2, 2), cpu_steps, 1)  # ~15 s of CPU work on 16 cores (2.2 s/step)
        line["cpu_baseline"] = {
            "value": v, "unit": "atom-steps/s", "cores": os.cpu_count(), "kind": "port",
            "sample": f"water 2x2x2 tiling ({n} atoms) of the workload, {cpu_steps} steps after 1 warm-up "
                      f"({dt:.2f} s/step)"}
    print(json.dumps(line))
    if failures:
        print("PARITY FAILURE: " + "; ".join(failures), file=sys.stderr)
        sys.exit(3)


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_petb200(args)


if __name__ == "__main__":
    main()
