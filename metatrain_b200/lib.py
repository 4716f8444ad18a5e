"""ctypes binding of ``libpetb200.so`` (the C ABI declared in ``include/petb200.h``).

There is deliberately no CPU or PyTorch fallback: if the shared library is missing or a
call fails, a ``RuntimeError`` is raised (the reference's CLI wraps such errors as
``ArchitectureError``, ``src/metatrain/utils/errors.py:1-19``).
"""
import ctypes
import os
import subprocess
import threading

import torch

_CSRC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc")
#: PETB200_LIB selects another build of the same library (e.g. the `make chaos` stress build)
_LIB_PATH = os.path.abspath(os.environ["PETB200_LIB"]) if os.environ.get("PETB200_LIB") else os.path.join(
    _CSRC, "libpetb200.so")
_lock = threading.Lock()
_lib = None

# enums of include/petb200.h
EPI_NONE, EPI_SILU, EPI_SWIGLU, EPI_MUL_DSILU, EPI_SWIGLU_BWD, EPI_RMS_BWD = range(6)
PREC_FP32, PREC_BF16X3, PREC_BF16 = range(3)
CUTOFF_BUMP, CUTOFF_COSINE = range(2)

_P = ctypes.c_void_p
_I64 = ctypes.c_int64
_I = ctypes.c_int
_F = ctypes.c_float
_SZ = ctypes.c_size_t

# name -> argument ctypes (return type is int unless listed in _RESTYPE)
_SIGNATURES = {
    "petb200_nl_filter_count": [_P, _P, _P, _P, _P, _P, _I64, _I64, _F, _P, _P, _P],
    "petb200_csr_build_workspace": [_I64, _I64],
    "petb200_csr_build": [_P, _P, _P, _I64, _I64, _P, _P, _P, _P, _SZ, _P],
    "petb200_csr_gather": [_P, _P, _P, _P, _I64, _P, _P, _P, _P],
    "petb200_reverse_map": [_P, _P, _P, _P, _I64, _I64, _P, _P, _P],
    "petb200_nl_num_bins": [_P, _I, _F, _I64],
    "petb200_nl_workspace": [_I64, _I64],
    "petb200_nl_count": [_P, _I64, _P, _P, _I, _F, _P, _SZ, _P, _P],
    "petb200_nl_fill": [_I64, _P, _P, _I, _F, _P, _SZ, _P, _P, _P, _P, _P],
    "petb200_csr_to_nef": [_P, _P, _I64, _I64, _I, _I, _P, _P],
    "petb200_nef_to_csr": [_P, _P, _P, _I64, _I64, _I, _I, _P, _P],
    "petb200_edges_fwd": [_P, _P, _P, _P, _P, _P, _I64, _F, _F, _I, _P, _P, _P, _P],
    "petb200_edges_bwd": [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I64, _I64, _F, _F, _I,
                          _P, _P, _P, _P],
    "petb200_edge_grad": [_P, _P, _P, _P, _P, _I64, _F, _F, _I, _P, _P],
    "petb200_adaptive_cutoff_solve": [_P, _P, _I64, _F, _F, _F, _P, _P, _P, _P, _P],
    "petb200_adaptive_grid_solve": [_P, _P, _I64, _F, _F, _F, _F, _I, _P, _P, _P],
    "petb200_adaptive_grid_bwd": [_P, _P, _P, _P, _I64, _P, _P, _P, _I64, _F, _F, _F, _I, _P, _P, _P],
    "petb200_adaptive_pair_mask": [_P, _P, _P, _P, _I64, _P, _P, _P, _P],
    "petb200_edges_fwd_rc": [_P, _P, _P, _P, _P, _P, _I64, _P, _F, _I, _P, _P, _P, _P],
    "petb200_edges_bwd_rc": [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I64, _I64, _P, _F, _I, _P, _P, _P,
                             _P, _P],
    "petb200_adaptive_cutoff_bwd": [_P, _P, _P, _P, _P, _I64, _P, _P, _P, _I64, _F, _P, _P, _P],
    "petb200_force_scatter": [_P, _P, _P, _P, _P, _P, _I64, _I64, _P, _P, _P],
    "petb200_gemm": [_P, _I64, _P, _I64, _P, _I64, _I64, _I, _I, _P, _P, _P, _I64, _P, _P,
                     _I64, _I, _I, _I, _P],
    "petb200_split_bf16": [_P, _I64, _I, _P, _P],
    "petb200_compress_gemm": [_P, _I64, _P, _P, _P, _P, _P, _P, _P, _I64, _I, _P, _P, _I, _P],
    "petb200_norm_linear_image_bytes": [_I],
    "petb200_norm_linear_pack": [_P, _I, _I, _P, _P],
    "petb200_norm_linear": [_P, _I64, _P, _P, _I64, _I, _I, _P, _I64, _P, _P],
    "petb200_mlp_image_bytes": [_I, _I],
    "petb200_mlp_pack": [_P, _P, _I, _I, _P, _P, _P],
    "petb200_mlp_fwd": [_P, _I64, _P, _P, _P, _I64, _I, _I, _P, _I64, _P],
    "petb200_mlp_bwd": [_P, _I64, _P, _I64, _P, _P, _I64, _I, _I, _P, _I64, _P],
    "petb200_embedding": [_P, _P, _I64, _I, _P, _I64, _P],
    "petb200_add_gathered_rows": [_P, _P, _I64, _I, _P, _I64, _P],
    "petb200_transpose_scale": [_P, _I, _I, _P, _P, _P, _P],
    "petb200_compress_input": [_P, _P, _P, _P, _P, _P, _P, _I64, _I, _P, _P],
    "petb200_geom_embed_bwd": [_P, _I64, _P, _I64, _I, _I, _P, _P, _P],
    "petb200_avg_reverse_fwd": [_P, _P, _P, _I64, _I, _P, _P],
    "petb200_avg_reverse_bwd": [_P, _P, _I64, _I, _P, _P, _P],
    "petb200_rms_rstd": [_P, _I64, _I, _P, _P],
    "petb200_rms_bwd": [_P, _P, _P, _P, _I64, _I, _P, _P],
    "petb200_layer_norm_fwd": [_P, _P, _P, _I64, _I, _P, _P, _P, _P],
    "petb200_layer_norm_bwd": [_P, _P, _P, _P, _P, _P, _I64, _I, _P, _P],
    "petb200_rms_norm_fwd": [_P, _P, _I64, _I, _P, _P, _P],
    "petb200_rms_norm_bwd": [_P, _P, _P, _P, _P, _I64, _I, _P, _P],
    "petb200_attention_fwd": [_P, _P, _P, _I64, _I64, _I, _I, _F, _I, _I, _P, _P, _P],
    "petb200_attention_bwd": [_P, _P, _P, _P, _P, _P, _I64, _I64, _I, _I, _F, _I, _I, _P, _P, _P, _P],
    "petb200_combine_ln_fwd": [_P, _P, _P, _P, _I64, _I, _P, _P, _P, _P],
    "petb200_combine_ln_bwd": [_P, _P, _P, _P, _P, _P, _I64, _I, _P, _P],
    "petb200_combine_scatter_bwd": [_P, _P, _P, _I64, _I, _P, _P],
    "petb200_combine_image_bytes": [_I, _I],
    "petb200_combine_pack": [_P, _P, _I, _P, _P, _P],
    "petb200_combine_fwd": [_P, _I64, _P, _P, _P, _P, _P, _I64, _I, _P, _I64, _P, _P, _P],
    "petb200_combine_bwd": [_P, _I64, _P, _P, _I64, _P, _P, _P, _P, _P, _I64, _I, _P, _P],
    "petb200_readout_fwd": [_P, _P, _P, _P, _P, _P, _P, _P, _I64, _I64, _I, _I, _P, _P, _P],
    "petb200_readout_bwd": [_P, _P, _P, _P, _P, _P, _P, _P, _I64, _I64, _I, _I, _P, _P, _P, _P],
    "petb200_chain_image_bytes": [_I],
    "petb200_chain_pack": [_P, _P, _I, _P, _P, _P],
    "petb200_edge_head_fwd": [_P, _I64, _P, _P, _P, _P, _F, _I64, _I, _P, _P, _P, _P],
    "petb200_edge_head_bwd": [_P, _P, _P, _P, _P, _P, _P, _P, _I64, _I, _P, _I64, _P, _P],
    "petb200_compress_fwd": [_P, _I64, _P, _P, _P, _P, _P, _P, _P, _P, _I64, _I, _P, _P, _I64, _P],
    "petb200_compress_bwd": [_P, _I64, _P, _P, _P, _I64, _I, _P, _I64, _I, _P, _P, _P],
    "petb200_gnn_saved_bytes": [_P, _P],
    "petb200_gnn_scratch_bytes": [_P, _P],
    "petb200_gnn_fwd": [_P, _P, _P, _P, _P, _P, _P, _P, _P, _I64, _P, _P, _P, _SZ, _P, _SZ, _P],
    "petb200_gnn_bwd": [_P, _P, _P, _P, _P, _P, _P, _P, _I64, _P, _P, _P, _P, _P, _SZ, _P],
    "petb200_sum_over_atoms": [_P, _P, _I64, _I, _P, _P],
    "petb200_last_error": [],
    "petb200_version": [],
}
_RESTYPE = {"petb200_csr_build_workspace": _SZ, "petb200_mlp_image_bytes": _SZ, "petb200_combine_image_bytes": _SZ, "petb200_chain_image_bytes": _SZ, "petb200_gnn_saved_bytes": _SZ, "petb200_gnn_scratch_bytes": _SZ, "petb200_norm_linear_image_bytes": _SZ, "petb200_last_error": ctypes.c_char_p,
            "petb200_nl_num_bins": _I64, "petb200_nl_workspace": _SZ}


# ---- structs of the stage-level schedule (include/petb200.h, "stage-level schedule")
class Mat(ctypes.Structure):
    _fields_ = [("w", _P), ("ld", _I64)]


class TLWeights(ctypes.Structure):
    _fields_ = [("qkv_image", _P), ("b_qkv", _P), ("w_qkv_t", Mat), ("w_o", Mat), ("w_o_t", Mat), ("b_o", _P),
                ("mlp_image_fwd", _P), ("mlp_image_bwd", _P), ("b_in", _P), ("b_out", _P), ("d_ff", _I),
                ("w_con", Mat), ("w_con_t", Mat), ("b_con", _P), ("w_exp", Mat), ("w_exp_t", Mat), ("b_exp", _P),
                ("wc_in", Mat), ("wc_in_t", Mat), ("bc_in", _P), ("wc_out", Mat), ("wc_out_t", Mat), ("bc_out", _P)]


class GNNWeights(ctypes.Structure):
    _fields_ = [("w1m", Mat), ("w1m_t", Mat), ("b_fold", _P), ("geo_fold", _P), ("nbr_fold", _P), ("w2", Mat),
                ("w2_t", Mat), ("b2", _P), ("compress_image_fwd", _P), ("compress_image_bwd", _P),
                ("n_tl", _I), ("tl", ctypes.POINTER(TLWeights))]


class Dims(ctypes.Structure):
    _fields_ = [("n_atoms", _I64), ("n_edges", _I64), ("n_ghost", _I64), ("d", _I), ("d_node", _I),
                ("num_heads", _I), ("max_row", _I), ("precision", _I), ("scale", _F)]


def library_path() -> str:
    return _LIB_PATH


def build(verbose: bool = False) -> str:
    """Compile ``libpetb200.so`` in-tree for sm_100a (``make`` in ``csrc/``)."""
    proc = subprocess.run(["make", "-C", _CSRC, "-j8"], capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("building libpetb200.so failed:\n" + proc.stdout + proc.stderr)
    if verbose:
        print(proc.stdout)
    return _LIB_PATH


def load() -> ctypes.CDLL:
    """Load the library (once).  Raises ``RuntimeError`` if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is None:
            if not os.path.isfile(_LIB_PATH):
                raise RuntimeError(
                    f"{_LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; "
                    "g.build()'` (or `make -C metatrain_b200/csrc`). There is no fallback path."
                )
            lib = ctypes.CDLL(_LIB_PATH)
            for name, argtypes in _SIGNATURES.items():
                fn = getattr(lib, name)
                fn.argtypes = argtypes
                fn.restype = _RESTYPE.get(name, ctypes.c_int)
            _lib = lib
    return _lib


def ptr(t):
    """Device pointer of a tensor (``None`` -> NULL)."""
    return None if t is None else t.data_ptr()


def stream_ptr() -> int:
    """Raw handle of torch's current stream on the current device (the fast private accessor when
    this torch has it: torch.cuda.current_stream() costs ~15 us per call, ~2 ms per step)."""
    try:
        return torch._C._cuda_getCurrentRawStream(torch.cuda.current_device())
    except AttributeError:  # pragma: no cover - older / newer torch without the private accessor
        return torch.cuda.current_stream().cuda_stream


# kernels launched by each entry point (for bench.py's ``gpu_launches`` claim)
_KERNELS_PER_CALL = {"nl_count": 6, "attention_bwd": 2, "edges_bwd": 3, "edges_bwd_rc": 3, "adaptive_cutoff_bwd": 2, "adaptive_grid_bwd": 2, "csr_build": 6, "readout_bwd": 2,
                     "force_scatter": 2, "mlp_pack": 2}
launch_count = 0
#: kernels enqueued by one petb200_gnn_fwd / _bwd call: (fixed part, per attention layer)
GNN_KERNELS = {"gnn_fwd": (1, 9), "gnn_bwd": (1, 9)}
#: optional profiler hook: ``hook(name, args) -> context manager`` wrapped around a call
profile_hook = None


_ENTRY = {}  # name -> bound ctypes function (saves a string build + getattr per launch)


def call(name: str, *args) -> None:
    """Invoke ``petb200_<name>`` on torch's current stream; raise on a non-zero status."""
    global launch_count
    fn = _ENTRY.get(name)
    if fn is None:
        fn = _ENTRY[name] = getattr(load(), "petb200_" + name)
    n_kernels = _KERNELS_PER_CALL.get(name, 1)
    if name == "attention_bwd" and args[12] != PREC_FP32 and args[11] + 1 <= 64:
        n_kernels = 1  # single tensor-core kernel (attention_tc.cu) instead of dQ + dKdV
    launch_count += n_kernels
    if profile_hook is not None:
        with profile_hook(name, args):
            status = fn(*args, stream_ptr())
    else:
        status = fn(*args, stream_ptr())
    if status != 0:
        msg = load().petb200_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"petb200_{name} failed ({status}): {msg}")
