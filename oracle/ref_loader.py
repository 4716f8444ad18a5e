"""TEST INFRASTRUCTURE ONLY — loads the *unmodified* reference PET backend.

Works only where ``/root/reference`` is mounted (the build container); it cannot
travel to the GPU box.  Used by ``tests/golden/make_golden.py`` to generate the
committed golden vectors and by CPU tests (skipped when the reference is absent)
to validate ``oracle/pet_oracle.py`` against the real thing.

Recipe = SURVEY.md Appendix C: the backend files
``src/metatrain/pet/modules/{utilities,nef,adaptive_cutoff,conditioning,
transformer,structures,backend}.py`` only import metatensor/metatomic for type
annotations and ``concatenate_structures``; empty stub modules satisfy them.
Nothing from the reference is copied: the files are executed where they lie.
"""
import importlib.util
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("PETB200_REFERENCE_ROOT", "/root/reference")
_PET_DIR = os.path.join(REFERENCE_ROOT, "src", "metatrain", "pet")

# pet/documentation.py:159-259 defaults (ModelHypers)
DEFAULT_HYPERS = dict(
    cutoff=4.5,
    num_neighbors_adaptive=None,
    adaptive_cutoff_method="solver",
    cutoff_function="Bump",
    cutoff_width=0.5,
    cutoff_width_adaptive=1.0,
    d_pet=128,
    d_head=128,
    d_node=256,
    d_feedforward=256,
    num_heads=8,
    num_attention_layers=2,
    num_gnn_layers=2,
    normalization="RMSNorm",
    activation="SwiGLU",
    attention_temperature=1.0,
    transformer_type="PreLN",
    featurizer_type="feedforward",
    zbl=False,
    long_range=dict(enable=False),
    system_conditioning=False,
    max_charge=10,
    max_spin_multiplicity=10,
)


def reference_available() -> bool:
    return os.path.isfile(os.path.join(_PET_DIR, "modules", "backend.py"))


class _Dummy:
    def __init__(self, *a, **k):
        pass


def _stub(name, **attrs):
    if name in sys.modules and not getattr(sys.modules[name], "_petb200_stub", False):
        return  # a real package is installed: leave it alone
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    m.__path__ = []
    m._petb200_stub = True
    sys.modules[name] = m


_LOADED = None


def load_reference_backend_class():
    """Return the reference ``PETBackend`` class (backend.py:12), unmodified."""
    global _LOADED
    if _LOADED is not None:
        return _LOADED
    if not reference_available():
        raise RuntimeError(f"reference not mounted at {REFERENCE_ROOT}")
    _stub("metatensor")
    _stub("metatensor.torch", Labels=_Dummy, TensorBlock=_Dummy, TensorMap=_Dummy)
    _stub("metatomic")
    _stub("metatomic.torch", NeighborListOptions=_Dummy, System=_Dummy)
    _stub("metatrain")
    _stub("metatrain.pet")
    _stub("metatrain.pet.modules")
    _stub("metatrain.pet.documentation", ModelHypers=dict)
    for n in [
        "utilities",
        "nef",
        "adaptive_cutoff",
        "conditioning",
        "transformer",
        "structures",
        "backend",
    ]:
        full = f"metatrain.pet.modules.{n}"
        spec = importlib.util.spec_from_file_location(
            full, os.path.join(_PET_DIR, "modules", f"{n}.py")
        )
        mod = importlib.util.module_from_spec(spec)
        sys.modules[full] = mod
        spec.loader.exec_module(mod)
    _LOADED = sys.modules["metatrain.pet.modules.backend"].PETBackend
    return _LOADED


def load_reference_finetuning():
    """The reference's ``pet/modules/finetuning.py`` module (LoRA injection), unmodified."""
    load_reference_backend_class()
    full = "metatrain.pet.modules.finetuning"
    if full not in sys.modules:
        _stub("metatrain.utils")
        _stub("metatrain.utils.data")
        _stub("metatrain.utils.data.target_info", TargetInfo=_Dummy)
        spec = importlib.util.spec_from_file_location(full, os.path.join(_PET_DIR, "modules", "finetuning.py"))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[full] = mod
        spec.loader.exec_module(mod)
    return sys.modules[full]


def build_reference_backend(atomic_types, target="energy", hypers=None, seed=0,
                            dtype=None, out_shape=(1,)):
    """Seeded construction in the RNG order of pet/model.py:115,145-147."""
    import random

    import numpy as np
    import torch

    cls = load_reference_backend_class()
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    h = dict(DEFAULT_HYPERS)
    if hypers:
        h.update({k: v for k, v in hypers.items() if not k.startswith("_")})
    be = cls(h, list(atomic_types))
    # key naming of pet/model.py:1045-1051: <target>_<keyname>_<keyvalue>; a scalar
    # target has the single key "_" = 0  -> "<target>___0"
    be.add_output(target, {f"{target}___0": list(out_shape)})
    gate_seed = (hypers or {}).get("_gate_seed")
    if gate_seed is not None:  # test-only pseudo hyper: activate the zero-initialised conditioning gate
        torch.manual_seed(gate_seed)
        gate = be.system_conditioning.project[2]
        with torch.no_grad():
            gate.weight.normal_(0.0, 0.05)
            gate.bias.normal_(0.0, 0.05)
    lora = (hypers or {}).get("_lora")
    if lora:  # test-only pseudo hyper: LoRA adapters injected by the reference's own function
        torch.manual_seed(lora["seed"])
        load_reference_finetuning().inject_lora_layers(
            be, tuple(lora["target_modules"]), rank=lora["rank"], alpha=lora["alpha"])
    if dtype is not None:
        be = be.to(dtype)
    return be
