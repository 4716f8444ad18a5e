"""petb200 — B200-native forward/backward engine for metatrain's PET hot path."""
from .backend import B200PETBackend  # noqa: F401
from .evaluate import evaluate, sum_over_atoms  # noqa: F401
from .eval_loop import eval_targets  # noqa: F401
from .md import GraphedEvaluator  # noqa: F401
