"""LoRA adapters on the backend's Linear layers — mirror of ``inject_lora_layers`` / ``LoRALinear``
(``src/metatrain/pet/modules/finetuning.py:322-378``), so that LoRA-finetuned PET checkpoints
(keys ``...<name>.linear.weight``, ``...<name>.lora_A.weight``, ``...<name>.lora_B.weight``) load
and evaluate.  The CUDA engine never runs the adapter as a separate low-rank product: when the
weights are packed (``engine.PackedWeights``) every adapted layer is merged into
``W + (alpha / rank) * B A``; any change of the adapter parameters re-packs.
"""
from typing import Sequence

from torch import nn


class LoRALinear(nn.Module):
    """``y = linear(x) + (alpha / rank) * lora_B(lora_A(x))`` (finetuning.py:370-378); same
    sub-module names as the reference so state-dict keys agree."""

    def __init__(self, linear_layer: nn.Linear, rank: int = 4, alpha: float = 1.0):
        super().__init__()
        self.linear = linear_layer
        self.lora_A = nn.Linear(linear_layer.in_features, rank, bias=False)
        self.lora_B = nn.Linear(rank, linear_layer.out_features, bias=False)
        self.scaling = alpha / rank

    @property
    def in_features(self) -> int:
        return self.linear.in_features

    @property
    def out_features(self) -> int:
        return self.linear.out_features


def inject_lora_layers(backend: nn.Module, target_modules: Sequence[str] = ("input_linear", "output_linear"),
                       rank: int = 4, alpha: float = 1.0) -> nn.Module:
    """Wrap every ``nn.Linear`` attribute whose name is in ``target_modules`` (default: the
    attention projections) — same traversal order as the reference (finetuning.py:346-353), hence
    the same RNG order for the adapter initialisation."""
    for _, module in list(backend.named_modules()):
        for attr in target_modules:
            child = getattr(module, attr, None)
            if isinstance(child, nn.Linear):
                ref = child.weight
                setattr(module, attr, LoRALinear(child, rank=rank, alpha=alpha).to(device=ref.device, dtype=ref.dtype))
    return backend
