"""Activation registry.  Names follow ppsci/arch/activation.py:139-154; an activation is usable
by the engine iff the jet kernels implement its Taylor coefficients (csrc/jet_math.h)."""
from __future__ import annotations

from ..engine.binding import ACT_IDS

# names the reference registers (activation.py:139-154)
REFERENCE_ACTS = ("elu", "relu", "selu", "gelu", "leaky_relu", "sigmoid", "silu", "sin", "cos", "swish",
                  "tanh", "identity", "siren", "stan")


_WARNED = {"swish": False}


def get_activation(act_name: str) -> str:
    """Validate ``act_name`` and return its canonical lower-case name.

    Raises ValueError for unknown names exactly like the reference (activation.py:166-167) and
    NotImplementedError for reference activations that have no jet kernel yet."""
    name = act_name.lower()
    if name not in REFERENCE_ACTS:
        raise ValueError(f"act_name({act_name}) not found in act_func_dict")
    if name not in ACT_IDS:
        raise NotImplementedError(
            f"activation '{act_name}' has no Taylor-jet kernel in this engine yet "
            f"(supported: {sorted(ACT_IDS)})")
    return name


def warn_fixed_swish():
    """activation.py:49-58, 149, 169-171: the reference instantiates one Swish per layer, each with a TRAINABLE scalar
    beta (x * sigmoid(beta x), beta = 1 at start).  ``arch.MLP`` trains it (PPSCI_ACT_SWISH_B); callers that cannot
    (the DeepONet sub-networks) keep beta fixed at 1 — identical to 'silu' — and say so once."""
    if not _WARNED["swish"]:
        _WARNED["swish"] = True
        import warnings

        warnings.warn("activation 'swish': the reference trains one scalar beta per layer (Swish, beta = 1 at start); "
                      "here beta stays fixed at 1 (identical to 'silu')", stacklevel=3)
