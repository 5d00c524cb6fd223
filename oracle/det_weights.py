"""TEST INFRASTRUCTURE ONLY.  Deterministic, RNG-free parameter values for end-to-end parity tests.

The end-to-end golden vectors (tests/golden/e2e_eemflow_cdc.npz) were produced by the REAL reference
model whose parameters were set with this function; the GPU test sets the drop-in model's parameters
the same way (the two models share parameter names, shapes and order), so no 6 MB state dict has to
be committed and no RNG stream has to be reproduced.
"""
import torch


def set_deterministic_weights(model: torch.nn.Module) -> None:
    with torch.no_grad():
        for i, (_, p) in enumerate(model.named_parameters()):
            idx = torch.arange(p.numel(), dtype=torch.float64)
            if p.dim() > 1:
                fan_in = p[0].numel()
                v = torch.sin(0.37 * idx + i) * (1.6 / fan_in ** 0.5)   # ~ Kaiming-normal scale
            else:
                v = 0.05 * torch.cos(0.11 * idx + i)
            p.copy_(v.view_as(p).float())


def _hash_uniform(n: int, stream: int) -> torch.Tensor:
    """n float64 values in [-1, 1): splitmix64 finaliser of (index, stream); exact on every machine."""
    import numpy as np
    with np.errstate(over="ignore"):
        x = np.arange(n, dtype=np.uint64) + np.uint64(stream + 1) * np.uint64(0x9E3779B97F4A7C15)
        x ^= x >> np.uint64(30)
        x *= np.uint64(0xBF58476D1CE4E5B9)
        x ^= x >> np.uint64(27)
        x *= np.uint64(0x94D049BB133111EB)
        x ^= x >> np.uint64(31)
    return torch.from_numpy((x >> np.uint64(11)).astype(np.float64) * (2.0 ** -52) - 1.0)


def set_hashed_weights(model: torch.nn.Module, weight_gain: float = 1.0, bias_scale: float = 0.02) -> None:
    """RNG-free parameters with white-noise statistics: conv/linear weights uniform with std
    weight_gain / sqrt(fan_in), biases uniform in +-bias_scale, normalisation weights 1 + that.  Used by the
    ERAFT golden, where the sinusoidal pattern above cancels over smooth inputs and leaves only the biases."""
    norm_w = {id(m.weight) for m in model.modules()
              if isinstance(m, (torch.nn.BatchNorm2d, torch.nn.GroupNorm, torch.nn.InstanceNorm2d)) and m.weight is not None}
    with torch.no_grad():
        for i, (_, p) in enumerate(model.named_parameters()):
            u = _hash_uniform(p.numel(), i)
            if p.dim() > 1:
                v = u * (3.0 ** 0.5) * weight_gain / p[0].numel() ** 0.5
            else:
                v = u * bias_scale + (1.0 if id(p) in norm_w else 0.0)
            p.copy_(v.view_as(p).float())
