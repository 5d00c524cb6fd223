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
