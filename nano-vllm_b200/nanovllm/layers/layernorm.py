"""RMSNorm module (reference nanovllm/layers/layernorm.py:5-50) on the CUDA kernels."""
from __future__ import annotations

import torch
from torch import nn

from .. import ops


class RMSNorm(nn.Module):
    def __init__(self, hidden_size: int, eps: float = 1e-6):
        super().__init__()
        self.eps = eps
        self.weight = nn.Parameter(torch.ones(hidden_size), requires_grad=False)

    def forward(self, x: torch.Tensor, residual: torch.Tensor | None = None):
        """``norm(x)``, or ``(norm(x + residual), bf16(x + residual))`` when a residual is given.

        The fused form updates ``residual`` in place and returns it (the reference returns a new tensor
        with the same values)."""
        if residual is None:
            return ops.rmsnorm(x, self.weight, self.eps)
        return ops.add_rmsnorm(x, residual, self.weight, self.eps)
