"""Sampler (reference nanovllm/layers/sampler.py:5-12) on the CUDA kernel, with a greedy branch."""
from __future__ import annotations

import torch
from torch import nn

from .. import ops


class Sampler(nn.Module):
    def __init__(self, seed: int = 0):
        super().__init__()
        self.seed = seed
        self.step = 0

    def forward(self, logits: torch.Tensor, temperatures: torch.Tensor) -> torch.Tensor:
        """logits [rows, vocab] (bf16 or fp32), temperatures [rows] fp32; 0 means argmax."""
        self.step += 1
        return ops.sample(logits, temperatures, self.seed, self.step)
