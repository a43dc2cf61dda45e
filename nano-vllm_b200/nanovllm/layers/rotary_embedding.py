"""NeoX rotary embedding table (reference nanovllm/layers/rotary_embedding.py:17-59).

The rotation itself is fused with q/k-norm and the KV scatter in ``ops.qknorm_rope_store``; this
module owns the fp32 ``[max_pos, head_dim] = cat(cos, sin)`` table the kernel gathers from.
"""
from __future__ import annotations

from functools import lru_cache

import torch


def build_cos_sin(head_size: int, max_position: int, base: float, device=None) -> torch.Tensor:
    exponent = torch.arange(0, head_size, 2, dtype=torch.float) / head_size
    inv_freq = 1.0 / (base ** exponent)
    angles = torch.outer(torch.arange(max_position, dtype=torch.float), inv_freq)
    table = torch.cat((angles.cos(), angles.sin()), dim=-1).contiguous()
    return table.to(device) if device is not None else table


@lru_cache(4)
def get_rope(head_size: int, rotary_dim: int, max_position: int, base: float, device: str = "cuda") -> torch.Tensor:
    assert rotary_dim == head_size
    return build_cos_sin(head_size, max_position, base, device)
