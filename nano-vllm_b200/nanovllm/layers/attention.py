"""The attention operator seam (reference nanovllm/layers/attention.py:33-75), backed by libb200attn.

Same module contract as the reference: ``Attention(num_heads, head_dim, scale, num_kv_heads)``,
attributes ``k_cache`` / ``v_cache`` (discovered with ``hasattr`` by the runner,
engine/model_runner.py:117-121), ``forward(q, k, v) -> o`` reading the process-global Context.

What differs is what the cache tensors look like: a layer's ``k_cache`` is the physical
page array ``[num_blocks, num_kv_heads, block_size, head_dim]`` (head-major pages), bound
to the library once by the runner; block ids and slot numbers mean exactly what they mean in the
reference.
"""
from __future__ import annotations

import torch
from torch import nn

from .. import ops
from ..utils.context import get_context


def store_kvcache(key: torch.Tensor, value: torch.Tensor, k_cache: torch.Tensor, v_cache: torch.Tensor,
                  slot_mapping: torch.Tensor, layer_id: int = 0):
    """Scatter new K/V rows into their slots (reference layers/attention.py:33-40).

    ``k_cache`` / ``v_cache`` must be the layer views of the cache bound with ``ops.bind_kv_cache``;
    they are accepted for signature compatibility and checked, the library addresses the bound cache.
    """
    n, num_heads, head_dim = key.shape
    assert key.stride(-1) == 1 and value.stride(-1) == 1
    assert key.stride(1) == head_dim and value.stride(1) == head_dim
    assert slot_mapping.numel() == n
    assert k_cache.dim() == 4 and k_cache.shape[1] == num_heads and k_cache.shape[3] == head_dim
    ops.store_kv(layer_id, key, value, slot_mapping)


class Attention(nn.Module):
    def __init__(self, num_heads: int, head_dim: int, scale: float, num_kv_heads: int):
        super().__init__()
        self.num_heads = num_heads
        self.head_dim = head_dim
        self.scale = scale
        self.num_kv_heads = num_kv_heads
        self.k_cache = self.v_cache = torch.tensor([])
        self.layer_id = 0                       # set by the runner when it binds the cache

    def forward(self, q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, kv_stored: bool = False) -> torch.Tensor:
        """q [T, Hq, D], k/v [T, Hkv, D] (views of the qkv GEMM output are fine).

        ``kv_stored=True`` tells the operator that the fused q/k-norm+RoPE kernel already scattered
        this step's K/V (the model's fast path); the default does the scatter itself, like the reference.
        """
        ctx = get_context()
        have_cache = self.k_cache.numel() > 0 and self.v_cache.numel() > 0
        if have_cache and not kv_stored:
            store_kvcache(k, v, self.k_cache, self.v_cache, ctx.slot_mapping, self.layer_id)
        if ctx.is_prefill:
            return ops.paged_prefill(self.layer_id, q, k, v, ctx.cu_seqlens_q, ctx.cu_seqlens_k,
                                     ctx.max_seqlen_q, ctx.max_seqlen_k, self.scale,
                                     block_tables=ctx.block_tables, num_kv_heads=self.num_kv_heads)
        o = ops.paged_decode(self.layer_id, q, ctx.block_tables, ctx.context_lens, self.scale)
        return o.unsqueeze(1)                   # [B, 1, Hq, D] like flash_attn_with_kvcache
