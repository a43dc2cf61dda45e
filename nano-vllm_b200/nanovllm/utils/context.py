"""Per-step attention metadata, handed to the operator out of band (reference nanovllm/utils/context.py:5-27).

The reference passes these through a process-global record rather than through function
arguments; ``Attention.forward`` keeps that contract, so the same names live here.
"""
from __future__ import annotations

from dataclasses import dataclass

import torch


@dataclass(slots=True)
class Context:
    is_prefill: bool = False
    cu_seqlens_q: torch.Tensor | None = None      # [S+1] int32
    cu_seqlens_k: torch.Tensor | None = None      # [S+1] int32
    max_seqlen_q: int = 0
    max_seqlen_k: int = 0
    slot_mapping: torch.Tensor | None = None      # [T] or [B] int32, -1 = skip
    context_lens: torch.Tensor | None = None      # [B] int32 (decode)
    block_tables: torch.Tensor | None = None      # [S or B, W] int32, -1 padded


_current = Context()


def get_context() -> Context:
    return _current


def set_context(is_prefill, cu_seqlens_q=None, cu_seqlens_k=None, max_seqlen_q=0, max_seqlen_k=0,
                slot_mapping=None, context_lens=None, block_tables=None) -> None:
    global _current
    _current = Context(is_prefill, cu_seqlens_q, cu_seqlens_k, max_seqlen_q, max_seqlen_k,
                       slot_mapping, context_lens, block_tables)


def reset_context() -> None:
    global _current
    _current = Context()
