"""Engine configuration (reference nanovllm/config.py:6-25): same fields, same defaults.

Differences, both widenings: ``kvcache_block_size`` may be any power of two in [16, 256]
(the reference's ``% 256`` assertion exists only for flash-attn's paged path), and
``hf_config`` may be supplied pre-built (tests) instead of being read from ``model``.
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Any


@dataclass(slots=True)
class Config:
    model: str
    max_num_batched_tokens: int = 16384
    max_num_seqs: int = 512
    max_model_len: int = 4096
    gpu_memory_utilization: float = 0.9
    tensor_parallel_size: int = 1
    enforce_eager: bool = False
    hf_config: Any = None
    eos: int = -1
    kvcache_block_size: int = 256
    num_kvcache_blocks: int = -1

    def __post_init__(self):
        bs = self.kvcache_block_size
        if bs < 16 or bs > 256 or bs & (bs - 1):
            raise ValueError("kvcache_block_size must be a power of two in [16, 256]")
        if not 1 <= self.tensor_parallel_size <= 8:
            raise ValueError("tensor_parallel_size must be in 1..8")
        if self.hf_config is None:
            if not os.path.isdir(self.model):
                raise FileNotFoundError(f"model directory not found: {self.model}")
            from transformers import AutoConfig
            self.hf_config = AutoConfig.from_pretrained(self.model)
        self.max_model_len = min(self.max_model_len, self.hf_config.max_position_embeddings)
