"""Checkpoint loading (reference nanovllm/utils/loader.py:12-28): every ``*.safetensors`` tensor in the
model directory is handed to the model, which slices its own tensor-parallel shard.

Unlike the reference, a directory without weights is an error unless ``allow_random`` is set, in
which case the model is filled from a seeded generator (the reference would silently run on
``torch.empty`` garbage).
"""
from __future__ import annotations

import os
from glob import glob

import torch


def load_model(model, path: str, allow_random: bool = False, seed: int = 0) -> int:
    files = sorted(glob(os.path.join(path, "*.safetensors")))
    if not files:
        if not allow_random:
            raise FileNotFoundError(f"no *.safetensors under {path}")
        return _random_fill(model, seed)
    from safetensors import safe_open
    n = 0
    for file in files:
        with safe_open(file, "pt", "cpu") as f:
            for name in f.keys():
                model.load_hf_tensor(name, f.get_tensor(name))
                n += 1
    return n


def _random_fill(model, seed: int, init_std: float = 0.02) -> int:
    """Seeded random weights for a directory that has none (benchmarks at 8B / 32B dimensions: no checkpoint exists
    offline).  On a CUDA model every tensor is drawn ON THE DEVICE from a generator keyed by (seed, parameter name) at
    its FULL shape and handed to ``load_hf_tensor``, which slices this rank's shard -- so all tensor-parallel ranks hold
    shards of one well-defined model without 64 GB of host memory or a broadcast.  On a CPU model the values are
    ``synthetic.random_weights`` (what the tests compare against)."""
    from .synthetic import random_weights, weight_shapes
    c = model.cfg
    dims = dict(hidden_size=c.hidden_size, num_hidden_layers=c.num_hidden_layers,
                num_attention_heads=c.num_attention_heads, num_key_value_heads=c.num_key_value_heads,
                head_dim=model.head_dim, intermediate_size=c.intermediate_size, vocab_size=c.vocab_size,
                tie_word_embeddings=model.tie)
    if model.device.type != "cuda":
        ws = random_weights(dims, seed)
        for name, w in ws.items():
            model.load_hf_tensor(name, w)
        return len(ws)
    import zlib
    gen = torch.Generator(device=model.device)
    shapes = weight_shapes(dims)
    for name, shape in shapes.items():
        gen.manual_seed((seed << 32) ^ zlib.crc32(name.encode()))
        w = torch.randn(shape, generator=gen, device=model.device, dtype=torch.float32)
        w = (1.0 + 0.1 * w) if len(shape) == 1 else init_std * w
        model.load_hf_tensor(name, w.to(torch.bfloat16))
        del w
    return len(shapes)
