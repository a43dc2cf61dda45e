"""Shared helpers for the GPU parity tests (whole-model comparisons against the CPU oracle)."""
from __future__ import annotations

import json
import os
from types import SimpleNamespace

import torch
import torch.nn.functional as F

from oracle.qwen3_ref import Qwen3Ref, RefDims

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def record(name: str, payload: dict):
    """Append measured parity numbers to gpurun_out/parity.jsonl so they can be read back after a GPU run."""
    d = os.path.join(ROOT, "gpurun_out")
    os.makedirs(d, exist_ok=True)
    with open(os.path.join(d, "parity.jsonl"), "a") as f:
        f.write(json.dumps(dict(test=name, **payload)) + "\n")


def oracle_full_logits(model: Qwen3Ref, tokens: list[int]) -> torch.Tensor:
    """fp32 logits at every position of one sequence (packed causal prefill, no cache)."""
    n = len(tokens)
    ctx = SimpleNamespace(is_prefill=True, cu_seqlens_q=torch.tensor([0, n], dtype=torch.int32),
                          cu_seqlens_k=torch.tensor([0, n], dtype=torch.int32), max_seqlen_q=n, max_seqlen_k=n,
                          slot_mapping=None, context_lens=None, block_tables=None)
    h = model.forward(torch.tensor(tokens, dtype=torch.int64), torch.arange(n, dtype=torch.int64), ctx, None)
    return F.linear(h, model.head).float()


def check_greedy_against_oracle(model: Qwen3Ref, prompt: list[int], completion: list[int], ulps: float = 6.0):
    """Teacher-forced: every generated token must be the oracle's argmax, or lose to it by less than
    `ulps` bf16 roundings of the logit scale (two bf16 pipelines cannot agree closer than that).
    Returns (#tokens, #tokens that differ from the oracle argmax, worst margin / tolerance)."""
    seq = prompt + completion
    lg = oracle_full_logits(model, seq[:-1])
    tol = ulps * 2 ** -8 * lg.abs().max().item()
    diff, worst = 0, 0.0
    for i, tok in enumerate(completion):
        row = lg[len(prompt) - 1 + i]
        margin = (row.max() - row[tok]).item()
        assert margin <= tol, f"token {i}: chose {tok} but oracle prefers {int(row.argmax())} by {margin} (tol {tol})"
        diff += int(row.argmax()) != tok
        worst = max(worst, margin / tol)
    return len(completion), diff, worst


def make_oracle(preset_dims: dict, weights: dict, rounding="fused", max_pos=4096) -> Qwen3Ref:
    from nanovllm.utils.synthetic import hf_config_dict
    return Qwen3Ref(RefDims.from_json(hf_config_dict(preset_dims)), weights, rounding=rounding, max_pos=max_pos)
