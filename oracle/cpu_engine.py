"""The reference's serving loop on host CPU cores.  TEST INFRASTRUCTURE / REPORTED BASELINE ONLY.

The reference cannot run without a GPU (NCCL, torch.cuda and flash-attn are hard-wired,
engine/model_runner.py:26-30, layers/attention.py:3-6), so "the reference on the box's CPU" is this
port: the reference's step loop (engine/llm_engine.py:49-90: schedule -> run -> postprocess) over
  * the scheduler / block manager semantics pinned bit-exactly to the reference's traces
    (tests/test_bookkeeping_golden.py),
  * oracle.qwen3_ref.Qwen3Ref in "eager" rounding mode, pinned bit-exactly to the reference's own
    nn.Modules on CPU (tests/test_oracle_golden.py),
  * oracle.paged_attention_ref for the attention core (flash-attn has no CPU build).
bench.py times it (cpu_baseline / --impl reference); nothing in the product imports it.
"""
from __future__ import annotations

import itertools
import os
import sys
import time
from types import SimpleNamespace

import numpy as np
import torch

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_PKG = os.path.join(_ROOT, "nano-vllm_b200")
if _PKG not in sys.path:
    sys.path.insert(0, _PKG)

from .qwen3_ref import Qwen3Ref, RefDims, alloc_logical_kv  # noqa: E402


class CpuEngine:
    def __init__(self, dims_json: dict, weights: dict, block_size: int = 256, num_blocks: int = 64,
                 max_num_seqs: int = 512, max_num_batched_tokens: int = 16384, eos: int = -1, threads: int | None = None):
        from nanovllm.engine.scheduler import Scheduler
        from nanovllm.engine.sequence import Sequence
        if threads:
            torch.set_num_threads(threads)
        self.threads = torch.get_num_threads()
        self.dims = RefDims.from_json(dims_json)
        self.model = Qwen3Ref(self.dims, weights, rounding="eager", max_pos=8192)
        self.block_size = block_size
        Sequence.block_size = block_size
        Sequence.counter = itertools.count()
        cfg = SimpleNamespace(max_num_seqs=max_num_seqs, max_num_batched_tokens=max_num_batched_tokens, eos=eos,
                              kvcache_block_size=block_size, num_kvcache_blocks=num_blocks)
        self.scheduler = Scheduler(cfg)
        self.kv = alloc_logical_kv(self.dims, num_blocks, block_size)
        self._Sequence = Sequence

    def _meta(self, seqs, is_prefill):
        from nanovllm.engine.model_runner import ModelRunner
        stub = SimpleNamespace(block_size=self.block_size)
        stub.prepare_block_tables = lambda s: ModelRunner.prepare_block_tables(stub, s)
        a = ModelRunner.prefill_arrays(stub, seqs) if is_prefill else ModelRunner.decode_arrays(stub, seqs)
        t = lambda x: None if x is None else torch.from_numpy(np.ascontiguousarray(x))
        ctx = SimpleNamespace(is_prefill=is_prefill, cu_seqlens_q=t(a.get("cu_seqlens_q")), cu_seqlens_k=t(a.get("cu_seqlens_k")),
                              max_seqlen_q=a.get("max_seqlen_q", 0), max_seqlen_k=a.get("max_seqlen_k", 0),
                              slot_mapping=t(a["slot_mapping"]), context_lens=t(a.get("context_lens")),
                              block_tables=t(a.get("block_tables")))
        return t(a["input_ids"]), t(a["positions"]), ctx

    @torch.inference_mode()
    def generate(self, prompts, sampling_params) -> tuple[list[list[int]], float]:
        """Greedy (temperature ignored: the CPU port is a throughput baseline).  Returns (completions, seconds)."""
        for p, sp in zip(prompts, sampling_params):
            self.scheduler.add(self._Sequence(p, sp))
        done = {}
        t0 = time.perf_counter()
        while not self.scheduler.is_finished():
            seqs, is_prefill = self.scheduler.schedule()
            ids, pos, ctx = self._meta(seqs, is_prefill)
            hidden = self.model.forward(ids, pos, ctx, self.kv)
            tokens = self.model.greedy(self.model.logits(hidden, ctx)).tolist()
            self.scheduler.postprocess(seqs, tokens, is_prefill)
            for s in seqs:
                if s.is_finished:
                    done[s.seq_id] = s.completion_token_ids
        dt = time.perf_counter() - t0
        return [done[k] for k in sorted(done)], dt

    @torch.inference_mode()
    def generate_bounded(self, prompts, sampling_params, budget_s: float):
        """Like generate() but stops after `budget_s` seconds: returns (output tokens produced, seconds, engine steps)."""
        for p, sp in zip(prompts, sampling_params):
            self.scheduler.add(self._Sequence(p, sp))
        produced = steps = 0
        t0 = time.perf_counter()
        while not self.scheduler.is_finished() and time.perf_counter() - t0 < budget_s:
            seqs, is_prefill = self.scheduler.schedule()
            ids, pos, ctx = self._meta(seqs, is_prefill)
            hidden = self.model.forward(ids, pos, ctx, self.kv)
            tokens = self.model.greedy(self.model.logits(hidden, ctx)).tolist()
            before = sum(s.num_completion_tokens for s in seqs)
            self.scheduler.postprocess(seqs, tokens, is_prefill)
            produced += sum(s.num_completion_tokens for s in seqs) - before
            steps += 1
        return produced, time.perf_counter() - t0, steps
