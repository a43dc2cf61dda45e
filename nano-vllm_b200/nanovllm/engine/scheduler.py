"""Continuous-batching scheduler (reference nanovllm/engine/scheduler.py:25-92), same decisions.

* prefill first: admit waiting sequences in FIFO order while sequence and token budgets allow;
  blocks for the whole prompt are reserved at first admission; only the first sequence of a
  step may be split into chunks (scheduler.py:30-53);
* a decode step runs only when no prefill was scheduled; if a sequence needs a block and none
  is free, the most recently admitted running sequence is preempted (its blocks freed, it goes
  back to the head of the waiting queue), else the sequence itself (54-73);
* postprocess publishes finished blocks to the prefix cache, advances the cached-token
  count, appends the sampled token unless the prompt is still incomplete, and retires sequences
  on EOS / max_tokens (81-92).
"""
from __future__ import annotations

from collections import deque

from ..config import Config
from .block_manager import BlockManager
from .sequence import Sequence, SequenceStatus


class Scheduler:
    def __init__(self, config: Config):
        self.max_num_seqs = config.max_num_seqs
        self.max_num_batched_tokens = config.max_num_batched_tokens
        self.eos = config.eos
        self.block_size = config.kvcache_block_size
        self.block_manager = BlockManager(config.num_kvcache_blocks, config.kvcache_block_size)
        self.waiting: deque[Sequence] = deque()
        self.running: deque[Sequence] = deque()

    def is_finished(self) -> bool:
        return not (self.waiting or self.running)

    def add(self, seq: Sequence) -> None:
        self.waiting.append(seq)

    # ---- one step ----------------------------------------------------------------------------
    def schedule(self) -> tuple[list[Sequence], bool]:
        batch = self._admit_prefill()
        if batch:
            return batch, True
        batch = self._pick_decode()
        assert batch, "nothing schedulable: KV cache too small for a single sequence"
        return batch, False

    def _admit_prefill(self) -> list[Sequence]:
        bm = self.block_manager
        batch: list[Sequence] = []
        budget = self.max_num_batched_tokens
        while self.waiting and len(batch) < self.max_num_seqs and budget > 0:
            seq = self.waiting[0]
            fresh = not seq.block_table
            if fresh:
                hits = bm.can_allocate(seq)
                if hits < 0:
                    break
                todo = seq.num_tokens - hits * self.block_size
            else:                                       # continuing a chunked prompt
                todo = seq.num_tokens - seq.num_cached_tokens
            if todo > budget and batch:                 # only the head of a step may be chunked
                break
            if fresh:
                bm.allocate(seq, hits)
            seq.num_scheduled_tokens = min(todo, budget)
            budget -= seq.num_scheduled_tokens
            if seq.num_cached_tokens + seq.num_scheduled_tokens == seq.num_tokens:
                seq.status = SequenceStatus.RUNNING
                self.waiting.popleft()
                self.running.append(seq)
            batch.append(seq)
        return batch

    def _pick_decode(self) -> list[Sequence]:
        bm = self.block_manager
        batch: list[Sequence] = []
        while self.running and len(batch) < self.max_num_seqs:
            seq = self.running.popleft()
            evicted_self = False
            while not bm.can_append(seq):
                if self.running:
                    self.preempt(self.running.pop())
                else:
                    self.preempt(seq)
                    evicted_self = True
                    break
            if evicted_self:
                continue
            seq.num_scheduled_tokens = 1
            seq.is_prefill = False
            bm.may_append(seq)
            batch.append(seq)
        self.running.extendleft(reversed(batch))        # keep admission order at the head
        return batch

    def preempt(self, seq: Sequence) -> None:
        seq.status = SequenceStatus.WAITING
        seq.is_prefill = True
        self.block_manager.deallocate(seq)
        self.waiting.appendleft(seq)

    def postprocess(self, seqs: list[Sequence], token_ids: list[int], is_prefill: bool) -> None:
        bm = self.block_manager
        for seq, tok in zip(seqs, token_ids):
            bm.hash_blocks(seq)
            seq.num_cached_tokens += seq.num_scheduled_tokens
            seq.num_scheduled_tokens = 0
            if is_prefill and seq.num_cached_tokens < seq.num_tokens:
                continue                                # prompt chunk: the sampled token is meaningless
            seq.append_token(tok)
            if (tok == self.eos and not seq.ignore_eos) or seq.num_completion_tokens == seq.max_tokens:
                seq.status = SequenceStatus.FINISHED
                bm.deallocate(seq)
                self.running.remove(seq)

    def abort_all(self) -> list[Sequence]:
        """Retire every queued and running sequence, free its blocks and forget the prefix cache: the state after a
        step whose results cannot be trusted.  Returns the aborted sequences."""
        dropped = list(self.running) + list(self.waiting)
        for seq in dropped:
            if seq.block_table:
                self.block_manager.deallocate(seq)
            seq.status = SequenceStatus.FINISHED
        self.running.clear()
        self.waiting.clear()
        self.block_manager.forget_prefix_cache()
        return dropped
