"""Paged KV block allocator with ref-counts and chained-xxh64 prefix caching.

Semantics are the reference's (nanovllm/engine/block_manager.py:26-120), reproduced bit for bit
because block ids and slot numbers are the contract with the GPU kernels:

* free list is FIFO (allocate from the left, freed blocks go to the right), blocks of a
  sequence are released in reverse table order (block_manager.py:43-56, 94-101);
* a block's hash = xxh64(8-byte LE hash of the previous block, if any  ||  its token ids as
  int64 LE) (block_manager.py:35-41); the hash -> block map keeps entries of freed blocks until
  the block is handed out again, so a freed block can be revived as a cache hit (47-48, 80-88);
* a lookup hit must also match the token ids themselves (62-70); the last block of a prompt is
  never taken from the cache, so at least one token is always computed;
* a decode step needs a fresh block exactly when len(seq) % block_size == 1 (103-108).

State is kept in flat per-block arrays (ref count, hash, token tuple) instead of one Python
object per block; `blocks[i]` exposes a read-only view with the reference's field names.
"""
from __future__ import annotations

from array import array
from collections import deque

import xxhash

from .sequence import Sequence


class _BlockView:
    __slots__ = ("_m", "block_id")

    def __init__(self, mgr: "BlockManager", block_id: int):
        self._m, self.block_id = mgr, block_id

    @property
    def ref_count(self) -> int:
        return self._m._ref[self.block_id]

    @property
    def hash(self) -> int:
        return self._m._hash[self.block_id]

    @property
    def token_ids(self) -> list[int]:
        t = self._m._tokens[self.block_id]
        return list(t) if t is not None else []


class BlockManager:
    def __init__(self, num_blocks: int, block_size: int):
        self.block_size = block_size
        self.num_blocks = num_blocks
        self._ref = [0] * num_blocks
        self._hash = [-1] * num_blocks
        self._tokens: list[tuple | None] = [None] * num_blocks
        self.hash_to_block_id: dict[int, int] = {}
        self.free_block_ids: deque[int] = deque(range(num_blocks))
        self.used_block_ids: set[int] = set()
        self.blocks = [_BlockView(self, i) for i in range(num_blocks)]

    # ---- hashing ----------------------------------------------------------------------------
    @staticmethod
    def compute_hash(token_ids, prefix: int = -1) -> int:
        h = xxhash.xxh64()
        if prefix != -1:
            h.update(prefix.to_bytes(8, "little"))
        h.update(array("q", token_ids).tobytes())
        return h.intdigest()

    # ---- single-block primitives ---------------------------------------------------------------
    def _take_free_block(self) -> int:
        bid = self.free_block_ids.popleft()
        assert self._ref[bid] == 0
        old = self._hash[bid]
        if old != -1 and self.hash_to_block_id.get(old) == bid:
            del self.hash_to_block_id[old]            # its cached content is about to be overwritten
        self._ref[bid], self._hash[bid], self._tokens[bid] = 1, -1, None
        self.used_block_ids.add(bid)
        return bid

    def _release_block(self, bid: int) -> None:
        assert self._ref[bid] == 0
        self.used_block_ids.remove(bid)
        self.free_block_ids.append(bid)               # hash entry survives until the block is reused

    # ---- prompt admission --------------------------------------------------------------------
    def can_allocate(self, seq: Sequence) -> int:
        """-1 if the prompt does not fit, else the number of leading blocks found in the cache."""
        nblocks = seq.num_blocks
        need = nblocks
        hits = 0
        h = -1
        for i in range(nblocks - 1):
            toks = seq.block(i)
            h = self.compute_hash(toks, h)
            bid = self.hash_to_block_id.get(h, -1)
            if bid == -1 or self._tokens[bid] != tuple(toks):
                break
            hits += 1
            if bid in self.used_block_ids:
                need -= 1                              # shared with a live sequence: no free block consumed
        return hits if len(self.free_block_ids) >= need else -1

    def allocate(self, seq: Sequence, num_cached_blocks: int) -> None:
        assert not seq.block_table
        table = seq.block_table
        h = -1
        for i in range(num_cached_blocks):
            h = self.compute_hash(seq.block(i), h)
            bid = self.hash_to_block_id[h]
            if bid in self.used_block_ids:
                self._ref[bid] += 1
            else:                                      # revive a freed block whose content is still valid
                self._ref[bid] = 1
                self.free_block_ids.remove(bid)
                self.used_block_ids.add(bid)
            table.append(bid)
        for _ in range(num_cached_blocks, seq.num_blocks):
            table.append(self._take_free_block())
        seq.num_cached_tokens = num_cached_blocks * self.block_size

    def deallocate(self, seq: Sequence) -> None:
        for bid in reversed(seq.block_table):
            self._ref[bid] -= 1
            if self._ref[bid] == 0:
                self._release_block(bid)
        seq.num_cached_tokens = 0
        seq.block_table.clear()

    # ---- decode growth -----------------------------------------------------------------------
    def can_append(self, seq: Sequence) -> bool:
        return len(self.free_block_ids) >= (1 if len(seq) % self.block_size == 1 else 0)

    def may_append(self, seq: Sequence) -> None:
        if len(seq) % self.block_size == 1:
            seq.block_table.append(self._take_free_block())

    # ---- publishing finished blocks to the prefix cache --------------------------------------------
    def hash_blocks(self, seq: Sequence) -> None:
        bs = self.block_size
        first = seq.num_cached_tokens // bs
        last = (seq.num_cached_tokens + seq.num_scheduled_tokens) // bs
        if first == last:
            return
        h = self._hash[seq.block_table[first - 1]] if first > 0 else -1
        for i in range(first, last):
            bid = seq.block_table[i]
            toks = seq.block(i)
            h = self.compute_hash(toks, h)
            self._hash[bid], self._tokens[bid] = h, tuple(toks)
            self.hash_to_block_id[h] = bid

    # ---- recovery ----------------------------------------------------------------------------------
    def forget_prefix_cache(self) -> None:
        """Drop every hash -> block entry (the KV bytes behind them can no longer be trusted, e.g. after a failed
        tensor-parallel exchange wrote undefined data).  Live block tables and ref counts are untouched."""
        self.hash_to_block_id.clear()
        for bid in range(self.num_blocks):
            self._hash[bid], self._tokens[bid] = -1, None
