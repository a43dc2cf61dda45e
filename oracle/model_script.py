"""A fixed, teacher-forced serving script used to compare whole-model forwards.  TEST INFRASTRUCTURE ONLY.

Seven steps over one small paged cache (block size 16) that visit every branch of the reference's
Attention.forward (layers/attention.py:59-75) with the metadata its runner would build
(engine/model_runner.py:129-188): packed prefill, decode (with a block boundary), a prefix-cache hit
batched with a fresh prompt (paged prefill, len_q < len_k), a two-chunk prompt, and a mixed decode.
Tokens are fixed in advance, so every implementation sees identical inputs at every step.

``run_script(torch, script, step_fn)`` calls ``step_fn(input_ids, positions, ctx_dict) -> logits``
and returns the list of logits; it is driven by
  * oracle/make_golden.py with the REFERENCE's nn.Modules        -> tests/golden/model_*.npz
  * tests with oracle.qwen3_ref.Qwen3Ref                           (must equal the golden bit for bit)
  * GPU tests with the product model                               (compared with Qwen3Ref "fused")
"""
from __future__ import annotations

import random

BLOCK = 16
NUM_BLOCKS = 24


def make_script(vocab: int, seed: int = 7, scale: int = 1, max_token: int | None = None) -> dict:
    """``scale`` multiplies every length and the block size (scale 16 -> 256-token blocks, the only page size the
    reference's flash-attn path accepts, config.py:22): the same seven steps at serving-size sequence lengths.
    ``max_token`` bounds the token ids (default: the whole vocabulary)."""
    rnd = random.Random(seed)
    hi = (max_token if max_token is not None else vocab) - 1
    tok = lambda n: [rnd.randint(0, hi) for _ in range(n)]
    A, B, C = tok(37 * scale), tok(16 * scale), tok(70 * scale)
    D = A[:32 * scale] + tok(9 * scale)           # shares A's first two full blocks
    E, F = tok(20 * scale), tok(50 * scale)
    forced = {name: tok(4) for name in "ABCDEF"}
    return dict(block_size=BLOCK * scale, num_blocks=NUM_BLOCKS, vocab=vocab, scale=scale,
                prompts=dict(A=A, B=B, C=C, D=D, E=E, F=F), forced=forced)


def _slots(table, start, end, bs=BLOCK):
    return [table[p // bs] * bs + p % bs for p in range(start, end)]


def script_steps(script: dict) -> list[dict]:
    """Plain-Python description of every step (lists of ints), independent of torch."""
    P, Fd = script["prompts"], script["forced"]
    sc = script.get("scale", 1)
    bs = script["block_size"]
    toks = {k: list(v) for k, v in P.items()}
    tables = dict(A=[0, 1, 2], B=[3], C=[4, 5, 6, 7, 8])
    steps = []

    def prefill(entries, paged):
        # entries: (name, start, end)
        ids, pos, slots, cu_q, cu_k = [], [], [], [0], [0]
        for name, s, e in entries:
            ids += toks[name][s:e]
            pos += list(range(s, e))
            slots += _slots(tables[name], s, e, bs)
            cu_q.append(cu_q[-1] + e - s)
            cu_k.append(cu_k[-1] + e)
        bt = None
        if paged:
            w = max(len(tables[n]) for n, _, _ in entries)
            bt = [tables[n] + [-1] * (w - len(tables[n])) for n, _, _ in entries]
        steps.append(dict(is_prefill=True, input_ids=ids, positions=pos, slot_mapping=slots, cu_seqlens_q=cu_q,
                          cu_seqlens_k=cu_k, max_seqlen_q=max(e - s for _, s, e in entries),
                          max_seqlen_k=max(e for _, _, e in entries), block_tables=bt))

    def decode(names, new_blocks):
        for n in names:
            toks[n].append(Fd[n][len(toks[n]) - len(P[n])])
            if n in new_blocks:
                tables[n].append(new_blocks[n])
        w = max(len(tables[n]) for n in names)
        steps.append(dict(is_prefill=False, input_ids=[toks[n][-1] for n in names],
                          positions=[len(toks[n]) - 1 for n in names],
                          slot_mapping=[_slots(tables[n], len(toks[n]) - 1, len(toks[n]), bs)[0] for n in names],
                          context_lens=[len(toks[n]) for n in names],
                          block_tables=[tables[n] + [-1] * (w - len(tables[n])) for n in names]))

    prefill([("A", 0, 37 * sc), ("B", 0, 16 * sc), ("C", 0, 70 * sc)], paged=False)   # 0: packed prefill
    decode(["A", "B", "C"], {"B": 9})                                          # 1: B crosses into a new block
    decode(["A", "B", "C"], {})                                                # 2
    tables["D"] = [0, 1, 10]
    tables["E"] = [11, 12]
    prefill([("D", 32 * sc, 41 * sc), ("E", 0, 20 * sc)], paged=True)          # 3: prefix hit + fresh prompt
    tables["F"] = [13, 14, 15, 16]
    prefill([("F", 0, 24 * sc)], paged=False)                                  # 4: first chunk
    prefill([("F", 24 * sc, 50 * sc)], paged=True)                             # 5: second chunk over the cache
    decode(["A", "D", "F"], {})                                                # 6: mixed lengths 40 / 42 / 51
    return steps


def run_script(torch, script: dict, step_fn, device="cpu") -> list:
    outs = []
    for st in script_steps(script):
        i32 = lambda x: None if x is None else torch.tensor(x, dtype=torch.int32, device=device)
        ctx = dict(is_prefill=st["is_prefill"], slot_mapping=i32(st["slot_mapping"]), block_tables=i32(st["block_tables"]))
        if st["is_prefill"]:
            ctx.update(cu_seqlens_q=i32(st["cu_seqlens_q"]), cu_seqlens_k=i32(st["cu_seqlens_k"]),
                       max_seqlen_q=st["max_seqlen_q"], max_seqlen_k=st["max_seqlen_k"])
        else:
            ctx.update(context_lens=i32(st["context_lens"]))
        ids = torch.tensor(st["input_ids"], dtype=torch.int64, device=device)
        pos = torch.tensor(st["positions"], dtype=torch.int64, device=device)
        outs.append(step_fn(ids, pos, ctx))
    return outs
