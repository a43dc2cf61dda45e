"""Whole-model and engine parity on the GPU: product forward (CUDA kernels + cuBLAS) vs the CPU oracle
that tests/test_oracle_golden.py pinned to the reference's own modules."""
import os
import random
from types import SimpleNamespace

import pytest
import torch

from gpu_helpers import check_greedy_against_oracle, make_oracle, record
from oracle.model_script import make_script, run_script
from oracle.qwen3_ref import alloc_logical_kv

pytestmark = pytest.mark.gpu


def build_product_model(preset, weights, nblk, bs):
    from nanovllm import ops
    from nanovllm.models.qwen3 import Qwen3ForCausalLM
    from nanovllm.utils.synthetic import PRESETS, hf_config_dict
    hf = SimpleNamespace(**hf_config_dict(PRESETS[preset]))
    model = Qwen3ForCausalLM(hf, 0, 1, "cuda", max_position=4096)
    for name, w in weights.items():
        model.load_hf_tensor(name, w)
    kv = torch.zeros(ops.kv_cache_shape(hf.num_hidden_layers, nblk, model.num_kv_heads, bs, model.head_dim),
                     dtype=torch.bfloat16, device="cuda")
    ops.bind_kv_cache(kv)
    for i, a in enumerate(model.modules()):
        a.k_cache, a.v_cache, a.layer_id = kv[0, i], kv[1, i], i
    return model, kv


@pytest.mark.parametrize("preset", ["tiny", "tiny-g4", "tiny-g1", "tiny-g8"])
def test_model_script_vs_oracle(preset):
    from nanovllm.utils.context import reset_context, set_context
    from nanovllm.utils.synthetic import PRESETS, random_weights
    weights = random_weights(PRESETS[preset], seed=1234)
    script = make_script(PRESETS[preset]["vocab_size"])
    model, _ = build_product_model(preset, weights, script["num_blocks"], script["block_size"])

    def gpu_step(ids, pos, c):
        set_context(c["is_prefill"], c.get("cu_seqlens_q"), c.get("cu_seqlens_k"), c.get("max_seqlen_q", 0),
                    c.get("max_seqlen_k", 0), c.get("slot_mapping"), c.get("context_lens"), c.get("block_tables"))
        out = model.compute_logits(model(ids, pos)).float().cpu()
        reset_context()
        return out

    got = run_script(torch, script, gpu_step, device="cuda")

    oracle = make_oracle(PRESETS[preset], weights, "fused")
    kv = alloc_logical_kv(oracle.d, script["num_blocks"], script["block_size"])
    from test_oracle_golden import oracle_step_fn
    want = [o.float() for o in run_script(torch, script, oracle_step_fn(oracle, kv))]

    worst_rel, flips, rows = 0.0, 0, 0
    for i, (g, w) in enumerate(zip(got, want)):
        assert g.shape == w.shape and torch.isfinite(g).all()
        rel = ((g - w).norm() / w.norm()).item()
        worst_rel = max(worst_rel, rel)
        assert rel < 2e-2, f"step {i}: logits relative L2 {rel}"
        tol = 6 * 2 ** -8 * w.abs().max().item()
        for r in range(g.shape[0]):
            tok = int(g[r].argmax())
            margin = (w[r].max() - w[r, tok]).item()
            assert margin <= tol, f"step {i} row {r}: greedy token {tok} loses by {margin} in the oracle (tol {tol})"
            flips += tok != int(w[r].argmax())
            rows += 1
    # and directly against the committed golden vectors: the logits the REFERENCE's own nn.Modules produced on CPU
    # (tests/golden/model_*.npz, eager rounding) -- same script, same weights
    import os
    import numpy as np
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", f"model_{preset}.npz"))
    worst_gold = 0.0
    for i, g in enumerate(got):
        w = torch.from_numpy(gold[f"logits_{i}"]).view(torch.bfloat16).float()
        rel = ((g - w).norm() / w.norm()).item()
        worst_gold = max(worst_gold, rel)
        # expected ~1.3e-2: 1.0e-2 separates the reference's eager CPU rounding from its fused GPU rounding (measured on
        # the oracle), 0.7e-2 separates our kernels from the fused oracle
        assert rel < 3.5e-2, f"step {i}: logits vs the reference modules' golden output: relative L2 {rel}"
    record("model_script", dict(preset=preset, worst_rel_l2=worst_rel, worst_rel_l2_vs_reference_golden=worst_gold,
                                greedy_rows=rows, greedy_differ=flips))


@pytest.fixture(scope="module")
def tiny_dir(tmp_path_factory):
    from nanovllm.utils.synthetic import make_model_dir
    return make_model_dir(str(tmp_path_factory.mktemp("models") / "tiny"), "tiny", seed=1234)


def _prompts(vocab, seed=0):
    rnd = random.Random(seed)
    shared = [rnd.randint(2, vocab - 1) for _ in range(40)]
    ps = [[rnd.randint(2, vocab - 1) for _ in range(rnd.randint(3, 90))] for _ in range(10)]
    ps += [shared + [rnd.randint(2, vocab - 1) for _ in range(rnd.randint(1, 20))] for _ in range(6)]
    return ps


@pytest.mark.parametrize("eager", [True, False])
def test_engine_greedy_generate_vs_oracle(tiny_dir, eager):
    from nanovllm import LLM, SamplingParams
    from nanovllm.utils.synthetic import PRESETS, random_weights
    llm = LLM(tiny_dir, enforce_eager=eager, max_model_len=256, max_num_seqs=8, max_num_batched_tokens=128,
              kvcache_block_size=16, num_kvcache_blocks=64)
    try:
        prompts = _prompts(PRESETS["tiny"]["vocab_size"])
        sps = [SamplingParams(temperature=0.0, max_tokens=8 + (i % 5) * 6, ignore_eos=True) for i in range(len(prompts))]
        outs = llm.generate(prompts, sps, use_tqdm=False)
    finally:
        llm.exit()
    assert len(outs) == len(prompts)
    oracle = make_oracle(PRESETS["tiny"], random_weights(PRESETS["tiny"], seed=1234), "fused")
    tot = diff = 0
    worst = 0.0
    for p, sp, o in zip(prompts, sps, outs):
        assert set(o) == {"text", "token_ids"} and len(o["token_ids"]) == sp.max_tokens
        n, d, w = check_greedy_against_oracle(oracle, p, o["token_ids"])
        tot, diff, worst = tot + n, diff + d, max(worst, w)
    record("engine_greedy", dict(eager=eager, tokens=tot, differ_from_oracle_argmax=diff, worst_margin_over_tol=worst))


@pytest.mark.parametrize("eager", [True, False])
def test_engine_two_stream_decode_vs_oracle(tiny_dir, eager, monkeypatch):
    """B200_DUAL=1: every decode step of >= 2 rows runs as two half batches on two streams (event chain between the two
    attention launches, tcgen05 projections with a 3-slot ring, captured into the decode graphs when not eager); greedy
    tokens teacher-forced against the oracle, and equal to the single-stream engine's wherever the oracle has a clear winner."""
    from nanovllm import LLM, SamplingParams
    from nanovllm.utils.synthetic import PRESETS, random_weights
    monkeypatch.setenv("B200_DUAL", "1")
    monkeypatch.setenv("B200_DUAL_MIN", "2")
    llm = LLM(tiny_dir, enforce_eager=eager, max_model_len=256, max_num_seqs=8, max_num_batched_tokens=128,
              kvcache_block_size=16, num_kvcache_blocks=64)
    try:
        assert llm.model_runner.model.dual
        prompts = _prompts(PRESETS["tiny"]["vocab_size"])
        sps = [SamplingParams(temperature=0.0, max_tokens=8 + (i % 5) * 6, ignore_eos=True) for i in range(len(prompts))]
        outs = llm.generate(prompts, sps, use_tqdm=False)
        again = llm.generate(prompts, sps, use_tqdm=False)             # replays of the same graphs / streams: deterministic
    finally:
        llm.exit()
    assert [o["token_ids"] for o in outs] == [o["token_ids"] for o in again]
    oracle = make_oracle(PRESETS["tiny"], random_weights(PRESETS["tiny"], seed=1234), "fused")
    tot = diff = 0
    worst = 0.0
    for p, sp, o in zip(prompts, sps, outs):
        assert len(o["token_ids"]) == sp.max_tokens
        n, d, w = check_greedy_against_oracle(oracle, p, o["token_ids"])
        tot, diff, worst = tot + n, diff + d, max(worst, w)
    record("engine_two_stream", dict(eager=eager, tokens=tot, differ_from_oracle_argmax=diff, worst_margin_over_tol=worst))


def test_engine_chunked_prefill_and_preemption(tiny_dir):
    """Prompts longer than max_num_batched_tokens (chunked prefill through the paged path) and a KV cache too small
    for the whole batch (preemption + re-prefill through the prefix cache), greedy, checked against the oracle."""
    from nanovllm import LLM, SamplingParams
    from nanovllm.utils.synthetic import PRESETS, random_weights
    rnd = random.Random(11)
    vocab = PRESETS["tiny"]["vocab_size"]
    prompts = [[rnd.randint(2, vocab - 1) for _ in range(n)] for n in (200, 150, 33, 90, 64, 17)]
    sps = [SamplingParams(temperature=0.0, max_tokens=40, ignore_eos=True) for _ in prompts]
    llm = LLM(tiny_dir, max_model_len=256, max_num_seqs=6, max_num_batched_tokens=64, kvcache_block_size=16, num_kvcache_blocks=40)
    try:
        outs = llm.generate(prompts, sps, use_tqdm=False)
        assert len(llm.scheduler.block_manager.used_block_ids) == 0
    finally:
        llm.exit()
    oracle = make_oracle(PRESETS["tiny"], random_weights(PRESETS["tiny"], seed=1234), "fused")
    tot = diff = 0
    for p, o in zip(prompts, outs):
        assert len(o["token_ids"]) == 40
        n, d, _ = check_greedy_against_oracle(oracle, p, o["token_ids"])
        tot, diff = tot + n, diff + d
    record("engine_chunked_preempt", dict(tokens=tot, differ_from_oracle_argmax=diff))


def test_prefill_bench_size_rows_vs_oracle():
    """BASELINE config-2 prefill shape (Hq16/Hkv8, prompts up to 1024 tokens, ~16k tokens per step): a few whole
    sequences against the oracle, the rest through a property -- every sequence's output is independent of its
    neighbours in the batch."""
    from nanovllm import ops
    from oracle.paged_attention_ref import varlen_prefill_ref
    from oracle.make_golden import workloads
    from test_gpu_kernels import assert_close_bf16
    lens = [len(p) for p in workloads()["bench"]["prompts"][:31]]
    tot = sum(lens)
    assert tot == 15705
    g = torch.Generator().manual_seed(5)
    qkv = torch.randn(tot, 32 * 128, generator=g).to(torch.bfloat16)
    cu = torch.tensor([0] + list(torch.tensor(lens).cumsum(0)), dtype=torch.int32)
    dev = qkv.cuda()
    q, k, v = dev[:, :2048].view(tot, 16, 128), dev[:, 2048:3072].view(tot, 8, 128), dev[:, 3072:].view(tot, 8, 128)
    out = ops.paged_prefill(0, q, k, v, cu.cuda(), cu.cuda(), max(lens), max(lens), 128 ** -0.5).cpu()
    for s in (0, 7, 30):
        a, b = int(cu[s]), int(cu[s + 1])
        c2 = torch.tensor([0, b - a], dtype=torch.int32)
        want = varlen_prefill_ref(qkv[a:b, :2048].view(-1, 16, 128), qkv[a:b, 2048:3072].view(-1, 8, 128),
                                  qkv[a:b, 3072:].view(-1, 8, 128), c2, c2, 128 ** -0.5, p_dtype=torch.bfloat16)
        assert_close_bf16(out[a:b], want, f"bench-size prefill seq {s}")
        alone = ops.paged_prefill(0, q[a:b], k[a:b], v[a:b], c2.cuda(), c2.cuda(), b - a, b - a, 128 ** -0.5).cpu()
        assert torch.equal(alone, out[a:b]), "a sequence's output must not depend on its batch neighbours"


def test_engine_sampling_and_eos(tiny_dir):
    from nanovllm import LLM, SamplingParams
    llm = LLM(tiny_dir, max_model_len=128, max_num_seqs=4, kvcache_block_size=32, num_kvcache_blocks=32)
    try:
        outs = llm.generate([[5, 6, 7, 8]] * 6, SamplingParams(temperature=1.0, max_tokens=20), use_tqdm=False)
        assert all(1 <= len(o["token_ids"]) <= 20 for o in outs)
        assert all(0 <= t < 2048 for o in outs for t in o["token_ids"])
        assert len({tuple(o["token_ids"]) for o in outs}) > 1, "temperature sampling should not be deterministic across rows"
        for o in outs:                                  # EOS (id 1) ends a sequence unless ignore_eos
            assert 1 not in o["token_ids"][:-1]
        text = llm.generate(["t5 t6 t7"], SamplingParams(temperature=0.0, max_tokens=3), use_tqdm=False)[0]
        assert isinstance(text["text"], str) and len(text["token_ids"]) <= 3
    finally:
        llm.exit()


def test_decode_bench_shape_vs_oracle_rows():
    """BASELINE config-2 decode shape (B=256, Hq16/Hkv8, page 256, context mix of bench step 0): sampled rows
    against the oracle, and the checksum property that every row only depends on its own pages."""
    from nanovllm import ops
    from oracle.make_golden import workloads
    from oracle.paged_attention_ref import paged_decode_ref, to_logical
    w = workloads()["bench"]
    lens = [len(p) + 1 for p in w["prompts"]]
    assert sum(lens) == 143083                           # sum of contexts at decode step 0
    hq, hkv, bs = 16, 8, 256
    nblk = sum((c + bs - 1) // bs for c in lens)
    g = torch.Generator(device="cuda").manual_seed(0)
    kv = torch.randn(ops.kv_cache_shape(1, nblk, hkv, bs, 128), generator=g, device="cuda", dtype=torch.float32).to(torch.bfloat16)
    ops.bind_kv_cache(kv)
    perm = torch.randperm(nblk, generator=torch.Generator().manual_seed(1)).to(torch.int32)
    tables = torch.full((256, 8), -1, dtype=torch.int32)
    used = 0
    for i, c in enumerate(lens):
        n = (c + bs - 1) // bs
        tables[i, :n] = perm[used:used + n]
        used += n
    ctx = torch.tensor(lens, dtype=torch.int32)
    q = torch.randn(256, hq, 128, generator=torch.Generator().manual_seed(2)).to(torch.bfloat16)
    out = ops.paged_decode(0, q.cuda(), tables.cuda(), ctx.cuda(), 128 ** -0.5).cpu()
    rows = [0, 17, 100, 255, max(range(256), key=lambda i: lens[i]), min(range(256), key=lambda i: lens[i])]
    kl, vl = to_logical(kv[0, 0].cpu()), to_logical(kv[1, 0].cpu())
    want = paged_decode_ref(q[rows], kl, vl, ctx[rows], tables[rows], 128 ** -0.5)
    from test_gpu_kernels import assert_close_bf16
    assert_close_bf16(out[rows], want, "bench-shape decode rows")
    sub = [3, 99, 200]                                    # a sub-batch must reproduce its rows bit for bit
    out2 = ops.paged_decode(0, q[sub].cuda(), tables[sub].cuda(), ctx[sub].cuda(), 128 ** -0.5).cpu()
    assert_close_bf16(out2, out[sub], "row independence", ulps=1.01, rel_l2=2e-3)
