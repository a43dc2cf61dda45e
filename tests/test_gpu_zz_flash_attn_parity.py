"""Our attention kernels against the reference's actual arithmetic: flash-attn, called exactly as the reference calls
it (nanovllm/layers/attention.py:67-74: flash_attn_varlen_func for prefill, flash_attn_with_kvcache for decode), on
the same tensors, on the GPU.

flash-attn is the reference's (unpinned, un-vendored) dependency; it cannot run on the CPU, so the CPU oracle restates
its published semantics.  This file closes the loop on the GPU box, where the image ships flash-attn 2.8.x: it is used
here as a checker only.  Both pipelines keep S and the accumulator in fp32 and round P to bf16, but tile and order
the sums differently, so the bar is the same as against the oracle: a few bf16 roundings of the output scale.
The cache is stored in both layouts from the same logical pages: ours head-major [nblk, Hkv, bs, D], flash-attn's
[nblk, bs, Hkv, D] (model_runner.py:115 of the reference).
"""
import random

import pytest
import torch

from gpu_helpers import record
from oracle.paged_attention_ref import to_physical
from test_gpu_kernels import bf, make_tables

pytestmark = pytest.mark.gpu


def flash():
    """Imported lazily (slow import; only needed on the GPU box); skip when the image has no flash-attn."""
    return pytest.importorskip("flash_attn", reason="flash-attn (the reference's attention library) is not installed")


def compare(got, want, what, ulps=6.0):
    got, want = got.float(), want.float()
    assert torch.isfinite(got).all(), what
    scale = max(want.abs().max().item(), 1e-3)
    err = (got - want).abs().max().item()
    l2 = ((got - want).norm() / max(want.norm().item(), 1e-6)).item()
    record("flash_attn_parity", dict(case=what, max_abs_err=err, scale=scale, err_in_bf16_ulps=err / (scale * 2 ** -8), rel_l2=l2))
    assert err <= ulps * scale * 2 ** -8, f"{what}: max abs err {err} vs scale {scale}"
    assert l2 <= 1e-2, f"{what}: relative L2 {l2}"


def bind_both_layouts(layers, nblk, hkv, bs, seed):
    """Logical pages -> (our bound cache, flash-attn's k caches, v caches) on the GPU."""
    from nanovllm import ops
    ks = [bf(nblk, bs, hkv, 128, seed=seed + 2 * l) for l in range(layers)]
    vs = [bf(nblk, bs, hkv, 128, seed=seed + 2 * l + 1) for l in range(layers)]
    kv = torch.stack([torch.stack([to_physical(k) for k in ks]), torch.stack([to_physical(v) for v in vs])]).cuda()
    ops.bind_kv_cache(kv)
    return kv, [k.cuda() for k in ks], [v.cuda() for v in vs]


@pytest.mark.parametrize("hq,hkv,lens", [
    (16, 8, [1, 15, 16, 17, 255, 256, 257, 1000, 2048, 3]),
    (16, 8, [random.Random(3).randint(100, 2048) for _ in range(64)]),            # benchmark-like rows
    (8, 2, [700, 64, 65, 1]),                                                     # G = 4: tensor-core decode variant
    (8, 1, [900, 31, 32, 33]),                                                    # G = 8
])
def test_decode_vs_flash_attn_with_kvcache(hq, hkv, lens):
    flash_attn_with_kvcache = flash().flash_attn_with_kvcache
    from nanovllm import ops
    bs = 256                                                                      # flash-attn pages are 256 tokens
    nblk = sum((c + bs - 1) // bs for c in lens) + 2
    _, ks, vs = bind_both_layouts(2, nblk, hkv, bs, seed=21)
    tables = make_tables(lens, bs, nblk, seed=4).cuda()
    ctx = torch.tensor(lens, dtype=torch.int32, device="cuda")
    q = bf(len(lens), hq, 128, seed=12).cuda()
    scale = 128 ** -0.5
    got = ops.paged_decode(1, q, tables, ctx, scale)
    want = flash_attn_with_kvcache(q.unsqueeze(1), ks[1], vs[1], cache_seqlens=ctx, block_table=tables,
                                   softmax_scale=scale, causal=True).squeeze(1)
    compare(got, want, f"decode hq={hq} hkv={hkv} n={len(lens)}")


@pytest.mark.parametrize("hq,hkv,lens", [(16, 8, [5, 128, 129, 700, 1024]), (8, 2, [300, 64, 1]), (4, 4, [257, 31])])
def test_prefill_packed_vs_flash_attn_varlen(hq, hkv, lens):
    flash_attn_varlen_func = flash().flash_attn_varlen_func
    from nanovllm import ops
    tot = sum(lens)
    row = bf(tot, (hq + 2 * hkv) * 128, seed=5).cuda()                            # q, k, v are views of the packed qkv row
    q = row[:, :hq * 128].view(tot, hq, 128)
    k = row[:, hq * 128:(hq + hkv) * 128].view(tot, hkv, 128)
    v = row[:, (hq + hkv) * 128:].view(tot, hkv, 128)
    cu = torch.tensor([0] + list(torch.tensor(lens).cumsum(0)), dtype=torch.int32, device="cuda")
    scale = 128 ** -0.5
    got = ops.paged_prefill(0, q, k, v, cu, cu, max(lens), max(lens), scale)
    want = flash_attn_varlen_func(q, k, v, max_seqlen_q=max(lens), cu_seqlens_q=cu, max_seqlen_k=max(lens), cu_seqlens_k=cu,
                                  softmax_scale=scale, causal=True, block_table=None)
    compare(got, want, f"packed prefill hq={hq} hkv={hkv}")


def test_prefill_paged_prefix_vs_flash_attn_varlen():
    """Prefix-cache / chunked-prefill shape: fewer queries than keys, keys read through the block table, mask aligned
    to the bottom right (the reference passes k_cache / v_cache and block_table, attention.py:64-70)."""
    flash_attn_varlen_func = flash().flash_attn_varlen_func
    from nanovllm import ops
    hq, hkv, bs = 16, 8, 256
    len_k = [600, 257, 1024, 40]
    len_q = [88, 257, 1, 13]
    nblk = sum((c + bs - 1) // bs for c in len_k) + 2
    _, ks, vs = bind_both_layouts(2, nblk, hkv, bs, seed=41)
    tables = make_tables(len_k, bs, nblk, seed=7).cuda()
    q = bf(sum(len_q), hq, 128, seed=6).cuda()
    cu_q = torch.tensor([0] + list(torch.tensor(len_q).cumsum(0)), dtype=torch.int32, device="cuda")
    cu_k = torch.tensor([0] + list(torch.tensor(len_k).cumsum(0)), dtype=torch.int32, device="cuda")
    scale = 128 ** -0.5
    got = ops.paged_prefill(1, q, None, None, cu_q, cu_k, max(len_q), max(len_k), scale, block_tables=tables, num_kv_heads=hkv)
    want = flash_attn_varlen_func(q, ks[1], vs[1], max_seqlen_q=max(len_q), cu_seqlens_q=cu_q, max_seqlen_k=max(len_k),
                                  cu_seqlens_k=cu_k, softmax_scale=scale, causal=True, block_table=tables)
    compare(got, want, "paged prefix prefill")
