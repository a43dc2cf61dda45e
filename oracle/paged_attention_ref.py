"""CPU restatement of the reference's attention operator.  TEST INFRASTRUCTURE ONLY.

Follows, function by function:

* ``store_kvcache`` / ``store_kvcache_kernel``   reference nanovllm/layers/attention.py:10-40
* ``Attention.forward`` dispatch                  reference nanovllm/layers/attention.py:59-75
* metadata meaning (slot_mapping, block_tables,
  cu_seqlens, context_lens)                       reference nanovllm/engine/model_runner.py:123-188
* mask alignment / paged layout of the third-party
  kernels the reference calls (flash-attn 2.8.3,
  unpinned, not vendored): ``flash_attn_varlen_func``
  (bottom-right aligned causal mask) and
  ``flash_attn_with_kvcache`` (paged KV,
  ``cache_seqlens``), flash_attn_interface.py:1394-1402,1526-1546.

Logical cache layout here is the reference's: ``[num_blocks, block_size,
num_kv_heads, head_dim]`` (model_runner.py:115).  The product stores pages
head-major in HBM; the tests convert with
``to_physical`` / ``to_logical`` below, so this file stays a restatement of
the reference and nothing else.

Arithmetic restated: S = scale * Q K^T in fp32, causal mask, fp32 softmax,
fp32 P V, output rounded once to the input dtype.  (flash-attn rounds P to
bf16 before the PV MMA; pass ``p_dtype=torch.bfloat16`` to mimic that.)

parity unpinned: the reference has no golden vectors for this arithmetic.
"""
from __future__ import annotations

import torch


# --------------------------------------------------------------------------
# layout helpers (logical reference layout <-> product physical layout)
# --------------------------------------------------------------------------
def to_physical(cache_logical: torch.Tensor) -> torch.Tensor:
    """[nblk, bs, Hkv, D] (reference) -> [nblk, Hkv, bs, D] (product HBM layout)."""
    return cache_logical.permute(0, 2, 1, 3).contiguous()


def to_logical(cache_physical: torch.Tensor) -> torch.Tensor:
    """[nblk, Hkv, bs, D] (product) -> [nblk, bs, Hkv, D] (reference)."""
    return cache_physical.permute(0, 2, 1, 3).contiguous()


# --------------------------------------------------------------------------
# K1: KV-cache scatter            reference layers/attention.py:10-40
# --------------------------------------------------------------------------
def store_kvcache_ref(key, value, k_cache, v_cache, slot_mapping):
    """cache.view(-1, Hkv*D)[slot] = kv[i]  for every i with slot != -1.

    key/value: [N, Hkv, D]; caches: [nblk, bs, Hkv, D]; slot_mapping: [N] int.
    Pure copy of bf16 bit patterns (attention.py:21-30).
    """
    n, hkv, d = key.shape
    flat_k = k_cache.view(-1, hkv * d)
    flat_v = v_cache.view(-1, hkv * d)
    slots = slot_mapping.to(torch.long)
    keep = slots >= 0
    flat_k[slots[keep]] = key.reshape(n, hkv * d)[keep]
    flat_v[slots[keep]] = value.reshape(n, hkv * d)[keep]


# --------------------------------------------------------------------------
# core: one sequence, all heads
# --------------------------------------------------------------------------
def _attend_one(q, k, v, scale, p_dtype=None):
    """q [Lq,Hq,D], k/v [Lk,Hkv,D] -> [Lq,Hq,D] fp32.

    Bottom-right aligned causal mask: query i sees keys j <= i + Lk - Lq.
    """
    lq, hq, d = q.shape
    lk, hkv, _ = k.shape
    g = hq // hkv
    qf = q.float().permute(1, 0, 2)                     # [Hq, Lq, D]
    kf = k.float().permute(1, 0, 2).repeat_interleave(g, dim=0)
    vf = v.float().permute(1, 0, 2).repeat_interleave(g, dim=0)
    s = torch.matmul(qf, kf.transpose(1, 2)) * scale    # [Hq, Lq, Lk]
    i = torch.arange(lq).view(-1, 1)
    j = torch.arange(lk).view(1, -1)
    s = s.masked_fill(j > i + (lk - lq), float("-inf"))
    m = s.amax(dim=-1, keepdim=True)
    p = torch.exp(s - m)
    l = p.sum(dim=-1, keepdim=True)
    if p_dtype is not None:
        p = p.to(p_dtype).float()
    o = torch.matmul(p, vf) / l
    return o.permute(1, 0, 2)                           # [Lq, Hq, D]


def gather_pages(cache, block_table_row, length):
    """Rows 0..length-1 of one sequence out of a paged cache [nblk,bs,Hkv,D]."""
    bs = cache.shape[1]
    nb = (length + bs - 1) // bs
    ids = block_table_row[:nb].to(torch.long)
    return cache[ids].reshape(nb * bs, cache.shape[2], cache.shape[3])[:length]


# --------------------------------------------------------------------------
# K2/K3: prefill                   reference layers/attention.py:64-70
# --------------------------------------------------------------------------
def varlen_prefill_ref(q, k, v, cu_seqlens_q, cu_seqlens_k, scale,
                       block_tables=None, k_cache=None, v_cache=None, p_dtype=None):
    """Causal varlen attention.

    Un-paged (block_tables is None): k, v are [T, Hkv, D] packed like q.
    Paged (prefix-cache hit / chunked prefill, attention.py:65-66): keys and
    values of sequence s are rows 0..len_k(s)-1 gathered through
    block_tables[s] out of k_cache / v_cache (which already hold this step's
    new tokens: store_kvcache ran first, attention.py:62-63).
    """
    out = torch.empty(q.shape, dtype=q.dtype)
    nseq = len(cu_seqlens_q) - 1
    for s in range(nseq):
        q0, q1 = int(cu_seqlens_q[s]), int(cu_seqlens_q[s + 1])
        k0, k1 = int(cu_seqlens_k[s]), int(cu_seqlens_k[s + 1])
        if q1 == q0:
            continue
        if block_tables is None:
            ks, vs = k[k0:k1], v[k0:k1]
        else:
            ks = gather_pages(k_cache, block_tables[s], k1 - k0)
            vs = gather_pages(v_cache, block_tables[s], k1 - k0)
        out[q0:q1] = _attend_one(q[q0:q1], ks, vs, scale, p_dtype).to(q.dtype)
    return out


# --------------------------------------------------------------------------
# K4: decode                       reference layers/attention.py:71-74
# --------------------------------------------------------------------------
def paged_decode_ref(q, k_cache, v_cache, context_lens, block_tables, scale, p_dtype=None):
    """q [B,Hq,D] (one new token per sequence) over keys 0..context_lens[b]-1.

    Rows with context_lens == 0 are CUDA-graph padding (model_runner.py:207);
    their output is defined as zeros here.
    """
    out = torch.zeros(q.shape, dtype=q.dtype)
    for b in range(q.shape[0]):
        n = int(context_lens[b])
        if n == 0:
            continue
        ks = gather_pages(k_cache, block_tables[b], n)
        vs = gather_pages(v_cache, block_tables[b], n)
        out[b] = _attend_one(q[b:b + 1], ks, vs, scale, p_dtype)[0].to(q.dtype)
    return out


# --------------------------------------------------------------------------
# a1: the operator as the model sees it    reference layers/attention.py:59-75
# --------------------------------------------------------------------------
def attention_forward_ref(q, k, v, k_cache, v_cache, ctx, scale, p_dtype=None):
    """ctx: any object with the fields of reference utils/context.py:5-14."""
    if k_cache is not None and k_cache.numel():
        store_kvcache_ref(k, v, k_cache, v_cache, ctx.slot_mapping)
    if ctx.is_prefill:
        return varlen_prefill_ref(q, k, v, ctx.cu_seqlens_q, ctx.cu_seqlens_k, scale,
                                  ctx.block_tables, k_cache, v_cache, p_dtype)
    return paged_decode_ref(q, k_cache, v_cache, ctx.context_lens, ctx.block_tables, scale, p_dtype)
