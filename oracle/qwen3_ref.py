"""CPU restatement of the reference's Qwen3 forward around the attention operator.
TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows:
* Qwen3Attention.forward          reference nanovllm/models/qwen3.py:72-88
* Qwen3MLP / Qwen3DecoderLayer    reference nanovllm/models/qwen3.py:91-159
* Qwen3Model / ForCausalLM        reference nanovllm/models/qwen3.py:162-216
* RMSNorm (plain and fused-add)   reference nanovllm/layers/layernorm.py:16-40
* NeoX RoPE + cos/sin cache       reference nanovllm/layers/rotary_embedding.py:6-48
* SiluAndMul                      reference nanovllm/layers/activation.py:8-11
* lm_head last-token selection    reference nanovllm/layers/embed_head.py:56-61
* greedy branch of the sampler    argmax over fp32 logits (the reference's
                                  Sampler forbids temperature 0, sampling_params.py:10-11;
                                  greedy == argmax of the logits it would have sampled from)

Two rounding models:
* ``rounding="eager"``  every ``.to(bf16)`` in the reference source is a real
  rounding — what the reference modules do on CPU with TORCH_COMPILE_DISABLE=1.
  ``tests/golden/model_*.npz`` (made by running the reference's own modules)
  pins this mode bit-exactly.
* ``rounding="fused"``  what the reference does on a GPU: each ``@torch.compile``
  site computes in fp32 and rounds once at its stored outputs.  The CUDA
  kernels are compared against this mode.

Weights are a dict of HF-named tensors (``model.layers.N.self_attn.q_proj.weight`` ...);
q/k/v and gate/up are concatenated exactly like the reference's packed
modules (qwen3.py:187-193) so the GEMMs see the same operands.
"""
from __future__ import annotations

from dataclasses import dataclass
from types import SimpleNamespace

import torch
import torch.nn.functional as F

from .paged_attention_ref import attention_forward_ref


@dataclass
class RefDims:
    hidden_size: int
    num_hidden_layers: int
    num_attention_heads: int
    num_key_value_heads: int
    head_dim: int
    intermediate_size: int
    vocab_size: int
    rms_norm_eps: float = 1e-6
    rope_theta: float = 1000000.0
    max_position_embeddings: int = 40960
    tie_word_embeddings: bool = True

    @classmethod
    def from_json(cls, cfg: dict) -> "RefDims":
        theta = cfg.get("rope_theta", 1000000.0)
        if isinstance(cfg.get("rope_scaling"), dict):
            theta = cfg["rope_scaling"].get("rope_theta", theta)
        return cls(cfg["hidden_size"], cfg["num_hidden_layers"], cfg["num_attention_heads"],
                   cfg["num_key_value_heads"], cfg.get("head_dim", cfg["hidden_size"] // cfg["num_attention_heads"]),
                   cfg["intermediate_size"], cfg["vocab_size"], cfg.get("rms_norm_eps", 1e-6), float(theta),
                   cfg.get("max_position_embeddings", 40960), cfg.get("tie_word_embeddings", False))


def rope_table(head_dim: int, max_pos: int, theta: float) -> torch.Tensor:
    """fp32 [max_pos, head_dim] = cat(cos, sin)   (rotary_embedding.py:29-35)."""
    inv_freq = 1.0 / (theta ** (torch.arange(0, head_dim, 2, dtype=torch.float) / head_dim))
    ang = torch.outer(torch.arange(max_pos, dtype=torch.float), inv_freq)
    return torch.cat((ang.cos(), ang.sin()), dim=-1)


# ---- the reference's elementwise ops as free functions (used by the class below and by kernel tests) ----
def rmsnorm_ref(x, weight, eps, rounding="fused"):
    """RMSNorm.rms_forward, layers/layernorm.py:16-26."""
    dt = x.dtype
    xf = x.float()
    var = xf.pow(2).mean(dim=-1, keepdim=True)
    xf = xf * torch.rsqrt(var + eps)
    if rounding == "eager":                              # layernorm.py:25: x.to(bf16).mul_(weight)
        return xf.to(dt).mul_(weight)
    return (xf * weight.float()).to(dt)


def add_rmsnorm_ref(x, residual, weight, eps, rounding="fused"):
    """RMSNorm.add_rms_forward, layers/layernorm.py:28-40 -> (normed, new_residual)."""
    dt = x.dtype
    xf = x.float() + residual.float()
    new_residual = xf.to(dt)
    var = xf.pow(2).mean(dim=-1, keepdim=True)
    xf = xf * torch.rsqrt(var + eps)
    if rounding == "eager":
        return xf.to(dt).mul_(weight), new_residual
    return (xf * weight.float()).to(dt), new_residual


def rope_ref(cos_sin, positions, x):
    """RotaryEmbedding.forward / apply_rotary_emb, layers/rotary_embedding.py:6-14,37-48.  x: [T, H, D]."""
    cs = cos_sin[positions].unsqueeze(1)
    cos, sin = cs.chunk(2, dim=-1)
    x1, x2 = torch.chunk(x.float(), 2, dim=-1)
    return torch.cat((x1 * cos - x2 * sin, x2 * cos + x1 * sin), dim=-1).to(x.dtype)


def silu_mul_ref(gate_up, rounding="fused"):
    """SiluAndMul.forward, layers/activation.py:8-11."""
    g, u = gate_up.chunk(2, -1)
    if rounding == "eager":
        return F.silu(g) * u
    return (F.silu(g.float()) * u.float()).to(gate_up.dtype)


class Qwen3Ref:
    def __init__(self, dims: RefDims, weights: dict, rounding: str = "fused", max_pos: int | None = None,
                 p_dtype=None):
        assert rounding in ("eager", "fused")
        self.d = dims
        self.rounding = rounding
        self.p_dtype = p_dtype
        w = weights
        self.embed = w["model.embed_tokens.weight"]
        self.head = self.embed if dims.tie_word_embeddings else w["lm_head.weight"]
        self.final_norm = w["model.norm.weight"]
        self.layers = []
        for i in range(dims.num_hidden_layers):
            p = f"model.layers.{i}."
            self.layers.append(SimpleNamespace(
                qkv=torch.cat([w[p + "self_attn.q_proj.weight"], w[p + "self_attn.k_proj.weight"],
                               w[p + "self_attn.v_proj.weight"]], 0),
                o=w[p + "self_attn.o_proj.weight"],
                q_norm=w[p + "self_attn.q_norm.weight"], k_norm=w[p + "self_attn.k_norm.weight"],
                gate_up=torch.cat([w[p + "mlp.gate_proj.weight"], w[p + "mlp.up_proj.weight"]], 0),
                down=w[p + "mlp.down_proj.weight"],
                ln1=w[p + "input_layernorm.weight"], ln2=w[p + "post_attention_layernorm.weight"]))
        self.cos_sin = rope_table(dims.head_dim, max_pos or dims.max_position_embeddings, dims.rope_theta)
        self.scale = dims.head_dim ** -0.5

    # ---- elementwise pieces ------------------------------------------------
    def rmsnorm(self, x, weight):
        return rmsnorm_ref(x, weight, self.d.rms_norm_eps, self.rounding)

    def add_rmsnorm(self, x, residual, weight):
        return add_rmsnorm_ref(x, residual, weight, self.d.rms_norm_eps, self.rounding)

    def rope(self, positions, x):
        return rope_ref(self.cos_sin, positions, x)

    def silu_mul(self, gate_up):
        return silu_mul_ref(gate_up, self.rounding)

    # ---- model -------------------------------------------------------------
    def forward(self, input_ids, positions, ctx, kv_caches):
        """kv_caches: list of (k_cache, v_cache) in the reference's logical layout, or None (warm-up)."""
        d = self.d
        h = F.embedding(input_ids, self.embed)
        residual = None
        for li, L in enumerate(self.layers):
            if residual is None:
                h, residual = self.rmsnorm(h, L.ln1), h
            else:
                h, residual = self.add_rmsnorm(h, residual, L.ln1)
            qkv = F.linear(h, L.qkv)
            qs, ks = d.num_attention_heads * d.head_dim, d.num_key_value_heads * d.head_dim
            q, k, v = qkv.split([qs, ks, ks], dim=-1)
            q = q.reshape(-1, d.num_attention_heads, d.head_dim)
            k = k.reshape(-1, d.num_key_value_heads, d.head_dim)
            v = v.reshape(-1, d.num_key_value_heads, d.head_dim)
            q = self.rope(positions, self.rmsnorm(q, L.q_norm))
            k = self.rope(positions, self.rmsnorm(k, L.k_norm))
            kc, vc = kv_caches[li] if kv_caches is not None else (None, None)
            o = attention_forward_ref(q, k, v, kc, vc, ctx, self.scale, self.p_dtype)
            h = F.linear(o.reshape(o.shape[0], -1), L.o)
            h, residual = self.add_rmsnorm(h, residual, L.ln2)
            h = F.linear(self.silu_mul(F.linear(h, L.gate_up)), L.down)
        h, _ = self.add_rmsnorm(h, residual, self.final_norm)
        return h

    def logits(self, hidden, ctx):
        if ctx.is_prefill:                               # embed_head.py:58-60
            last = (torch.as_tensor(ctx.cu_seqlens_q)[1:] - 1).to(torch.long)
            hidden = hidden[last].contiguous()
        return F.linear(hidden, self.head)

    @staticmethod
    def greedy(logits):
        return logits.float().argmax(dim=-1)


def alloc_logical_kv(dims: RefDims, num_blocks: int, block_size: int, dtype=torch.bfloat16):
    """Per-layer (k, v) caches in the reference layout (model_runner.py:115)."""
    shape = (num_blocks, block_size, dims.num_key_value_heads, dims.head_dim)
    return [(torch.zeros(shape, dtype=dtype), torch.zeros(shape, dtype=dtype))
            for _ in range(dims.num_hidden_layers)]


def model_weight_bytes(dims: RefDims, itemsize: int = 2) -> int:
    """Bytes of weights one decode step reads (tied head counted once)."""
    d = dims
    per_layer = (d.hidden_size * (d.num_attention_heads + 2 * d.num_key_value_heads) * d.head_dim
                 + d.num_attention_heads * d.head_dim * d.hidden_size
                 + 3 * d.hidden_size * d.intermediate_size
                 + 2 * d.hidden_size + 2 * d.head_dim)
    total = d.num_hidden_layers * per_layer + d.hidden_size + d.vocab_size * d.hidden_size
    if not d.tie_word_embeddings:
        total += d.vocab_size * d.hidden_size
    return total * itemsize

