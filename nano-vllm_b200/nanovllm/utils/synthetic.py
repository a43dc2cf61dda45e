"""Random-init Qwen3 checkpoints for benchmarking and tests (no network, no real weights here).

Writes what ``Config`` / ``LLMEngine`` / the loader read from a model directory (the same three
things the reference reads): ``config.json``, a WordLevel ``tokenizer.json``
covering the whole vocabulary, and ``model.safetensors`` with HF parameter names.  Weights are
N(0, init_std) from a seeded CPU generator, norm weights 1 + 0.1 N(0, 1); the same file is loaded
by the product, by the CPU oracle and (when present) by the reference itself, so all three see
identical bytes.
"""
from __future__ import annotations

import json
import os

import torch

PRESETS = {
    # public HF configs (dims only)
    "qwen3-0.6b": dict(hidden_size=1024, num_hidden_layers=28, num_attention_heads=16, num_key_value_heads=8,
                       head_dim=128, intermediate_size=3072, vocab_size=151936, tie_word_embeddings=True),
    "qwen3-8b": dict(hidden_size=4096, num_hidden_layers=36, num_attention_heads=32, num_key_value_heads=8,
                     head_dim=128, intermediate_size=12288, vocab_size=151936, tie_word_embeddings=False),
    "qwen3-32b": dict(hidden_size=5120, num_hidden_layers=64, num_attention_heads=64, num_key_value_heads=8,
                      head_dim=128, intermediate_size=25600, vocab_size=151936, tie_word_embeddings=False),
    # small shapes for parity tests
    "tiny": dict(hidden_size=256, num_hidden_layers=2, num_attention_heads=4, num_key_value_heads=2,
                 head_dim=128, intermediate_size=512, vocab_size=2048, tie_word_embeddings=True),
    "tiny-g4": dict(hidden_size=256, num_hidden_layers=3, num_attention_heads=8, num_key_value_heads=2,
                    head_dim=128, intermediate_size=768, vocab_size=4096, tie_word_embeddings=False),
    "tiny-g1": dict(hidden_size=192, num_hidden_layers=2, num_attention_heads=2, num_key_value_heads=2,
                    head_dim=128, intermediate_size=320, vocab_size=1024, tie_word_embeddings=True),
    "tiny-g8": dict(hidden_size=320, num_hidden_layers=2, num_attention_heads=8, num_key_value_heads=1,
                    head_dim=128, intermediate_size=448, vocab_size=3072, tie_word_embeddings=False),
}


def hf_config_dict(dims: dict) -> dict:
    cfg = dict(architectures=["Qwen3ForCausalLM"], model_type="qwen3", max_position_embeddings=40960,
               rms_norm_eps=1e-6, rope_theta=1000000, hidden_act="silu", attention_bias=False,
               torch_dtype="bfloat16", bos_token_id=0, eos_token_id=1)
    cfg.update(dims)
    return cfg


def weight_shapes(d: dict) -> dict[str, tuple]:
    H, D = d["hidden_size"], d["head_dim"]
    hq, hkv, I, V = d["num_attention_heads"], d["num_key_value_heads"], d["intermediate_size"], d["vocab_size"]
    shapes = {"model.embed_tokens.weight": (V, H), "model.norm.weight": (H,)}
    if not d.get("tie_word_embeddings", False):
        shapes["lm_head.weight"] = (V, H)
    for i in range(d["num_hidden_layers"]):
        p = f"model.layers.{i}."
        shapes.update({
            p + "self_attn.q_proj.weight": (hq * D, H), p + "self_attn.k_proj.weight": (hkv * D, H),
            p + "self_attn.v_proj.weight": (hkv * D, H), p + "self_attn.o_proj.weight": (H, hq * D),
            p + "self_attn.q_norm.weight": (D,), p + "self_attn.k_norm.weight": (D,),
            p + "mlp.gate_proj.weight": (I, H), p + "mlp.up_proj.weight": (I, H), p + "mlp.down_proj.weight": (H, I),
            p + "input_layernorm.weight": (H,), p + "post_attention_layernorm.weight": (H,)})
    return shapes


def random_weights(dims: dict, seed: int = 0, init_std: float = 0.02) -> dict[str, torch.Tensor]:
    gen = torch.Generator(device="cpu").manual_seed(seed)
    out = {}
    for name, shape in weight_shapes(dims).items():
        if len(shape) == 1:
            w = 1.0 + 0.1 * torch.randn(shape, generator=gen)
        else:
            w = init_std * torch.randn(shape, generator=gen)
        out[name] = w.to(torch.bfloat16)
    return out


def write_tokenizer(path: str, vocab_size: int, eos_id: int = 1) -> None:
    vocab = {f"t{i}": i for i in range(vocab_size)}
    vocab[f"t{eos_id}"] = eos_id
    tok = {
        "version": "1.0", "truncation": None, "padding": None, "added_tokens": [], "normalizer": None,
        "pre_tokenizer": {"type": "Whitespace"}, "post_processor": None, "decoder": None,
        "model": {"type": "WordLevel", "vocab": vocab, "unk_token": "t0"},
    }
    with open(os.path.join(path, "tokenizer.json"), "w") as f:
        json.dump(tok, f)
    with open(os.path.join(path, "tokenizer_config.json"), "w") as f:
        json.dump({"tokenizer_class": "PreTrainedTokenizerFast", "eos_token": f"t{eos_id}", "unk_token": "t0"}, f)


def make_model_dir(path: str, preset: str | dict = "qwen3-0.6b", seed: int = 0, init_std: float = 0.02,
                   weights: bool = True, tokenizer: bool = True) -> str:
    """Create (idempotently) a loadable model directory; returns ``path``."""
    dims = dict(PRESETS[preset]) if isinstance(preset, str) else dict(preset)
    os.makedirs(path, exist_ok=True)
    stamp = os.path.join(path, ".synthetic.json")
    want = dict(dims=dims, seed=seed, init_std=init_std, weights=weights, tokenizer=tokenizer, v=2)
    if os.path.exists(stamp):
        with open(stamp) as f:
            if json.load(f) == want:
                return path
    with open(os.path.join(path, "config.json"), "w") as f:
        json.dump(hf_config_dict(dims), f, indent=1)
    if tokenizer:
        write_tokenizer(path, dims["vocab_size"])
    if weights:
        from safetensors.torch import save_file
        save_file(random_weights(dims, seed, init_std), os.path.join(path, "model.safetensors"))
    with open(stamp, "w") as f:
        json.dump(want, f)
    return path
