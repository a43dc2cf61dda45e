"""CPU-side checks of the host plumbing: synthetic checkpoints, Config / LLM keyword handling, loader sharding."""
import json
import os
from types import SimpleNamespace

import pytest
import torch


def test_synthetic_model_dir_roundtrip(tmp_path):
    from safetensors import safe_open
    from nanovllm.config import Config
    from nanovllm.utils.synthetic import PRESETS, make_model_dir, random_weights, weight_shapes
    d = make_model_dir(str(tmp_path / "tiny"), "tiny", seed=3)
    assert make_model_dir(d, "tiny", seed=3) == d                      # idempotent
    cfg = Config(d, kvcache_block_size=16, max_model_len=100000)
    assert cfg.hf_config.num_hidden_layers == PRESETS["tiny"]["num_hidden_layers"]
    assert cfg.max_model_len == cfg.hf_config.max_position_embeddings   # clamped like reference config.py:25
    want = random_weights(PRESETS["tiny"], seed=3)
    with safe_open(os.path.join(d, "model.safetensors"), "pt", "cpu") as f:
        assert set(f.keys()) == set(weight_shapes(PRESETS["tiny"]))
        for k in f.keys():
            assert torch.equal(f.get_tensor(k), want[k])
    from transformers import AutoTokenizer
    tok = AutoTokenizer.from_pretrained(d, use_fast=True)
    assert tok.encode("t5 t6 t2047") == [5, 6, 2047] and tok.eos_token_id == 1
    assert tok.decode([7, 8]) == "t7 t8"


def test_loader_fills_model_on_cpu(tmp_path):
    from nanovllm.models.qwen3 import Qwen3ForCausalLM
    from nanovllm.utils.loader import load_model
    from nanovllm.utils.synthetic import PRESETS, hf_config_dict, make_model_dir, random_weights
    d = make_model_dir(str(tmp_path / "g4"), "tiny-g4", seed=9, tokenizer=False)
    hf = SimpleNamespace(**hf_config_dict(PRESETS["tiny-g4"]))
    m = Qwen3ForCausalLM(hf, 0, 1, device="cpu", max_position=32)
    n = load_model(m, d)
    w = random_weights(PRESETS["tiny-g4"], seed=9)
    assert n == len(w)
    assert torch.equal(m.layers[2].qkv[:m.q_size], w["model.layers.2.self_attn.q_proj.weight"])
    assert torch.equal(m.lm_head, w["lm_head.weight"]) and torch.equal(m.norm, w["model.norm.weight"])
    empty = str(tmp_path / "empty")
    os.makedirs(empty)
    with pytest.raises(FileNotFoundError):
        load_model(m, empty)                                            # the reference would silently run on garbage
    assert load_model(m, empty, allow_random=True, seed=9) == len(w)
    assert torch.equal(m.layers[0].down, w["model.layers.0.mlp.down_proj.weight"])


def test_llm_ignores_unknown_keywords_and_needs_a_gpu(tmp_path):
    """LLM(**kwargs) keeps only Config fields (reference llm_engine.py:18-20); without CUDA it refuses to start."""
    from nanovllm import LLM
    from nanovllm._native import B200Error
    from nanovllm.utils.synthetic import make_model_dir
    d = make_model_dir(str(tmp_path / "tiny"), "tiny", seed=3)
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(B200Error):
        LLM(d, enforce_eager=True, not_a_config_field=123, kvcache_block_size=16)


def test_public_api_matches_reference_surface(golden_dir):
    """tests/golden/api_surface.json is the reference's public surface (oracle/make_golden.py --api, by introspection of
    the reference package): exports, Config / SamplingParams fields with defaults, the engine's public methods with
    parameter names and defaults.  The product must offer all of it unchanged; it may add optional extras at the end."""
    import json
    import os
    import nanovllm
    from oracle.make_golden import api_surface
    ref = json.load(open(os.path.join(golden_dir, "api_surface.json")))
    got = api_surface(nanovllm)
    assert got["exports"] == ref["exports"] and got["llm_is_engine"] == ref["llm_is_engine"]
    for section in ("config", "sampling_params"):
        mine = {name: default for name, default in got[section]}
        for name, default in ref[section]:
            assert name in mine, f"{section}.{name} missing"
            if (section, name) == ("config", "num_kvcache_blocks"):
                assert mine[name] in ("-1", "None")          # both mean "size the cache from free memory"
            else:
                assert mine[name] == default, f"{section}.{name}: default {mine[name]} vs reference {default}"
        # field ORDER matters for positional construction
        names = [n for n, _ in got[section]]
        assert names[:len(ref[section])] == [n for n, _ in ref[section]], section
    for method, params in ref["engine_methods"].items():
        mine = got["engine_methods"][method]
        assert mine[:len(params)] == params, f"{method}: {mine} vs reference {params}"
        for extra in mine[len(params):]:
            assert extra[1] != "<required>" or extra[2] == "VAR_KEYWORD", f"{method}: extra required parameter {extra}"
