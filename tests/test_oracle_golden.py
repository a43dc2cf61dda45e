"""Pin the oracle: oracle/qwen3_ref.py ("eager" rounding) must reproduce, bit for bit, the logits the
REFERENCE's own nn.Modules produced on CPU for the fixed serving script (tests/golden/model_*.npz)."""
import os

import numpy as np
import pytest
import torch

from oracle.model_script import make_script, run_script, script_steps
from oracle.paged_attention_ref import (attention_forward_ref, paged_decode_ref, store_kvcache_ref, to_logical,
                                        to_physical, varlen_prefill_ref)
from oracle.qwen3_ref import Qwen3Ref, RefDims, alloc_logical_kv

from nanovllm.utils.synthetic import PRESETS, hf_config_dict, random_weights


def oracle_step_fn(model, kv):
    from types import SimpleNamespace

    def step(ids, pos, c):
        ctx = SimpleNamespace(is_prefill=c["is_prefill"], cu_seqlens_q=c.get("cu_seqlens_q"), cu_seqlens_k=c.get("cu_seqlens_k"),
                              max_seqlen_q=c.get("max_seqlen_q", 0), max_seqlen_k=c.get("max_seqlen_k", 0),
                              slot_mapping=c.get("slot_mapping"), context_lens=c.get("context_lens"),
                              block_tables=c.get("block_tables"))
        return model.logits(model.forward(ids, pos, ctx, kv), ctx)
    return step


@pytest.mark.parametrize("preset", ["tiny", "tiny-g4", "tiny-g1", "tiny-g8"])
def test_qwen3_ref_equals_reference_modules(preset, golden_dir):
    gold = np.load(os.path.join(golden_dir, f"model_{preset}.npz"))
    dims = RefDims.from_json(hf_config_dict(PRESETS[preset]))
    model = Qwen3Ref(dims, random_weights(PRESETS[preset], seed=1234), rounding="eager", max_pos=4096)
    script = make_script(dims.vocab_size)
    kv = alloc_logical_kv(dims, script["num_blocks"], script["block_size"])
    outs = run_script(torch, script, oracle_step_fn(model, kv))
    assert len(outs) == len(gold.files)
    for i, o in enumerate(outs):
        want = torch.from_numpy(gold[f"logits_{i}"]).view(torch.bfloat16)
        assert o.dtype == torch.bfloat16 and o.shape == want.shape
        assert torch.equal(o.view(torch.int16), want.view(torch.int16)), f"step {i} differs from the reference modules"


def test_fused_rounding_is_close_to_eager():
    """The GPU rounding model differs from eager by at most a few bf16 ulps at the logits."""
    preset = "tiny"
    dims = RefDims.from_json(hf_config_dict(PRESETS[preset]))
    w = random_weights(PRESETS[preset], seed=1234)
    script = make_script(dims.vocab_size)
    res = []
    for mode in ("eager", "fused"):
        m = Qwen3Ref(dims, w, rounding=mode, max_pos=4096)
        kv = alloc_logical_kv(dims, script["num_blocks"], script["block_size"])
        res.append(torch.cat([o.float() for o in run_script(torch, script, oracle_step_fn(m, kv))]))
    rel = (res[0] - res[1]).norm() / res[0].norm()
    assert rel < 2e-2


def _rand(*shape):
    return torch.randn(*shape, dtype=torch.float32).to(torch.bfloat16)


def test_attention_oracle_paged_equals_packed():
    """Same keys through pages or packed rows give identical results; decode == 1-token prefill."""
    torch.manual_seed(0)
    hq, hkv, d, bs = 8, 2, 128, 16
    lens = [5, 16, 37]
    tot = sum(lens)
    q, k, v = _rand(tot, hq, d), _rand(tot, hkv, d), _rand(tot, hkv, d)
    cu = torch.tensor([0, 5, 21, 58], dtype=torch.int32)
    packed = varlen_prefill_ref(q, k, v, cu, cu, d ** -0.5)
    kc, vc = torch.zeros(12, bs, hkv, d, dtype=torch.bfloat16), torch.zeros(12, bs, hkv, d, dtype=torch.bfloat16)
    tables = torch.tensor([[7, -1, -1], [2, -1, -1], [9, 0, 4]], dtype=torch.int32)
    slots = []
    for s, n in enumerate(lens):
        slots += [int(tables[s, p // bs]) * bs + p % bs for p in range(n)]
    store_kvcache_ref(k, v, kc, vc, torch.tensor(slots, dtype=torch.int32))
    paged = varlen_prefill_ref(q, None, None, cu, cu, d ** -0.5, tables, kc, vc)
    assert torch.equal(packed, paged)
    last = (cu[1:] - 1).long()
    dec = paged_decode_ref(q[last], kc, vc, torch.tensor(lens, dtype=torch.int32), tables, d ** -0.5)
    assert torch.equal(dec, packed[last])
    assert torch.equal(to_logical(to_physical(kc)), kc)


def test_store_skips_minus_one():
    k, v = _rand(4, 2, 128), _rand(4, 2, 128)
    kc, vc = torch.zeros(2, 16, 2, 128, dtype=torch.bfloat16), torch.zeros(2, 16, 2, 128, dtype=torch.bfloat16)
    store_kvcache_ref(k, v, kc, vc, torch.tensor([3, -1, 17, -1], dtype=torch.int32))
    assert torch.equal(kc[0, 3], k[0]) and torch.equal(kc[1, 1], k[2])
    assert torch.equal(vc[0, 3], v[0]) and torch.equal(vc[1, 1], v[2])
    touched = (kc.float().abs().sum(dim=(2, 3)) > 0).sum().item()
    assert touched == 2


def test_script_covers_all_branches():
    steps = script_steps(make_script(2048))
    kinds = [(s["is_prefill"], s["block_tables"] is not None) for s in steps]
    assert (True, False) in kinds and (True, True) in kinds and (False, True) in kinds
    assert any(s["is_prefill"] and s["cu_seqlens_q"][-1] < s["cu_seqlens_k"][-1] for s in steps)
