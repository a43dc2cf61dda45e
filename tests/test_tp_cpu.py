"""Tensor-parallel host logic on CPU with gloo (world_size 2): weight shard slicing and the row-parallel
all-reduce reproduce the unsharded layer; SPMD scheduler replicas stay in lock-step."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path[:0] = [ROOT, os.path.join(ROOT, "nano-vllm_b200")]
    from types import SimpleNamespace
    from nanovllm.models.qwen3 import Qwen3ForCausalLM
    from nanovllm.utils.synthetic import PRESETS, hf_config_dict, random_weights
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", world_size=world, rank=rank)
    preset = "tiny-g4"
    dims = PRESETS[preset]
    w = random_weights(dims, seed=5)
    hf = SimpleNamespace(**hf_config_dict(dims))
    m = Qwen3ForCausalLM(hf, rank, world, device="cpu", max_position=64)
    for name, t in w.items():
        m.load_hf_tensor(name, t)
    L, p = m.layers[1], "model.layers.1."
    D, hq, hkv, I = 128, dims["num_attention_heads"], dims["num_key_value_heads"], dims["intermediate_size"]
    ok = True
    # column-parallel q/k/v and gate/up, row-parallel o/down (reference linear.py:65-70,87-93,114-128,142-150)
    ok &= torch.equal(L.qkv[:m.q_size], w[p + "self_attn.q_proj.weight"][rank * m.q_size:(rank + 1) * m.q_size])
    ok &= torch.equal(L.qkv[m.q_size:m.q_size + m.kv_size], w[p + "self_attn.k_proj.weight"][rank * m.kv_size:(rank + 1) * m.kv_size])
    ok &= torch.equal(L.qkv[m.q_size + m.kv_size:], w[p + "self_attn.v_proj.weight"][rank * m.kv_size:(rank + 1) * m.kv_size])
    ok &= torch.equal(L.o, w[p + "self_attn.o_proj.weight"][:, rank * m.q_size:(rank + 1) * m.q_size])
    ok &= torch.equal(L.gate_up[:m.inter], w[p + "mlp.gate_proj.weight"][rank * m.inter:(rank + 1) * m.inter])
    ok &= torch.equal(L.gate_up[m.inter:], w[p + "mlp.up_proj.weight"][rank * m.inter:(rank + 1) * m.inter])
    ok &= torch.equal(L.down, w[p + "mlp.down_proj.weight"][:, rank * m.inter:(rank + 1) * m.inter])
    ok &= torch.equal(m.embed, w["model.embed_tokens.weight"])                                  # replicated
    ok &= torch.equal(m.lm_head, w["lm_head.weight"][rank * m.vocab_shard:(rank + 1) * m.vocab_shard])
    # sharded MLP + all-reduce == full MLP (fp32 to compare exactly up to summation order)
    x = torch.randn(5, dims["hidden_size"], generator=torch.Generator().manual_seed(1))
    gu = F.linear(x, L.gate_up.float())
    part = F.linear(F.silu(gu[:, :m.inter]) * gu[:, m.inter:], L.down.float())
    dist.all_reduce(part)
    full_gu = torch.cat([w[p + "mlp.gate_proj.weight"], w[p + "mlp.up_proj.weight"]]).float()
    g = F.linear(x, full_gu)
    full = F.linear(F.silu(g[:, :I]) * g[:, I:], w[p + "mlp.down_proj.weight"].float())
    ok &= torch.allclose(part, full, rtol=1e-4, atol=1e-5)
    # sampled-token agreement: all-reduce(MAX) over packed keys picks one winner on every rank
    keys = torch.tensor([(3 << 32) | (0xffffffff - (100 + rank)), (7 + rank) << 32 | 5], dtype=torch.int64)
    dist.all_reduce(keys, op=dist.ReduceOp.MAX)
    ok &= (0xffffffff - (keys & 0xffffffff)).tolist() == [100, 0xffffffff - 5]
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_tp2_sharding_and_allreduce_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 300
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]


def test_spmd_scheduler_replicas_agree():
    """Two scheduler replicas fed the same requests and the same sampled tokens make identical decisions
    (what lets every TP rank run its own copy instead of receiving the batch over RPC)."""
    import itertools
    import types
    from oracle.make_golden import drive, workloads
    from nanovllm.engine.scheduler import Scheduler
    from nanovllm.engine.sequence import Sequence
    from nanovllm.sampling_params import SamplingParams
    from test_bookkeeping_golden import product_meta_builder
    w = workloads()["eos64"]
    runs = []
    for _ in range(2):
        cfg = types.SimpleNamespace(eos=w["eos"], **w["cfg"])
        Sequence.block_size = cfg.kvcache_block_size
        Sequence.counter = itertools.count()
        make = lambda p, t, mt, ie: Sequence(p, SamplingParams(temperature=t, max_tokens=mt, ignore_eos=ie))
        runs.append(drive(make, Scheduler(cfg), cfg.kvcache_block_size, product_meta_builder(cfg.kvcache_block_size), w))
    assert runs[0] == runs[1]
