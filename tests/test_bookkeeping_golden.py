"""Host integer work vs traces of the REFERENCE's own classes (tests/golden/trace_*.json, made by
oracle/make_golden.py).  Bit-exact: block ids, slot numbers, schedule order, hashes, metadata arrays."""
import itertools
import json
import os
import types

import numpy as np
import pytest

from oracle.make_golden import drive, workloads

from nanovllm.engine.block_manager import BlockManager
from nanovllm.engine.model_runner import ModelRunner
from nanovllm.engine.scheduler import Scheduler
from nanovllm.engine.sequence import Sequence
from nanovllm.sampling_params import SamplingParams


def test_hash_known_answers(golden_dir):
    cases = json.load(open(os.path.join(golden_dir, "hash_kat.json")))
    assert BlockManager.compute_hash(list(range(256))) == 5218229187174952702
    assert BlockManager.compute_hash(list(range(256)), 12345) == 3063373928522455270
    for c in cases:
        assert BlockManager.compute_hash(c["tokens"], c["prefix"]) == c["hash"]


def product_meta_builder(block_size):
    """The product runner's array builders, unbound (constructing a ModelRunner needs a GPU)."""
    stub = types.SimpleNamespace(block_size=block_size)
    stub.prepare_block_tables = lambda seqs: ModelRunner.prepare_block_tables(stub, seqs)

    def build(seqs, is_prefill):
        return ModelRunner.prefill_arrays(stub, seqs) if is_prefill else ModelRunner.decode_arrays(stub, seqs)
    return build


@pytest.mark.parametrize("name", ["prefix16", "chunked32", "eos64", "bench", "bench_tight", "mixed1024", "longctx128"])
def test_trace_matches_reference(name, golden_dir):
    gold = json.load(open(os.path.join(golden_dir, f"trace_{name}.json")))
    w = workloads()[name]
    cfg = types.SimpleNamespace(eos=w["eos"], **w["cfg"])
    Sequence.block_size = cfg.kvcache_block_size
    Sequence.counter = itertools.count()
    make = lambda p, t, mt, ie: Sequence(p, SamplingParams(temperature=t, max_tokens=mt, ignore_eos=ie))
    got = drive(make, Scheduler(cfg), cfg.kvcache_block_size, product_meta_builder(cfg.kvcache_block_size), w)
    assert got["num_steps"] == gold["num_steps"]
    for i, (a, b) in enumerate(zip(got["steps"], gold["steps"])):
        assert a == b, f"step {i}: product {a} != reference {b}"
    for k in ("final_state", "outputs", "num_prefill_steps", "sum_decode_batch"):
        assert got[k] == gold[k], k


def test_bench_workload_invariants(golden_dir):
    """The numbers BASELINE.md section 3 quotes for the benchmark request mix."""
    gold = json.load(open(os.path.join(golden_dir, "trace_bench.json")))
    w = workloads()["bench"]
    assert sum(len(p) for p in w["prompts"]) == 142827
    assert sum(mt for _, mt, _ in w["sps"]) == 133966
    assert gold["num_prefill_steps"] == 9 and gold["num_steps"] == 1032 and gold["sum_decode_batch"] == 133710
    assert [(s[1], s[2]) for s in gold["steps"][:9]] == [(31, 15705), (30, 16364), (28, 16114), (26, 16130), (30, 16314),
                                                          (30, 16167), (31, 16047), (29, 15817), (21, 14169)]


def test_worked_example_block4():
    """SURVEY.md section 4 worked example (block_size 4): prefix hit, decode slots, chunked prefill."""
    Sequence.block_size = 4
    Sequence.counter = itertools.count()
    bm = BlockManager(16, 4)
    build = product_meta_builder(4)
    A = Sequence(list(range(100, 110)))
    assert bm.can_allocate(A) == 0
    bm.allocate(A, 0)
    A.num_scheduled_tokens = 10
    m = build([A], True)
    assert A.block_table == [0, 1, 2] and m["slot_mapping"].tolist() == list(range(10)) and m["block_tables"] is None
    assert m["cu_seqlens_q"].tolist() == [0, 10] == m["cu_seqlens_k"].tolist()
    bm.hash_blocks(A); A.num_cached_tokens += 10; A.num_scheduled_tokens = 0
    B = Sequence(list(range(100, 108)) + [7, 8, 9])
    assert bm.can_allocate(B) == 2
    bm.allocate(B, 2)
    assert B.block_table == [0, 1, 3] and B.num_cached_tokens == 8
    B.num_scheduled_tokens = 3
    m = build([B], True)
    assert m["input_ids"].tolist() == [7, 8, 9] and m["positions"].tolist() == [8, 9, 10]
    assert m["cu_seqlens_q"].tolist() == [0, 3] and m["cu_seqlens_k"].tolist() == [0, 11]
    assert (m["max_seqlen_q"], m["max_seqlen_k"]) == (3, 11)
    assert m["slot_mapping"].tolist() == [12, 13, 14] and m["block_tables"].tolist() == [[0, 1, 3]]
    assert bm.blocks[0].ref_count == 2 and bm.blocks[1].ref_count == 2
    bm.hash_blocks(B); B.num_cached_tokens += 3; B.num_scheduled_tokens = 0
    A.append_token(555); B.append_token(666)
    for s in (A, B):
        assert bm.can_append(s)
        bm.may_append(s)
    m = build([A, B], False)
    assert m["input_ids"].tolist() == [555, 666] and m["positions"].tolist() == [10, 11]
    assert m["context_lens"].tolist() == [11, 12] and m["slot_mapping"].tolist() == [10, 15]
    assert m["block_tables"].tolist() == [[0, 1, 2], [0, 1, 3]]
    C = Sequence(list(range(200, 210)))
    bm.allocate(C, bm.can_allocate(C))
    assert C.block_table == [4, 5, 6]
    C.num_scheduled_tokens = 6
    m = build([C], True)
    assert m["slot_mapping"].tolist() == list(range(16, 22)) and m["cu_seqlens_k"].tolist() == [0, 6] and m["block_tables"] is None
    bm.hash_blocks(C); C.num_cached_tokens += 6
    C.num_scheduled_tokens = 4
    m = build([C], True)
    assert m["input_ids"].tolist() == [206, 207, 208, 209] and m["positions"].tolist() == [6, 7, 8, 9]
    assert m["slot_mapping"].tolist() == [22, 23, 24, 25]
    assert m["cu_seqlens_q"].tolist() == [0, 4] and m["cu_seqlens_k"].tolist() == [0, 10] and m["block_tables"].tolist() == [[4, 5, 6]]


def test_sequence_pickle_roundtrip():
    import pickle
    Sequence.block_size = 16
    s = Sequence([1, 2, 3, 4], SamplingParams(temperature=0.0, max_tokens=5))
    s.block_table = [3]
    t = pickle.loads(pickle.dumps(s))
    assert t.token_ids == [1, 2, 3, 4] and t.last_token == 4 and t.block_table == [3]
    s.is_prefill = False
    t = pickle.loads(pickle.dumps(s))
    assert t.token_ids == [] and t.last_token == 4 and t.num_tokens == 4


def test_sampling_params_and_config_contract():
    sp = SamplingParams()
    assert (sp.temperature, sp.max_tokens, sp.ignore_eos) == (1.0, 64, False)
    assert SamplingParams(temperature=0.0).greedy
    with pytest.raises(ValueError):
        SamplingParams(temperature=-1)
    from nanovllm.config import Config
    hf = types.SimpleNamespace(max_position_embeddings=40960)
    c = Config("unused", hf_config=hf)
    assert (c.max_num_batched_tokens, c.max_num_seqs, c.max_model_len, c.gpu_memory_utilization,
            c.tensor_parallel_size, c.enforce_eager, c.kvcache_block_size) == (16384, 512, 4096, 0.9, 1, False, 256)
    assert Config("unused", hf_config=hf, kvcache_block_size=16).kvcache_block_size == 16
    with pytest.raises(ValueError):
        Config("unused", hf_config=hf, kvcache_block_size=48)
