"""The overlapped engine loop (host work of step N+1 under the GPU time of step N) must make exactly the
decisions of the synchronous loop, i.e. of the reference: replay the golden traces through
LLMEngine._run_overlapped with a fake runner that 'samples' the deterministic fake tokens."""
import itertools
import json
import os
import types

import pytest

from oracle.make_golden import digest, fake_token, step_record, workloads

import numpy as np


class FakeRunner:
    def __init__(self, block_size, vocab):
        from test_bookkeeping_golden import product_meta_builder
        self.build = product_meta_builder(block_size)
        self.vocab = vocab
        self.steps = []
        self.queue = []
        self.staged = 0
        self.early = 0

    def call(self, name, *a):
        return getattr(self, name)(*a)

    def stage_decode(self, seqs):
        self.staged += 1
        return 0

    def launch(self, seqs, is_prefill, staged=None, src=None):
        # every token value a step consumes must be known at launch, or named by `src` as a row of the step
        # still in flight (the device-side gather); emulate the gather to record what the model would see
        meta = self.build(seqs, is_prefill)
        ids = meta["input_ids"].copy()
        if src is not None:
            assert not is_prefill and len(self.queue) == 1, "early launch only for a decode step behind one running step"
            prev = self.queue[0]
            for i, r in enumerate(src):
                if r >= 0:
                    assert ids[i] == -1
                    ids[i] = prev[r]
            self.early += 1
        assert (ids >= 0).all(), "placeholder token reached the model input"
        meta["input_ids"] = ids
        self.steps.append(step_record(seqs, is_prefill, meta))
        # the fake model's token depends on the sequence length only, which is already final
        self.queue.append([fake_token(s.seq_id, len(s), self.vocab) for s in seqs])
        assert len(self.queue) <= 2

    def collect(self):
        return self.queue.pop(0)


@pytest.mark.parametrize("name", ["prefix16", "chunked32", "eos64", "bench_tight", "mixed1024", "longctx128"])
def test_overlapped_loop_equals_reference_trace(name, golden_dir):
    from nanovllm.engine.llm_engine import LLMEngine
    from nanovllm.engine.scheduler import Scheduler
    from nanovllm.engine.sequence import Sequence
    from nanovllm.sampling_params import SamplingParams
    gold = json.load(open(os.path.join(golden_dir, f"trace_{name}.json")))
    w = workloads()[name]
    cfg = types.SimpleNamespace(eos=w["eos"], **w["cfg"])
    Sequence.block_size = cfg.kvcache_block_size
    Sequence.counter = itertools.count()
    eng = object.__new__(LLMEngine)
    eng.scheduler = Scheduler(cfg)
    eng.model_runner = FakeRunner(cfg.kvcache_block_size, w["vocab"])
    for p, (t, mt, ie) in zip(w["prompts"], w["sps"]):
        eng.scheduler.add(Sequence(p, SamplingParams(temperature=t, max_tokens=mt, ignore_eos=ie)))
    outputs = {}

    def on_step(finished, num_tokens, dt):
        for s in finished:
            assert -1 not in s.completion_token_ids
            outputs[s.seq_id] = list(s.completion_token_ids)

    eng._run_overlapped(on_step)
    got = eng.model_runner.steps
    assert len(got) == gold["num_steps"]
    for i, (a, b) in enumerate(zip(got, gold["steps"])):
        assert a == b, f"step {i}: overlapped {a} != reference {b}"
    assert digest(*[np.asarray(outputs[k]) for k in sorted(outputs)]) == gold["outputs"]
    if w["sps"][0][2]:          # ignore_eos workloads enqueue decode steps early
        assert eng.model_runner.early > 0
    else:
        assert eng.model_runner.early == 0


@pytest.mark.parametrize("seed", [3])
def test_overlapped_loop_equals_synchronous_loop_on_random_workloads(seed):
    """The same property on 120 random workloads (oracle.make_golden.fuzz_workloads: mixed ignore_eos / EOS batches,
    preemption, chunking, prefix hits): the overlapped loop must issue exactly the steps of the synchronous
    schedule -> run -> postprocess loop.  Needs no reference: the synchronous loop is pinned to it elsewhere."""
    from oracle.make_golden import drive, fuzz_workloads
    from nanovllm.engine.llm_engine import LLMEngine
    from nanovllm.engine.scheduler import Scheduler
    from nanovllm.engine.sequence import Sequence
    from nanovllm.sampling_params import SamplingParams
    from test_bookkeeping_golden import product_meta_builder
    early = 0
    for i, w in enumerate(fuzz_workloads(seed, 120)):
        cfg = types.SimpleNamespace(eos=w["eos"], **w["cfg"])
        make = lambda p, t, mt, ie: Sequence(p, SamplingParams(temperature=t, max_tokens=mt, ignore_eos=ie))
        Sequence.block_size = cfg.kvcache_block_size
        Sequence.counter = itertools.count()
        sync = drive(make, Scheduler(cfg), cfg.kvcache_block_size, product_meta_builder(cfg.kvcache_block_size), w)

        Sequence.counter = itertools.count()
        eng = object.__new__(LLMEngine)
        eng.scheduler = Scheduler(cfg)
        eng.model_runner = FakeRunner(cfg.kvcache_block_size, w["vocab"])
        for p, (t, mt, ie) in zip(w["prompts"], w["sps"]):
            eng.scheduler.add(make(p, t, mt, ie))
        outputs = {}

        def on_step(finished, num_tokens, dt, outputs=outputs):
            for s in finished:
                assert -1 not in s.completion_token_ids
                outputs[s.seq_id] = list(s.completion_token_ids)

        eng._run_overlapped(on_step)
        got = eng.model_runner.steps
        assert len(got) == sync["num_steps"], f"workload {i}: {w['cfg']}"
        for k, (a, b) in enumerate(zip(got, sync["steps"])):
            assert a == b, f"workload {i} ({w['cfg']}), step {k}: overlapped {a} != synchronous {b}"
        assert digest(*[np.asarray(outputs[k]) for k in sorted(outputs)]) == sync["outputs"]
        early += eng.model_runner.early
    assert early > 500          # the early-launch path was exercised, not just the synchronous fallback
