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
        self.pending = None
        self.staged = 0

    def call(self, name, *a):
        return getattr(self, name)(*a)

    def stage_decode(self, seqs):
        self.staged += 1
        return 0

    def launch(self, seqs, is_prefill, staged=None):
        # at launch every token value a step consumes must be known (no placeholder left in its inputs)
        meta = self.build(seqs, is_prefill)
        assert (meta["input_ids"] >= 0).all(), "placeholder token reached the model input"
        self.steps.append(step_record(seqs, is_prefill, meta))
        self.pending = [fake_token(s.seq_id, len(s), self.vocab) for s in seqs]

    def collect(self):
        return self.pending


@pytest.mark.parametrize("name", ["prefix16", "chunked32", "eos64", "bench_tight"])
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
    if w["sps"][0][2]:          # ignore_eos workloads take the early-staging path
        assert eng.model_runner.staged > 0
    else:
        assert eng.model_runner.staged == 0
