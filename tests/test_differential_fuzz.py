"""Differential fuzzing of the host bookkeeping against the reference's own classes, where the reference is available.

The committed golden traces (tests/golden/trace_*.json) pin seven hand-picked workloads.  In the build container the
reference can also be executed directly, so this test draws a few hundred random small workloads (page sizes 2..256,
tight caches that force preemption, shared prefixes, duplicates, EOS stops, tiny token budgets that force chunking),
runs the REFERENCE's Scheduler / BlockManager / Sequence / prepare_* over them in a subprocess (both packages are called
`nanovllm`, so they cannot share an interpreter) and requires the product to produce the same step-by-step digests.
Skipped where /root/reference does not exist (e.g. on the GPU box) -- the golden traces cover that case.
"""
import itertools
import json
import os
import subprocess
import sys
import types

import pytest

from oracle.make_golden import REF, drive, fuzz_workloads

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "nanovllm")), reason="reference tree not present on this machine")
@pytest.mark.parametrize("seed", [1, 2])
def test_random_workloads_match_reference(seed, tmp_path):
    from nanovllm.engine.scheduler import Scheduler
    from nanovllm.engine.sequence import Sequence
    from nanovllm.sampling_params import SamplingParams
    from test_bookkeeping_golden import product_meta_builder
    n = 150
    out = tmp_path / "ref.json"
    env = dict(os.environ, TORCH_COMPILE_DISABLE="1", PYTHONPATH="")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "oracle", "make_golden.py"), "--fuzz", str(seed), str(n), str(out)],
                       env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-3000:]
    ref = json.load(open(out))
    assert len(ref) == n
    preempting = prefix_hits = 0
    for i, (w, gold) in enumerate(zip(fuzz_workloads(seed, n), ref)):
        cfg = types.SimpleNamespace(eos=w["eos"], **w["cfg"])
        Sequence.block_size = cfg.kvcache_block_size
        Sequence.counter = itertools.count()
        make = lambda p, t, mt, ie: Sequence(p, SamplingParams(temperature=t, max_tokens=mt, ignore_eos=ie))
        sched = Scheduler(cfg)
        count = [0]
        orig = sched.preempt

        def counting(seq, orig=orig, count=count):
            count[0] += 1
            return orig(seq)
        sched.preempt = counting
        hits = [0]
        orig_alloc = sched.block_manager.allocate

        def counting_alloc(seq, num_cached, orig_alloc=orig_alloc, hits=hits):
            hits[0] += num_cached > 0
            return orig_alloc(seq, num_cached)
        sched.block_manager.allocate = counting_alloc
        got = drive(make, sched, cfg.kvcache_block_size, product_meta_builder(cfg.kvcache_block_size), w)
        assert got["num_steps"] == gold["num_steps"], f"workload {i}: {w['cfg']}"
        for k, (a, b) in enumerate(zip(got["steps"], gold["steps"])):
            assert a == b, f"workload {i} ({w['cfg']}), step {k}: product {a} != reference {b}"
        for key in ("final_state", "outputs", "num_prefill_steps", "sum_decode_batch"):
            assert got[key] == gold[key], f"workload {i}: {key}"
        preempting += count[0] > 0
        prefix_hits += hits[0] > 0
    # the fuzzer must actually reach the interesting branches (measured: ~50 preempting, ~130 prefix-hit workloads of 150)
    assert preempting >= 25 and prefix_hits >= 60, (preempting, prefix_hits)
