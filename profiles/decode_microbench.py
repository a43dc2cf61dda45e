"""Paged-decode kernel alone on the benchmark's decode shapes (no model needed).

    python profiles/decode_microbench.py [--layers 28] [--reps 5] [--steps 0,500,900,1020] [--ncu]

Context lengths are those of the benchmark schedule at the given decode steps (step 0: batch 256,
sum of contexts 143 083); every layer has its own KV (>= 0.5 GB each at step 0, far beyond L2).
Prints achieved GB/s = algorithmic bytes (K+V once, q, out, metadata) / CUDA-event time.
"""
import argparse
import itertools
import json
import os
import sys
from types import SimpleNamespace

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "nano-vllm_b200")]

import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402


def schedule_decode_steps(wanted, block_size=256, num_blocks=800):
    from nanovllm.engine.model_runner import ModelRunner
    from nanovllm.engine.scheduler import Scheduler
    from nanovllm.engine.sequence import Sequence
    from nanovllm.sampling_params import SamplingParams
    Sequence.block_size = block_size
    Sequence.counter = itertools.count()
    sched = Scheduler(SimpleNamespace(max_num_seqs=512, max_num_batched_tokens=16384, eos=-1,
                                      kvcache_block_size=block_size, num_kvcache_blocks=num_blocks))
    prompts, max_tokens = bench.bench_requests(0)
    for p, mt in zip(prompts, max_tokens):
        sched.add(Sequence(p, SamplingParams(temperature=0.6, max_tokens=mt, ignore_eos=True)))
    stub = SimpleNamespace(block_size=block_size)
    stub.prepare_block_tables = lambda s: ModelRunner.prepare_block_tables(stub, s)
    out, d = {}, 0
    while not sched.is_finished() and len(out) < len(wanted):
        seqs, is_prefill = sched.schedule()
        if not is_prefill:
            if d in wanted:
                out[d] = ModelRunner.decode_arrays(stub, seqs)
            d += 1
        sched.postprocess(seqs, [0] * len(seqs), is_prefill)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--layers", type=int, default=28)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--steps", default="0,500,900,1020")
    ap.add_argument("--hq", type=int, default=16)
    ap.add_argument("--hkv", type=int, default=8)
    ap.add_argument("--ncu", action="store_true", help="one launch per step inside a cudaProfiler window")
    ap.add_argument("--fused", action="store_true", help="time b200_paged_decode_fused (raw qkv in) against qknorm_rope_store + paged_decode")
    args = ap.parse_args()
    from nanovllm import ops
    steps = [int(x) for x in args.steps.split(",")]
    metas = schedule_decode_steps(set(steps))
    nblk, bs, L = 800, 256, args.layers
    kv = torch.empty(ops.kv_cache_shape(L, nblk, args.hkv, bs, 128), dtype=torch.bfloat16, device="cuda")
    kv.normal_()
    ops.bind_kv_cache(kv)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
    peaks = bench.load_peaks()
    results = []
    for st in steps:
        a = metas[st]
        n = len(a["context_lens"])
        ctx = torch.from_numpy(a["context_lens"]).cuda()
        bt = torch.from_numpy(np.ascontiguousarray(a["block_tables"])).cuda()
        q = torch.randn(n, args.hq, 128, device="cuda").to(torch.bfloat16)
        out = torch.empty_like(q)
        if args.ncu:
            torch.cuda.synchronize()
            torch.cuda.profiler.start()
            ops.paged_decode(0, q, bt, ctx, 0.0884, out=out)
            torch.cuda.synchronize()
            torch.cuda.profiler.stop()
            continue
        if args.fused:
            from nanovllm.layers.rotary_embedding import build_cos_sin
            qkv = torch.randn(n, (args.hq + 2 * args.hkv) * 128, device="cuda").to(torch.bfloat16)
            qw = torch.ones(128, device="cuda", dtype=torch.bfloat16)
            cs = build_cos_sin(128, 4096, 1e6, "cuda")
            pos = (ctx.long() - 1)
            slots = torch.from_numpy(a["slot_mapping"]).cuda()
            qv = qkv[:, :args.hq * 128].view(n, args.hq, 128)

            def run_two():
                for layer in range(L):
                    ops.qknorm_rope_store(layer, qkv, args.hq, args.hkv, pos, qw, qw, cs, 1e-6, slots)
                    ops.paged_decode(layer, qv, bt, ctx, 0.0884, out=out)

            def run_fused():
                for layer in range(L):
                    ops.paged_decode_fused(layer, qkv, args.hq, qw, qw, cs, 1e-6, bt, ctx, 0.0884, out=out)

            row = dict(step=st, batch=n)
            for name, fn in (("two_kernels_us", run_two), ("fused_us", run_fused)):
                fn(); fn()
                best = 1e9
                for _ in range(args.reps):
                    flush.fill_(1)
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    torch.cuda.synchronize()
                    e0.record(); fn(); e1.record()
                    torch.cuda.synchronize()
                    best = min(best, e0.elapsed_time(e1) / L)
                row[name] = round(best * 1000, 2)
            print(json.dumps(row))
            continue
        for _ in range(2):
            for layer in range(L):
                ops.paged_decode(layer, q, bt, ctx, 0.0884, out=out)
        best = 1e9
        for _ in range(args.reps):
            flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            for layer in range(L):
                ops.paged_decode(layer, q, bt, ctx, 0.0884, out=out)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) / L)
        sctx = int(a["context_lens"].sum())
        nbytes = sctx * 2 * args.hkv * 128 * 2 + n * args.hq * 128 * 2 * 2 + n * (bt.shape[1] + 1) * 4
        gbs = nbytes / (best * 1e-3) / 1e9
        results.append(dict(step=st, batch=n, sum_ctx=sctx, us=best * 1000, GBps=gbs, frac_measured=gbs / peaks["hbm_gbs"], frac_8TBs=gbs / 8000))
        print(json.dumps(results[-1]))
    return results


if __name__ == "__main__":
    main()
