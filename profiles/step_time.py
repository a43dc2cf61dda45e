"""Device time of ONE captured decode step (CUDA graph replay: embedding ... LM head + sampling) per batch size, on the
benchmark's own request mix.  CUDA events around each step's GPU work (ModelRunner.begin_profile), median over steps.

    python profiles/step_time.py [batches, default 256,128,64,16,1] [steps per batch, default 24]

Use it to compare library flavours / env switches (B200ATTN_LIB, B200_LINEAR, B200_LM_HEAD ...) step for step:
    B200_LINEAR=tc python profiles/step_time.py
Prints one JSON line {"step_us": {batch: median}, "attn_us": {batch: decode-attention launches of that step alone}, ...}.
"""
import json
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "nano-vllm_b200")]

import torch  # noqa: E402

import bench  # noqa: E402


def main():
    batches = [int(b) for b in (sys.argv[1] if len(sys.argv) > 1 else "256,128,64,16,1").split(",")]
    n_steps = int(sys.argv[2]) if len(sys.argv) > 2 else 24
    from nanovllm import LLM, SamplingParams, ops
    mdir = bench.ensure_model_dir(0)
    llm = LLM(mdir, enforce_eager=False, max_model_len=4096)
    runner = llm.model_runner
    m = runner.model
    prompts, max_tokens = bench.bench_requests(0)
    out, attn = {}, {}
    for B in batches:
        for p in prompts[:B]:
            llm.add_request(p, SamplingParams(temperature=0.6, ignore_eos=True, max_tokens=n_steps + 8))
        while True:                                    # prefill steps + 3 warm decode steps, untimed
            _, nt = llm.step()
            if nt < 0:
                break
        for _ in range(3):
            llm.step()
        per = []
        for _ in range(n_steps):
            runner.begin_profile()
            llm.step()
            per.append(runner.end_profile()["device_ms"] * 1000.0)
        # the attention launches of this batch alone (same block tables / context lengths), for the share
        seqs = list(llm.scheduler.running)
        a = runner.decode_arrays(seqs)
        import numpy as np
        ctx = torch.from_numpy(a["context_lens"] - 1).cuda()      # tokens whose K/V are stored (the newest has no page yet)
        bt = torch.from_numpy(np.ascontiguousarray(a["block_tables"])).cuda()
        q = torch.randn(len(seqs), m.num_heads, m.head_dim, device="cuda").to(torch.bfloat16)
        o = torch.empty_like(q)
        L = len(m.layers)
        for _ in range(2):
            for layer in range(L):
                ops.paged_decode(layer, q, bt, ctx, 0.0884, out=o)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for layer in range(L):
            ops.paged_decode(layer, q, bt, ctx, 0.0884, out=o)
        e1.record()
        torch.cuda.synchronize()
        attn[B] = round(e0.elapsed_time(e1) * 1000.0, 1)
        out[B] = round(statistics.median(per), 1)
        while not llm.is_finished():
            llm.step()
    print(json.dumps({"step_us": out, "attn_us_28_launches": attn, "rest_us": {b: round(out[b] - attn[b], 1) for b in out},
                      "env": {k: v for k, v in os.environ.items() if k.startswith("B200")}}))
    llm.exit()


if __name__ == "__main__":
    main()
