"""Drive a few engine steps of the benchmark workload eagerly so ncu can list / profile every kernel.

    ncu --profile-from-start off ... python profiles/step_trace.py decode|prefill [n_steps]

The profiled window (cudaProfilerStart/Stop) covers `n_steps` steps of the requested kind: the first decode
steps at batch 256 (sum of contexts 143 083 at step 0) or the first prefill steps (~16k tokens each).
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "nano-vllm_b200")]

import torch  # noqa: E402

import bench  # noqa: E402


def main():
    kind = sys.argv[1] if len(sys.argv) > 1 else "decode"
    n_steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    from nanovllm import LLM, SamplingParams
    mdir = bench.ensure_model_dir(0)
    llm = LLM(mdir, enforce_eager=True, max_model_len=4096, num_kvcache_blocks=800)
    prompts, max_tokens = bench.bench_requests(0)
    for p, mt in zip(prompts, max_tokens):
        llm.add_request(p, SamplingParams(temperature=0.6, ignore_eos=True, max_tokens=mt))
    done = 0
    if kind == "prefill":
        llm.step()                                   # one un-profiled step to warm cuBLAS
    else:
        while True:                                  # run all prefill steps + one decode step un-profiled
            _, num_tokens = llm.step()
            if num_tokens < 0:
                break
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    while done < n_steps:
        _, num_tokens = llm.step()
        if (num_tokens < 0) == (kind == "decode"):
            done += 1
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    llm.exit()


if __name__ == "__main__":
    main()
