"""`LLM(path, tensor_parallel_size=2)` from ONE process, the way the reference is used (llm_engine.py:24-30): the
engine spawns the second rank itself and mirrors generate() to it.  Run by tests/test_gpu_tp.py (needs 2 GPUs)."""
import json
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "nano-vllm_b200"), os.path.join(ROOT, "tests")]


def main():
    from nanovllm import LLM, SamplingParams
    from nanovllm.utils.synthetic import PRESETS, make_model_dir, random_weights
    preset = "tiny-g4"
    mdir = make_model_dir(f"/tmp/tp_spawn_model_{preset}", preset, seed=1234)
    llm = LLM(mdir, tensor_parallel_size=2, max_model_len=256, max_num_seqs=8, max_num_batched_tokens=256,
              kvcache_block_size=16, num_kvcache_blocks=96)
    rnd = random.Random(5)
    vocab = PRESETS[preset]["vocab_size"]
    prompts = [[rnd.randint(2, vocab - 1) for _ in range(rnd.randint(4, 60))] for _ in range(9)]
    sps = [SamplingParams(temperature=0.0, max_tokens=10, ignore_eos=True) for _ in prompts]
    outs = llm.generate(prompts, sps, use_tqdm=False)
    outs2 = llm.generate(prompts[:3], sps[:3], use_tqdm=False)          # a second mirrored call
    llm.exit()
    from gpu_helpers import check_greedy_against_oracle, make_oracle
    oracle = make_oracle(PRESETS[preset], random_weights(PRESETS[preset], seed=1234), "fused")
    n = d = 0
    for p, o in zip(prompts + prompts[:3], outs + outs2):
        assert len(o["token_ids"]) == 10
        a, b, _ = check_greedy_against_oracle(oracle, p, o["token_ids"])
        n, d = n + a, d + b
    print("SPAWN_RESULT " + json.dumps({"ok": True, "tokens": n, "differ_from_oracle_argmax": d}), flush=True)


if __name__ == "__main__":
    main()
