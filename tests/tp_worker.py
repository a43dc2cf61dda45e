"""Tensor-parallel engine check, one process per GPU (launched by tests/test_gpu_tp.py or by hand):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 tests/tp_worker.py

Every rank builds LLM(..., tensor_parallel_size=world) (SPMD replicas), generates greedily, and rank 0 checks the
tokens against the CPU oracle (teacher-forced margin test) and against what the other ranks produced.
"""
import json
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "nano-vllm_b200"), os.path.join(ROOT, "tests")]


def main():
    import torch
    import torch.distributed as dist
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    backend = os.environ.get("B200_TP_BACKEND", "nccl")
    if os.environ.get("B200_TP_ONE_DEVICE") == "1":          # every rank on GPU 0 (single-GPU boxes; gloo collectives)
        local = 0
    torch.cuda.set_device(local)
    if backend == "nccl":
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        dist.init_process_group(backend)
    from nanovllm import LLM, SamplingParams
    from nanovllm.utils.synthetic import PRESETS, make_model_dir, random_weights
    preset = os.environ.get("TP_PRESET", "tiny-g4")
    mdir = f"/tmp/tp_model_{preset}"
    if local == 0:
        make_model_dir(mdir, preset, seed=1234)
    dist.barrier()
    eager = os.environ.get("TP_EAGER", "0") == "1"
    llm = LLM(mdir, tensor_parallel_size=world, enforce_eager=eager, max_model_len=256, max_num_seqs=8,
              max_num_batched_tokens=256, kvcache_block_size=16, num_kvcache_blocks=96)
    rnd = random.Random(3)
    vocab = PRESETS[preset]["vocab_size"]
    prompts = [[rnd.randint(2, vocab - 1) for _ in range(rnd.randint(4, 70))] for _ in range(12)]
    sps = [SamplingParams(temperature=0.0 if i % 3 else 0.8, max_tokens=6 + i, ignore_eos=True) for i in range(12)]
    outs = llm.generate(prompts, sps, use_tqdm=False)
    toks = [o["token_ids"] for o in outs]
    gathered = [None] * world
    dist.all_gather_object(gathered, toks)
    result = {"world": world, "eager": eager}
    if rank == 0:
        assert all(g == toks for g in gathered), "ranks disagree on the sampled tokens"
        from gpu_helpers import check_greedy_against_oracle, make_oracle
        oracle = make_oracle(PRESETS[preset], random_weights(PRESETS[preset], seed=1234), "fused")
        n = d = 0
        for p, sp, t in zip(prompts, sps, toks):
            assert len(t) == sp.max_tokens
            if sp.temperature == 0.0:
                a, b, _ = check_greedy_against_oracle(oracle, p, t)
                n, d = n + a, d + b
        result.update(greedy_tokens=n, differ_from_oracle_argmax=d, ok=True, backend=backend,
                      peer_exchange=llm.model_runner.model.peer is not None)
        print("TP_RESULT " + json.dumps(result), flush=True)
    llm.exit()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
