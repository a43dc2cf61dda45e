"""One decoder layer's non-attention work at Qwen3-0.6B decode shapes: the chain of library GEMMs + elementwise kernels the
step issues today against the persistent layer-tail kernel (csrc/layer_tail.cu).

    python profiles/tail_microbench.py [splits_o,splits_down]

Each measurement is one CUDA-graph replay of the chain over 16 distinct weight sets (a real step: every layer has its own
weights, 31 MB each, far beyond L2 residency of a single set), CUDA events around the replay, best of 5, us per layer.
"""
import json
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, "nano-vllm_b200")
from nanovllm import ops  # noqa: E402

NSETS = 16
HID, INTER, QKV, OIN = 1024, 3072, 4096, 2048


def timed(fn_per_set, nsets=NSETS):
    for i in range(nsets):
        fn_per_set(i)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(nsets):
            fn_per_set(i)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / nsets)
    return round(best * 1000, 2)


def rnd(*shape, sc=1.0):
    return (torch.randn(*shape, device="cuda") * sc).to(torch.bfloat16)


def main():
    so, sd = [int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "8,8").split(",")]
    res = {}
    w_o = [rnd(HID, OIN, sc=OIN ** -0.5) for _ in range(NSETS)]
    w_gu = [rnd(2 * INTER, HID, sc=HID ** -0.5) for _ in range(NSETS)]
    w_dn = [rnd(HID, INTER, sc=INTER ** -0.5) for _ in range(NSETS)]
    w_qkv = [rnd(QKV, HID, sc=HID ** -0.5) for _ in range(NSETS)]
    wn = torch.ones(HID, dtype=torch.bfloat16, device="cuda")
    ws = ops.layer_tail_workspace(256, HID, INTER, max(so, sd))
    for M in (256, 128, 64, 16, 1):
        attn = rnd(M, OIN)
        resid = rnd(M, HID)

        def lib_chain(i):
            x, r = ops.add_rmsnorm(F.linear(attn, w_o[i]), resid, wn, 1e-6)
            a = ops.silu_mul(F.linear(x, w_gu[i]))
            x2, r = ops.add_rmsnorm(F.linear(a, w_dn[i]), resid, wn, 1e-6)
            return F.linear(x2, w_qkv[i])

        def tail(i):
            return ops.layer_tail(attn, resid, w_o[i], wn, w_gu[i], w_dn[i], wn, 1e-6, ws, w_qkv_next=w_qkv[i], splits_o=so, splits_down=sd)

        def tail_noqkv(i):
            return ops.layer_tail(attn, resid, w_o[i], wn, w_gu[i], w_dn[i], wn, 1e-6, ws, splits_o=so, splits_down=sd)

        r = {"library_chain_7_kernels": timed(lib_chain), "layer_tail_with_qkv": timed(tail), "layer_tail_without_qkv": timed(tail_noqkv)}
        torch.cuda.synchronize()
        r["barrier_timeout"] = ops.layer_tail_error(ws)
        res[f"M{M}"] = r
        print(f"M={M} " + json.dumps(r), file=sys.stderr, flush=True)
    print(json.dumps({"unit": "us per layer (graph replay over 16 weight sets)", "splits": [so, sd], "results": res}))


if __name__ == "__main__":
    main()
