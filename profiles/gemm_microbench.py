"""cuBLAS GEMM latency at the decode shapes of Qwen3-0.6B (library calls on the path; context for DESIGN.md)."""
import json
import os
import sys
import torch
import torch.nn.functional as F

shapes = {"qkv": (4096, 1024), "o": (1024, 2048), "gate_up": (6144, 1024), "down": (1024, 3072), "lm_head": (151936, 1024)}
res = {}
for M in (256, 128, 32, 1):
    for name, (N, K) in shapes.items():
        x = torch.randn(M, K, device="cuda").to(torch.bfloat16)
        ws = [torch.randn(N, K, device="cuda").to(torch.bfloat16) for _ in range(8 if N < 100000 else 2)]
        for w in ws:
            F.linear(x, w)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for w in ws:
                F.linear(x, w)
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) / len(ws))
        res[f"M{M}_{name}"] = round(best * 1000, 2)
print(json.dumps({"tunable": os.environ.get("PYTORCH_TUNABLEOP_ENABLED", "0"), "us": res}))
