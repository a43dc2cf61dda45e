"""Pure-read HBM bandwidth on this GPU (context for the decode roofline: the driver's peak is a read+write copy)."""
import json
import torch
x = torch.empty(8 << 30, dtype=torch.uint8, device="cuda").view(torch.float32)
x.zero_()
res = {}
for name, fn in (("torch.sum_fp32_8GiB", lambda: x.sum()), ("torch.max_fp32_8GiB", lambda: x.max())):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    res[name] = round(x.numel() * 4 / (best * 1e-3) / 1e9, 1)
y = torch.empty_like(x[: x.numel() // 2])
for _ in range(2):
    y.copy_(x[: x.numel() // 2])
torch.cuda.synchronize()
best = 1e9
for _ in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); y.copy_(x[: x.numel() // 2]); e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
res["torch.copy_4GiB_read_plus_write"] = round(2 * y.numel() * 4 / (best * 1e-3) / 1e9, 1)
print(json.dumps({"GB/s": res}))
