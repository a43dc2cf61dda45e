"""The persistent layer-tail kernel (csrc/layer_tail.cu) against a plain PyTorch restatement of the ops it replaces:
o_proj -> add+RMSNorm -> gate_up+SiluAndMul -> down_proj -> add+RMSNorm -> next qkv_proj, with the reference's rounding
points (every F.linear output is bf16: layers/linear.py:51,73,153; the norms use the un-rounded fp32 sum:
layers/layernorm.py:28-40; SiluAndMul in fp32 on bf16 inputs: layers/activation.py:8-11)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref(attn, resid, wo, ln_mid, wgu, wd, ln_next, wqkv, eps):
    f = lambda t: t.float()
    h1 = (f(attn) @ f(wo).t()).to(torch.bfloat16)
    s1 = f(h1) + f(resid)
    r1 = s1.to(torch.bfloat16)
    x = (s1 * torch.rsqrt(s1.pow(2).mean(-1, keepdim=True) + eps) * f(ln_mid)).to(torch.bfloat16)
    gu = (f(x) @ f(wgu).t()).to(torch.bfloat16)
    inter = wd.shape[1]
    g, u = f(gu[:, :inter]), f(gu[:, inter:])
    act = (g * torch.sigmoid(g) * u).to(torch.bfloat16)
    h2 = (f(act) @ f(wd).t()).to(torch.bfloat16)
    s2 = f(h2) + f(r1)
    r2 = s2.to(torch.bfloat16)
    xn = (s2 * torch.rsqrt(s2.pow(2).mean(-1, keepdim=True) + eps) * f(ln_next)).to(torch.bfloat16)
    qkv = (f(xn) @ f(wqkv).t()).to(torch.bfloat16) if wqkv is not None else None
    return r2, xn, qkv


def _close(got, want, what, ulps=4.0):
    got, want = got.float(), want.float()
    assert torch.isfinite(got).all(), f"{what}: non-finite"
    scale = max(want.abs().max().item(), 1e-3)
    err = (got - want).abs().max().item()
    rel = ((got - want).norm() / want.norm().clamp_min(1e-6)).item()
    assert err <= ulps * 2 ** -8 * scale and rel < 1e-2, f"{what}: max abs {err} (scale {scale}), rel L2 {rel}"


def _make(rows, hidden, q_size, inter, qkv_n, seed):
    g = torch.Generator(device="cpu").manual_seed(seed)
    rn = lambda *s, sc=1.0: (torch.randn(*s, generator=g) * sc).to(torch.bfloat16).cuda()
    attn, resid = rn(rows, q_size), rn(rows, hidden)
    wo, wgu, wd = rn(hidden, q_size, sc=q_size ** -0.5), rn(2 * inter, hidden, sc=hidden ** -0.5), rn(hidden, inter, sc=inter ** -0.5)
    ln_mid = (1 + 0.1 * torch.randn(hidden, generator=g)).to(torch.bfloat16).cuda()
    ln_next = (1 + 0.1 * torch.randn(hidden, generator=g)).to(torch.bfloat16).cuda()
    wqkv = rn(qkv_n, hidden, sc=hidden ** -0.5) if qkv_n else None
    return attn, resid, wo, ln_mid, wgu, wd, ln_next, wqkv


@pytest.mark.parametrize("rows", [1, 37, 128, 200, 256])
@pytest.mark.parametrize("dims", [(1024, 2048, 3072, 4096), (256, 512, 512, 768), (1024, 2048, 3072, 0)])
def test_layer_tail_vs_torch(rows, dims):
    from nanovllm import ops
    hidden, q_size, inter, qkv_n = dims
    attn, resid, wo, ln_mid, wgu, wd, ln_next, wqkv = _make(rows, hidden, q_size, inter, qkv_n, seed=rows + hidden)
    want_r, want_x, want_qkv = _ref(attn, resid, wo, ln_mid, wgu, wd, ln_next, wqkv, 1e-6)
    ws = ops.layer_tail_workspace(256, hidden, inter, 8)
    r = resid.clone()
    splits_o = 8 if (q_size // 64) % 8 == 0 else 4
    splits_d = 8 if (inter // 64) % 8 == 0 else 4
    x_next, qkv = ops.layer_tail(attn, r, wo, ln_mid, wgu, wd, ln_next, 1e-6, ws, w_qkv_next=wqkv, splits_o=splits_o, splits_down=splits_d)
    torch.cuda.synchronize()
    assert not ops.layer_tail_error(ws), "a grid barrier timed out"
    _close(r, want_r, "residual")
    _close(x_next, want_x, "x_next")
    if qkv_n:
        _close(qkv, want_qkv, "next qkv")


def test_layer_tail_relaunch_is_bit_identical_and_graph_capturable():
    """The barrier generation carries over between launches (and CUDA-graph replays) on the same workspace."""
    from nanovllm import ops
    hidden, q_size, inter, qkv_n, rows = 1024, 2048, 3072, 4096, 77
    attn, resid, wo, ln_mid, wgu, wd, ln_next, wqkv = _make(rows, hidden, q_size, inter, qkv_n, seed=5)
    ws = ops.layer_tail_workspace(256, hidden, inter, 8)
    outs = []
    for _ in range(3):
        r = resid.clone()
        x, q = ops.layer_tail(attn, r, wo, ln_mid, wgu, wd, ln_next, 1e-6, ws, w_qkv_next=wqkv)
        outs.append((r.clone(), x.clone(), q.clone()))
    torch.cuda.synchronize()
    for o in outs[1:]:
        assert all(torch.equal(a, b) for a, b in zip(o, outs[0]))
    r_static, x_static, q_static = resid.clone(), torch.empty_like(resid), torch.empty(rows, qkv_n, dtype=torch.bfloat16, device="cuda")
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        ops.layer_tail(attn, r_static, wo, ln_mid, wgu, wd, ln_next, 1e-6, ws, w_qkv_next=wqkv, x_next=x_static, qkv_out=q_static)
    torch.cuda.current_stream().wait_stream(s)
    r_static.copy_(resid)
    with torch.cuda.graph(g):
        ops.layer_tail(attn, r_static, wo, ln_mid, wgu, wd, ln_next, 1e-6, ws, w_qkv_next=wqkv, x_next=x_static, qkv_out=q_static)
    for _ in range(2):
        r_static.copy_(resid)
        g.replay()
        torch.cuda.synchronize()
        assert torch.equal(r_static, outs[0][0]) and torch.equal(x_static, outs[0][1]) and torch.equal(q_static, outs[0][2])
    assert not ops.layer_tail_error(ws)
