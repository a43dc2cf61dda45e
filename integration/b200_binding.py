"""INTEGRATION.md option B, executable: the binding a maintainer of the reference adds to put libb200attn.so behind
``nanovllm.layers.attention.Attention`` -- and nothing else.  This file imports NOTHING from this repository's
``nano-vllm_b200`` package: it is the ctypes stub of INTEGRATION.md verbatim plus ``patch()``, which applies the two
edits the document describes (Attention.forward body; head-major cache allocation + bind) to an imported reference
package.  tests/test_gpu_reference_forward.py runs the otherwise unmodified reference through it on the GPU and
compares logits and greedy tokens with the unmodified reference (flash-attn) on the same weights and script.
"""
import ctypes as C
import os

import torch

_LIB = os.environ.get("B200ATTN_LIB") or os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                      "nano-vllm_b200", "lib", "libb200attn.so")
_lib = C.CDLL(_LIB)
_vp, _i, _i64, _f, _sz = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_size_t
_lib.b200_init.argtypes = [_i, C.POINTER(_vp)]
_lib.b200_kv_bind.argtypes = [_vp, _vp, _vp, _i, _i64, _i, _i, _i]
_lib.b200_decode_workspace_bytes.restype = _sz
_lib.b200_decode_workspace_bytes.argtypes = [_vp, _i, _i]
_lib.b200_store_kv.argtypes = [_vp, _i, _vp, _i64, _vp, _i64, _vp, _i, _vp]
_lib.b200_paged_decode.argtypes = [_vp, _i, _vp, _i64, _vp, _i, _vp, _vp, _i64, _i, _i, _f, _vp, _sz, _vp]
_lib.b200_paged_prefill.argtypes = [_vp, _i, _vp, _i64, _vp, _i64, _vp, _i64, _vp, _vp, _vp, _i, _vp, _i64,
                                    _i, _i, _i, _i, _i, _i, _f, _vp]
_lib.b200_strerror.restype = C.c_char_p
_lib.b200_strerror.argtypes = [_i]
_ctx = _vp()
_state = {"ws": None}


def _ok(code):
    if code != 0:
        raise RuntimeError("libb200attn: " + _lib.b200_strerror(code).decode())


def init(device: int):
    _ok(_lib.b200_init(device, C.byref(_ctx)))


def stream():
    return torch.cuda.current_stream().cuda_stream


def allocate_kv_cache(model, num_layers, num_blocks, block_size, num_kv_heads, head_dim, num_heads, max_batch=512):
    """engine/model_runner.py:115-121 with the one changed line: pages are head-major, and the cache is bound once."""
    kv_cache = torch.zeros(2, num_layers, num_blocks, num_kv_heads, block_size, head_dim, dtype=torch.bfloat16, device="cuda")
    _ok(_lib.b200_kv_bind(_ctx, kv_cache[0].data_ptr(), kv_cache[1].data_ptr(), num_layers, num_blocks, block_size, num_kv_heads, head_dim))
    _state["ws"] = torch.zeros(_lib.b200_decode_workspace_bytes(_ctx, max_batch, num_heads), dtype=torch.uint8, device="cuda")
    layer_id = 0
    for module in model.modules():
        if hasattr(module, "k_cache") and hasattr(module, "v_cache"):
            module.k_cache = kv_cache[0, layer_id]
            module.v_cache = kv_cache[1, layer_id]
            module.layer_id = layer_id
            layer_id += 1
    return kv_cache


def forward(self, q: torch.Tensor, k: torch.Tensor, v: torch.Tensor):
    """layers/attention.py:59-75 -- the new Attention.forward body."""
    from nanovllm.utils.context import get_context          # the REFERENCE's context module
    ctx = get_context()
    s = stream()
    if self.k_cache.numel():                                  # store_kvcache, attention.py:33-40
        _ok(_lib.b200_store_kv(_ctx, self.layer_id, k.data_ptr(), k.stride(0), v.data_ptr(), v.stride(0),
                               ctx.slot_mapping.data_ptr(), k.shape[0], s))
    o = torch.empty_like(q)
    if ctx.is_prefill:                                        # flash_attn_varlen_func, attention.py:67-70
        bt = ctx.block_tables
        _ok(_lib.b200_paged_prefill(_ctx, self.layer_id, q.data_ptr(), q.stride(0), k.data_ptr(), k.stride(0), v.data_ptr(), v.stride(0),
                                    ctx.cu_seqlens_q.data_ptr(), ctx.cu_seqlens_k.data_ptr(),
                                    bt.data_ptr() if bt is not None else None, bt.stride(0) if bt is not None else 0,
                                    o.data_ptr(), o.stride(0), q.shape[0], ctx.cu_seqlens_q.numel() - 1,
                                    ctx.max_seqlen_q, ctx.max_seqlen_k, self.num_heads, self.num_kv_heads, self.scale, s))
    else:                                                     # flash_attn_with_kvcache, attention.py:72-74
        ws = _state["ws"]
        _ok(_lib.b200_paged_decode(_ctx, self.layer_id, q.data_ptr(), q.stride(0), ctx.block_tables.data_ptr(), ctx.block_tables.stride(0),
                                   ctx.context_lens.data_ptr(), o.data_ptr(), o.stride(0), q.shape[0], self.num_heads, self.scale,
                                   ws.data_ptr(), ws.numel(), s))
    return o


def patch():
    """Apply the operator edit to the imported reference package (the maintainer would edit the file instead)."""
    from nanovllm.layers import attention as ref_attention
    ref_attention.Attention.forward = forward
