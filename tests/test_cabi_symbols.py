"""The C-ABI library loads and exports every symbol include/b200_paged_attn.h declares (no GPU needed)."""
import ctypes
import os
import re

import pytest

from nanovllm import _native as nat

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "b200_paged_attn.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b200_[a-z0-9_]+)\s*\(", text)))


def test_library_built():
    assert nat.lib_path().exists(), "run `make -C nano-vllm_b200/csrc` (or __graft_entry__.build())"


def test_exports_match_header():
    names = declared_symbols()
    assert len(names) >= 17
    lib = ctypes.CDLL(str(nat.lib_path()))
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
    assert set(nat.SIGNATURES) == set(names), "ctypes table and header disagree"


def test_host_only_entry_points():
    lib = nat.load()
    assert lib.b200_abi_version() == 1
    assert lib.b200_strerror(0) == b"ok"
    assert b"not bound" in lib.b200_strerror(-4)
    assert lib.b200_sm_count(None) == 0
    assert lib.b200_decode_workspace_bytes(None, 4, 4) == 0


def test_no_cpu_fallback(monkeypatch):
    """Without a CUDA device the product refuses to run instead of falling back."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    nat.reset_handle()
    with pytest.raises(nat.B200Error):
        nat.handle()
    from nanovllm import ops
    with pytest.raises(nat.B200Error):
        ops.silu_mul(torch.zeros(2, 16, dtype=torch.bfloat16))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "nano-vllm_b200", "nanovllm")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f"{f} imports the oracle"
