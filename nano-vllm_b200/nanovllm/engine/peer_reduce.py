"""Symmetric-memory workspace for the fused NVLink all-reduce + add + RMSNorm (csrc/tp_allreduce.cu).

Tensor-parallel decode steps exchange ``[batch, hidden]`` partial sums 2 x layers times per step; those messages
are at most a few hundred KB, so the exchange is latency bound.  Instead of ``dist.all_reduce`` (NCCL) followed by the
add+RMSNorm kernel, the row-parallel GEMM writes its partial straight into this rank's slice of a peer-mapped
allocation and ONE kernel does handshake + peer reads + reduction + residual update + normalisation.

PyTorch provides the plumbing (``torch.distributed._symmetric_memory``: allocation, handle exchange, the device array
of peer base pointers); the kernel and its flag protocol are ours.  If symmetric memory cannot be set up on the box,
``PeerReduce.create`` returns None and the model keeps NCCL.
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist

from .. import _native as nat
from .. import ops


class PeerReduce:
    def __init__(self, rows_cap: int, hidden: int, rank: int, world: int, device: torch.device):
        import torch.distributed._symmetric_memory as symm
        self.rows_cap, self.hidden, self.rank, self.world = rows_cap, hidden, rank, world
        self.buf_bytes = (rows_cap * hidden * 2 + 255) // 256 * 256
        self.flag_off = 2 * self.buf_bytes
        self.raw = symm.empty(self.flag_off + 256, dtype=torch.uint8, device=device)
        self.raw.zero_()
        torch.cuda.synchronize()
        self.handle = symm.rendezvous(self.raw, dist.group.WORLD)
        self.bases_dev = int(self.handle.buffer_ptrs_dev)
        self.state = torch.zeros(2, dtype=torch.int32, device=device)           # [epoch, done counter]
        self.views = [self.raw[i * self.buf_bytes:i * self.buf_bytes + rows_cap * hidden * 2].view(torch.bfloat16).view(rows_cap, hidden)
                      for i in range(2)]
        self.turn = 0
        self.calls = 0
        torch.cuda.synchronize()
        dist.barrier()

    @classmethod
    def create(cls, rows_cap, hidden, rank, world, device):
        if os.environ.get("B200_TP_ALLREDUCE", "peer") != "peer":
            return None
        try:
            return cls(rows_cap, hidden, rank, world, device)
        except Exception as e:                       # no P2P / fabric handles on this box: NCCL stays in charge
            if rank == 0:
                print(f"[nanovllm] peer-memory all-reduce unavailable ({type(e).__name__}: {e}); using NCCL", flush=True)
            return None

    def next_out(self, rows: int) -> torch.Tensor:
        """Where the row-parallel GEMM of the next exchange must write its partial (the two buffers alternate)."""
        self.turn ^= 1
        return self.views[self.turn][:rows]

    def reduce_add_norm(self, rows: int, residual: torch.Tensor, weight: torch.Tensor, eps: float, out: torch.Tensor | None = None):
        """residual <- bf16(residual + sum over ranks of the partials in the current buffer); returns (normed, residual)."""
        if out is None:
            out = torch.empty_like(residual)
        lib = nat.load()
        nat.check(lib.b200_allreduce_add_rmsnorm(self.bases_dev, self.turn * self.buf_bytes, self.flag_off,
                                                 self.state.data_ptr(), self.state.data_ptr() + 4, self.rank, self.world,
                                                 residual.data_ptr(), weight.data_ptr(), out.data_ptr(), rows, self.hidden, eps,
                                                 ops._stream()))
        self.calls += 1
        return out, residual
