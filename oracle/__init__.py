"""CPU oracle for the paged-attention hot path.  TEST INFRASTRUCTURE ONLY.

Everything under ``oracle/`` is a checker: a CPU restatement of what the
reference (GeeeekExplorer/nano-vllm @ bb823b3e) computes on the path named by
BASELINE.json's north_star.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s cpu_baseline / ``--impl reference`` legs may import it.  The
product package (``nano-vllm_b200/nanovllm``) never does, and fails loudly if
its CUDA library is missing.

Pinning status
--------------
* integer bookkeeping (Scheduler / BlockManager / Sequence / prepare_*):
  pinned — ``oracle/make_golden.py`` runs the reference's own classes in the
  build container and commits their traces under ``tests/golden/``.
* model arithmetic around the kernel (RMSNorm, RoPE, SwiGLU, linears,
  sampler-greedy): pinned — the reference's own ``nn.Module``s are run on CPU
  by ``make_golden.py`` and ``oracle/qwen3_ref.py`` reproduces them bit-exactly.
* QK^T / softmax / PV: the arithmetic lives in flash-attn (unpinned dependency,
  not vendored; installed here 2.8.3).  The reference holds no golden vectors
  for it => **parity unpinned** for the floating-point attention core; the
  oracle follows flash-attn's published semantics (mask alignment, paged
  layout, fp32 softmax) and is additionally cross-checked on the GPU box
  against the installed flash-attn binary when it is importable.
"""
