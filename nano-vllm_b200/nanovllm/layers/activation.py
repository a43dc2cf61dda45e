"""SwiGLU gate (reference nanovllm/layers/activation.py:6-11) on the CUDA kernel."""
from torch import nn

from .. import ops


class SiluAndMul(nn.Module):
    def forward(self, x):
        return ops.silu_mul(x)
