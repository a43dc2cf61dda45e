"""Per-request sampling knobs (reference nanovllm/sampling_params.py:4-11).

Field names and defaults are the reference's.  One deliberate widening: temperature == 0 is
accepted and means greedy argmax (the reference asserts temperature > 1e-10; the north-star's
parity criterion needs greedy token ids).
"""
from dataclasses import dataclass


@dataclass(slots=True)
class SamplingParams:
    temperature: float = 1.0
    max_tokens: int = 64
    ignore_eos: bool = False

    def __post_init__(self):
        if self.temperature < 0:
            raise ValueError("temperature must be >= 0 (0 selects greedy decoding)")
        if self.max_tokens < 1:
            raise ValueError("max_tokens must be >= 1")

    @property
    def greedy(self) -> bool:
        return self.temperature <= 1e-10
