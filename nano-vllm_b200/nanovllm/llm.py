"""``LLM`` is the engine (reference nanovllm/llm.py:4)."""
from .engine.llm_engine import LLMEngine


class LLM(LLMEngine):
    """User-facing handle: ``LLM(model_dir, **config).generate(prompts, sampling_params)``."""
