"""ctransformers-b200: B200-native drop-in for ctransformers' quantized eval hot path."""
from .hub import AutoConfig, AutoModelForCausalLM
from .llm import LLM, Config

__all__ = ["AutoConfig", "AutoModelForCausalLM", "LLM", "Config"]
__version__ = "0.1.0"
