"""Host-side mirror of paddlemix.models.qwen2_vl (prefill, KV-cache decode, greedy generate)."""
from .modeling_qwen2_vl import GraphedDecodeStep, KVCache, Qwen2VLConfig, Qwen2VLForConditionalGeneration  # noqa: F401
