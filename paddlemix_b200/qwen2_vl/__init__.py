"""Host-side mirror of paddlemix.models.qwen2_vl (prefill path)."""
from .modeling_qwen2_vl import Qwen2VLConfig, Qwen2VLForConditionalGeneration  # noqa: F401
