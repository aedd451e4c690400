from .modeling_llava import LlamaForCausalLM, LlavaLlamaForCausalLM  # noqa: F401
