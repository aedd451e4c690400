"""Pins oracle/clip.py and oracle/llama.py on HuggingFace transformers (the implementations the reference's CLIP and
PaddleNLP's Llama were ported from): identical random weights, identical inputs, every output compared in fp32.
Tolerance 2e-4 absolute on O(1) activations (two fp32 CPU implementations with different summation orders)."""
import pytest
import torch

from oracle import clip as OC
from oracle import llama as OL

transformers = pytest.importorskip("transformers")


def _to_oracle(sd, shapes, strip=""):
    """HF state dict ([out, in] Linear weights) -> the oracle's reference-named dict ([in, out])."""
    P = {}
    for k, shp in shapes.items():
        t = sd[strip + k].detach().float()
        if k.endswith(".weight") and t.ndim == 2 and "embedding" not in k and "embed_tokens" not in k:
            t = t.t()
        assert tuple(t.shape) == tuple(shp), (k, t.shape, shp)
        P[k] = t.contiguous()
    return P


def test_clip_text_vs_hf():
    from transformers import CLIPTextConfig, CLIPTextModelWithProjection
    cfg = OC.CLIP_TEXT_CONFIGS["tiny"]
    hf_cfg = CLIPTextConfig(**{k: cfg[k] for k in ("vocab_size", "hidden_size", "intermediate_size", "num_hidden_layers",
                                                   "num_attention_heads", "max_position_embeddings", "hidden_act",
                                                   "layer_norm_eps", "projection_dim")},
                            eos_token_id=2, bos_token_id=0, pad_token_id=1, attn_implementation="eager")
    torch.manual_seed(0)
    m = CLIPTextModelWithProjection(hf_cfg).eval()
    P = _to_oracle(m.state_dict(), OC.clip_text_param_shapes(cfg, with_projection=True))
    ids = torch.randint(3, 900, (3, 20))
    ids[:, -1] = 999  # the eos token = the largest id (legacy eos_token_id == 2 convention, :800-806)
    am = torch.ones(3, 20, dtype=torch.long)
    am[1, 14:] = 0
    for mask in (None, am):
        with torch.no_grad():
            ref = m(input_ids=ids, attention_mask=mask, output_hidden_states=True)
        out = OC.clip_text_forward(cfg, P, ids, attention_mask=mask)
        assert torch.allclose(out["last_hidden_state"], ref.last_hidden_state, atol=2e-4)
        assert torch.allclose(out["text_embeds"], ref.text_embeds, atol=2e-4)
        assert len(out["hidden_states"]) == len(ref.hidden_states)
        for a, b in zip(out["hidden_states"], ref.hidden_states):
            assert torch.allclose(a, b, atol=2e-4)


def test_clip_vision_vs_hf():
    from transformers import CLIPVisionConfig, CLIPVisionModel
    cfg = OC.CLIP_VISION_CONFIGS["tiny"]
    hf_cfg = CLIPVisionConfig(**cfg, attn_implementation="eager")
    torch.manual_seed(1)
    m = CLIPVisionModel(hf_cfg).eval()
    P = _to_oracle(m.state_dict(), OC.clip_vision_param_shapes(cfg))
    x = torch.randn(2, 3, 56, 56)
    with torch.no_grad():
        ref = m(pixel_values=x, output_hidden_states=True)
    out = OC.clip_vision_forward(cfg, P, x)
    assert torch.allclose(out["last_hidden_state"], ref.last_hidden_state, atol=2e-4)
    assert torch.allclose(out["pooler_output"], ref.pooler_output, atol=2e-4)
    assert torch.allclose(OC.llava_feature_select(out["hidden_states"]), ref.hidden_states[-2][:, 1:], atol=2e-4)


def test_llama_vs_hf():
    from transformers import LlamaConfig, LlamaForCausalLM
    cfg = OL.LLAMA_CONFIGS["tiny"]
    hf_cfg = LlamaConfig(**cfg, attention_bias=False, mlp_bias=False, tie_word_embeddings=False, max_position_embeddings=512,
                         attn_implementation="eager")
    torch.manual_seed(2)
    m = LlamaForCausalLM(hf_cfg).eval()
    sd = {k.replace("model.", "llama.", 1) if k.startswith("model.") else k: v for k, v in m.state_dict().items()}
    P = _to_oracle(sd, OL.llama_param_shapes(cfg))
    ids = torch.randint(0, 1000, (2, 33))
    with torch.no_grad():
        ref = m(input_ids=ids).logits
    out = OL.llama_forward(cfg, P, ids)
    assert torch.allclose(out, ref, atol=3e-4), (out - ref).abs().max()


def test_llava_composition_shapes():
    llm, vis = OL.LLAMA_CONFIGS["tiny"], OC.CLIP_VISION_CONFIGS["tiny"]
    P = OL.init_params(OL.llava_param_shapes(llm, vis), seed=3)
    ids = torch.randint(0, 1000, (2, 12))
    ids[:, 4] = OL.IMAGE_TOKEN_INDEX
    out = OL.llava_forward(llm, vis, P, ids, torch.randn(2, 3, 56, 56))
    assert out.shape == (2, 11 + 16, 1000) and torch.isfinite(out).all()
