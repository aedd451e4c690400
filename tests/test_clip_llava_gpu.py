"""CLIP text encoders, CLIP vision tower, Llama prefill and the LLaVA composition on the GPU against the CPU fp32 oracles
(oracle/clip.py, oracle/llama.py - both pinned on HuggingFace transformers in tests/test_oracle_clip_llama_vs_hf.py).
Stated tolerance (bf16 kernels vs fp32 oracle): cosine >= 0.999, max |err| <= 4 % of the output's max magnitude."""
import pytest
import torch

from oracle import clip as OC
from oracle import llama as OL

pytestmark = pytest.mark.gpu
bf16 = torch.bfloat16


def check(out, ref, what):
    o, r = out.float().cpu(), ref.float()
    assert o.shape == r.shape, (what, o.shape, r.shape)
    cos = torch.nn.functional.cosine_similarity(o.flatten().double(), r.flatten().double(), dim=0).item()
    err = (o - r).abs().max().item() / r.abs().max().item()
    assert cos >= 0.999 and err <= 0.04, (what, cos, err)


@pytest.mark.parametrize("name,layers", [("tiny", None), ("clip_l", 2), ("clip_bigg", 2)])
def test_clip_text_parity(name, layers):
    from paddlemix_b200.clip import CLIPTextModelWithProjection
    cfg = dict(OC.CLIP_TEXT_CONFIGS[name])
    if layers:
        cfg["num_hidden_layers"] = layers  # the real widths / head counts / activations, depth cut for the CPU oracle
    P = OC.init_clip_params(OC.clip_text_param_shapes(cfg, with_projection=True), seed=1)
    model = CLIPTextModelWithProjection(cfg).load_state_dict(P, device=0)
    assert model.state_dict_shapes() == OC.clip_text_param_shapes(cfg, with_projection=True)
    g = torch.Generator().manual_seed(0)
    ids = torch.randint(3, cfg["vocab_size"] - 1, (3, 77), generator=g)
    ids[:, 60] = cfg["vocab_size"] - 1  # eos = largest id; everything after it is padding the causal mask leaves visible
    am = torch.ones(3, 77, dtype=torch.long)
    am[1, 61:] = 0
    for mask in (None, am):
        ref = OC.clip_text_forward(cfg, P, ids, attention_mask=mask)
        out = model(input_ids=ids, attention_mask=mask, output_hidden_states=True)
        check(out.last_hidden_state, ref["last_hidden_state"], f"{name} last_hidden_state mask={mask is not None}")
        check(out.text_embeds, ref["text_embeds"], f"{name} text_embeds")
        check(out.hidden_states[-2], ref["hidden_states"][-2], f"{name} penultimate hidden state (SDXL prompt_embeds)")
        assert len(out.hidden_states) == len(ref["hidden_states"])


@pytest.mark.parametrize("name,layers", [("tiny", None), ("clip_l_336", 3)])
def test_clip_vision_parity(name, layers):
    from paddlemix_b200.clip import CLIPVisionModel
    cfg = dict(OC.CLIP_VISION_CONFIGS[name])
    if layers:
        cfg["num_hidden_layers"] = layers
    P = OC.init_clip_params(OC.clip_vision_param_shapes(cfg), seed=2)
    model = CLIPVisionModel(cfg).load_state_dict(P, device=0)
    assert model.state_dict_shapes() == OC.clip_vision_param_shapes(cfg)
    x = torch.randn(2, 3, cfg["image_size"], cfg["image_size"], generator=torch.Generator().manual_seed(1)).to(bf16).float()
    ref = OC.clip_vision_forward(cfg, P, x)
    out = model(pixel_values=x.cuda(), output_hidden_states=True)
    check(out.last_hidden_state, ref["last_hidden_state"], f"{name} last_hidden_state")
    check(out.pooler_output, ref["pooler_output"], f"{name} pooler_output")
    check(out.hidden_states[-2][:, 1:], OC.llava_feature_select(ref["hidden_states"]), f"{name} LLaVA features")


@pytest.mark.parametrize("name,layers,S", [("tiny", None, 40), ("vicuna_7b", 2, 200)])
def test_llama_prefill_parity(name, layers, S):
    from paddlemix_b200.llava import LlamaForCausalLM
    cfg = dict(OL.LLAMA_CONFIGS[name])
    if layers:
        cfg["num_hidden_layers"] = layers
    P = OL.init_params(OL.llama_param_shapes(cfg), seed=3)
    model = LlamaForCausalLM(cfg).load_state_dict(P, device=0)
    assert model.state_dict_shapes() == OL.llama_param_shapes(cfg)
    ids = torch.randint(0, cfg["vocab_size"], (2, S), generator=torch.Generator().manual_seed(2))
    check(model(input_ids=ids).logits, OL.llama_forward(cfg, P, ids), f"llama {name}")


def test_llava_forward_parity():
    from paddlemix_b200.llava import LlavaLlamaForCausalLM
    llm, vis = OL.LLAMA_CONFIGS["tiny"], OC.CLIP_VISION_CONFIGS["tiny"]
    P = OL.init_params(OL.llava_param_shapes(llm, vis), seed=4)
    model = LlavaLlamaForCausalLM(llm, vision_config=vis).load_state_dict(P, device=0)
    assert model.state_dict_shapes() == OL.llava_param_shapes(llm, vis)
    g = torch.Generator().manual_seed(3)
    ids = torch.randint(0, 1000, (2, 24), generator=g)
    ids[:, 5] = OL.IMAGE_TOKEN_INDEX
    img = torch.randn(2, 3, 56, 56, generator=g).to(bf16).float()
    out = model(input_ids=ids, images=img.cuda()).logits
    check(out, OL.llava_forward(llm, vis, P, ids, img), "llava logits")
    assert out.shape == (2, 23 + 16, 1000)
