"""ORACLE (test infrastructure, not product code): CPU fp32 restatement of the CLIP transformers the hot path's callers use:

  * CLIPTextModel / CLIPTextModelWithProjection - the text encoders of StableDiffusionPipeline / StableDiffusionXLPipeline
    (ppdiffusers/ppdiffusers/transformers/clip/modeling.py: embeddings :199-231, attention :234-335, MLP :338-350,
    encoder layer :353-400, text transformer :726-842, projection head :1231-1304);
  * CLIPVisionModel - LLaVA's vision tower (same file :162-196, :900-953; paddlemix/models/llava/clip_model.py:945-1078 is
    the same block structure; feature selection paddlemix/models/llava/clip_encoder.py:49-57).

Parameters use the reference's state-dict names with Paddle's Linear layout ([in, out]).
Pinned: tests/test_oracle_clip_llama_vs_hf.py loads identical weights into HuggingFace transformers' CLIPTextModel /
CLIPVisionModel (the implementation the reference was ported from) and compares every output.
"""
import math
from typing import Dict

import torch
import torch.nn.functional as F

Params = Dict[str, torch.Tensor]

CLIP_TEXT_CONFIGS = {
    # openai/clip-vit-large-patch14 text tower (SD1.5 / SDXL text_encoder)
    "clip_l": dict(vocab_size=49408, hidden_size=768, intermediate_size=3072, num_hidden_layers=12, num_attention_heads=12,
                   max_position_embeddings=77, hidden_act="quick_gelu", layer_norm_eps=1e-5, eos_token_id=2, projection_dim=768),
    # laion/CLIP-ViT-bigG-14 text tower (SDXL text_encoder_2)
    "clip_bigg": dict(vocab_size=49408, hidden_size=1280, intermediate_size=5120, num_hidden_layers=32, num_attention_heads=20,
                      max_position_embeddings=77, hidden_act="gelu", layer_norm_eps=1e-5, eos_token_id=2, projection_dim=1280),
    "tiny": dict(vocab_size=1000, hidden_size=128, intermediate_size=256, num_hidden_layers=3, num_attention_heads=2,
                 max_position_embeddings=77, hidden_act="quick_gelu", layer_norm_eps=1e-5, eos_token_id=2, projection_dim=64),
}
CLIP_VISION_CONFIGS = {
    # openai/clip-vit-large-patch14-336 (LLaVA-1.5 vision tower)
    "clip_l_336": dict(hidden_size=1024, intermediate_size=4096, num_hidden_layers=24, num_attention_heads=16, image_size=336,
                       patch_size=14, num_channels=3, hidden_act="quick_gelu", layer_norm_eps=1e-5),
    "tiny": dict(hidden_size=128, intermediate_size=256, num_hidden_layers=3, num_attention_heads=2, image_size=56,
                 patch_size=14, num_channels=3, hidden_act="quick_gelu", layer_norm_eps=1e-5),
}


def _encoder_shapes(S, prefix, cfg):
    D, I = cfg["hidden_size"], cfg["intermediate_size"]
    for i in range(cfg["num_hidden_layers"]):
        b = f"{prefix}.encoder.layers.{i}"
        for n in ("q_proj", "k_proj", "v_proj", "out_proj"):
            S[f"{b}.self_attn.{n}.weight"], S[f"{b}.self_attn.{n}.bias"] = (D, D), (D,)
        for n in ("layer_norm1", "layer_norm2"):
            S[f"{b}.{n}.weight"], S[f"{b}.{n}.bias"] = (D,), (D,)
        S[f"{b}.mlp.fc1.weight"], S[f"{b}.mlp.fc1.bias"] = (D, I), (I,)
        S[f"{b}.mlp.fc2.weight"], S[f"{b}.mlp.fc2.bias"] = (I, D), (D,)


def clip_text_param_shapes(cfg, with_projection=False) -> Dict[str, tuple]:
    D = cfg["hidden_size"]
    S = {"text_model.embeddings.token_embedding.weight": (cfg["vocab_size"], D),
         "text_model.embeddings.position_embedding.weight": (cfg["max_position_embeddings"], D),
         "text_model.final_layer_norm.weight": (D,), "text_model.final_layer_norm.bias": (D,)}
    _encoder_shapes(S, "text_model", cfg)
    if with_projection:
        S["text_projection.weight"] = (D, cfg["projection_dim"])
    return S


def clip_vision_param_shapes(cfg) -> Dict[str, tuple]:
    D, p = cfg["hidden_size"], cfg["patch_size"]
    n_pos = (cfg["image_size"] // p) ** 2 + 1
    S = {"vision_model.embeddings.class_embedding": (D,),
         "vision_model.embeddings.patch_embedding.weight": (D, cfg["num_channels"], p, p),
         "vision_model.embeddings.position_embedding.weight": (n_pos, D),
         "vision_model.pre_layrnorm.weight": (D,), "vision_model.pre_layrnorm.bias": (D,),
         "vision_model.post_layernorm.weight": (D,), "vision_model.post_layernorm.bias": (D,)}
    _encoder_shapes(S, "vision_model", cfg)
    return S


def init_clip_params(shapes, seed=1) -> Params:
    g = torch.Generator().manual_seed(seed)
    P = {}
    for name in sorted(shapes):
        shp = shapes[name]
        if "embedding" in name and len(shp) == 2 and "patch" not in name:
            t = 0.3 * torch.randn(shp, generator=g)
        elif name.endswith("class_embedding"):
            t = 0.3 * torch.randn(shp, generator=g)
        elif name.endswith(".weight") and len(shp) >= 2:
            fan_in = shp[0] if len(shp) == 2 else math.prod(shp[1:])
            t = (torch.rand(shp, generator=g) * 2 - 1) / math.sqrt(fan_in)
        elif name.endswith(".weight"):
            t = 1.0 + 0.1 * torch.randn(shp, generator=g)
        else:
            t = 0.05 * torch.randn(shp, generator=g)
        P[name] = t.to(torch.bfloat16).float()
    return P


def _act(x, name):
    if name == "quick_gelu":  # activations.py / clip_model.py:60-61: x * sigmoid(1.702 x)
        return x * torch.sigmoid(1.702 * x)
    if name == "gelu":
        return F.gelu(x)
    raise ValueError(name)


def _lin(x, P, n):
    return x @ P[n + ".weight"] + P[n + ".bias"]


def _ln(x, P, n, eps):
    return F.layer_norm(x, (x.shape[-1],), P[n + ".weight"], P[n + ".bias"], eps)


def clip_attention(x, P, p, heads, mask):
    """CLIPAttention.forward, modeling.py:259-335: q scaled by head_dim**-0.5, additive causal / padding masks."""
    B, S, D = x.shape
    d = D // heads
    q = (_lin(x, P, p + ".q_proj") * d ** -0.5).reshape(B, S, heads, d).transpose(1, 2)
    k = _lin(x, P, p + ".k_proj").reshape(B, S, heads, d).transpose(1, 2)
    v = _lin(x, P, p + ".v_proj").reshape(B, S, heads, d).transpose(1, 2)
    w = q @ k.transpose(-1, -2)
    if mask is not None:
        w = w + mask
    o = torch.softmax(w, -1) @ v
    return _lin(o.transpose(1, 2).reshape(B, S, D), P, p + ".out_proj")


def clip_encoder(x, P, prefix, cfg, mask):
    """CLIPEncoder.forward :644-723: returns (last_hidden_state, all hidden states [input of every layer + final])."""
    hidden = []
    for i in range(cfg["num_hidden_layers"]):
        hidden.append(x)
        b = f"{prefix}.encoder.layers.{i}"
        x = x + clip_attention(_ln(x, P, b + ".layer_norm1", cfg["layer_norm_eps"]), P, b + ".self_attn",
                               cfg["num_attention_heads"], mask)  # :362-400
        h = _ln(x, P, b + ".layer_norm2", cfg["layer_norm_eps"])
        x = x + _lin(_act(_lin(h, P, b + ".mlp.fc1"), cfg["hidden_act"]), P, b + ".mlp.fc2")
    hidden.append(x)
    return x, hidden


def clip_text_forward(cfg, P: Params, input_ids, attention_mask=None):
    """CLIPTextTransformer.forward :745-835 (+ text_projection :1286-1288 when present): returns dict with
    last_hidden_state [B,S,D], pooler_output [B,D], hidden_states (list), text_embeds (if projection)."""
    B, S = input_ids.shape
    x = P["text_model.embeddings.token_embedding.weight"][input_ids] + P["text_model.embeddings.position_embedding.weight"][:S]
    mask = torch.full((S, S), torch.finfo(torch.float32).min).triu(1)[None, None]  # _build_causal_attention_mask :837-842
    if attention_mask is not None:  # _expand_mask: (1 - m) * finfo.min, [B,1,S,S]
        inv = 1.0 - attention_mask[:, None, None, :].to(torch.float32)
        mask = mask + inv.masked_fill(inv.bool(), torch.finfo(torch.float32).min)
    x, hidden = clip_encoder(x, P, "text_model", cfg, mask)
    last = _ln(x, P, "text_model.final_layer_norm", cfg["layer_norm_eps"])
    if cfg["eos_token_id"] == 2:  # :800-806: legacy configs: the eos token is the largest id
        eos = input_ids.argmax(-1)
    else:
        eos = (input_ids == cfg["eos_token_id"]).int().argmax(-1)
    pooled = last[torch.arange(B), eos]
    out = dict(last_hidden_state=last, pooler_output=pooled, hidden_states=hidden)
    if "text_projection.weight" in P:
        out["text_embeds"] = pooled @ P["text_projection.weight"]
    return out


def clip_vision_forward(cfg, P: Params, pixel_values):
    """CLIPVisionTransformer.forward :911-953 with CLIPVisionEmbeddings :187-196."""
    B = pixel_values.shape[0]
    p = cfg["patch_size"]
    patches = F.conv2d(pixel_values, P["vision_model.embeddings.patch_embedding.weight"], None, stride=p)  # no bias
    patches = patches.flatten(2).transpose(1, 2)
    cls = P["vision_model.embeddings.class_embedding"].expand(B, 1, -1)
    x = torch.cat([cls, patches], 1) + P["vision_model.embeddings.position_embedding.weight"][None]
    x = _ln(x, P, "vision_model.pre_layrnorm", cfg["layer_norm_eps"])
    x, hidden = clip_encoder(x, P, "vision_model", cfg, None)
    pooled = _ln(x[:, 0], P, "vision_model.post_layernorm", cfg["layer_norm_eps"])
    return dict(last_hidden_state=x, pooler_output=pooled, hidden_states=hidden)


def llava_feature_select(hidden_states, select_layer=-2, select_feature="patch"):
    """CLIPVisionTower.feature_select, paddlemix/models/llava/clip_encoder.py:49-57."""
    f = hidden_states[select_layer]
    return f[:, 1:] if select_feature == "patch" else f
