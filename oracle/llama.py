"""ORACLE (test infrastructure, not product code): CPU fp32 restatement of the Llama decoder LLaVA runs its prefill on, and
of the LLaVA composition around it:

  * LlamaRMSNorm / rotary embedding / LlamaAttention / LlamaMLP / LlamaDecoderLayer / LlamaModel / LlamaLMHead
    (PaddleNLP/paddlenlp/transformers/llama/modeling.py: RMSNorm :386-420, rotate_half + apply_rotary_pos_emb :534-555,
    attention :197-330 (scaled_dot_product_attention), MLP, decoder layer, model forward, lm head);
  * LlavaLlamaForCausalLM (paddlemix/models/llava/modeling.py:47-120): CLIP vision tower -> feature_select
    (clip_encoder.py:49-57) -> mm_projector mlp2x_gelu (mm_projector.py:45-58) -> the image features replace the
    IMAGE_TOKEN_INDEX placeholder of each sample (base_model.py:136-330, `prepare_inputs_labels_for_multimodal`).

Parameter names / layouts are the reference's (Paddle Linear weights are [in, out]).
Pinned: tests/test_oracle_clip_llama_vs_hf.py compares the Llama part with HuggingFace transformers' LlamaForCausalLM on
identical weights (the implementation PaddleNLP's was ported from).
"""
import math
from typing import Dict

import torch
import torch.nn.functional as F

from . import clip as OC

Params = Dict[str, torch.Tensor]
IMAGE_TOKEN_INDEX = -200  # paddlemix/models/llava/constants.py

LLAMA_CONFIGS = {
    # vicuna-7b-v1.5 (LLaVA-1.5-7B language model)
    "vicuna_7b": dict(vocab_size=32000, hidden_size=4096, intermediate_size=11008, num_hidden_layers=32,
                      num_attention_heads=32, num_key_value_heads=32, rms_norm_eps=1e-5, rope_theta=10000.0),
    "tiny": dict(vocab_size=1000, hidden_size=256, intermediate_size=512, num_hidden_layers=2, num_attention_heads=4,
                 num_key_value_heads=2, rms_norm_eps=1e-5, rope_theta=10000.0),
}


def llama_param_shapes(cfg, prefix="llama") -> Dict[str, tuple]:
    H, I = cfg["hidden_size"], cfg["intermediate_size"]
    hd = H // cfg["num_attention_heads"]
    kv = cfg["num_key_value_heads"] * hd
    S = {f"{prefix}.embed_tokens.weight": (cfg["vocab_size"], H), f"{prefix}.norm.weight": (H,),
         "lm_head.weight": (H, cfg["vocab_size"])}
    for i in range(cfg["num_hidden_layers"]):
        b = f"{prefix}.layers.{i}"
        S[b + ".input_layernorm.weight"] = S[b + ".post_attention_layernorm.weight"] = (H,)
        S[b + ".self_attn.q_proj.weight"], S[b + ".self_attn.o_proj.weight"] = (H, H), (H, H)
        S[b + ".self_attn.k_proj.weight"] = S[b + ".self_attn.v_proj.weight"] = (H, kv)
        S[b + ".mlp.gate_proj.weight"] = S[b + ".mlp.up_proj.weight"] = (H, I)
        S[b + ".mlp.down_proj.weight"] = (I, H)
    return S


def llava_param_shapes(llm_cfg, vis_cfg) -> Dict[str, tuple]:
    S = llama_param_shapes(llm_cfg)
    for k, v in OC.clip_vision_param_shapes(vis_cfg).items():
        S["llama.vision_tower.vision_tower." + k] = v
    Dv, H = vis_cfg["hidden_size"], llm_cfg["hidden_size"]
    S["llama.mm_projector.0.weight"], S["llama.mm_projector.0.bias"] = (Dv, H), (H,)
    S["llama.mm_projector.2.weight"], S["llama.mm_projector.2.bias"] = (H, H), (H,)
    return S


def init_params(shapes, seed=1) -> Params:
    g = torch.Generator().manual_seed(seed)
    P = {}
    for name in sorted(shapes):
        shp = shapes[name]
        if name.endswith("embed_tokens.weight") or "position_embedding" in name or name.endswith("class_embedding"):
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


def rms_norm(x, w, eps):  # :386-420: fp32 statistics, x * rsqrt(mean(x^2) + eps), then * weight
    v = x.float().pow(2).mean(-1, keepdim=True)
    return (x.float() * torch.rsqrt(v + eps)) * w


def rotate_half(x):  # :534-538
    x1, x2 = x[..., : x.shape[-1] // 2], x[..., x.shape[-1] // 2:]
    return torch.cat([-x2, x1], -1)


def rope_cos_sin(cfg, positions):
    """LlamaRotaryEmbedding: inv_freq = theta^(-2i/d); emb = cat(freqs, freqs); cos / sin [S, d]."""
    hd = cfg["hidden_size"] // cfg["num_attention_heads"]
    inv = 1.0 / (cfg["rope_theta"] ** (torch.arange(0, hd, 2, dtype=torch.float32) / hd))
    fr = positions.to(torch.float32)[:, None] * inv[None]
    emb = torch.cat([fr, fr], -1)
    return emb.cos(), emb.sin()


def llama_decoder(cfg, P: Params, x, positions, prefix="llama"):
    """LlamaModel.forward on input embeddings x [B, S, H] with a causal mask; positions [S]. Returns final-norm output."""
    B, S, H = x.shape
    nh, nkv = cfg["num_attention_heads"], cfg["num_key_value_heads"]
    hd = H // nh
    cos, sin = rope_cos_sin(cfg, positions)
    causal = torch.full((S, S), float("-inf")).triu(1)
    for i in range(cfg["num_hidden_layers"]):
        b = f"{prefix}.layers.{i}"
        h = rms_norm(x, P[b + ".input_layernorm.weight"], cfg["rms_norm_eps"])
        q = (h @ P[b + ".self_attn.q_proj.weight"]).reshape(B, S, nh, hd)
        k = (h @ P[b + ".self_attn.k_proj.weight"]).reshape(B, S, nkv, hd)
        v = (h @ P[b + ".self_attn.v_proj.weight"]).reshape(B, S, nkv, hd)
        q = q * cos[None, :, None] + rotate_half(q) * sin[None, :, None]  # :541-555
        k = k * cos[None, :, None] + rotate_half(k) * sin[None, :, None]
        q, k, v = (t.transpose(1, 2) for t in (q, k, v))
        k, v = k.repeat_interleave(nh // nkv, 1), v.repeat_interleave(nh // nkv, 1)
        w = (q @ k.transpose(-1, -2)) / math.sqrt(hd) + causal
        o = (torch.softmax(w, -1) @ v).transpose(1, 2).reshape(B, S, H)
        x = x + o @ P[b + ".self_attn.o_proj.weight"]
        h = rms_norm(x, P[b + ".post_attention_layernorm.weight"], cfg["rms_norm_eps"])
        x = x + (F.silu(h @ P[b + ".mlp.gate_proj.weight"]) * (h @ P[b + ".mlp.up_proj.weight"])) @ P[b + ".mlp.down_proj.weight"]
    return rms_norm(x, P[f"{prefix}.norm.weight"], cfg["rms_norm_eps"])


def llama_forward(cfg, P: Params, input_ids):
    """LlamaForCausalLM.forward: logits [B, S, vocab] in fp32."""
    x = P["llama.embed_tokens.weight"][input_ids]
    h = llama_decoder(cfg, P, x, torch.arange(input_ids.shape[1]))
    return h @ P["lm_head.weight"]


def llava_forward(llm_cfg, vis_cfg, P: Params, input_ids, images, select_layer=-2, select_feature="patch"):
    """LlavaLlamaForCausalLM.forward for batches whose samples hold exactly one IMAGE_TOKEN_INDEX placeholder at the same
    index (equal final lengths, so the padding branch of base_model.py:294-330 is the identity). Returns logits."""
    Pv = {k[len("llama.vision_tower.vision_tower."):]: v for k, v in P.items() if k.startswith("llama.vision_tower.vision_tower.")}
    feats = OC.llava_feature_select(OC.clip_vision_forward(vis_cfg, Pv, images)["hidden_states"], select_layer, select_feature)
    h = feats @ P["llama.mm_projector.0.weight"] + P["llama.mm_projector.0.bias"]  # mlp2x_gelu
    feats = F.gelu(h) @ P["llama.mm_projector.2.weight"] + P["llama.mm_projector.2.bias"]
    rows = []
    for b in range(input_ids.shape[0]):
        ids = input_ids[b]
        pos = (ids == IMAGE_TOKEN_INDEX).nonzero().flatten().tolist()
        assert len(pos) == 1, "oracle covers one image per sample"
        emb = P["llama.embed_tokens.weight"]
        rows.append(torch.cat([emb[ids[:pos[0]]], feats[b], emb[ids[pos[0] + 1:]]], 0))
    x = torch.stack(rows, 0)
    h = llama_decoder(llm_cfg, P, x, torch.arange(x.shape[1]))
    return h @ P["lm_head.weight"]
