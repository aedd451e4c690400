"""ORACLE (test infrastructure, not product code): CPU fp32 restatement of the reference's Qwen2-VL prefill,
paddlemix/models/qwen2_vl/modeling_qwen2_vl.py: Qwen2VisionTransformerPretrainedModel (:916-986) with PatchEmbed
(:252-277), VisionRotaryEmbedding (:242-249), rot_pos_emb (:940-971), Qwen2VLVisionBlock (:384-399), VisionAttention
(:307-343), VisionMlp (:296-304), PatchMerger (:280-293); Qwen2VLModel / Qwen2VLDecoderLayer (:813-889),
Qwen2RMSNorm (:454-478), Qwen2VLAttention (:509-624), apply_multimodal_rotary_pos_emb (:179-224), Qwen2MLP
(:482-493), get_rope_index (:1217-1360) and Qwen2VLForConditionalGeneration.forward (:1382-1503, logits fp32).

PARITY UNPINNED: the reference has no test for qwen2_vl; the only independent cross-check available
offline is transformers' Qwen2-VL implementation (tests/test_oracle_qwen2vl_vs_hf.py, when importable).
All math is fp32 (the "reference Paddle CPU forward"); the eager attention's hard bf16 cast of P and V (:611) is a
GPU-dtype artefact and is not reproduced. Paddle Linear weights are [in, out].
"""
import math
from typing import Dict, List

import torch
import torch.nn.functional as F

from .unet import init_params, linear

Params = Dict[str, torch.Tensor]

QWEN2VL_CONFIGS = {
    # Qwen/Qwen2-VL-7B-Instruct config.json (loaded by name in paddlemix/examples/qwen2_vl/README.md:5)
    "qwen2vl_7b": dict(hidden_size=3584, intermediate_size=18944, num_hidden_layers=28, num_attention_heads=28,
                       num_key_value_heads=4, vocab_size=152064, rms_norm_eps=1e-6, rope_theta=1000000.0,
                       mrope_section=(16, 24, 24), image_token_id=151655, video_token_id=151656,
                       vision_start_token_id=151652, vision_end_token_id=151653,
                       vision=dict(depth=32, embed_dim=1280, num_heads=16, mlp_ratio=4, in_channels=3, patch_size=14,
                                   temporal_patch_size=2, spatial_merge_size=2, hidden_act="quick_gelu")),
    "tiny": dict(hidden_size=256, intermediate_size=512, num_hidden_layers=2, num_attention_heads=4,
                 num_key_value_heads=2, vocab_size=1024, rms_norm_eps=1e-6, rope_theta=1000000.0,
                 mrope_section=(8, 12, 12), image_token_id=1001, video_token_id=1002, vision_start_token_id=1003,
                 vision_end_token_id=1004,
                 vision=dict(depth=2, embed_dim=160, num_heads=2, mlp_ratio=4, in_channels=3, patch_size=14,
                             temporal_patch_size=2, spatial_merge_size=2, hidden_act="quick_gelu")),
}


def qwen2vl_param_shapes(cfg) -> Dict[str, tuple]:
    S: Dict[str, tuple] = {}
    v = cfg["vision"]
    E, H = v["embed_dim"], cfg["hidden_size"]

    def lin(name, i, o, bias=True):
        S[name + ".weight"] = (i, o)
        if bias:
            S[name + ".bias"] = (o,)

    S["visual.patch_embed.proj.weight"] = (E, v["in_channels"], v["temporal_patch_size"], v["patch_size"], v["patch_size"])
    for i in range(v["depth"]):
        b = f"visual.blocks.{i}"
        S[b + ".norm1.weight"], S[b + ".norm1.bias"] = (E,), (E,)
        S[b + ".norm2.weight"], S[b + ".norm2.bias"] = (E,), (E,)
        lin(b + ".attn.qkv", E, 3 * E), lin(b + ".attn.proj", E, E)
        lin(b + ".mlp.fc1", E, E * v["mlp_ratio"]), lin(b + ".mlp.fc2", E * v["mlp_ratio"], E)
    m = E * v["spatial_merge_size"] ** 2
    S["visual.merger.ln_q.weight"], S["visual.merger.ln_q.bias"] = (E,), (E,)
    lin("visual.merger.mlp.0", m, m), lin("visual.merger.mlp.2", m, H)
    S["model.embed_tokens.weight"] = (cfg["vocab_size"], H)
    hd = H // cfg["num_attention_heads"]
    kv = cfg["num_key_value_heads"] * hd
    for i in range(cfg["num_hidden_layers"]):
        b = f"model.layers.{i}"
        S[b + ".input_layernorm.weight"] = (H,)
        S[b + ".post_attention_layernorm.weight"] = (H,)
        lin(b + ".self_attn.q_proj", H, H), lin(b + ".self_attn.k_proj", H, kv), lin(b + ".self_attn.v_proj", H, kv)
        lin(b + ".self_attn.o_proj", H, H, bias=False)
        lin(b + ".mlp.gate_proj", H, cfg["intermediate_size"], bias=False)
        lin(b + ".mlp.up_proj", H, cfg["intermediate_size"], bias=False)
        lin(b + ".mlp.down_proj", cfg["intermediate_size"], H, bias=False)
    S["model.norm.weight"] = (H,)
    lin("lm_head", H, cfg["vocab_size"], bias=False)
    return S


def init_qwen2vl_params(cfg, seed=1) -> Params:
    shapes = qwen2vl_param_shapes(cfg)
    P = init_params({k: v for k, v in shapes.items() if k != "model.embed_tokens.weight" and len(v) != 5}, seed)
    g = torch.Generator().manual_seed(seed + 1000)
    P["model.embed_tokens.weight"] = (0.5 * torch.randn(shapes["model.embed_tokens.weight"], generator=g)).to(torch.bfloat16).float()
    w = shapes["visual.patch_embed.proj.weight"]
    fan = w[1] * w[2] * w[3] * w[4]
    P["visual.patch_embed.proj.weight"] = ((torch.rand(w, generator=g) * 2 - 1) / math.sqrt(fan)).to(torch.bfloat16).float()
    return P


def rotate_half(x):  # :169-173
    x1, x2 = x[..., : x.shape[-1] // 2], x[..., x.shape[-1] // 2:]
    return torch.cat([-x2, x1], dim=-1)


# ------------------------------------------------------------------------------------------------------------------
# vision tower
# ------------------------------------------------------------------------------------------------------------------
def rot_pos_emb(cfg, grid_thw: List[List[int]]):
    """:940-971 — per-patch (h, w) position ids in merge-window order, looked up in the 1-D frequency table."""
    v = cfg["vision"]
    m = v["spatial_merge_size"]
    pos_ids = []
    for t, h, w in grid_thw:
        hpos = torch.arange(h).unsqueeze(1).expand(-1, w).reshape(h // m, m, w // m, m).permute(0, 2, 1, 3).flatten()
        wpos = torch.arange(w).unsqueeze(0).expand(h, -1).reshape(h // m, m, w // m, m).permute(0, 2, 1, 3).flatten()
        pos_ids.append(torch.stack([hpos, wpos], dim=-1).repeat(t, 1))
    pos_ids = torch.cat(pos_ids, 0)
    max_grid = max(max(h, w) for _, h, w in grid_thw)
    dim = (v["embed_dim"] // v["num_heads"]) // 2  # VisionRotaryEmbedding(head_dim // 2), :931-932
    inv_freq = 1.0 / 10000.0 ** (torch.arange(0, dim, 2, dtype=torch.float32) / dim)
    freqs = torch.outer(torch.arange(max_grid, dtype=torch.float32), inv_freq)
    return freqs[pos_ids].flatten(1)  # [T, head_dim/2]


def vision_cu_seqlens(grid_thw):
    cu = [0]
    for t, h, w in grid_thw:
        for _ in range(t):
            cu.append(cu[-1] + h * w)
    return cu


def vision_forward(cfg, P: Params, pixel_values, grid_thw):
    """Qwen2VisionTransformerPretrainedModel.forward :973-986. pixel_values [T, C*tp*p*p]; returns [T/m^2, hidden]."""
    v = cfg["vision"]
    E, nh = v["embed_dim"], v["num_heads"]
    hd = E // nh
    w = P["visual.patch_embed.proj.weight"].reshape(E, -1)  # Conv3D k = s = (tp, p, p), no bias == GEMM (:269-277)
    x = pixel_values @ w.t()
    freqs = rot_pos_emb(cfg, grid_thw)
    cos = freqs.cos().unsqueeze(1).repeat(1, 1, 2)  # apply_rotary_pos_emb_vision :227-238 ([T,1,hd])
    sin = freqs.sin().unsqueeze(1).repeat(1, 1, 2)
    cu = vision_cu_seqlens(grid_thw)
    T = x.shape[0]
    mask = torch.full((T, T), float("-inf"))
    for i in range(1, len(cu)):
        mask[cu[i - 1]:cu[i], cu[i - 1]:cu[i]] = 0.0
    for i in range(v["depth"]):
        b = f"visual.blocks.{i}"
        h = F.layer_norm(x, (E,), P[b + ".norm1.weight"], P[b + ".norm1.bias"], 1e-6)
        qkv = linear(h, P, b + ".attn.qkv").reshape(T, 3, nh, hd).permute(1, 0, 2, 3)
        q, k, vv = qkv[0], qkv[1], qkv[2]
        q = q * cos + rotate_half(q) * sin
        k = k * cos + rotate_half(k) * sin
        q, k, vv = q.transpose(0, 1), k.transpose(0, 1), vv.transpose(0, 1)  # [nh, T, hd]
        a = torch.softmax(q @ k.transpose(1, 2) / math.sqrt(hd) + mask, dim=-1) @ vv
        x = x + linear(a.transpose(0, 1).reshape(T, E), P, b + ".attn.proj")
        h = F.layer_norm(x, (E,), P[b + ".norm2.weight"], P[b + ".norm2.bias"], 1e-6)
        h = linear(h, P, b + ".mlp.fc1")
        h = h * torch.sigmoid(1.702 * h)  # quick_gelu (configuration_qwen2_vl.py:34)
        x = x + linear(h, P, b + ".mlp.fc2")
    m2 = v["spatial_merge_size"] ** 2
    h = F.layer_norm(x, (E,), P["visual.merger.ln_q.weight"], P["visual.merger.ln_q.bias"], 1e-6).reshape(-1, E * m2)
    return linear(F.gelu(linear(h, P, "visual.merger.mlp.0")), P, "visual.merger.mlp.2")  # PatchMerger :291-293


# ------------------------------------------------------------------------------------------------------------------
# language model
# ------------------------------------------------------------------------------------------------------------------
def get_rope_index(cfg, input_ids, image_grid_thw, attention_mask=None):
    """:1217-1360 (images only). Returns position_ids [3, B, S] (int64) and rope deltas [B, 1]."""
    m = cfg["vision"]["spatial_merge_size"]
    B, S = input_ids.shape
    position_ids = torch.ones(3, B, S, dtype=torch.long)
    deltas = []
    image_index = 0
    for i in range(B):
        ids = input_ids[i]
        keep = torch.ones(S, dtype=torch.bool) if attention_mask is None else attention_mask[i] == 1
        toks = ids[keep].tolist()
        starts = [j for j, t in enumerate(toks) if t == cfg["vision_start_token_id"]]
        image_nums = sum(1 for j in starts if j + 1 < len(toks) and toks[j + 1] == cfg["image_token_id"])
        pos_list, st = [], 0
        for _ in range(image_nums):
            ed = toks.index(cfg["image_token_id"], st)
            t, h, w = image_grid_thw[image_index]
            image_index += 1
            gt, gh, gw = t, h // m, w // m
            text_len = ed - st
            st_idx = int(pos_list[-1].max()) + 1 if pos_list else 0
            pos_list.append(torch.arange(text_len).reshape(1, -1).expand(3, -1) + st_idx)
            t_index = torch.arange(gt).reshape(-1, 1).expand(-1, gh * gw).flatten()
            h_index = torch.arange(gh).reshape(1, -1, 1).expand(gt, -1, gw).flatten()
            w_index = torch.arange(gw).reshape(1, 1, -1).expand(gt, gh, -1).flatten()
            pos_list.append(torch.stack([t_index, h_index, w_index]) + text_len + st_idx)
            st = ed + gt * gh * gw
        if st < len(toks):
            st_idx = int(pos_list[-1].max()) + 1 if pos_list else 0
            pos_list.append(torch.arange(len(toks) - st).reshape(1, -1).expand(3, -1) + st_idx)
        llm_positions = torch.cat(pos_list, dim=1).reshape(3, -1)
        position_ids[:, i, keep] = llm_positions
        deltas.append(int(llm_positions.max()) + 1 - S)
    return position_ids, torch.tensor(deltas).unsqueeze(1)


def mrope_cos_sin(cfg, position_ids):
    """Qwen2RotaryEmbedding (:136-166) + the section gather of apply_multimodal_rotary_pos_emb (:212-220):
    returns cos, sin [B, S, head_dim]."""
    hd = cfg["hidden_size"] // cfg["num_attention_heads"]
    inv_freq = 1.0 / (cfg["rope_theta"] ** (torch.arange(0, hd, 2, dtype=torch.float32) / hd))
    n = int(position_ids.max()) + 1
    freqs = torch.einsum("i,j->ij", torch.arange(n, dtype=torch.float32), inv_freq)
    emb = torch.cat([freqs, freqs], dim=-1)
    cos, sin = emb.cos()[position_ids], emb.sin()[position_ids]  # [3, B, S, hd]
    sec = list(cfg["mrope_section"]) * 2
    cos = torch.cat([c[i % 3] for i, c in enumerate(cos.split(sec, dim=-1))], dim=-1)
    sin = torch.cat([s[i % 3] for i, s in enumerate(sin.split(sec, dim=-1))], dim=-1)
    return cos, sin


def rms_norm(x, w, eps):  # Qwen2RMSNorm :467-478
    var = x.float().pow(2).mean(-1, keepdim=True)
    return (torch.rsqrt(var + eps) * x) * w


def causal_padding_mask(attention_mask, S):
    """_prepare_4d_causal_attention_mask_with_cache_position (:403-444) for a prefill (cache_position = arange(S)):
    additive [B, 1, S, S] fp32, finfo.min above the diagonal and on padded keys (attention_mask [B, S], 1 = token)."""
    mn = torch.finfo(torch.float32).min
    m = torch.full((S, S), mn).triu(1)[None, None].expand(attention_mask.shape[0], 1, -1, -1).clone()
    pad = (m + attention_mask[:, None, None, :].to(torch.float32)) == 0  # unmasked by causality but a padded key
    return m.masked_fill(pad, mn)


def decoder_forward(cfg, P: Params, inputs_embeds, position_ids, attention_mask=None):
    """Qwen2VLModel.forward (:1037-1130), eager attention (:604-607): causal mask, plus the key-padding mask of a padded
    batch when attention_mask [B, S] has zeros."""
    B, S, H = inputs_embeds.shape
    nh, nkv = cfg["num_attention_heads"], cfg["num_key_value_heads"]
    hd = H // nh
    cos, sin = mrope_cos_sin(cfg, position_ids)
    cos, sin = cos.unsqueeze(1), sin.unsqueeze(1)
    causal = torch.full((S, S), float("-inf")).triu(1)
    if attention_mask is not None:
        causal = causal_padding_mask(attention_mask, S)
    x = inputs_embeds
    for i in range(cfg["num_hidden_layers"]):
        b = f"model.layers.{i}"
        h = rms_norm(x, P[b + ".input_layernorm.weight"], cfg["rms_norm_eps"])
        q = linear(h, P, b + ".self_attn.q_proj").reshape(B, S, nh, hd).transpose(1, 2)
        k = linear(h, P, b + ".self_attn.k_proj").reshape(B, S, nkv, hd).transpose(1, 2)
        v = linear(h, P, b + ".self_attn.v_proj").reshape(B, S, nkv, hd).transpose(1, 2)
        q = q * cos + rotate_half(q) * sin
        k = k * cos + rotate_half(k) * sin
        k, v = k.repeat_interleave(nh // nkv, 1), v.repeat_interleave(nh // nkv, 1)  # repeat_kv :497-506
        a = torch.softmax(q @ k.transpose(-1, -2) / math.sqrt(hd) + causal, dim=-1) @ v
        x = x + linear(a.transpose(1, 2).reshape(B, S, H), P, b + ".self_attn.o_proj")
        h = rms_norm(x, P[b + ".post_attention_layernorm.weight"], cfg["rms_norm_eps"])
        x = x + linear(F.silu(linear(h, P, b + ".mlp.gate_proj")) * linear(h, P, b + ".mlp.up_proj"), P, b + ".mlp.down_proj")
    return rms_norm(x, P["model.norm.weight"], cfg["rms_norm_eps"])


def qwen2vl_prefill(cfg, P: Params, input_ids, pixel_values, image_grid_thw, position_ids=None, attention_mask=None):
    """Qwen2VLForConditionalGeneration.forward :1382-1503 (prefill, no cache): logits fp32 [B, S, vocab]. With a padded
    batch (attention_mask [B, S] with zeros) only the rows of real tokens are meaningful."""
    embeds = P["model.embed_tokens.weight"][input_ids]
    if pixel_values is not None:
        image_embeds = vision_forward(cfg, P, pixel_values, image_grid_thw)
        embeds = embeds.clone()
        embeds[input_ids == cfg["image_token_id"]] = image_embeds
    if position_ids is None:
        position_ids, _ = get_rope_index(cfg, input_ids, image_grid_thw, attention_mask)
    hidden = decoder_forward(cfg, P, embeds, position_ids, attention_mask)
    return linear(hidden, P, "lm_head").float()


def qwen2vl_flops(cfg, B, S, grid_thw):
    v = cfg["vision"]
    E, H, I = v["embed_dim"], cfg["hidden_size"], cfg["intermediate_size"]
    T = sum(t * h * w for t, h, w in grid_thw)
    mac = T * (v["in_channels"] * v["temporal_patch_size"] * v["patch_size"] ** 2) * E
    per_seq = [h * w for t, h, w in grid_thw for _ in range(t)]
    mac += v["depth"] * (T * (4 * E * E + 2 * E * E * v["mlp_ratio"]) + sum(2 * n * n * E for n in per_seq))
    m = E * v["spatial_merge_size"] ** 2
    mac += (T // v["spatial_merge_size"] ** 2) * (m * m + m * H)
    hd = H // cfg["num_attention_heads"]
    kv = cfg["num_key_value_heads"] * hd
    mac += cfg["num_hidden_layers"] * B * (S * (2 * H * H + 2 * H * kv + 3 * H * I) + 2 * S * S * H)
    mac += B * S * H * cfg["vocab_size"]
    return 2 * mac
