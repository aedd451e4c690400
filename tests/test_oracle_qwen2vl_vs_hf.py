"""Independent cross-check of the Qwen2-VL oracle (oracle/qwen2vl.py, a restatement of paddlemix's
modeling_qwen2_vl.py) against HuggingFace transformers' Qwen2-VL on a tiny config with identical weights and
inputs. The reference has no test for this model (parity unpinned), so this is the only external
anchor available offline. Skipped when transformers' Qwen2-VL cannot be constructed."""
import pytest
import torch

from oracle import qwen2vl as Q


def _hf_model(cfg):
    transformers = pytest.importorskip("transformers")
    try:
        from transformers import Qwen2VLConfig, Qwen2VLForConditionalGeneration
        v = cfg["vision"]
        hc = Qwen2VLConfig(
            vocab_size=cfg["vocab_size"], hidden_size=cfg["hidden_size"], intermediate_size=cfg["intermediate_size"],
            num_hidden_layers=cfg["num_hidden_layers"], num_attention_heads=cfg["num_attention_heads"],
            num_key_value_heads=cfg["num_key_value_heads"], rms_norm_eps=cfg["rms_norm_eps"], rope_theta=cfg["rope_theta"],
            rope_scaling={"type": "mrope", "mrope_section": list(cfg["mrope_section"])}, image_token_id=cfg["image_token_id"],
            video_token_id=cfg["video_token_id"], vision_start_token_id=cfg["vision_start_token_id"],
            vision_end_token_id=cfg["vision_end_token_id"], tie_word_embeddings=False, attn_implementation="eager",
            vision_config=dict(depth=v["depth"], embed_dim=v["embed_dim"], hidden_size=cfg["hidden_size"],
                               num_heads=v["num_heads"], mlp_ratio=v["mlp_ratio"], in_channels=3, patch_size=14,
                               temporal_patch_size=2, spatial_merge_size=2, hidden_act="quick_gelu"))
        return Qwen2VLForConditionalGeneration(hc).eval().float()
    except Exception as ex:  # API drift across transformers versions
        pytest.skip(f"transformers Qwen2-VL not constructible: {ex}")


def test_oracle_matches_hf_qwen2vl():
    cfg = Q.QWEN2VL_CONFIGS["tiny"]
    P = Q.init_qwen2vl_params(cfg, seed=3)
    m = _hf_model(cfg)
    sd = m.state_dict()
    new = {}
    for k, t in sd.items():
        name = k.replace("model.visual.", "visual.").replace("model.language_model.", "model.")
        if name not in P:
            pytest.skip(f"unexpected HF parameter name {k}")
        w = P[name]
        if w.ndim == 2 and "embed_tokens" not in name:
            w = w.t()  # Paddle [in, out] -> torch [out, in]
        assert tuple(w.shape) == tuple(t.shape), (k, w.shape, t.shape)
        new[k] = w.contiguous()
    m.load_state_dict(new)
    g = torch.Generator().manual_seed(0)
    grid = [[1, 8, 8], [1, 4, 8]]
    pv = torch.randn(64 + 32, 3 * 2 * 14 * 14, generator=g)

    def seq(n_img, n_txt):
        return [cfg["vision_start_token_id"]] + [cfg["image_token_id"]] * n_img + [cfg["vision_end_token_id"]] + \
            torch.randint(0, 1000, (n_txt,), generator=g).tolist()
    input_ids = torch.tensor([seq(16, 20), seq(8, 28)])
    ours = Q.qwen2vl_prefill(cfg, P, input_ids, pv, grid)
    pos, _ = Q.get_rope_index(cfg, input_ids, grid)
    mm = (input_ids == cfg["image_token_id"]).int()
    with torch.no_grad():
        # (1) HF computes the M-RoPE indices itself from the token types: checks get_rope_index as well
        try:
            ref_auto = m(input_ids=input_ids, pixel_values=pv, image_grid_thw=torch.tensor(grid),
                         attention_mask=torch.ones_like(input_ids), mm_token_type_ids=mm).logits
        except (TypeError, ValueError):
            ref_auto = None
        # (2) HF with our position ids
        ref = m(input_ids=input_ids, pixel_values=pv, image_grid_thw=torch.tensor(grid),
                attention_mask=torch.ones_like(input_ids), position_ids=pos).logits
    assert ours.shape == ref.shape
    tol = 2e-4 * max(1.0, ref.abs().max().item())
    assert (ours - ref).abs().max().item() < tol, (ours - ref).abs().max().item()
    if ref_auto is not None:
        assert (ours - ref_auto).abs().max().item() < tol, "M-RoPE position ids differ from transformers'"


@pytest.mark.parametrize("side", ["left", "right"])
def test_oracle_matches_hf_qwen2vl_padded_batch(side):
    """Padded batches (attention_mask with zeros): causal + key-padding mask and the mask-aware M-RoPE positions against
    transformers, compared on the rows of real tokens."""
    cfg = Q.QWEN2VL_CONFIGS["tiny"]
    P = Q.init_qwen2vl_params(cfg, seed=4)
    m = _hf_model(cfg)
    new = {}
    for k, t in m.state_dict().items():
        name = k.replace("model.visual.", "visual.").replace("model.language_model.", "model.")
        if name not in P:
            pytest.skip(f"unexpected HF parameter name {k}")
        w = P[name]
        if w.ndim == 2 and "embed_tokens" not in name:
            w = w.t()
        new[k] = w.contiguous()
    m.load_state_dict(new)
    g = torch.Generator().manual_seed(1)
    grid = [[1, 8, 8], [1, 4, 8]]
    pv = torch.randn(64 + 32, 3 * 2 * 14 * 14, generator=g)
    seqs = [[cfg["vision_start_token_id"]] + [cfg["image_token_id"]] * n_img + [cfg["vision_end_token_id"]] +
            torch.randint(0, 1000, (n_txt,), generator=g).tolist() for n_img, n_txt in ((16, 20), (8, 11))]
    S = max(len(x) for x in seqs)
    pad = 0
    ids = torch.full((2, S), pad, dtype=torch.long)
    am = torch.zeros(2, S, dtype=torch.long)
    for i, x in enumerate(seqs):
        sl = slice(S - len(x), S) if side == "left" else slice(0, len(x))
        ids[i, sl] = torch.tensor(x)
        am[i, sl] = 1
    ours = Q.qwen2vl_prefill(cfg, P, ids, pv, grid, attention_mask=am)
    pos, _ = Q.get_rope_index(cfg, ids, grid, am)
    with torch.no_grad():
        ref = m(input_ids=ids, pixel_values=pv, image_grid_thw=torch.tensor(grid), attention_mask=am, position_ids=pos).logits
    keep = am.bool()
    tol = 2e-4 * max(1.0, ref[keep].abs().max().item())
    assert (ours[keep] - ref[keep]).abs().max().item() < tol, (ours[keep] - ref[keep]).abs().max().item()
