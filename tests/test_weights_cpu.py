"""Checkpoint readers / layout mapping (paddlemix_b200/weights.py): host logic only, runs without a GPU."""
import json
import os
import pickle
import struct

import numpy as np
import pytest
import torch

from oracle import unet as O
from paddlemix_b200 import weights as W


def test_safetensors_reader_on_a_handwritten_archive(tmp_path):
    # built straight from the format description (8-byte LE header size, JSON header, raw LE data), not by our writer
    a = np.arange(6, dtype="<f4").reshape(2, 3)
    b = np.array([1, -2, 3], dtype="<i8")
    bf = torch.tensor([1.5, -0.25, 3.0e4], dtype=torch.bfloat16)
    raw_bf = bf.view(torch.int16).numpy().astype("<i2").tobytes()
    data = a.tobytes() + b.tobytes() + raw_bf
    header = {"__metadata__": {"format": "pt"},
              "w": {"dtype": "F32", "shape": [2, 3], "data_offsets": [0, 24]},
              "idx": {"dtype": "I64", "shape": [3], "data_offsets": [24, 48]},
              "h": {"dtype": "BF16", "shape": [3], "data_offsets": [48, 54]}}
    hj = json.dumps(header).encode()
    p = tmp_path / "x.safetensors"
    p.write_bytes(struct.pack("<Q", len(hj)) + hj + data)
    sd, meta = W.read_safetensors(str(p))
    assert meta == {"format": "pt"}
    assert torch.equal(sd["w"], torch.from_numpy(a.copy())) and torch.equal(sd["idx"], torch.from_numpy(b.copy()))
    assert sd["h"].dtype == torch.bfloat16 and torch.equal(sd["h"], bf)
    only, _ = W.read_safetensors(str(p), keys=["idx"])
    assert list(only) == ["idx"]
    with pytest.raises(W.CheckpointError):
        W.read_safetensors(str(p), keys=["nope"])


def test_safetensors_round_trip_and_corruption(tmp_path):
    g = torch.Generator().manual_seed(0)
    t = {"a.weight": torch.randn(5, 7, generator=g), "b": torch.randn(3, generator=g).to(torch.bfloat16),
         "c": torch.randn(2, 2, 2, generator=g).to(torch.float16), "empty": torch.empty(0, 4), "flag": torch.tensor([True, False]),
         "scalar": torch.tensor(3.5)}
    p = str(tmp_path / "m.safetensors")
    W.write_safetensors(p, t, {"format": "pd"})
    sd, meta = W.read_safetensors(p)
    assert meta["format"] == "pd" and set(sd) == set(t)
    for k in t:
        assert sd[k].dtype == t[k].dtype and sd[k].shape == t[k].shape and torch.equal(sd[k], t[k])
    assert W.read_checkpoint(p)[1] == "paddle"
    blob = open(p, "rb").read()
    bad = str(tmp_path / "bad.safetensors")
    open(bad, "wb").write(blob[:-5])  # truncated data section
    with pytest.raises(W.CheckpointError):
        W.read_safetensors(bad)
    open(bad, "wb").write(struct.pack("<Q", 1 << 40) + b"{}")
    with pytest.raises(W.CheckpointError):
        W.read_safetensors(bad)


def test_sharded_safetensors_index(tmp_path):
    t1, t2 = {"x.weight": torch.ones(2, 3)}, {"y.bias": torch.zeros(4), "z": torch.arange(3)}
    W.write_safetensors(str(tmp_path / "model-00001-of-00002.safetensors"), t1, {"format": "pt"})
    W.write_safetensors(str(tmp_path / "model-00002-of-00002.safetensors"), t2, {"format": "pt"})
    idx = {"metadata": {}, "weight_map": {"x.weight": "model-00001-of-00002.safetensors", "y.bias": "model-00002-of-00002.safetensors",
                                          "z": "model-00002-of-00002.safetensors"}}
    (tmp_path / "model.safetensors.index.json").write_text(json.dumps(idx))
    sd, layout = W.read_checkpoint(str(tmp_path))  # a directory: the index wins over the shard files
    assert layout == "torch" and set(sd) == {"x.weight", "y.bias", "z"} and torch.equal(sd["z"], torch.arange(3))


def test_pdparams_reader(tmp_path):
    bf = torch.tensor([[1.0, -2.5], [0.125, 300.0]], dtype=torch.bfloat16)
    obj = {"lin.weight": np.arange(6, dtype=np.float32).reshape(3, 2), "lin.bias": np.zeros(2, np.float16),
           "h.weight": bf.view(torch.int16).numpy().view(np.uint16), "StructuredToParameterName@@": {"lin.weight": "linear_0.w_0"}}
    p = str(tmp_path / "model_state.pdparams")
    with open(p, "wb") as f:
        pickle.dump(obj, f, protocol=4)
    sd, layout = W.read_checkpoint(p)
    assert layout == "paddle" and set(sd) == {"lin.weight", "lin.bias", "h.weight"}
    assert sd["h.weight"].dtype == torch.bfloat16 and torch.equal(sd["h.weight"], bf)
    assert torch.equal(sd["lin.weight"], torch.arange(6, dtype=torch.float32).reshape(3, 2))

    class Evil:
        def __reduce__(self):
            return (os.system, ("true",))

    with open(p, "wb") as f:
        pickle.dump({"w": Evil()}, f)
    with pytest.raises(W.CheckpointError):
        W.read_pdparams(p)


def test_pdparams_reader_refuses_dotted_and_submodule_globals(tmp_path):
    """Protocol-4 STACK_GLOBAL resolves dotted names through module attributes: ('numpy._core._methods', 'os.getcwd')
    reaches os via a numpy submodule. The allow-list is exact (module, name) pairs, so all of these are refused."""
    def payload(module, name):  # PROTO 4, SHORT_BINUNICODE module, SHORT_BINUNICODE name, STACK_GLOBAL, EMPTY_TUPLE, REDUCE, STOP
        m, n = module.encode(), name.encode()
        return b"\x80\x04" + b"\x8c" + bytes([len(m)]) + m + b"\x8c" + bytes([len(n)]) + n + b"\x93" + b")" + b"R" + b"."

    for module, name in (("numpy._core._methods", "os.getcwd"), ("numpy.core._methods", "os.getcwd"),
                         ("builtins", "getattr"), ("builtins", "eval"), ("numpy", "load"),
                         ("numpy.lib.npyio", "load"), ("collections", "OrderedDict.fromkeys"), ("copyreg", "_reconstructor")):
        p = str(tmp_path / "evil.pdparams")
        with open(p, "wb") as f:
            f.write(payload(module, name))
        with pytest.raises(W.CheckpointError):
            W.read_pdparams(p)


def _tiny_unet():
    from paddlemix_b200.ppdiffusers.unet_2d_condition import UNet2DConditionModel
    cfg = O.UNET_CONFIGS["tiny_xl"]
    keys = ("in_channels", "out_channels", "flip_sin_to_cos", "freq_shift", "down_block_types", "up_block_types",
            "block_out_channels", "layers_per_block", "norm_num_groups", "norm_eps", "cross_attention_dim",
            "transformer_layers_per_block", "attention_head_dim", "use_linear_projection", "addition_embed_type",
            "addition_time_embed_dim", "projection_class_embeddings_input_dim", "resnet_out_scale_factor")
    return cfg, UNet2DConditionModel(**{k: cfg[k] for k in keys})


def test_torch_layout_conversion_matches_reference_rule():
    """convert_pytorch_state_dict_to_paddle (modeling_pytorch_paddle_utils.py:27-64): exactly the nn.Linear weights are
    transposed; convs, norms, biases are untouched; bookkeeping keys are dropped."""
    cfg, model = _tiny_unet()
    P = O.init_params(O.unet_param_shapes(cfg), seed=3)  # reference names, paddle layouts
    lin = W.linear_weight_keys(model)
    assert "time_embedding.linear_1.weight" in lin and "conv_in.weight" not in lin
    assert any(k.endswith("attn1.to_q.weight") for k in lin) and any(k.endswith("ff.net.0.proj.weight") for k in lin)
    assert all(len(model.state_dict_shapes()[k]) == 2 for k in lin)
    torch_sd = {k: (v.t().contiguous() if k in lin else v) for k, v in P.items()}  # what a diffusers checkpoint holds
    torch_sd["some.position_ids"] = torch.arange(4)
    back = W.torch_to_paddle_layout(model, torch_sd)
    assert set(back) == set(P)
    for k in P:
        assert torch.equal(back[k], P[k]), k
    rep = W.check_against_model(model, back)
    assert rep["missing"] == [] and rep["bad_shape"] == []
    # handing the torch layout over as if it were paddle's is caught by the shape check (non-square Linears exist)
    torch_sd.pop("some.position_ids")
    with pytest.raises(W.CheckpointError, match="wrong shape"):
        W.check_against_model(model, torch_sd)
    del back["conv_in.weight"]
    with pytest.raises(W.CheckpointError, match="missing"):
        W.check_against_model(model, back)


def test_embedding_tables_are_not_transposed():
    from paddlemix_b200.qwen2_vl.modeling_qwen2_vl import Qwen2VLConfig, Qwen2VLForConditionalGeneration
    cfg = Qwen2VLConfig(vocab_size=64, hidden_size=128, intermediate_size=128, num_hidden_layers=1, num_attention_heads=2,
                        num_key_value_heads=1, vision=dict(depth=1, embed_dim=64, num_heads=2, mlp_ratio=2, in_channels=3,
                                                           patch_size=2, temporal_patch_size=2, spatial_merge_size=2))
    model = Qwen2VLForConditionalGeneration(cfg)
    lin = W.linear_weight_keys(model)
    assert "model.embed_tokens.weight" not in lin and "lm_head.weight" in lin
    assert "visual.patch_embed.proj.weight" not in lin and "model.layers.0.self_attn.q_proj.weight" in lin


def test_safetensors_interop_with_the_reference_library(tmp_path):
    """Cross-check with the `safetensors` wheel when it is installed (it is in this image; the product does not depend
    on it): archives written by either side are read identically by the other, incl. bf16, metadata and 0-d tensors."""
    st = pytest.importorskip("safetensors.torch")
    g = torch.Generator().manual_seed(1)
    t = {"lin.weight": torch.randn(6, 10, generator=g), "lin.bias": torch.randn(10, generator=g).to(torch.bfloat16),
         "conv.weight": torch.randn(4, 3, 3, 3, generator=g).to(torch.float16), "step": torch.tensor(7), "mask": torch.tensor([True, False, True])}
    theirs, ours = str(tmp_path / "theirs.safetensors"), str(tmp_path / "ours.safetensors")
    st.save_file(t, theirs, metadata={"format": "pt"})
    sd, meta = W.read_safetensors(theirs)
    assert meta == {"format": "pt"} and set(sd) == set(t)
    for k in t:
        assert sd[k].dtype == t[k].dtype and torch.equal(sd[k], t[k]), k
    W.write_safetensors(ours, t, {"format": "pd"})
    back = st.load_file(ours)
    for k in t:
        assert back[k].dtype == t[k].dtype and torch.equal(back[k], t[k]), k
    from safetensors import safe_open
    with safe_open(ours, framework="pt") as f:
        assert f.metadata() == {"format": "pd"}
