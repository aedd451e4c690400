"""STDiT2 (Open-Sora) parity on the GPU against the CPU fp32 oracle (oracle/stdit2.py); bf16-vs-fp32 tolerance as in
tests/test_unet_gpu.py."""
import pytest
import torch

from oracle import stdit2 as O

pytestmark = pytest.mark.gpu
bf16 = torch.bfloat16


@pytest.mark.parametrize("use_mask", [False, True])
def test_stdit2_tiny_parity(use_mask):
    from paddlemix_b200.opensora import STDiT2
    cfg = O.STDIT2_CONFIGS["tiny"]
    P = O.init_stdit2_params(cfg, seed=1)
    model = STDiT2(cfg).load_state_dict(P, device=0)
    assert model.state_dict_shapes() == O.stdit2_param_shapes(cfg)
    g = torch.Generator().manual_seed(0)
    B, T, H = 2, 4, 16
    x = torch.randn(B, 4, T, H, H, generator=g).to(bf16).float()
    y = torch.randn(B, 1, 12, cfg["caption_channels"], generator=g).to(bf16).float()
    kw = dict(num_frames=torch.tensor([4., 4.]), height=torch.tensor([128., 128.]), width=torch.tensor([128., 128.]),
              ar=torch.tensor([1., 1.]), fps=torch.tensor([24., 24.]))
    mask = None
    if use_mask:
        mask = torch.ones(B, 12, dtype=torch.long)
        mask[1, 7:] = 0
    ts = torch.tensor([500., 500.])
    ref = O.stdit2_forward(cfg, P, x, ts, y, mask, **kw)
    out = model(x.cuda(), ts.cuda(), y.cuda(), mask=None if mask is None else mask.cuda(), **kw)
    assert out.shape == ref.shape and out.dtype == torch.float32
    o = out.cpu()
    cos = torch.nn.functional.cosine_similarity(o.flatten(), ref.flatten(), dim=0).item()
    err = (o - ref).abs().max().item() / ref.abs().max().item()
    assert cos >= 0.999 and err <= 0.04, (cos, err)
