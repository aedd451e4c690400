"""AutoencoderKL.decode on the GPU against the CPU fp32 oracle (oracle/vae.py), incl. the row-softmax kernel and the
one-head mid-block attention built from GEMMs. Tolerance: cosine >= 0.999, max |err| <= 4 % of the output range."""
import pytest
import torch

from oracle import vae as OV

pytestmark = pytest.mark.gpu
bf16 = torch.bfloat16
KEYS = ("latent_channels", "out_channels", "block_out_channels", "layers_per_block", "norm_num_groups", "scaling_factor")


def test_softmax_rows():
    from paddlemix_b200 import ops
    ops.init(0)
    g = torch.Generator().manual_seed(0)
    for M, N, scale in ((7, 64, 1.0), (300, 4096, 0.044), (64, 16384, 0.0442), (3, 51200, 0.1)):
        x = (torch.randn(M, N, generator=g) * 6).cuda()
        x[0, : N // 2] = float("-inf")
        y = ops.softmax_rows(x, scale=scale)
        ref = torch.softmax(x.double() * scale, -1)
        assert y.dtype == bf16 and (y.float() - ref.float()).abs().max().item() <= 2 ** -8 * ref.max().item() + 1e-6
        assert torch.allclose(y.float().sum(-1), torch.ones(M, device="cuda"), atol=2e-2)
        assert torch.equal(y, ops.softmax_rows(x, scale=scale))


@pytest.mark.parametrize("name,h,B", [("tiny", 16, 2), ("tiny", 40, 1), ("sd_vae", 32, 1)])
def test_vae_decode_parity(name, h, B):
    from paddlemix_b200.ppdiffusers.autoencoder_kl import AutoencoderKL
    cfg = OV.VAE_CONFIGS[name]
    P = OV.init_vae_params(cfg, seed=1)
    vae = AutoencoderKL(**{k: cfg[k] for k in KEYS}).load_state_dict(P, device=0)
    assert vae.state_dict_shapes() == OV.vae_decoder_param_shapes(cfg)
    z = torch.randn(B, 4, h, h, generator=torch.Generator().manual_seed(2)).to(bf16).float()
    up = 2 ** (len(cfg["block_out_channels"]) - 1)
    out = vae.decode(z.cuda()).sample
    with torch.no_grad():
        ref = OV.vae_decode(cfg, P, z)
    assert out.shape == ref.shape == (B, 3, h * up, h * up) and out.dtype == torch.float32
    o = out.cpu()
    cos = torch.nn.functional.cosine_similarity(o.flatten().double(), ref.flatten().double(), dim=0).item()
    err = (o - ref).abs().max().item() / ref.abs().max().item()
    assert cos >= 0.999 and err <= 0.04, (name, cos, err)
    assert torch.equal(vae.decode(z.cuda(), return_dict=False)[0], out)  # determinism


@pytest.mark.slow
def test_sdxl_vae_decode_1024_vs_oracle():
    """The real decoder (49.5 M params) on a 128x128 latent -> 1024x1024 image, B = 1: 10.5 TFLOP incl. the 16384-token
    one-head attention (1 GB fp32 score scratch)."""
    from paddlemix_b200.ppdiffusers.autoencoder_kl import AutoencoderKL
    cfg = OV.VAE_CONFIGS["sdxl_vae"]
    P = OV.init_vae_params(cfg, seed=3)
    vae = AutoencoderKL(**{k: cfg[k] for k in KEYS}).load_state_dict(P, device=0)
    z = torch.randn(1, 4, 128, 128, generator=torch.Generator().manual_seed(4)).to(bf16).float()
    out = vae.decode(z.cuda()).sample.cpu()
    with torch.no_grad():
        ref = OV.vae_decode(cfg, P, z)
    cos = torch.nn.functional.cosine_similarity(out.flatten().double(), ref.flatten().double(), dim=0).item()
    err = (out - ref).abs().max().item() / ref.abs().max().item()
    assert out.shape == (1, 3, 1024, 1024) and cos >= 0.999 and err <= 0.04, (cos, err)


def test_pipeline_decodes_to_pixels():
    """StableDiffusionPipeline(output_type='pt'): loop -> latents / scaling_factor -> vae.decode -> [0, 1] image."""
    from oracle import unet as OU
    from paddlemix_b200.ppdiffusers.autoencoder_kl import AutoencoderKL
    from paddlemix_b200.ppdiffusers.pipelines import StableDiffusionPipeline
    from paddlemix_b200.ppdiffusers.schedulers import DDIMScheduler
    from paddlemix_b200.ppdiffusers.unet_2d_condition import UNet2DConditionModel
    ucfg = OU.UNET_CONFIGS["tiny_sd"]
    keys = ("in_channels", "out_channels", "down_block_types", "up_block_types", "block_out_channels", "layers_per_block",
            "cross_attention_dim", "transformer_layers_per_block", "attention_head_dim", "use_linear_projection")
    unet = UNet2DConditionModel(**{k: ucfg[k] for k in keys}).load_state_dict(OU.init_params(OU.unet_param_shapes(ucfg)), device=0)
    vcfg = OV.VAE_CONFIGS["tiny"]
    Pv = OV.init_vae_params(vcfg, seed=5)
    vae = AutoencoderKL(**{k: vcfg[k] for k in KEYS}).load_state_dict(Pv, device=0)
    SD = dict(beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear", clip_sample=False, set_alpha_to_one=False, steps_offset=1)
    pipe = StableDiffusionPipeline(unet, DDIMScheduler(**SD), use_cuda_graph=False, vae=vae)
    g = torch.Generator().manual_seed(6)
    emb = torch.randn(1, 77, ucfg["cross_attention_dim"], generator=g)
    lat0 = torch.randn(1, 4, 16, 16, generator=g)
    lat = pipe(prompt_embeds=emb, latents=lat0, num_inference_steps=3, guidance_scale=5.0, output_type="latent")
    img = pipe(prompt_embeds=emb, latents=lat0, num_inference_steps=3, guidance_scale=5.0, output_type="pt")
    assert img.shape == (1, 3, 32, 32) and img.min() >= 0 and img.max() <= 1
    ref = (OV.vae_decode(vcfg, Pv, lat.cpu() / vcfg["scaling_factor"]) / 2 + 0.5).clamp(0, 1)
    assert (img.cpu() - ref).abs().max().item() <= 0.03
