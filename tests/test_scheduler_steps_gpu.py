"""Device scheduler steps added in round 2, through the C ABI, against the CPU oracle (torch fp32 restatement of the
reference): DDIM for every prediction_type / clip_sample (scheduling_ddim.py:424-452), LCMScheduler.step
(scheduling_lcm.py:461-549) and rescale_noise_cfg (pipeline_stable_diffusion.py:69-80).
DDIM / LCM: bit-exact (each fp32 operation is rounded individually in the reference's order).
rescale_noise_cfg: the per-sample std is a reduction (order differs from Paddle's): relative 1e-5."""
import pytest
import torch

from oracle import schedulers as O

pytestmark = pytest.mark.gpu
bf16 = torch.bfloat16
SD = dict(beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear", set_alpha_to_one=False, steps_offset=1)


@pytest.mark.parametrize("kw", [dict(SD, clip_sample=False, prediction_type="v_prediction"),
                                dict(SD, clip_sample=False, prediction_type="sample"),
                                dict(SD, clip_sample=True, prediction_type="epsilon"),
                                dict(SD, clip_sample=True, clip_sample_range=0.7, prediction_type="v_prediction"),
                                dict(clip_sample=True)])  # the reference's own test config (test_scheduler_ddim.py:25-35)
def test_ddim_step_all_prediction_types_bit_exact(kw):
    from paddlemix_b200.ppdiffusers.schedulers import DDIMScheduler
    o, s = O.DDIMScheduler(**kw), DDIMScheduler(**kw)
    o.set_timesteps(10), s.set_timesteps(10)
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 4, 32, 32, generator=g)
    for t in s.timesteps[:4]:
        m = torch.randn(2, 4, 32, 32, generator=g)
        nxt = s.step(m.cuda(), int(t), x.cuda())
        ref = o.step(m, t, x)
        assert torch.equal(nxt.cpu(), ref), (kw, int(t), (nxt.cpu() - ref).abs().max().item())
        x = ref
    # fused CFG combine in the same kernel: eps = u + g*(c - u), computed in fp32 exactly like the pipeline does
    u, c = torch.randn(2, 4, 32, 32, generator=g), torch.randn(2, 4, 32, 32, generator=g)
    t = s.timesteps[4]
    nxt = s.step(u.cuda(), int(t), x.cuda(), model_output_cond=c.cuda(), guidance_scale=7.5)
    assert torch.equal(nxt.cpu(), o.step(u + 7.5 * (c - u), t, x))


def test_ddim_full_loop_goldens_on_device():
    """The reference's RNG-free full-loop goldens (test_scheduler_ddim.py:128-190) reproduced through the DEVICE step:
    default config 172.0067 / 0.223967 and v_prediction 52.5302 / 0.0684."""
    import json
    import os
    from paddlemix_b200.ppdiffusers.schedulers import DDIMScheduler
    gold = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ddim_goldens.json")))
    n = 4 * 3 * 8 * 8
    deter = (torch.arange(n).reshape(3, 8, 8, 4) / n).permute(3, 0, 1, 2).contiguous()
    for case in gold["full_loop"]:
        s = DDIMScheduler(**dict(gold["config"], **case["config"]))
        s.set_timesteps(gold["num_inference_steps"])
        x = deter.cuda()
        for t in s.timesteps:
            x = s.step((x * float(t) / (float(t) + 1)).contiguous(), int(t), x)
        assert abs(x.abs().sum().item() - case["sum"]) < gold["sum_atol"], case
        assert abs(x.abs().mean().item() - case["mean"]) < gold["mean_atol"], case


@pytest.mark.parametrize("kw", [dict(), dict(prediction_type="v_prediction", clip_sample=True), dict(prediction_type="sample")])
@pytest.mark.parametrize("n", [1, 4])
def test_lcm_step_bit_exact(kw, n):
    from paddlemix_b200.ppdiffusers.schedulers import LCMScheduler
    o, s = O.LCMScheduler(**kw), LCMScheduler(**kw)
    o.set_timesteps(n), s.set_timesteps(n)
    g = torch.Generator().manual_seed(1)
    x = torch.randn(2, 4, 32, 32, generator=g)
    for t in s.timesteps:
        m, z = torch.randn(2, 4, 32, 32, generator=g), torch.randn(2, 4, 32, 32, generator=g)
        nxt, den = s.step(m.cuda(), int(t), x.cuda(), noise=z.cuda(), return_denoised=True)
        ref, den_ref = o.step(m, t, x, noise=z)
        assert torch.equal(nxt.cpu(), ref) and torch.equal(den.cpu(), den_ref), (kw, int(t))
        x = ref


def test_lcm_one_step_golden_on_device():
    """test_scheduler_lcm.py:239-247 (RNG-free): one-step full loop through the device kernel: 18.7097 / 0.0244."""
    from paddlemix_b200.ppdiffusers.schedulers import LCMScheduler
    s = LCMScheduler(num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear")
    s.set_timesteps(1)
    n = 4 * 3 * 8 * 8
    x = (torch.arange(n).reshape(3, 8, 8, 4) / n).permute(3, 0, 1, 2).contiguous().cuda()
    for t in s.timesteps:
        x = s.step((x * float(t) / (float(t) + 1)).contiguous(), int(t), x)
    assert abs(x.abs().sum().item() - 18.7097) < 1e-3 and abs(x.abs().mean().item() - 0.0244) < 1e-3


@pytest.mark.parametrize("dtype", [torch.float32, bf16])
@pytest.mark.parametrize("gr", [0.0, 0.7, 1.0])
def test_rescale_noise_cfg(dtype, gr):
    from paddlemix_b200 import ops
    ops.init(0)
    g = torch.Generator().manual_seed(2)
    B = 3
    u = (torch.randn(B, 4, 128, 128, generator=g) * 0.8).to(dtype)
    c = (torch.randn(B, 4, 128, 128, generator=g) * torch.tensor([0.5, 1.0, 2.0]).reshape(B, 1, 1, 1) + 0.1).to(dtype)
    out = ops.cfg_combine(u.cuda(), c.cuda(), 7.5, guidance_rescale=gr).cpu()
    cfg = u.float() + 7.5 * (c.float() - u.float())
    ref = O.rescale_noise_cfg(cfg, c.float(), gr) if gr > 0 else cfg
    assert out.dtype == torch.float32
    assert (out - ref).abs().max().item() <= 1e-5 * ref.abs().max().item() + 1e-6
    if gr == 1.0:  # fully rescaled: the per-sample std equals the text branch's
        assert torch.allclose(out.flatten(1).std(1), c.float().flatten(1).std(1), rtol=1e-4)
