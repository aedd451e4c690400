"""Host scheduler math of the product (numpy fp32) against the oracle (torch fp32 restatement of the reference):
bit-exact timesteps, alphas_cumprod and per-step scalars."""
import numpy as np
import pytest
import torch

from oracle import schedulers as O
from paddlemix_b200.ppdiffusers import schedulers as S

SD = dict(beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear", clip_sample=False, set_alpha_to_one=False,
          steps_offset=1)


@pytest.mark.parametrize("kw", [SD, dict(), dict(beta_schedule="squaredcos_cap_v2"), dict(SD, timestep_spacing="trailing"),
                                dict(SD, set_alpha_to_one=True, steps_offset=0)])
@pytest.mark.parametrize("n", [5, 10, 50])
def test_ddim_bit_exact(kw, n):
    o, s = O.DDIMScheduler(**kw), S.DDIMScheduler(**kw)
    assert np.array_equal(o.betas.numpy(), s.betas)
    assert np.array_equal(o.alphas_cumprod.numpy(), s.alphas_cumprod)
    o.set_timesteps(n), s.set_timesteps(n)
    assert o.timesteps.tolist() == s.timesteps.tolist()
    for t in s.timesteps:
        assert o.step_scalars(int(t)) == s.step_scalars(int(t)), (kw, int(t))


def test_ddim_scalars_reproduce_step():
    # applying the four scalars in the kernel's operation order == the oracle's step() on fp32 tensors, bit for bit
    o = O.DDIMScheduler(**SD)
    o.set_timesteps(50)
    g = torch.Generator().manual_seed(0)
    x, e = torch.randn(2, 4, 16, 16, generator=g), torch.randn(2, 4, 16, 16, generator=g)
    for t in o.timesteps[:5]:
        sa_t, sb_t, sa_p, sb_p = (torch.tensor(v, dtype=torch.float32) for v in o.step_scalars(int(t)))
        mine = sa_p * ((x - sb_t * e) / sa_t) + sb_p * e
        assert torch.equal(mine, o.step(e, t, x))


def test_sd_timesteps():
    s = S.DDIMScheduler(**SD)
    s.set_timesteps(50)
    assert s.timesteps[0] == 981 and s.timesteps[-1] == 1 and len(s.timesteps) == 50


@pytest.mark.parametrize("shift", [1.0, 3.0])
def test_flow_match_bit_exact(shift):
    o, s = O.FlowMatchEulerDiscreteScheduler(shift=shift), S.FlowMatchEulerDiscreteScheduler(shift=shift)
    assert np.array_equal(o.sigmas.numpy(), s.sigmas)
    o.set_timesteps(28), s.set_timesteps(28)
    assert np.array_equal(o.sigmas.numpy(), s.sigmas) and np.array_equal(o.timesteps.numpy(), s.timesteps)
    g = torch.Generator().manual_seed(0)
    x, v = torch.randn(2, 16, 8, 8, generator=g), torch.randn(2, 16, 8, 8, generator=g)
    for t in s.timesteps[:4]:
        sigma, dt = (torch.tensor(z, dtype=torch.float32) for z in s.step_scalars(t))
        den = x - v * sigma
        mine = x + (x - den) / sigma * dt
        ref = o.step(v, torch.tensor(t), x)
        assert torch.equal(mine, ref)
        x = ref


SDXL_EULER = dict(beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear", timestep_spacing="leading", steps_offset=1)


@pytest.mark.parametrize("kw", [dict(num_train_timesteps=1100), SDXL_EULER, dict(SDXL_EULER, use_karras_sigmas=True),
                                dict(timestep_spacing="trailing"), dict(interpolation_type="log_linear"),
                                dict(use_karras_sigmas=True, sigma_min=0.02, sigma_max=700.0)])
@pytest.mark.parametrize("n", [10, 30])
def test_euler_discrete_bit_exact(kw, n):
    """Host EulerDiscreteScheduler (numpy) == oracle (torch fp32 restatement of scheduling_euler_discrete.py): sigmas,
    timesteps, init_noise_sigma, the scale_model_input denominator and the (sigma, dt) handed to the step kernel."""
    o, s = O.EulerDiscreteScheduler(**kw), S.EulerDiscreteScheduler(**kw)
    assert np.array_equal(o.sigmas.numpy(), s.sigmas) and np.array_equal(o.timesteps.numpy(), s.timesteps)
    o.set_timesteps(n), s.set_timesteps(n)
    assert np.array_equal(o.sigmas.numpy(), s.sigmas) and np.array_equal(o.timesteps.numpy(), s.timesteps)
    assert float(o.init_noise_sigma) == s.init_noise_sigma
    g = torch.Generator().manual_seed(0)
    x, e = torch.randn(2, 4, 8, 8, generator=g) * s.init_noise_sigma, torch.randn(2, 4, 8, 8, generator=g)
    for t in s.timesteps[:5]:
        den = torch.tensor(s.input_scale_denominator(t), dtype=torch.float32)
        xs = o.scale_model_input(x, torch.tensor(t))
        assert torch.equal(x / den, xs)
        assert o.step_scalars() == tuple(np.float32(v) for v in s.step_scalars(t))
        sigma, dt = (torch.tensor(v, dtype=torch.float32) for v in (s.sigmas[s._step_index - 1], s.sigmas[s._step_index] - s.sigmas[s._step_index - 1]))
        mine = x + (x - (x - e * sigma)) / sigma * dt  # b200mix_euler_step's operation order
        x = o.step(e, torch.tensor(t), x)
        assert torch.equal(mine, x)


def test_euler_rejects_what_has_no_device_path():
    with pytest.raises(NotImplementedError):
        S.EulerDiscreteScheduler(prediction_type="v_prediction")
    s = S.EulerDiscreteScheduler()
    s.set_timesteps(5)
    with pytest.raises(ValueError):
        s.step_scalars(3)


SD_DPM = dict(beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear", timestep_spacing="leading", steps_offset=1)


@pytest.mark.parametrize("kw", [dict(), SD_DPM, dict(SD_DPM, use_karras_sigmas=True), dict(use_lu_lambdas=True),
                                dict(solver_order=1), dict(lower_order_final=False), dict(euler_at_final=True),
                                dict(timestep_spacing="trailing"), dict(beta_schedule="squaredcos_cap_v2", lambda_min_clipped=-5.1)])
@pytest.mark.parametrize("n", [10, 25])
def test_dpm_solver_pp_2m_bit_exact(kw, n):
    """Host DPMSolverMultistepScheduler (numpy) == oracle: sigmas, timesteps, and the per-step scalars applied in the
    kernel's operation order reproduce the oracle's step() bit for bit through a whole trajectory (orders 1 and 2,
    lower_order_final / euler_at_final switching)."""
    o, s = O.DPMSolverMultistepScheduler(**kw), S.DPMSolverMultistepScheduler(**kw)
    o.set_timesteps(n), s.set_timesteps(n)
    assert np.array_equal(o.sigmas.numpy(), s.sigmas) and o.timesteps.tolist() == s.timesteps.tolist()
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 4, 8, 8, generator=g)
    m_prev = None
    orders = []
    for t in s.timesteps:
        e = torch.randn(2, 4, 8, 8, generator=g)
        order, sigma_cur, alpha_cur, A, C, halfC, inv_r0 = (torch.tensor(v, dtype=torch.float32) if i else v
                                                            for i, v in enumerate(s.step_scalars(int(t))))
        orders.append(order)
        m0 = (x - sigma_cur * e) / alpha_cur
        mine = A * x - C * m0
        if order == 2:
            mine = mine - halfC * (inv_r0 * (m0 - m_prev))
        x = o.step(e, torch.tensor(int(t)), x)
        assert torch.equal(mine, x), (kw, int(t), order)
        m_prev = m0
    assert orders[0] == 1
    if kw.get("solver_order", 2) == 2:
        assert 2 in orders
        if kw.get("euler_at_final") or (kw.get("lower_order_final", True) and n < 15):
            assert orders[-1] == 1
        else:
            assert orders[-1] == 2


def test_dpm_solver_rejects_what_has_no_device_path():
    for bad in (dict(algorithm_type="dpmsolver"), dict(solver_type="heun"), dict(solver_order=3), dict(prediction_type="v_prediction"),
                dict(thresholding=True)):
        with pytest.raises(NotImplementedError):
            S.DPMSolverMultistepScheduler(**bad)


LCM = dict(num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear", prediction_type="epsilon")


@pytest.mark.parametrize("kw", [LCM, dict(LCM, prediction_type="v_prediction", clip_sample=True), dict(LCM, set_alpha_to_one=False)])
@pytest.mark.parametrize("n", [1, 4, 10, 50])
def test_lcm_bit_exact(kw, n):
    """Host LCMScheduler (numpy fp32) vs the oracle restatement of scheduling_lcm.py: timesteps and per-step scalars."""
    o, s = O.LCMScheduler(**kw), S.LCMScheduler(**kw)
    assert np.array_equal(o.alphas_cumprod.numpy(), s.alphas_cumprod)
    o.set_timesteps(n), s.set_timesteps(n)
    assert o.timesteps.tolist() == s.timesteps.tolist()
    for i, t in enumerate(s.timesteps):
        o._step_index = s._step_index = i
        assert o.step_scalars(int(t)) == s.step_scalars(int(t)), (kw, int(t))


def test_lcm_scalars_reproduce_step():
    o = O.LCMScheduler(**LCM)
    o.set_timesteps(4)
    g = torch.Generator().manual_seed(0)
    x, e, z = (torch.randn(2, 4, 16, 16, generator=g) for _ in range(3))
    for t in o.timesteps:
        sa_t, sb_t, c_skip, c_out, sa_p, sb_p, last = o.step_scalars(int(t))
        f = lambda v: torch.tensor(v, dtype=torch.float32)  # noqa: E731
        den = f(c_out) * ((x - f(sb_t) * e) / f(sa_t)) + f(c_skip) * x
        mine = den if last else f(sa_p) * den + f(sb_p) * z
        ref, den_ref = o.step(e, t, x, noise=z)
        assert torch.equal(mine, ref) and torch.equal(den, den_ref)
        x = ref


def test_lcm_errors_mirror_reference():
    s = S.LCMScheduler(**LCM)
    with pytest.raises(ValueError, match="exactly one"):
        s.set_timesteps()
    with pytest.raises(ValueError, match="Can only pass one"):
        s.set_timesteps(num_inference_steps=2, timesteps=[999, 499])
    with pytest.raises(ValueError, match="descending"):
        s.set_timesteps(timesteps=[499, 999])
    s.set_timesteps(timesteps=[999, 759, 499, 259])
    assert s.timesteps.tolist() == [999, 759, 499, 259] and s.num_inference_steps == 4
    with pytest.raises(ValueError):
        S.DDIMScheduler(prediction_type="nope")
