"""Schedulers on the hot path, mirroring ppdiffusers' interface (set_timesteps / scale_model_input / step).

The index and scalar math runs on the host in numpy float32, operation by operation as the reference does on 0-d
fp32 tensors (ppdiffusers/schedulers/scheduling_ddim.py:205-238, 305-348, 410-457;
scheduling_flow_match_euler_discrete.py:64-83, 140-163, 244-275); the elementwise update of the latent runs in one
fused CUDA kernel (b200mix_ddim_step / b200mix_euler_step) that also folds in the classifier-free-guidance combine.
tests/test_schedulers.py checks the host scalars bit-for-bit against the oracle.
"""
import math
from typing import Optional

import numpy as np

f32 = np.float32


def _betas_for_alpha_bar(n, max_beta=0.999):
    def alpha_bar(t):
        return math.cos((t + 0.008) / 1.008 * math.pi / 2) ** 2
    return np.array([min(1 - alpha_bar((i + 1) / n) / alpha_bar(i / n), max_beta) for i in range(n)], dtype=f32)


class DDIMScheduler:
    """ppdiffusers.DDIMScheduler (scheduling_ddim.py:131) restricted to the deterministic sampling path."""

    order = 1
    init_noise_sigma = 1.0

    def __init__(self, num_train_timesteps: int = 1000, beta_start: float = 0.0001, beta_end: float = 0.02,
                 beta_schedule: str = "linear", clip_sample: bool = True, set_alpha_to_one: bool = True,
                 steps_offset: int = 0, prediction_type: str = "epsilon", timestep_spacing: str = "leading",
                 clip_sample_range: float = 1.0, thresholding: bool = False):
        if thresholding:
            raise NotImplementedError("dynamic thresholding (pixel-space models) is outside the latent-diffusion hot path")
        if prediction_type not in ("epsilon", "sample", "v_prediction"):
            raise ValueError(f"prediction_type given as {prediction_type} must be one of `epsilon`, `sample`, or"
                             " `v_prediction`")
        self.clip_sample_range = clip_sample_range
        if beta_schedule == "linear":
            self.betas = _linspace_f32(beta_start, beta_end, num_train_timesteps)
        elif beta_schedule == "scaled_linear":
            self.betas = _linspace_f32(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps) ** 2
        elif beta_schedule == "squaredcos_cap_v2":
            self.betas = _betas_for_alpha_bar(num_train_timesteps)
        else:
            raise NotImplementedError(f"{beta_schedule} does is not implemented for {self.__class__}")
        self.alphas = (f32(1.0) - self.betas).astype(f32)
        self.alphas_cumprod = _cumprod_f32(self.alphas)
        self.final_alpha_cumprod = f32(1.0) if set_alpha_to_one else self.alphas_cumprod[0]
        self.num_train_timesteps = num_train_timesteps
        self.clip_sample = clip_sample
        self.steps_offset = steps_offset
        self.prediction_type = prediction_type
        self.timestep_spacing = timestep_spacing
        self.num_inference_steps: Optional[int] = None
        self.timesteps = np.arange(0, num_train_timesteps)[::-1].copy().astype(np.int64)

    def scale_model_input(self, sample, timestep=None):
        return sample

    def set_timesteps(self, num_inference_steps: int):
        if num_inference_steps > self.num_train_timesteps:
            raise ValueError(
                f"`num_inference_steps`: {num_inference_steps} cannot be larger than `self.config.train_timesteps`:"
                f" {self.num_train_timesteps} as the unet model trained with this scheduler can only handle"
                f" maximal {self.num_train_timesteps} timesteps.")
        self.num_inference_steps = num_inference_steps
        if self.timestep_spacing == "linspace":
            t = np.linspace(0, self.num_train_timesteps - 1, num_inference_steps).round()[::-1].copy().astype(np.int64)
        elif self.timestep_spacing == "leading":
            step_ratio = self.num_train_timesteps // num_inference_steps
            t = (np.arange(0, num_inference_steps) * step_ratio).round()[::-1].copy().astype(np.int64)
            t += self.steps_offset
        elif self.timestep_spacing == "trailing":
            step_ratio = self.num_train_timesteps / num_inference_steps
            t = np.round(np.arange(self.num_train_timesteps, 0, -step_ratio)).astype(np.int64)
            t -= 1
        else:
            raise ValueError(f"{self.timestep_spacing} is not supported. Please make sure to choose one of 'leading' "
                             "or 'trailing'.")
        self.timesteps = t

    def _get_variance(self, timestep, prev_timestep):
        a_t = self.alphas_cumprod[timestep]
        a_p = self.alphas_cumprod[prev_timestep] if prev_timestep >= 0 else self.final_alpha_cumprod
        b_t, b_p = f32(1) - a_t, f32(1) - a_p
        return f32(f32(b_p / b_t) * f32(f32(1) - f32(a_t / a_p)))

    def step_scalars(self, timestep: int):
        """(sqrt(a_t), sqrt(1-a_t), sqrt(a_prev), sqrt(1-a_prev-std^2)) as fp32, eta = 0 (:410-457)."""
        if self.num_inference_steps is None:
            raise ValueError("Number of inference steps is 'None', you need to run 'set_timesteps' after creating the "
                             "scheduler")
        timestep = int(timestep)
        prev_timestep = timestep - self.num_train_timesteps // self.num_inference_steps
        a_t = self.alphas_cumprod[timestep]
        a_p = self.alphas_cumprod[prev_timestep] if prev_timestep >= 0 else self.final_alpha_cumprod
        b_t = f32(f32(1) - a_t)
        var = self._get_variance(timestep, prev_timestep)
        std = f32(f32(0.0) * np.sqrt(var, dtype=f32))
        dir_coef = np.sqrt(f32(f32(f32(1) - a_p) - f32(std * std)), dtype=f32)
        return float(np.sqrt(a_t, dtype=f32)), float(np.sqrt(b_t, dtype=f32)), float(np.sqrt(a_p, dtype=f32)), float(dir_coef)

    def step(self, model_output, timestep, sample, eta: float = 0.0, model_output_cond=None, guidance_scale=0.0,
             out=None):
        """sample, result: fp32 CUDA tensors; model_output: bf16/fp32 CUDA tensor of the same numel (any layout that
        matches `sample` elementwise). With model_output_cond the CFG combine is fused into the same kernel."""
        if eta != 0.0:
            raise NotImplementedError("b200mix DDIM step covers eta = 0 (the deterministic sampler the pipelines use)")
        from .. import ops
        sa_t, sb_t, sa_p, sb_p = self.step_scalars(timestep)
        if self.prediction_type == "epsilon" and not self.clip_sample:  # the Stable Diffusion configuration
            return ops.ddim_step(model_output, model_output_cond, guidance_scale, sample, sa_t, sb_t, sa_p, sb_p, out=out)
        # v_prediction (SD 2.x), sample prediction, clip_sample (scheduling_ddim.py:424-452)
        return ops.ddim_step_ex(model_output, model_output_cond, guidance_scale, sample, sa_t, sb_t, sa_p, sb_p,
                                self.prediction_type, self.clip_sample_range if self.clip_sample else 0.0, out=out)


class LCMScheduler:
    """ppdiffusers.LCMScheduler (scheduling_lcm.py:133-560) for latent consistency models: host-side fp32 scalars
    (numpy, bit-identical to the reference's 0-d fp32 tensors), one fused device kernel per step
    (b200mix_lcm_step: CFG combine, x0, clip, boundary-condition blend, noise injection)."""

    order = 1
    init_noise_sigma = 1.0

    def __init__(self, num_train_timesteps: int = 1000, beta_start: float = 0.00085, beta_end: float = 0.012,
                 beta_schedule: str = "scaled_linear", original_inference_steps: int = 50, clip_sample: bool = False,
                 clip_sample_range: float = 1.0, set_alpha_to_one: bool = True, steps_offset: int = 0,
                 prediction_type: str = "epsilon", thresholding: bool = False, timestep_spacing: str = "leading",
                 timestep_scaling: float = 10.0, rescale_betas_zero_snr: bool = False):
        if thresholding or rescale_betas_zero_snr:
            raise NotImplementedError("LCMScheduler(b200): thresholding / rescale_betas_zero_snr are outside the hot path")
        if prediction_type not in ("epsilon", "sample", "v_prediction"):
            raise ValueError(f"prediction_type given as {prediction_type} must be one of `epsilon`, `sample` or"
                             " `v_prediction` for `LCMScheduler`.")
        if beta_schedule == "linear":
            self.betas = _linspace_f32(beta_start, beta_end, num_train_timesteps)
        elif beta_schedule == "scaled_linear":
            self.betas = _linspace_f32(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps) ** 2
        elif beta_schedule == "squaredcos_cap_v2":
            self.betas = _betas_for_alpha_bar(num_train_timesteps)
        else:
            raise NotImplementedError(f"{beta_schedule} does is not implemented for {self.__class__}")
        self.alphas = (f32(1.0) - self.betas).astype(f32)
        self.alphas_cumprod = _cumprod_f32(self.alphas)
        self.final_alpha_cumprod = f32(1.0) if set_alpha_to_one else self.alphas_cumprod[0]
        self.num_train_timesteps, self.original_inference_steps = num_train_timesteps, original_inference_steps
        self.clip_sample, self.clip_sample_range = clip_sample, clip_sample_range
        self.prediction_type, self.timestep_scaling = prediction_type, timestep_scaling
        self.num_inference_steps: Optional[int] = None
        self.timesteps = np.arange(0, num_train_timesteps)[::-1].copy().astype(np.int64)
        self._step_index = None

    @property
    def step_index(self):
        return self._step_index

    def scale_model_input(self, sample, timestep=None):
        return sample

    def set_timesteps(self, num_inference_steps: Optional[int] = None, original_inference_steps: Optional[int] = None,
                      timesteps=None, strength: float = 1.0):
        if num_inference_steps is None and timesteps is None:
            raise ValueError("Must pass exactly one of `num_inference_steps` or `custom_timesteps`.")
        if num_inference_steps is not None and timesteps is not None:
            raise ValueError("Can only pass one of `num_inference_steps` or `custom_timesteps`.")
        original_steps = original_inference_steps if original_inference_steps is not None else self.original_inference_steps
        if original_steps > self.num_train_timesteps:
            raise ValueError(f"`original_steps`: {original_steps} cannot be larger than `self.config.train_timesteps`:"
                             f" {self.num_train_timesteps} as the unet model trained with this scheduler can only handle"
                             f" maximal {self.num_train_timesteps} timesteps.")
        k = self.num_train_timesteps // original_steps
        origin = np.asarray(list(range(1, int(original_steps * strength) + 1))) * k - 1
        if timesteps is not None:
            for i in range(1, len(timesteps)):
                if timesteps[i] >= timesteps[i - 1]:
                    raise ValueError("`custom_timesteps` must be in descending order.")
            if timesteps[0] >= self.num_train_timesteps:
                raise ValueError(f"`timesteps` must start before `self.config.train_timesteps`: {self.num_train_timesteps}.")
            ts = np.array(timesteps, dtype=np.int64)
            self.num_inference_steps = len(ts)
            init_timestep = min(int(self.num_inference_steps * strength), self.num_inference_steps)
            ts = ts[max(self.num_inference_steps - init_timestep, 0) * self.order:]
        else:
            if num_inference_steps > self.num_train_timesteps:
                raise ValueError(f"`num_inference_steps`: {num_inference_steps} cannot be larger than "
                                 f"`self.config.train_timesteps`: {self.num_train_timesteps}")
            if len(origin) // num_inference_steps < 1:
                raise ValueError(f"The combination of `original_steps x strength`: {original_steps} x {strength} is smaller "
                                 f"than `num_inference_steps`: {num_inference_steps}.")
            if num_inference_steps > original_steps:
                raise ValueError(f"`num_inference_steps`: {num_inference_steps} cannot be larger than "
                                 f"`original_inference_steps`: {original_steps}")
            self.num_inference_steps = num_inference_steps
            origin = origin[::-1].copy()
            idx = np.floor(np.linspace(0, len(origin), num=num_inference_steps, endpoint=False)).astype(np.int64)
            ts = origin[idx]
        self.timesteps = ts.astype(np.int64)
        self._step_index = None

    def _init_step_index(self, timestep):
        cand = np.nonzero(self.timesteps == int(timestep))[0]
        self._step_index = int(cand[1] if len(cand) > 1 else cand[0])

    def get_scalings_for_boundary_condition_discrete(self, timestep):
        """scheduling_lcm.py:453-459 in fp32 (an int64 0-d tensor times a Python float is an fp32 tensor in Paddle)."""
        sd2 = f32(0.5 ** 2)
        st = f32(f32(int(timestep)) * f32(self.timestep_scaling))
        den = f32(f32(st * st) + sd2)
        return f32(sd2 / den), f32(st / np.sqrt(den, dtype=f32))

    def step_scalars(self, timestep):
        """(sqrt(a_t), sqrt(1-a_t), c_skip, c_out, sqrt(a_prev), sqrt(1-a_prev), is_last_step) as fp32 values."""
        if self.num_inference_steps is None:
            raise ValueError("Number of inference steps is 'None', you need to run 'set_timesteps' after creating the "
                             "scheduler")
        timestep = int(timestep)
        if self._step_index is None:
            self._init_step_index(timestep)
        nxt = self._step_index + 1
        prev_timestep = int(self.timesteps[nxt]) if nxt < len(self.timesteps) else timestep
        a_t = self.alphas_cumprod[timestep]
        a_p = self.alphas_cumprod[prev_timestep] if prev_timestep >= 0 else self.final_alpha_cumprod
        c_skip, c_out = self.get_scalings_for_boundary_condition_discrete(timestep)
        return (float(np.sqrt(a_t, dtype=f32)), float(np.sqrt(f32(f32(1) - a_t), dtype=f32)), float(c_skip), float(c_out),
                float(np.sqrt(a_p, dtype=f32)), float(np.sqrt(f32(f32(1) - a_p), dtype=f32)),
                self._step_index == self.num_inference_steps - 1)

    def step(self, model_output, timestep, sample, generator=None, model_output_cond=None, guidance_scale=0.0,
             noise=None, out=None, return_denoised=False):
        """sample / result: fp32 CUDA tensors. Multi-step sampling injects z ~ N(0, I) (scheduling_lcm.py:538-545): pass
        `noise` (fp32 CUDA tensor) or a torch `generator` (a CUDA generator of the sample's device); the final step is
        deterministic. With return_denoised also returns LCMSchedulerOutput.denoised."""
        import torch

        from .. import ops
        sa_t, sb_t, c_skip, c_out, sa_p, sb_p, last = self.step_scalars(timestep)
        if not last and noise is None:
            noise = torch.randn(sample.shape, device=sample.device, dtype=torch.float32, generator=generator)
        den = torch.empty_like(sample) if return_denoised else None
        res = ops.lcm_step(model_output, model_output_cond, guidance_scale, sample, None if last else noise, sa_t, sb_t,
                           c_skip, c_out, sa_p, sb_p, self.prediction_type,
                           self.clip_sample_range if self.clip_sample else 0.0, out=out, denoised=den)
        self._step_index += 1
        return (res, den) if return_denoised else res


def _linspace_f32(start, end, n):
    """float32 linspace as the CPU kernels of Paddle / torch define it: start + i*step for the first half,
    end - (n-1-i)*step for the second half, step = (end-start)/(n-1), everything in fp32."""
    start, end = f32(start), f32(end)
    step = f32((end - start) / f32(n - 1))
    i = np.arange(n)
    half = n // 2
    lo = (start + step * i.astype(f32)).astype(f32)
    hi = (end - step * (n - 1 - i).astype(f32)).astype(f32)
    return np.where(i < half, lo, hi).astype(f32)


def _cumprod_f32(a):
    out = np.empty_like(a, dtype=f32)
    acc = f32(1.0)
    for i, v in enumerate(a):
        acc = f32(acc * v)
        out[i] = acc
    return out


class FlowMatchEulerDiscreteScheduler:
    """ppdiffusers.FlowMatchEulerDiscreteScheduler (scheduling_flow_match_euler_discrete.py:44)."""

    order = 1

    def __init__(self, num_train_timesteps: int = 1000, shift: float = 1.0):
        self.num_train_timesteps, self.shift = num_train_timesteps, shift
        t = np.linspace(1, num_train_timesteps, num_train_timesteps, dtype=f32)[::-1].copy()
        sigmas = (t / f32(num_train_timesteps)).astype(f32)
        sigmas = (f32(shift) * sigmas / (f32(1) + f32(shift - 1) * sigmas)).astype(f32)
        self.timesteps = (sigmas * f32(num_train_timesteps)).astype(f32)
        self.sigmas = sigmas
        self.sigma_min, self.sigma_max = float(sigmas[-1]), float(sigmas[0])
        self._step_index = None

    def set_timesteps(self, num_inference_steps: int):
        self.num_inference_steps = num_inference_steps
        t = np.linspace(self.sigma_max * self.num_train_timesteps, self.sigma_min * self.num_train_timesteps,
                        num_inference_steps)
        sigmas = t / self.num_train_timesteps
        sigmas = self.shift * sigmas / (1 + (self.shift - 1) * sigmas)
        sigmas = sigmas.astype(f32)
        self.timesteps = (sigmas * f32(self.num_train_timesteps)).astype(f32)
        self.sigmas = np.concatenate([sigmas, np.zeros(1, dtype=f32)])
        self._step_index = None

    def step_scalars(self, timestep):
        if self._step_index is None:
            idx = np.nonzero(self.timesteps == f32(timestep))[0]
            self._step_index = int(idx[0])
        sigma = self.sigmas[self._step_index]
        dt = f32(self.sigmas[self._step_index + 1] - sigma)
        self._step_index += 1
        return float(sigma), float(dt)

    def step(self, model_output, timestep, sample, model_output_cond=None, guidance_scale=0.0, out=None):
        from .. import ops
        sigma, dt = self.step_scalars(timestep)
        return ops.euler_step(model_output, model_output_cond, guidance_scale, sample, sigma, dt, out=out)


class EulerDiscreteScheduler:
    """ppdiffusers.EulerDiscreteScheduler (scheduling_euler_discrete.py:135-503; SDXL's default sampler) restricted to
    the deterministic path the pipelines use (s_churn = 0 => gamma = 0, sigma_hat = sigma; epsilon prediction). All
    schedule arithmetic runs on the host in numpy with the reference's dtypes (fp32 tensors -> np.float32, its
    np.interp / Karras ramp / _sigma_to_t pieces in float64); the per-step update is the same fp32 expression as the
    reference's (`pred = x - sigma*eps; derivative = (x - pred)/sigma; x + derivative*dt`), which is exactly
    `b200mix_euler_step`; `scale_model_input` is `b200mix_scale_model_input` (x / sqrt(sigma^2 + 1) with an IEEE fp32
    division)."""

    order = 1

    def __init__(self, num_train_timesteps: int = 1000, beta_start: float = 0.0001, beta_end: float = 0.02,
                 beta_schedule: str = "linear", prediction_type: str = "epsilon", interpolation_type: str = "linear",
                 use_karras_sigmas: bool = False, sigma_min: Optional[float] = None, sigma_max: Optional[float] = None,
                 timestep_spacing: str = "linspace", timestep_type: str = "discrete", steps_offset: int = 0):
        if beta_schedule == "linear":
            self.betas = _linspace_f32(beta_start, beta_end, num_train_timesteps)
        elif beta_schedule == "scaled_linear":
            self.betas = _linspace_f32(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps) ** 2
        elif beta_schedule == "squaredcos_cap_v2":
            self.betas = _betas_for_alpha_bar(num_train_timesteps)
        else:
            raise NotImplementedError(f"{beta_schedule} does is not implemented for {self.__class__}")
        if prediction_type != "epsilon":
            raise NotImplementedError("EulerDiscreteScheduler(b200): only prediction_type='epsilon' has a device step")
        if timestep_type != "discrete":
            raise NotImplementedError("EulerDiscreteScheduler(b200): only timestep_type='discrete'")
        self.alphas = (f32(1.0) - self.betas).astype(f32)
        self.alphas_cumprod = _cumprod_f32(self.alphas)
        self.num_train_timesteps = num_train_timesteps
        self.interpolation_type, self.use_karras_sigmas = interpolation_type, use_karras_sigmas
        self.sigma_min, self.sigma_max = sigma_min, sigma_max
        self.timestep_spacing, self.steps_offset = timestep_spacing, steps_offset
        sig = self._train_sigmas()
        self.timesteps = np.linspace(0, num_train_timesteps - 1, num_train_timesteps, dtype=float)[::-1].astype(f32)
        self.sigmas = np.concatenate([sig[::-1], np.zeros(1, dtype=f32)]).astype(f32)
        self.num_inference_steps: Optional[int] = None
        self._step_index: Optional[int] = None

    def _train_sigmas(self):
        ac = self.alphas_cumprod
        return np.sqrt(((f32(1.0) - ac) / ac).astype(f32), dtype=f32)

    @property
    def init_noise_sigma(self) -> float:
        m = self.sigmas.max()
        if self.timestep_spacing in ("linspace", "trailing"):
            return float(m)
        return float(np.sqrt(f32(f32(m * m) + f32(1.0)), dtype=f32))

    def set_timesteps(self, num_inference_steps: int):
        N = self.num_train_timesteps
        self.num_inference_steps = num_inference_steps
        if self.timestep_spacing == "linspace":
            timesteps = np.linspace(0, N - 1, num_inference_steps, dtype=f32)[::-1].copy()
        elif self.timestep_spacing == "leading":
            step_ratio = N // num_inference_steps
            timesteps = (np.arange(0, num_inference_steps) * step_ratio).round()[::-1].copy().astype(f32)
            timesteps += self.steps_offset
        elif self.timestep_spacing == "trailing":
            step_ratio = N / num_inference_steps
            timesteps = (np.arange(N, 0, -step_ratio)).round().copy().astype(f32)
            timesteps -= 1
        else:
            raise ValueError(f"{self.timestep_spacing} is not supported. Please make sure to choose one of 'linspace', "
                             "'leading' or 'trailing'.")
        sigmas = self._train_sigmas()
        log_sigmas = np.log(sigmas)
        if self.interpolation_type == "linear":
            sigmas = np.interp(timesteps, np.arange(0, len(sigmas)), sigmas)
        elif self.interpolation_type == "log_linear":
            sigmas = np.exp(np.linspace(np.log(sigmas[-1]), np.log(sigmas[0]), num_inference_steps + 1, dtype=f32))
        else:
            raise ValueError(f"{self.interpolation_type} is not implemented. Please specify interpolation_type to either"
                             " 'linear' or 'log_linear'")
        if self.use_karras_sigmas:
            smin = self.sigma_min if self.sigma_min is not None else sigmas[-1].item()
            smax = self.sigma_max if self.sigma_max is not None else sigmas[0].item()
            rho = 7.0
            ramp = np.linspace(0, 1, num_inference_steps)
            sigmas = (smax ** (1 / rho) + ramp * (smin ** (1 / rho) - smax ** (1 / rho))) ** rho
            timesteps = np.array([self._sigma_to_t(s, log_sigmas) for s in sigmas])
        self.timesteps = np.asarray(timesteps).astype(f32)
        self.sigmas = np.concatenate([np.asarray(sigmas).astype(f32), np.zeros(1, dtype=f32)])
        self._step_index = None

    @staticmethod
    def _sigma_to_t(sigma, log_sigmas):
        log_sigma = np.log(np.maximum(sigma, 1e-10))
        dists = log_sigma - log_sigmas[:, np.newaxis]
        low_idx = np.cumsum((dists >= 0), axis=0).argmax(axis=0).clip(max=log_sigmas.shape[0] - 2)
        high_idx = low_idx + 1
        low, high = log_sigmas[low_idx], log_sigmas[high_idx]
        w = np.clip((low - log_sigma) / (low - high), 0, 1)
        return ((1 - w) * low_idx + w * high_idx).reshape(np.shape(sigma))

    def _init_step_index(self, timestep):
        cand = np.nonzero(self.timesteps == f32(timestep))[0]
        if len(cand) == 0:
            raise ValueError(f"timestep {timestep} is not one of scheduler.timesteps")
        self._step_index = int(cand[1] if len(cand) > 1 else cand[0])

    def input_scale_denominator(self, timestep) -> float:
        """(sigma^2 + 1) ** 0.5 in fp32: what scale_model_input divides the sample by."""
        if self._step_index is None:
            self._init_step_index(timestep)
        s = self.sigmas[self._step_index]
        return float(np.sqrt(f32(f32(s * s) + f32(1.0)), dtype=f32))

    def scale_model_input(self, sample, timestep, out=None):
        from .. import ops
        return ops.scale_model_input(sample, self.input_scale_denominator(timestep), out=out)

    def step_scalars(self, timestep):
        if isinstance(timestep, (int, np.integer)):
            raise ValueError("Passing integer indices (e.g. from `enumerate(timesteps)`) as timesteps to "
                             "`EulerDiscreteScheduler.step()` is not supported. Make sure to pass one of the "
                             "`scheduler.timesteps` as a timestep.")
        if self._step_index is None:
            self._init_step_index(timestep)
        sigma = self.sigmas[self._step_index]
        dt = f32(self.sigmas[self._step_index + 1] - sigma)
        self._step_index += 1
        return float(sigma), float(dt)

    def step(self, model_output, timestep, sample, model_output_cond=None, guidance_scale=0.0, out=None):
        from .. import ops
        sigma, dt = self.step_scalars(timestep)
        return ops.euler_step(model_output, model_output_cond, guidance_scale, sample, sigma, dt, out=out)

    def add_noise_sigma(self, timestep) -> float:
        """sigma of `timestep`: add_noise is original + noise * sigma (scheduling_euler_discrete.py:472-497)."""
        idx = np.nonzero(self.timesteps == f32(timestep))[0]
        return float(self.sigmas[int(idx[0])])


class DPMSolverMultistepScheduler:
    """ppdiffusers.DPMSolverMultistepScheduler (scheduling_dpmsolver_multistep.py:36-919) restricted to what has a device
    step: algorithm "dpmsolver++", solver_type "midpoint", order 1 or 2, epsilon prediction (the "DPM-Solver++ 2M"
    sampler of SD / SDXL), incl. Karras sigmas / Lu lambdas, lower_order_final, euler_at_final. The schedule and the
    per-step scalars are computed on the host in numpy with the reference's dtypes; `step` runs ONE fused kernel
    (`b200mix_dpmpp_2m_step`): CFG combine, x0 = (x - sigma_t*eps)/alpha_t, the first / second order update in the
    reference's operation order, and the x0 history for the next step (kept on the device)."""

    order = 1
    init_noise_sigma = 1.0

    def __init__(self, num_train_timesteps: int = 1000, beta_start: float = 0.0001, beta_end: float = 0.02,
                 beta_schedule: str = "linear", solver_order: int = 2, prediction_type: str = "epsilon",
                 thresholding: bool = False, sample_max_value: float = 1.0, algorithm_type: str = "dpmsolver++",
                 solver_type: str = "midpoint", lower_order_final: bool = True, euler_at_final: bool = False,
                 use_karras_sigmas: bool = False, use_lu_lambdas: bool = False, lambda_min_clipped: float = -float("inf"),
                 variance_type: Optional[str] = None, timestep_spacing: str = "linspace", steps_offset: int = 0):
        if beta_schedule == "linear":
            self.betas = _linspace_f32(beta_start, beta_end, num_train_timesteps)
        elif beta_schedule == "scaled_linear":
            self.betas = _linspace_f32(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps) ** 2
        elif beta_schedule == "squaredcos_cap_v2":
            self.betas = _betas_for_alpha_bar(num_train_timesteps)
        else:
            raise NotImplementedError(f"{beta_schedule} does is not implemented for {self.__class__}")
        if (algorithm_type != "dpmsolver++" or solver_type != "midpoint" or solver_order not in (1, 2) or thresholding or
                prediction_type != "epsilon" or variance_type):
            raise NotImplementedError("DPMSolverMultistepScheduler(b200): the device step covers dpmsolver++ / midpoint / "
                                      "order 1-2 / epsilon prediction without thresholding")
        self.alphas = (f32(1.0) - self.betas).astype(f32)
        self.alphas_cumprod = _cumprod_f32(self.alphas)
        self.num_train_timesteps, self.solver_order = num_train_timesteps, solver_order
        self.lower_order_final, self.euler_at_final = lower_order_final, euler_at_final
        self.use_karras_sigmas, self.use_lu_lambdas = use_karras_sigmas, use_lu_lambdas
        self.lambda_min_clipped = lambda_min_clipped
        self.timestep_spacing, self.steps_offset = timestep_spacing, steps_offset
        self.num_inference_steps: Optional[int] = None
        self.timesteps = np.linspace(0, num_train_timesteps - 1, num_train_timesteps, dtype=f32)[::-1].copy()
        self.lower_order_nums = 0
        self._step_index: Optional[int] = None
        self._m_prev = None  # x0 prediction of the previous step (device fp32)

    def scale_model_input(self, sample, timestep=None):
        return sample

    def set_timesteps(self, num_inference_steps: int):
        N = self.num_train_timesteps
        ac = self.alphas_cumprod
        if np.isinf(self.lambda_min_clipped):
            clipped_idx = 0
        else:
            lam = np.log(np.sqrt(ac, dtype=f32), dtype=f32) - np.log(np.sqrt((f32(1.0) - ac).astype(f32), dtype=f32), dtype=f32)
            clipped_idx = int(np.searchsorted(np.flip(lam), f32(self.lambda_min_clipped)))
        last_timestep = N - clipped_idx
        if self.timestep_spacing == "linspace":
            timesteps = np.linspace(0, last_timestep - 1, num_inference_steps + 1).round()[::-1][:-1].copy().astype(np.int64)
        elif self.timestep_spacing == "leading":
            step_ratio = last_timestep // (num_inference_steps + 1)
            timesteps = (np.arange(0, num_inference_steps + 1) * step_ratio).round()[::-1][:-1].copy().astype(np.int64)
            timesteps += self.steps_offset
        elif self.timestep_spacing == "trailing":
            step_ratio = N / num_inference_steps
            timesteps = np.arange(last_timestep, 0, -step_ratio).round().copy().astype(np.int64)
            timesteps -= 1
        else:
            raise ValueError(f"{self.timestep_spacing} is not supported. Please make sure to choose one of 'linspace', "
                             "'leading' or 'trailing'.")
        sigmas = np.sqrt(((f32(1.0) - ac) / ac).astype(f32), dtype=f32)
        log_sigmas = np.log(sigmas)
        if self.use_karras_sigmas:
            sigmas = np.flip(sigmas).copy()
            smin, smax, rho = sigmas[-1].item(), sigmas[0].item(), 7.0
            ramp = np.linspace(0, 1, num_inference_steps)
            sigmas = (smax ** (1 / rho) + ramp * (smin ** (1 / rho) - smax ** (1 / rho))) ** rho
            timesteps = np.array([EulerDiscreteScheduler._sigma_to_t(s, log_sigmas) for s in sigmas]).round()
            sigmas = np.concatenate([sigmas, sigmas[-1:]]).astype(f32)
        elif self.use_lu_lambdas:
            lambdas = np.flip(log_sigmas.copy())
            lmin, lmax = lambdas[-1].item(), lambdas[0].item()
            ramp = np.linspace(0, 1, num_inference_steps)
            sigmas = np.exp(lmax + ramp * (lmin - lmax))
            timesteps = np.array([EulerDiscreteScheduler._sigma_to_t(s, log_sigmas) for s in sigmas]).round()
            sigmas = np.concatenate([sigmas, sigmas[-1:]]).astype(f32)
        else:
            sigmas = np.interp(timesteps, np.arange(0, len(sigmas)), sigmas)
            sigma_last = np.sqrt(f32((f32(1.0) - ac[0]) / ac[0]), dtype=f32)
            sigmas = np.concatenate([sigmas, [sigma_last]]).astype(f32)
        self.sigmas = sigmas
        self.timesteps = np.asarray(timesteps).astype(np.int64)
        self.num_inference_steps = len(timesteps)
        self.lower_order_nums = 0
        self._step_index = None
        self._m_prev = None

    @staticmethod
    def _alpha_sigma_lambda(sigma):
        """_sigma_to_alpha_sigma_t (:362-366) + lambda = log(alpha) - log(sigma), all fp32."""
        sigma = f32(sigma)
        alpha_t = f32(f32(1.0) / np.sqrt(f32(f32(sigma * sigma) + f32(1.0)), dtype=f32))
        sigma_t = f32(sigma * alpha_t)
        return alpha_t, sigma_t, f32(np.log(alpha_t, dtype=f32) - np.log(sigma_t, dtype=f32))

    def step_scalars(self, timestep):
        """(order, sigma_cur, alpha_cur, A, C, halfC, inv_r0): x0 = (x - sigma_cur*eps)/alpha_cur;
        order 1: x' = A*x - C*x0; order 2: x' = A*x - C*x0 - halfC*(inv_r0*(x0 - x0_prev))."""
        if self.num_inference_steps is None:
            raise ValueError("Number of inference steps is 'None', you need to run 'set_timesteps' after creating the scheduler")
        if self._step_index is None:
            cand = np.nonzero(self.timesteps == int(timestep))[0]
            self._step_index = len(self.timesteps) - 1 if len(cand) == 0 else int(cand[1] if len(cand) > 1 else cand[0])
        i, n = self._step_index, len(self.timesteps)
        lower_order_final = (i == n - 1) and (self.euler_at_final or (self.lower_order_final and n < 15))
        alpha_s0, sigma_s0, lam_s0 = self._alpha_sigma_lambda(self.sigmas[i])
        alpha_t, sigma_t, lam_t = self._alpha_sigma_lambda(self.sigmas[i + 1])
        h = f32(lam_t - lam_s0)
        A = f32(sigma_t / sigma_s0)
        C = f32(alpha_t * f32(np.exp(f32(-h), dtype=f32) - f32(1.0)))
        first = self.solver_order == 1 or self.lower_order_nums < 1 or lower_order_final
        inv_r0 = f32(0.0)
        if not first:
            _, _, lam_s1 = self._alpha_sigma_lambda(self.sigmas[i - 1])
            # Karras / Lu schedules repeat the last sigma: h == 0 there, r0 = inf, 1/r0 = 0 and C = 0 (x' = x), exactly
            # like the reference's 0-d tensor arithmetic
            with np.errstate(divide="ignore", invalid="ignore"):
                r0 = f32(f32(lam_s0 - lam_s1) / h)
                inv_r0 = f32(f32(1.0) / r0)
        if self.lower_order_nums < self.solver_order:
            self.lower_order_nums += 1
        self._step_index += 1
        return (1 if first else 2, float(sigma_s0), float(alpha_s0), float(A), float(C), float(f32(f32(0.5) * C)), float(inv_r0))

    def step(self, model_output, timestep, sample, model_output_cond=None, guidance_scale=0.0, out=None):
        from .. import ops
        order, sigma_cur, alpha_cur, A, C, halfC, inv_r0 = self.step_scalars(timestep)
        if self._m_prev is None:
            import torch
            self._m_prev = [torch.empty_like(sample), torch.empty_like(sample)]
            self._m_slot = 0
        m_out, m_prev = self._m_prev[self._m_slot], self._m_prev[self._m_slot ^ 1]
        self._m_slot ^= 1
        return ops.dpmpp_2m_step(model_output, model_output_cond, guidance_scale, sample, m_prev if order == 2 else None,
                                 m_out, sigma_cur, alpha_cur, A, C, halfC, inv_r0, out=out)
