"""ORACLE (test infrastructure, not product code): CPU restatement of the two schedulers on the hot path,
DDIMScheduler (ppdiffusers/schedulers/scheduling_ddim.py:131-475) and FlowMatchEulerDiscreteScheduler
(ppdiffusers/schedulers/scheduling_flow_match_euler_discrete.py:44-282), with torch fp32 tensors standing in for
paddle fp32 tensors (0-d fp32 tensor ** 0.5 etc.).

Pinned: every RNG-free golden of ppdiffusers/tests/schedulers/test_scheduler_ddim.py (timesteps :68, variances
:116-126, full-loop sums :128-190) is asserted in tests/test_oracle_goldens.py.
"""
import math

import numpy as np
import torch


def linspace_f32(start, end, steps):
    """fp32 linspace evaluated the way the CPU kernels of Paddle and torch define it: step = (end-start)/(steps-1) in
    fp32; element i is start + step*i in the first half and end - step*(steps-1-i) in the second half. Written out
    explicitly (rather than calling torch.linspace, whose vectorised CPU kernel re-bases every 8/16 lanes) so that
    the value of every beta is defined by this file alone."""
    start, end = torch.tensor(start, dtype=torch.float32), torch.tensor(end, dtype=torch.float32)
    step = (end - start) / torch.tensor(steps - 1, dtype=torch.float32)
    i = torch.arange(steps, dtype=torch.float32)
    lo = start + step * i
    hi = end - step * (torch.tensor(steps - 1, dtype=torch.float32) - i)
    return torch.where(torch.arange(steps) < steps // 2, lo, hi)


def pow_half(t):
    """`t ** 0.5` on an fp32 tensor the way Paddle's CPU pow kernel evaluates it (std::pow(float, 0.5f), i.e. the
    correctly rounded square root). torch's CPU `** 0.5` / sqrt on fp32 is NOT correctly rounded (e.g.
    0.33933258**0.5 -> 0.58252257 instead of 0.58252263), so numpy's IEEE sqrt is used instead."""
    a = np.sqrt(np.asarray(t.detach().numpy(), dtype=np.float32), dtype=np.float32)
    return torch.from_numpy(np.ascontiguousarray(a).reshape(-1)).reshape(t.shape)


def cumprod_f32(x):
    """Running product with an fp32 accumulator (out[i] = fl32(out[i-1] * x[i])), the definition of an fp32 cumprod.
    Spelled out because torch's CPU cumprod accumulates in double and rounds each output, which differs in the last
    bit from a true fp32 scan for about a third of the 1000 alphas."""
    out = torch.empty_like(x)
    acc = torch.tensor(1.0, dtype=torch.float32)
    for i in range(x.numel()):
        acc = acc * x[i]
        out[i] = acc
    return out


def betas_for_alpha_bar(num_diffusion_timesteps, max_beta=0.999):
    """scheduling_ddim.py:52-93 (cosine)."""
    def alpha_bar_fn(t):
        return math.cos((t + 0.008) / 1.008 * math.pi / 2) ** 2
    betas = []
    for i in range(num_diffusion_timesteps):
        t1 = i / num_diffusion_timesteps
        t2 = (i + 1) / num_diffusion_timesteps
        betas.append(min(1 - alpha_bar_fn(t2) / alpha_bar_fn(t1), max_beta))
    return torch.tensor(betas, dtype=torch.float32)


class DDIMScheduler:
    order = 1

    def __init__(self, num_train_timesteps=1000, beta_start=0.0001, beta_end=0.02, beta_schedule="linear",
                 clip_sample=True, set_alpha_to_one=True, steps_offset=0, prediction_type="epsilon",
                 clip_sample_range=1.0, timestep_spacing="leading"):
        self.num_train_timesteps = num_train_timesteps
        self.clip_sample, self.clip_sample_range = clip_sample, clip_sample_range
        self.steps_offset, self.prediction_type, self.timestep_spacing = steps_offset, prediction_type, timestep_spacing
        if beta_schedule == "linear":  # :208-209
            self.betas = linspace_f32(beta_start, beta_end, num_train_timesteps)
        elif beta_schedule == "scaled_linear":  # :210-214
            self.betas = linspace_f32(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps) ** 2
        elif beta_schedule == "squaredcos_cap_v2":
            self.betas = betas_for_alpha_bar(num_train_timesteps)
        else:
            raise NotImplementedError(beta_schedule)
        self.alphas = 1.0 - self.betas
        self.alphas_cumprod = cumprod_f32(self.alphas)  # :225
        self.final_alpha_cumprod = torch.tensor(1.0) if set_alpha_to_one else self.alphas_cumprod[0]  # :231
        self.init_noise_sigma = 1.0
        self.num_inference_steps = None
        self.timesteps = torch.from_numpy(np.arange(0, num_train_timesteps)[::-1].copy().astype(np.int64))

    def scale_model_input(self, sample, timestep=None):  # :240-255
        return sample

    def _get_variance(self, timestep, prev_timestep):  # :257-265
        alpha_prod_t = self.alphas_cumprod[timestep]
        alpha_prod_t_prev = self.alphas_cumprod[prev_timestep] if prev_timestep >= 0 else self.final_alpha_cumprod
        beta_prod_t = 1 - alpha_prod_t
        beta_prod_t_prev = 1 - alpha_prod_t_prev
        return (beta_prod_t_prev / beta_prod_t) * (1 - alpha_prod_t / alpha_prod_t_prev)

    def set_timesteps(self, num_inference_steps):  # :305-348
        self.num_inference_steps = num_inference_steps
        if self.timestep_spacing == "linspace":
            timesteps = np.linspace(0, self.num_train_timesteps - 1, num_inference_steps).round()[::-1].copy().astype(np.int64)
        elif self.timestep_spacing == "leading":
            step_ratio = self.num_train_timesteps // num_inference_steps
            timesteps = (np.arange(0, num_inference_steps) * step_ratio).round()[::-1].copy().astype(np.int64)
            timesteps += self.steps_offset
        elif self.timestep_spacing == "trailing":
            step_ratio = self.num_train_timesteps / num_inference_steps
            timesteps = np.round(np.arange(self.num_train_timesteps, 0, -step_ratio)).astype(np.int64)
            timesteps -= 1
        else:
            raise ValueError(self.timestep_spacing)
        self.timesteps = torch.from_numpy(timesteps)

    def step(self, model_output, timestep, sample, eta=0.0):  # :350-475
        timestep = int(timestep)
        prev_timestep = timestep - self.num_train_timesteps // self.num_inference_steps
        alpha_prod_t = self.alphas_cumprod[timestep]
        alpha_prod_t_prev = self.alphas_cumprod[prev_timestep] if prev_timestep >= 0 else self.final_alpha_cumprod
        beta_prod_t = 1 - alpha_prod_t
        if self.prediction_type == "epsilon":
            pred_original_sample = (sample - pow_half(beta_prod_t) * model_output) / pow_half(alpha_prod_t)
            pred_epsilon = model_output
        elif self.prediction_type == "sample":
            pred_original_sample = model_output
            pred_epsilon = (sample - pow_half(alpha_prod_t) * pred_original_sample) / pow_half(beta_prod_t)
        elif self.prediction_type == "v_prediction":
            pred_original_sample = (pow_half(alpha_prod_t)) * sample - (pow_half(beta_prod_t)) * model_output
            pred_epsilon = (pow_half(alpha_prod_t)) * model_output + (pow_half(beta_prod_t)) * sample
        else:
            raise ValueError(self.prediction_type)
        if self.clip_sample:
            pred_original_sample = pred_original_sample.clip(-self.clip_sample_range, self.clip_sample_range)
        variance = self._get_variance(timestep, prev_timestep)
        std_dev_t = eta * pow_half(variance)
        pred_sample_direction = pow_half(1 - alpha_prod_t_prev - std_dev_t ** 2) * pred_epsilon
        prev_sample = pow_half(alpha_prod_t_prev) * pred_original_sample + pred_sample_direction
        assert eta == 0.0, "oracle covers the deterministic path (eta = 0) used by the pipelines"
        return prev_sample

    def step_scalars(self, timestep):
        """The four fp32 scalars of the eta=0 epsilon step, each computed exactly as step() computes it."""
        timestep = int(timestep)
        prev_timestep = timestep - self.num_train_timesteps // self.num_inference_steps
        a_t = self.alphas_cumprod[timestep]
        a_p = self.alphas_cumprod[prev_timestep] if prev_timestep >= 0 else self.final_alpha_cumprod
        b_t = 1 - a_t
        std = 0.0 * pow_half(self._get_variance(timestep, prev_timestep))
        return pow_half(a_t).item(), pow_half(b_t).item(), pow_half(a_p).item(), pow_half(1 - a_p - std ** 2).item()

    def add_noise(self, original_samples, noise, timesteps):  # :478-500
        sqrt_alpha_prod = pow_half(self.alphas_cumprod[timesteps])
        sqrt_one_minus = pow_half(1 - self.alphas_cumprod[timesteps])
        while sqrt_alpha_prod.ndim < original_samples.ndim:
            sqrt_alpha_prod = sqrt_alpha_prod.unsqueeze(-1)
            sqrt_one_minus = sqrt_one_minus.unsqueeze(-1)
        return sqrt_alpha_prod * original_samples + sqrt_one_minus * noise


def rescale_noise_cfg(noise_cfg, noise_pred_text, guidance_rescale=0.0):
    """pipeline_stable_diffusion.py:69-80: rescale the guided prediction to the per-sample std of the text branch
    (Paddle's Tensor.std is the unbiased estimator, like torch's) and blend with the unrescaled one."""
    dims = list(range(1, noise_pred_text.ndim))
    std_text = noise_pred_text.std(dim=dims, keepdim=True)
    std_cfg = noise_cfg.std(dim=dims, keepdim=True)
    noise_pred_rescaled = noise_cfg * (std_text / std_cfg)
    return guidance_rescale * noise_pred_rescaled + (1 - guidance_rescale) * noise_cfg


class LCMScheduler:
    """scheduling_lcm.py:133-560: __init__ :197-253, set_timesteps :330-450, boundary scalings :453-459, step :461-549.
    Restated for the default `leading` spacing without thresholding; the multi-step noise is supplied by the caller
    (the reference draws it from a paddle Generator, which cannot be reproduced here)."""
    order = 1

    def __init__(self, num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear",
                 original_inference_steps=50, clip_sample=False, clip_sample_range=1.0, set_alpha_to_one=True,
                 steps_offset=0, prediction_type="epsilon", timestep_scaling=10.0):
        self.num_train_timesteps, self.original_inference_steps = num_train_timesteps, original_inference_steps
        self.clip_sample, self.clip_sample_range = clip_sample, clip_sample_range
        self.prediction_type, self.timestep_scaling = prediction_type, timestep_scaling
        if beta_schedule == "linear":
            self.betas = linspace_f32(beta_start, beta_end, num_train_timesteps)
        elif beta_schedule == "scaled_linear":
            self.betas = linspace_f32(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps) ** 2
        elif beta_schedule == "squaredcos_cap_v2":
            self.betas = betas_for_alpha_bar(num_train_timesteps)
        else:
            raise NotImplementedError(beta_schedule)
        self.alphas = 1.0 - self.betas
        self.alphas_cumprod = cumprod_f32(self.alphas)
        self.final_alpha_cumprod = torch.tensor(1.0) if set_alpha_to_one else self.alphas_cumprod[0]
        self.init_noise_sigma = 1.0
        self.num_inference_steps = None
        self.timesteps = torch.from_numpy(np.arange(0, num_train_timesteps)[::-1].copy().astype(np.int64))
        self._step_index = None

    def scale_model_input(self, sample, timestep=None):
        return sample

    def set_timesteps(self, num_inference_steps, original_inference_steps=None, strength=1.0):  # :330-450 (2.2)
        original_steps = original_inference_steps if original_inference_steps is not None else self.original_inference_steps
        if original_steps > self.num_train_timesteps:
            raise ValueError("original_steps larger than num_train_timesteps")
        k = self.num_train_timesteps // original_steps
        lcm_origin_timesteps = np.asarray(list(range(1, int(original_steps * strength) + 1))) * k - 1
        skipping_step = len(lcm_origin_timesteps) // num_inference_steps
        if skipping_step < 1 or num_inference_steps > original_steps:
            raise ValueError("num_inference_steps too large for the original schedule")
        self.num_inference_steps = num_inference_steps
        lcm_origin_timesteps = lcm_origin_timesteps[::-1].copy()
        inference_indices = np.linspace(0, len(lcm_origin_timesteps), num=num_inference_steps, endpoint=False)
        inference_indices = np.floor(inference_indices).astype(np.int64)
        self.timesteps = torch.from_numpy(lcm_origin_timesteps[inference_indices].astype(np.int64))
        self._step_index = None

    def get_scalings_for_boundary_condition_discrete(self, timestep):  # :453-459 (int64 tensor * float -> fp32 tensor)
        sigma_data = 0.5
        scaled_timestep = torch.as_tensor(timestep).to(torch.float32) * self.timestep_scaling
        c_skip = sigma_data ** 2 / (scaled_timestep ** 2 + sigma_data ** 2)
        c_out = scaled_timestep / pow_half(scaled_timestep ** 2 + sigma_data ** 2)
        return c_skip, c_out

    def step(self, model_output, timestep, sample, noise=None):  # :461-549
        timestep = int(timestep)
        if self._step_index is None:
            cand = (self.timesteps == timestep).nonzero()
            self._step_index = (cand[1] if len(cand) > 1 else cand[0]).item()
        prev_step_index = self._step_index + 1
        prev_timestep = int(self.timesteps[prev_step_index]) if prev_step_index < len(self.timesteps) else timestep
        alpha_prod_t = self.alphas_cumprod[timestep]
        alpha_prod_t_prev = self.alphas_cumprod[prev_timestep] if prev_timestep >= 0 else self.final_alpha_cumprod
        beta_prod_t, beta_prod_t_prev = 1 - alpha_prod_t, 1 - alpha_prod_t_prev
        c_skip, c_out = self.get_scalings_for_boundary_condition_discrete(timestep)
        if self.prediction_type == "epsilon":
            x0 = (sample - pow_half(beta_prod_t) * model_output) / pow_half(alpha_prod_t)
        elif self.prediction_type == "sample":
            x0 = model_output
        elif self.prediction_type == "v_prediction":
            x0 = pow_half(alpha_prod_t) * sample - pow_half(beta_prod_t) * model_output
        else:
            raise ValueError(self.prediction_type)
        if self.clip_sample:
            x0 = x0.clip(-self.clip_sample_range, self.clip_sample_range)
        denoised = c_out * x0 + c_skip * sample
        if self._step_index != self.num_inference_steps - 1:
            assert noise is not None, "multi-step LCM needs the caller's noise tensor"
            prev_sample = pow_half(alpha_prod_t_prev) * denoised + pow_half(beta_prod_t_prev) * noise
        else:
            prev_sample = denoised
        self._step_index += 1
        return prev_sample, denoised

    def step_scalars(self, timestep):
        """(sqrt(a_t), sqrt(1-a_t), c_skip, c_out, sqrt(a_prev), sqrt(1-a_prev), is_last) exactly as step() uses them."""
        timestep = int(timestep)
        idx = self._step_index
        if idx is None:
            cand = (self.timesteps == timestep).nonzero()
            idx = (cand[1] if len(cand) > 1 else cand[0]).item()
        prev_timestep = int(self.timesteps[idx + 1]) if idx + 1 < len(self.timesteps) else timestep
        a_t = self.alphas_cumprod[timestep]
        a_p = self.alphas_cumprod[prev_timestep] if prev_timestep >= 0 else self.final_alpha_cumprod
        c_skip, c_out = self.get_scalings_for_boundary_condition_discrete(timestep)
        return (pow_half(a_t).item(), pow_half(1 - a_t).item(), c_skip.item(), c_out.item(), pow_half(a_p).item(),
                pow_half(1 - a_p).item(), idx == self.num_inference_steps - 1)


class FlowMatchEulerDiscreteScheduler:
    """scheduling_flow_match_euler_discrete.py: __init__ :64-83, set_timesteps :140-163, step :187-282."""
    order = 1

    def __init__(self, num_train_timesteps=1000, shift=1.0):
        self.num_train_timesteps, self.shift = num_train_timesteps, shift
        timesteps = np.linspace(1, num_train_timesteps, num_train_timesteps, dtype=np.float32)[::-1].copy()
        timesteps = torch.from_numpy(timesteps).to(torch.float32)
        sigmas = timesteps / num_train_timesteps
        sigmas = shift * sigmas / (1 + (shift - 1) * sigmas)
        self.timesteps = sigmas * num_train_timesteps
        self.sigmas = sigmas
        self.sigma_min, self.sigma_max = self.sigmas[-1].item(), self.sigmas[0].item()
        self._step_index = None

    def _sigma_to_t(self, sigma):
        return sigma * self.num_train_timesteps

    def set_timesteps(self, num_inference_steps):
        self.num_inference_steps = num_inference_steps
        timesteps = np.linspace(self._sigma_to_t(self.sigma_max), self._sigma_to_t(self.sigma_min), num_inference_steps)
        sigmas = timesteps / self.num_train_timesteps
        sigmas = self.shift * sigmas / (1 + (self.shift - 1) * sigmas)
        sigmas = torch.from_numpy(sigmas).to(torch.float32)
        self.timesteps = sigmas * self.num_train_timesteps
        self.sigmas = torch.cat([sigmas, torch.zeros(1)])
        self._step_index = None

    def step(self, model_output, timestep, sample):
        if self._step_index is None:
            self._step_index = int((self.timesteps == timestep).nonzero()[0].item())
        sample = sample.to(torch.float32)
        sigma = self.sigmas[self._step_index]
        sigma_next = self.sigmas[self._step_index + 1]
        # :244-275 with s_churn = 0: gamma = 0, sigma_hat = sigma, denoised = sample - model_output*sigma,
        # derivative = (sample - denoised)/sigma_hat, prev = sample + derivative*(sigma_next - sigma_hat)
        denoised = sample - model_output * sigma
        derivative = (sample - denoised) / sigma
        dt = sigma_next - sigma
        prev_sample = sample + derivative * dt
        self._step_index += 1
        return prev_sample.to(model_output.dtype)


class EulerDiscreteScheduler:
    """ppdiffusers/schedulers/scheduling_euler_discrete.py:135-503 (SDXL's default sampler), s_churn = 0 (the
    deterministic path: gamma = 0, the drawn noise is multiplied by zero). fp32 torch tensors stand in for paddle fp32
    tensors; the numpy float64 pieces (np.interp, the Karras ramp, _sigma_to_t) are numpy float64 here as well.
    Pinned by the RNG-free goldens of ppdiffusers/tests/schedulers/test_scheduler_euler.py:63-200
    (tests/golden/euler_goldens.json)."""
    order = 1

    def __init__(self, num_train_timesteps=1000, beta_start=0.0001, beta_end=0.02, beta_schedule="linear",
                 prediction_type="epsilon", interpolation_type="linear", use_karras_sigmas=False, sigma_min=None,
                 sigma_max=None, timestep_spacing="linspace", timestep_type="discrete", steps_offset=0):  # :135-203
        if beta_schedule == "linear":
            self.betas = linspace_f32(beta_start, beta_end, num_train_timesteps)
        elif beta_schedule == "scaled_linear":
            self.betas = linspace_f32(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps) ** 2
        elif beta_schedule == "squaredcos_cap_v2":
            self.betas = betas_for_alpha_bar(num_train_timesteps)
        else:
            raise NotImplementedError(f"{beta_schedule} does is not implemented for {self.__class__}")
        self.alphas = 1.0 - self.betas
        self.alphas_cumprod = cumprod_f32(self.alphas)
        self.config = dict(num_train_timesteps=num_train_timesteps, prediction_type=prediction_type,
                           interpolation_type=interpolation_type, sigma_min=sigma_min, sigma_max=sigma_max,
                           timestep_spacing=timestep_spacing, timestep_type=timestep_type, steps_offset=steps_offset)
        sigmas = self._train_sigmas()
        timesteps = np.linspace(0, num_train_timesteps - 1, num_train_timesteps, dtype=float)[::-1].copy()
        self.timesteps = torch.tensor(timesteps, dtype=torch.float32)
        self.sigmas = torch.cat([torch.from_numpy(sigmas[::-1].copy()), torch.zeros(1)])
        self.use_karras_sigmas = use_karras_sigmas
        self.num_inference_steps = None
        self._step_index = None

    def _train_sigmas(self):  # np.array(((1 - alphas_cumprod) / alphas_cumprod) ** 0.5): fp32 tensor math -> float32 array
        return pow_half((1 - self.alphas_cumprod) / self.alphas_cumprod).numpy()

    @property
    def init_noise_sigma(self):  # :205-211
        if self.config["timestep_spacing"] in ("linspace", "trailing"):
            return self.sigmas.max()
        return pow_half(self.sigmas.max() ** 2 + 1)

    def set_timesteps(self, num_inference_steps):  # :243-310
        c = self.config
        self.num_inference_steps = num_inference_steps
        N = c["num_train_timesteps"]
        if c["timestep_spacing"] == "linspace":
            timesteps = np.linspace(0, N - 1, num_inference_steps, dtype=np.float32)[::-1].copy()
        elif c["timestep_spacing"] == "leading":
            step_ratio = N // num_inference_steps
            timesteps = (np.arange(0, num_inference_steps) * step_ratio).round()[::-1].copy().astype(np.float32)
            timesteps += c["steps_offset"]
        elif c["timestep_spacing"] == "trailing":
            step_ratio = N / num_inference_steps
            timesteps = (np.arange(N, 0, -step_ratio)).round().copy().astype(np.float32)
            timesteps -= 1
        else:
            raise ValueError(f"{c['timestep_spacing']} is not supported.")
        sigmas = self._train_sigmas()
        log_sigmas = np.log(sigmas)
        if c["interpolation_type"] == "linear":
            sigmas = np.interp(timesteps, np.arange(0, len(sigmas)), sigmas)
        elif c["interpolation_type"] == "log_linear":
            sigmas = np.exp(np.linspace(np.log(sigmas[-1]), np.log(sigmas[0]), num_inference_steps + 1, dtype=np.float32))
        else:
            raise ValueError(f"{c['interpolation_type']} is not implemented.")
        if self.use_karras_sigmas:
            sigmas = self._convert_to_karras(sigmas, num_inference_steps)
            timesteps = np.array([self._sigma_to_t(sigma, log_sigmas) for sigma in sigmas])
        sigmas = torch.from_numpy(np.asarray(sigmas)).to(torch.float32)
        self.timesteps = torch.from_numpy(timesteps.astype(np.float32))
        self.sigmas = torch.cat([sigmas, torch.zeros(1)])
        self._step_index = None

    @staticmethod
    def _sigma_to_t(sigma, log_sigmas):  # :312-332
        log_sigma = np.log(np.maximum(sigma, 1e-10))
        dists = log_sigma - log_sigmas[:, np.newaxis]
        low_idx = np.cumsum((dists >= 0), axis=0).argmax(axis=0).clip(max=log_sigmas.shape[0] - 2)
        high_idx = low_idx + 1
        low, high = log_sigmas[low_idx], log_sigmas[high_idx]
        w = np.clip((low - log_sigma) / (low - high), 0, 1)
        t = (1 - w) * low_idx + w * high_idx
        return t.reshape(np.shape(sigma))

    def _convert_to_karras(self, in_sigmas, num_inference_steps):  # :335-358
        sigma_min = self.config["sigma_min"] if self.config["sigma_min"] is not None else in_sigmas[-1].item()
        sigma_max = self.config["sigma_max"] if self.config["sigma_max"] is not None else in_sigmas[0].item()
        rho = 7.0
        ramp = np.linspace(0, 1, num_inference_steps)
        min_inv_rho, max_inv_rho = sigma_min ** (1 / rho), sigma_max ** (1 / rho)
        return (max_inv_rho + ramp * (min_inv_rho - max_inv_rho)) ** rho

    def _init_step_index(self, timestep):  # :360-373
        cand = (self.timesteps == timestep).nonzero()
        self._step_index = (cand[1] if len(cand) > 1 else cand[0]).item()

    def scale_model_input(self, sample, timestep):  # :218-241
        if self._step_index is None:
            self._init_step_index(timestep)
        sigma = self.sigmas[self._step_index]
        return sample / pow_half(sigma ** 2 + 1)

    def step(self, model_output, timestep, sample):  # :375-470 with s_churn = 0 (gamma = 0, sigma_hat = sigma)
        if self._step_index is None:
            self._init_step_index(timestep)
        sigma = self.sigmas[self._step_index]
        sigma_hat = sigma * (0.0 + 1)
        pt = self.config["prediction_type"]
        if pt in ("original_sample", "sample"):
            pred_original_sample = model_output
        elif pt == "epsilon":
            pred_original_sample = sample - sigma_hat * model_output
        elif pt == "v_prediction":
            pred_original_sample = model_output * (-sigma / pow_half(sigma ** 2 + 1)) + (sample / (sigma ** 2 + 1))
        else:
            raise ValueError(f"prediction_type given as {pt} must be one of `epsilon`, or `v_prediction`")
        derivative = (sample - pred_original_sample) / sigma_hat
        dt = self.sigmas[self._step_index + 1] - sigma_hat
        prev_sample = sample + derivative * dt
        self._step_index += 1
        return prev_sample

    def step_scalars(self):
        """(sigma, dt) of the NEXT step() call as python floats (what the product's host scheduler hands the kernel)."""
        sigma = self.sigmas[self._step_index]
        return float(sigma), float(self.sigmas[self._step_index + 1] - sigma)

    def add_noise(self, original_samples, noise, timesteps):  # :472-497
        idx = [(self.timesteps == t).nonzero().item() for t in timesteps]
        sigma = self.sigmas[idx].flatten()
        while sigma.ndim < original_samples.ndim:
            sigma = sigma.unsqueeze(-1)
        return original_samples + noise * sigma


def log_f32(t):
    """fp32 natural log, evaluated with numpy's float32 kernel so that the oracle and the product's host scheduler (numpy)
    share one definition (a libm / SIMD kernel may differ from another in the last bit)."""
    a = np.log(np.asarray(t.detach().numpy(), dtype=np.float32), dtype=np.float32)
    return torch.from_numpy(np.ascontiguousarray(a).reshape(-1)).reshape(t.shape)


def exp_f32(t):
    a = np.exp(np.asarray(t.detach().numpy(), dtype=np.float32), dtype=np.float32)
    return torch.from_numpy(np.ascontiguousarray(a).reshape(-1)).reshape(t.shape)


class DPMSolverMultistepScheduler:
    """ppdiffusers/schedulers/scheduling_dpmsolver_multistep.py:36-919, the deterministic solvers ("dpmsolver++" and
    "dpmsolver", orders 1-2, midpoint / heun, epsilon / sample / v_prediction, Karras sigmas, Lu lambdas); the SDE variants
    (they draw noise), order 3 and dynamic thresholding (a quantile) are not restated. fp32 torch tensors stand in for
    paddle fp32 tensors. Pinned by the RNG-free goldens of tests/schedulers/test_scheduler_dpm_multi.py:229-284
    (tests/golden/dpm_multistep_goldens.json)."""
    order = 1
    init_noise_sigma = 1.0

    def __init__(self, num_train_timesteps=1000, beta_start=0.0001, beta_end=0.02, beta_schedule="linear", solver_order=2,
                 prediction_type="epsilon", thresholding=False, sample_max_value=1.0, algorithm_type="dpmsolver++",
                 solver_type="midpoint", lower_order_final=True, euler_at_final=False, use_karras_sigmas=False,
                 use_lu_lambdas=False, lambda_min_clipped=-float("inf"), variance_type=None, timestep_spacing="linspace",
                 steps_offset=0):  # :147-218
        if beta_schedule == "linear":
            self.betas = linspace_f32(beta_start, beta_end, num_train_timesteps)
        elif beta_schedule == "scaled_linear":
            self.betas = linspace_f32(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps) ** 2
        elif beta_schedule == "squaredcos_cap_v2":
            self.betas = betas_for_alpha_bar(num_train_timesteps)
        else:
            raise NotImplementedError(f"{beta_schedule} does is not implemented for {self.__class__}")
        if thresholding or algorithm_type not in ("dpmsolver++", "dpmsolver") or solver_order not in (1, 2) or variance_type:
            raise NotImplementedError("oracle: deterministic dpmsolver(++) of order 1-2 without thresholding only")
        if solver_type not in ("midpoint", "heun"):
            raise NotImplementedError(f"{solver_type} does is not implemented for {self.__class__}")
        self.alphas = 1.0 - self.betas
        self.alphas_cumprod = cumprod_f32(self.alphas)
        self.config = dict(num_train_timesteps=num_train_timesteps, solver_order=solver_order, prediction_type=prediction_type,
                           algorithm_type=algorithm_type, solver_type=solver_type, lower_order_final=lower_order_final,
                           euler_at_final=euler_at_final, use_karras_sigmas=use_karras_sigmas, use_lu_lambdas=use_lu_lambdas,
                           lambda_min_clipped=lambda_min_clipped, timestep_spacing=timestep_spacing, steps_offset=steps_offset)
        self.num_inference_steps = None
        self.timesteps = torch.from_numpy(np.linspace(0, num_train_timesteps - 1, num_train_timesteps, dtype=np.float32)[::-1].copy())
        self.model_outputs = [None] * solver_order
        self.lower_order_nums = 0
        self._step_index = None

    def _lambda_t(self):  # :188-192
        alpha_t, sigma_t = pow_half(self.alphas_cumprod), pow_half(1 - self.alphas_cumprod)
        return log_f32(alpha_t) - log_f32(sigma_t)

    def set_timesteps(self, num_inference_steps):  # :226-296
        c = self.config
        N = c["num_train_timesteps"]
        if np.isinf(c["lambda_min_clipped"]):
            clipped_idx = 0
        else:
            clipped_idx = int(np.searchsorted(np.flip(self._lambda_t().numpy()), np.float32(c["lambda_min_clipped"])))
        last_timestep = N - clipped_idx
        if c["timestep_spacing"] == "linspace":
            timesteps = np.linspace(0, last_timestep - 1, num_inference_steps + 1).round()[::-1][:-1].copy().astype(np.int64)
        elif c["timestep_spacing"] == "leading":
            step_ratio = last_timestep // (num_inference_steps + 1)
            timesteps = (np.arange(0, num_inference_steps + 1) * step_ratio).round()[::-1][:-1].copy().astype(np.int64)
            timesteps += c["steps_offset"]
        elif c["timestep_spacing"] == "trailing":
            step_ratio = N / num_inference_steps
            timesteps = np.arange(last_timestep, 0, -step_ratio).round().copy().astype(np.int64)
            timesteps -= 1
        else:
            raise ValueError(f"{c['timestep_spacing']} is not supported.")
        sigmas = pow_half((1 - self.alphas_cumprod) / self.alphas_cumprod).numpy()
        log_sigmas = np.log(sigmas)
        if c["use_karras_sigmas"]:
            sigmas = np.flip(sigmas).copy()
            smin, smax, rho = sigmas[-1].item(), sigmas[0].item(), 7.0
            ramp = np.linspace(0, 1, num_inference_steps)
            sigmas = (smax ** (1 / rho) + ramp * (smin ** (1 / rho) - smax ** (1 / rho))) ** rho
            timesteps = np.array([EulerDiscreteScheduler._sigma_to_t(s, log_sigmas) for s in sigmas]).round()
            sigmas = np.concatenate([sigmas, sigmas[-1:]]).astype(np.float32)
        elif c["use_lu_lambdas"]:
            lambdas = np.flip(log_sigmas.copy())
            lmin, lmax = lambdas[-1].item(), lambdas[0].item()
            ramp = np.linspace(0, 1, num_inference_steps)
            lambdas = lmax + ramp * (lmin - lmax)  # rho = 1
            sigmas = np.exp(lambdas)
            timesteps = np.array([EulerDiscreteScheduler._sigma_to_t(s, log_sigmas) for s in sigmas]).round()
            sigmas = np.concatenate([sigmas, sigmas[-1:]]).astype(np.float32)
        else:
            sigmas = np.interp(timesteps, np.arange(0, len(sigmas)), sigmas)
            sigma_last = pow_half((1 - self.alphas_cumprod[0]) / self.alphas_cumprod[0]).item()
            sigmas = np.concatenate([sigmas, [sigma_last]]).astype(np.float32)
        self.sigmas = torch.from_numpy(sigmas)
        self.timesteps = torch.from_numpy(np.asarray(timesteps).astype(np.int64))
        self.num_inference_steps = len(timesteps)
        self.model_outputs = [None] * c["solver_order"]
        self.lower_order_nums = 0
        self._step_index = None

    @staticmethod
    def _sigma_to_alpha_sigma_t(sigma):  # :362-366
        alpha_t = 1 / pow_half(sigma ** 2 + 1)
        return alpha_t, sigma * alpha_t

    def scale_model_input(self, sample, *a, **k):
        return sample

    def _init_step_index(self, timestep):  # :785-799
        cand = (self.timesteps == timestep).nonzero()
        if len(cand) == 0:
            self._step_index = len(self.timesteps) - 1
        else:
            self._step_index = (cand[1] if len(cand) > 1 else cand[0]).item()

    def convert_model_output(self, model_output, sample):  # :408-505
        c = self.config
        sigma = self.sigmas[self._step_index]
        alpha_t, sigma_t = self._sigma_to_alpha_sigma_t(sigma)
        if c["algorithm_type"] == "dpmsolver++":
            if c["prediction_type"] == "epsilon":
                return (sample - sigma_t * model_output) / alpha_t
            if c["prediction_type"] == "sample":
                return model_output
            if c["prediction_type"] == "v_prediction":
                return alpha_t * sample - sigma_t * model_output
        else:
            if c["prediction_type"] == "epsilon":
                return model_output
            if c["prediction_type"] == "sample":
                return (sample - alpha_t * model_output) / sigma_t
            if c["prediction_type"] == "v_prediction":
                return alpha_t * model_output + sigma_t * sample
        raise ValueError(f"prediction_type given as {c['prediction_type']} must be one of `epsilon`, `sample`, or `v_prediction`")

    def _lambdas(self, *idx):
        out = []
        for i in idx:
            a, s = self._sigma_to_alpha_sigma_t(self.sigmas[i])
            out.append((a, s, log_f32(a) - log_f32(s)))
        return out

    def first_order_update(self, m0, sample):  # :507-580
        (alpha_t, sigma_t, lam_t), (alpha_s, sigma_s, lam_s) = self._lambdas(self._step_index + 1, self._step_index)
        h = lam_t - lam_s
        if self.config["algorithm_type"] == "dpmsolver++":
            return (sigma_t / sigma_s) * sample - (alpha_t * (exp_f32(-h) - 1.0)) * m0
        return (alpha_t / alpha_s) * sample - (sigma_t * (exp_f32(h) - 1.0)) * m0

    def second_order_update(self, outs, sample):  # :582-700
        (alpha_t, sigma_t, lam_t), (alpha_s0, sigma_s0, lam_s0), (_, _, lam_s1) = self._lambdas(
            self._step_index + 1, self._step_index, self._step_index - 1)
        m0, m1 = outs[-1], outs[-2]
        h, h_0 = lam_t - lam_s0, lam_s0 - lam_s1
        r0 = h_0 / h
        D0, D1 = m0, (1.0 / r0) * (m0 - m1)
        pp, mid = self.config["algorithm_type"] == "dpmsolver++", self.config["solver_type"] == "midpoint"
        if pp and mid:
            return (sigma_t / sigma_s0) * sample - (alpha_t * (exp_f32(-h) - 1.0)) * D0 - 0.5 * (alpha_t * (exp_f32(-h) - 1.0)) * D1
        if pp:
            return (sigma_t / sigma_s0) * sample - (alpha_t * (exp_f32(-h) - 1.0)) * D0 + (alpha_t * ((exp_f32(-h) - 1.0) / h + 1.0)) * D1
        if mid:
            return (alpha_t / alpha_s0) * sample - (sigma_t * (exp_f32(h) - 1.0)) * D0 - 0.5 * (sigma_t * (exp_f32(h) - 1.0)) * D1
        return (alpha_t / alpha_s0) * sample - (sigma_t * (exp_f32(h) - 1.0)) * D0 - (sigma_t * ((exp_f32(h) - 1.0) / h - 1.0)) * D1

    def step(self, model_output, timestep, sample):  # :801-873
        if self.num_inference_steps is None:
            raise ValueError("Number of inference steps is 'None', you need to run 'set_timesteps' after creating the scheduler")
        if self._step_index is None:
            self._init_step_index(timestep)
        c, n = self.config, len(self.timesteps)
        lower_order_final = (self._step_index == n - 1) and (c["euler_at_final"] or (c["lower_order_final"] and n < 15))
        m = self.convert_model_output(model_output, sample)
        for i in range(c["solver_order"] - 1):
            self.model_outputs[i] = self.model_outputs[i + 1]
        self.model_outputs[-1] = m
        if c["solver_order"] == 1 or self.lower_order_nums < 1 or lower_order_final:
            prev = self.first_order_update(m, sample)
        else:
            prev = self.second_order_update(self.model_outputs, sample)
        if self.lower_order_nums < c["solver_order"]:
            self.lower_order_nums += 1
        self._step_index += 1
        return prev

    def add_noise(self, original_samples, noise, timesteps):  # :893-916
        idx = [(self.timesteps == t).nonzero().item() for t in timesteps]
        sigma = self.sigmas[idx].flatten()
        while sigma.ndim < original_samples.ndim:
            sigma = sigma.unsqueeze(-1)
        alpha_t, sigma_t = self._sigma_to_alpha_sigma_t(sigma)
        return alpha_t * original_samples + sigma_t * noise
