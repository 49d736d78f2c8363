"""ORACLE (test infrastructure). CPU restatement of the two diffusers schedulers the reference calls:

  DDPMScheduler.add_noise / get_velocity      mixofshow/pipelines/trainer_edlora.py:218,243
  DPMSolverMultistepScheduler.set_timesteps / scale_model_input / step
        mixofshow/pipelines/pipeline_edlora.py:249,274,290 ; gradient_fusion.py:601-622 ;
        mixofshow/pipelines/pipeline_regionally_t2iadapter.py:550-574

Third-party (diffusers, un-vendored, recommended ==0.19.3): restated from the published algorithm
(DPM-Solver++ 2M, midpoint, lower_order_final, epsilon prediction, scaled_linear betas 0.00085..0.012).
Parity unpinned against diffusers itself (not installable here); pinned instead against mathematics in
tests/test_oracle_golden.py: closed-form coefficients == step(), and convergence of the sampler to the analytic
probability-flow solution for Gaussian data at better than first order (a first-order update fails that test).
"""
import numpy as np
import torch


def sd15_alphas_cumprod(num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012):
    betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
    return torch.cumprod(1.0 - betas, dim=0)


class DDPMScheduler:
    def __init__(self):
        self.alphas_cumprod = sd15_alphas_cumprod()
        self.config = type('C', (), dict(num_train_timesteps=1000, prediction_type='epsilon'))()

    def add_noise(self, x0, noise, timesteps):
        ac = self.alphas_cumprod.to(x0.device, x0.dtype)
        a = ac[timesteps].sqrt().flatten()
        s = (1 - ac[timesteps]).sqrt().flatten()
        while a.ndim < x0.ndim:
            a, s = a.unsqueeze(-1), s.unsqueeze(-1)
        return a * x0 + s * noise

    def get_velocity(self, x0, noise, timesteps):
        ac = self.alphas_cumprod.to(x0.device, x0.dtype)
        a = ac[timesteps].sqrt().flatten()
        s = (1 - ac[timesteps]).sqrt().flatten()
        while a.ndim < x0.ndim:
            a, s = a.unsqueeze(-1), s.unsqueeze(-1)
        return a * noise - s * x0


class DPMSolverMultistepScheduler:
    """dpmsolver++, solver_order 2, midpoint, lower_order_final, no thresholding."""
    order = 1
    init_noise_sigma = 1.0

    def __init__(self):
        ac = sd15_alphas_cumprod()
        self.alphas_cumprod = ac
        self.alpha_t = ac.sqrt()
        self.sigma_t = (1 - ac).sqrt()
        self.lambda_t = self.alpha_t.log() - self.sigma_t.log()
        self.num_train_timesteps = 1000
        self.timesteps = None
        self.model_outputs = [None, None]
        self.lower_order_nums = 0

    def set_timesteps(self, num_inference_steps, device=None):
        ts = np.linspace(0, self.num_train_timesteps - 1, num_inference_steps + 1).round()[::-1][:-1].copy()
        ts = ts.astype(np.int64)
        _, uniq = np.unique(ts, return_index=True)
        ts = ts[np.sort(uniq)]
        self.timesteps = torch.from_numpy(ts)
        self.num_inference_steps = len(ts)
        self.model_outputs = [None, None]
        self.lower_order_nums = 0

    def scale_model_input(self, sample, t=None):
        return sample

    def coefficients(self, step_index):
        """(c_x, c_m0, c_m1, alpha_s, sigma_s) such that, with x0 = (x - sigma_s*eps)/alpha_s,
        prev = c_x*x + c_m0*x0_cur + c_m1*x0_prev  (c_m1 == 0 on first-order steps). Pure function of the
        schedule: this is what the fused device kernel consumes."""
        ts = self.timesteps
        n = len(ts)
        t = int(ts[step_index])
        prev_t = 0 if step_index == n - 1 else int(ts[step_index + 1])
        lower_final = (step_index == n - 1) and n < 15
        first_order = step_index == 0 or lower_final
        lam_t, lam_s = self.lambda_t[prev_t], self.lambda_t[t]
        a_t = self.alpha_t[prev_t]
        s_t, s_s = self.sigma_t[prev_t], self.sigma_t[t]
        h = lam_t - lam_s
        c_x = s_t / s_s
        e = a_t * (torch.exp(-h) - 1.0)
        if first_order:
            return float(c_x), float(-e), 0.0, float(self.alpha_t[t]), float(s_s)
        s1 = int(ts[step_index - 1])
        h0 = lam_s - self.lambda_t[s1]
        r0 = h0 / h
        # x = c_x x - e D0 - 0.5 e D1, D0 = m0, D1 = (m0 - m1)/r0
        c_m0 = -e - 0.5 * e / r0
        c_m1 = 0.5 * e / r0
        return float(c_x), float(c_m0), float(c_m1), float(self.alpha_t[t]), float(s_s)

    def step(self, model_output, timestep, sample):
        t = int(timestep)
        step_index = int((self.timesteps == t).nonzero()[0])
        n = len(self.timesteps)
        prev_t = 0 if step_index == n - 1 else int(self.timesteps[step_index + 1])
        lower_order_final = (step_index == n - 1) and n < 15
        lower_order_second = (step_index == n - 2) and n < 15
        a_s, s_s = self.alpha_t[t], self.sigma_t[t]
        x0 = (sample - s_s * model_output) / a_s
        self.model_outputs = [self.model_outputs[1], x0]
        lam_t, lam_s = self.lambda_t[prev_t], self.lambda_t[t]
        a_t, s_t = self.alpha_t[prev_t], self.sigma_t[prev_t]
        h = lam_t - lam_s
        if self.lower_order_nums < 1 or lower_order_final:
            prev = (s_t / s_s) * sample - (a_t * (torch.exp(-h) - 1.0)) * x0
        else:  # second order (solver_order == 2, or lower_order_second which is the same branch)
            s1 = int(self.timesteps[step_index - 1])
            m0, m1 = self.model_outputs[-1], self.model_outputs[-2]
            h0 = lam_s - self.lambda_t[s1]
            r0 = h0 / h
            d0, d1 = m0, (1.0 / r0) * (m0 - m1)
            prev = (s_t / s_s) * sample - (a_t * (torch.exp(-h) - 1.0)) * d0 \
                - 0.5 * (a_t * (torch.exp(-h) - 1.0)) * d1
        del lower_order_second
        if self.lower_order_nums < 2:
            self.lower_order_nums += 1
        return type('Out', (), dict(prev_sample=prev))()
