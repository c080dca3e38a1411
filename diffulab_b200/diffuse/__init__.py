from .diffuser import Diffuser
from .diffusion import Diffusion, SamplingOutput
from .flow import Flow
from .gaussian import GaussianDiffusion, space_timesteps
from .samplers import DDIM, DDPM, Euler, EulerMaruyama, FlowSampler, Heun, GaussianSampler, Sampler, StepResult

__all__ = ["Diffuser", "Diffusion", "SamplingOutput", "Flow", "GaussianDiffusion", "space_timesteps", "Euler", "EulerMaruyama", "Heun", "DDPM", "DDIM",
           "FlowSampler", "GaussianSampler", "Sampler", "StepResult"]
