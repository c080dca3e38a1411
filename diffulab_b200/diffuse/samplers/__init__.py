from .common import FlowSampler, Sampler, StepResult
from .euler import Euler
from .euler_maruyama import EulerMaruyama
from .gaussian import DDIM, DDPM, GaussianSampler
from .heun import Heun

__all__ = ["Sampler", "FlowSampler", "GaussianSampler", "StepResult", "Euler", "EulerMaruyama", "Heun", "DDPM", "DDIM"]
