from .common import FlowSampler, Sampler, StepResult
from .euler import Euler
from .gaussian import DDIM, DDPM, GaussianSampler

__all__ = ["Sampler", "FlowSampler", "GaussianSampler", "StepResult", "Euler", "DDPM", "DDIM"]
