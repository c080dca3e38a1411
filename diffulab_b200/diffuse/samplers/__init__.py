from .common import FlowSampler, Sampler, StepResult
from .euler import Euler

__all__ = ["Sampler", "FlowSampler", "StepResult", "Euler"]
