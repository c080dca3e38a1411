from .diffuser import Diffuser
from .diffusion import Diffusion, SamplingOutput
from .flow import Flow
from .samplers import Euler, FlowSampler, Sampler, StepResult

__all__ = ["Diffuser", "Diffusion", "SamplingOutput", "Flow", "Euler", "FlowSampler", "Sampler", "StepResult"]
