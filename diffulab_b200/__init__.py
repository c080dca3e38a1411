"""diffulab_b200: B200-native (sm_100a) implementation of DiffuLab's denoiser training / sampling hot path.

Same-named drop-ins for the reference's plugin surface (select them through `_target_: diffulab_b200.MMDiT` etc.
in the Hydra YAML, or `register_into_reference()` to patch the reference registries):
denoisers MMDiT / SprintDiT / DDT, the Flow formalisation, the Euler sampler, RepaLoss, PrecomputedEmbedder, Diffuser.
Everything on the device runs in libdiffulab_b200.so (hand-written CUDA behind a C ABI); there is no CPU fallback.
"""

__version__ = "0.1.0"

from .denoisers import DDT, Denoiser, MMDiT, ModelInput, ModelOutput, SprintDiT
from .diffuse import DDIM, DDPM, Diffuser, Diffusion, Euler, EulerMaruyama, Flow, Heun, GaussianDiffusion, SamplingOutput, StepResult
from .embedders import ContextEmbedder, PrecomputedEmbedder
from .losses import LossFunction, RepaLoss

__all__ = [
    "MMDiT", "SprintDiT", "DDT", "Denoiser", "ModelInput", "ModelOutput", "Diffuser", "Diffusion", "Flow", "GaussianDiffusion", "Euler", "EulerMaruyama", "Heun", "DDPM", "DDIM",
    "SamplingOutput", "StepResult", "ContextEmbedder", "PrecomputedEmbedder", "LossFunction", "RepaLoss",
]
