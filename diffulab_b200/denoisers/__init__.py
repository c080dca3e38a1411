from .common import Denoiser, ModelInput, ModelOutput
from .ddt import DDT
from .mmdit import MMDiT
from .sprint import SprintDiT

__all__ = ["Denoiser", "ModelInput", "ModelOutput", "MMDiT", "SprintDiT", "DDT"]
