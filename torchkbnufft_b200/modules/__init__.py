from .kbinterp import KbInterp, KbInterpAdjoint
from .kbnufft import KbNufft, KbNufftAdjoint, ToepNufft

__all__ = ["KbInterp", "KbInterpAdjoint", "KbNufft", "KbNufftAdjoint", "ToepNufft"]
