"""diffulab_b200: B200-native (sm_100a) implementation of DiffuLab's denoiser training / sampling hot path."""

__version__ = "0.1.0"
