"""adsorbdiff_b200: B200-native (sm_100a) implementation of AdsorbDiff's PaiNN denoising hot path.

Public surface (mirrors the reference's interfaces for this path):
    PaiNN                    adsorbdiff.models.painn.painn_denoising.PaiNN
    Denoiser, DiffTorchCalc  adsorbdiff.relaxation.diffusers.denoising_torch
    ml_diffuse               adsorbdiff.relaxation.ml_relaxation.ml_diffuse
"""
from .painn import PaiNN  # noqa: F401
from .denoiser import Denoiser, DiffTorchCalc, ml_diffuse  # noqa: F401

__all__ = ["PaiNN", "Denoiser", "DiffTorchCalc", "ml_diffuse"]
