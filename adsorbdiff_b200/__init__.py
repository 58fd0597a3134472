"""adsorbdiff_b200: B200-native (sm_100a) implementation of AdsorbDiff's PaiNN denoising hot path.

Public surface (mirrors the reference's interfaces for this path):
    PaiNN                    adsorbdiff.models.painn.painn_denoising.PaiNN
    Denoiser, DiffTorchCalc  adsorbdiff.relaxation.diffusers.denoising_torch
    ml_diffuse               adsorbdiff.relaxation.ml_relaxation.ml_diffuse
    run_diffusion_batches    DenoisingTrainer.run_relaxations (the sampling job of one rank over its batches)
    run_diffusion            AdsorbDiffCalculator.run_diffusion (one structure, optionally many placements)
    PackedSystems / PackedLoader   packed, memory-mappable input format + pinned, prefetching batch loader
"""
from .painn import PaiNN  # noqa: F401
from .denoiser import Denoiser, DiffTorchCalc, ml_diffuse  # noqa: F401
from .packed import PackedLoader, PackedSystems  # noqa: F401
from .runner import run_diffusion, run_diffusion_batches  # noqa: F401

__all__ = ["PaiNN", "Denoiser", "DiffTorchCalc", "ml_diffuse", "run_diffusion", "run_diffusion_batches",
           "PackedSystems", "PackedLoader"]
