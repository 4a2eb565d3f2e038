from .autoencoder_kl import AutoencoderKL
from .distributions import DiagonalGaussianDistribution

__all__ = ["AutoencoderKL", "DiagonalGaussianDistribution"]
