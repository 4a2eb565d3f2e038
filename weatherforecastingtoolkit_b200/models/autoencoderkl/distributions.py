"""Posterior object returned by ``AutoencoderKL.encode``.

Same surface as the reference ``DiagonalGaussianDistribution``
(pipeline/models/autoencoderkl/distributions.py:26-71): attributes ``parameters, mean, logvar,
std, var, deterministic`` and methods ``sample(generator=None), mode(), kl(other=None),
nll(sample, dims)``. The arithmetic the rollout uses -- clamp, std, var and ``mean + std * noise`` -- is one
``wfk_gaussian_posterior`` launch; ``kl`` / ``nll`` are training-loss helpers (out of scope) and stay torch
reductions. The noise itself comes from the device RNG like the reference's ``torch.randn`` (SURVEY hazard H3).
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

from ... import _cabi


class DiagonalGaussianDistribution(object):
    def __init__(self, parameters: torch.Tensor, deterministic: bool = False):
        if not parameters.is_cuda:
            raise RuntimeError("this path runs on a B200 only (no CPU fallback): pass CUDA tensors")
        parameters = parameters.detach().to(torch.float32).contiguous()
        self.parameters = parameters
        self.mean = parameters[:, : parameters.shape[1] // 2]          # view, like torch.chunk
        self.deterministic = deterministic
        n, c2 = parameters.shape[0], parameters.shape[1]
        self._lc, self._hw = c2 // 2, int(parameters[0, 0].numel())
        self.logvar = torch.empty_like(self.mean, memory_format=torch.contiguous_format)
        self.std = torch.empty_like(self.logvar)
        self.var = torch.empty_like(self.logvar)
        self._lib = _cabi.init(parameters.device.index if parameters.device.index is not None else 0)
        _cabi.check(self._lib.wfk_gaussian_posterior(parameters.data_ptr(), n, self._lc, self._hw, self.logvar.data_ptr(),
                                                     self.std.data_ptr(), self.var.data_ptr(), None, None,
                                                     torch.cuda.current_stream(parameters.device).cuda_stream),
                    "wfk_gaussian_posterior")
        if self.deterministic:
            self.var = self.std = torch.zeros_like(self.logvar)

    def sample(self, generator: Optional[torch.Generator] = None) -> torch.Tensor:
        noise = torch.randn(self.mean.shape, generator=generator, device=self.parameters.device,
                            dtype=self.parameters.dtype)
        return self.sample_with_noise(noise)

    def sample_with_noise(self, noise: torch.Tensor) -> torch.Tensor:
        """mean + std * noise for a caller-supplied noise tensor (the reproducible form of ``sample``)."""
        if self.deterministic:
            return self.mean.clone()
        noise = noise.detach().to(device=self.parameters.device, dtype=torch.float32).contiguous()
        out = torch.empty_like(self.logvar)
        _cabi.check(self._lib.wfk_gaussian_posterior(self.parameters.data_ptr(), self.parameters.shape[0], self._lc, self._hw,
                                                     None, None, None, noise.data_ptr(), out.data_ptr(),
                                                     torch.cuda.current_stream(self.parameters.device).cuda_stream),
                    "wfk_gaussian_posterior")
        return out

    def mode(self) -> torch.Tensor:
        return self.mean

    def kl(self, other=None):
        if self.deterministic:
            return torch.Tensor([0.0])
        if other is None:
            return 0.5 * torch.sum(torch.pow(self.mean, 2) + self.var - 1.0 - self.logvar, dim=[1, 2, 3])
        return 0.5 * torch.sum(
            torch.pow(self.mean - other.mean, 2) / other.var + self.var / other.var - 1.0 - self.logvar + other.logvar,
            dim=[1, 2, 3])

    def nll(self, sample, dims=[1, 2, 3]):
        if self.deterministic:
            return torch.Tensor([0.0])
        logtwopi = np.log(2.0 * np.pi)
        return 0.5 * torch.sum(logtwopi + self.logvar + torch.pow(sample - self.mean, 2) / self.var, dim=dims)
