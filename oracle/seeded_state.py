"""ORACLE / TEST INFRASTRUCTURE ONLY.

Seed-reproducible parameter initialisation shared by oracle/make_golden.py (applied to
the UNMODIFIED reference modules) and by the tests (applied to the drop-in modules): the
latent-128 fixtures then need not carry multi-megabyte state dicts, only a checksum that
proves both sides hold the same parameters. Values depend only on the seed and on the
order / shapes of ``named_parameters()`` — which the state-dict contract fixes.
"""
import math

import torch


def seeded_init(module, seed):
    """Matrices as the reference's kaiming_init (training_utils.py:48-58); 1-D parameters
    are perturbed (biases ~ 0.1 N(0,1), LayerNorm / BatchNorm weights ~ 1 + 0.2 N(0,1)) so that
    every affine term carries signal. Returns the checksum of the resulting parameters."""
    g = torch.Generator().manual_seed(int(seed))
    with torch.no_grad():
        for name, p in module.named_parameters():
            if p.dim() >= 2:
                std = (1.0 if name.endswith("0.weight") else math.sqrt(2.0)) / math.sqrt(p.shape[1])
                p.copy_(torch.randn(p.shape, generator=g) * std)
            elif name.endswith(".bias"):
                p.copy_(0.1 * torch.randn(p.shape, generator=g))
            else:
                p.copy_(1.0 + 0.2 * torch.randn(p.shape, generator=g))
    return checksum(module)


def checksum(module):
    """Order-sensitive fp64 digest of the parameters (position-weighted sums)."""
    tot = 0.0
    for i, (_, p) in enumerate(module.named_parameters()):
        v = p.detach().double().cpu().flatten()
        w = torch.arange(1, v.numel() + 1, dtype=torch.float64) % 251 + 1.0
        tot += (i + 1) * float((v * w).sum())
    return tot
