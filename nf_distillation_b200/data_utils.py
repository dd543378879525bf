"""Image pre/post-processing at the boundary of the flow path (reference: data/src/utils.py:1-25). Plain tensor ops on
whatever device the batch lives on; kept bit-for-bit identical to the reference so checkpoints and samples match."""
from __future__ import annotations

import torch

n_bits = 8


def preprocess(x: torch.Tensor) -> torch.Tensor:
    """[0, 1] floats (torchvision ToTensor) -> [-0.5, 0.5) with 2**n_bits levels (data/src/utils.py:7-18)."""
    x = x * 255
    n_bins = 2 ** n_bits
    if n_bits < 8:
        x = torch.floor(x / 2 ** (8 - n_bits))
    return x / n_bins - 0.5


def postprocess(x: torch.Tensor) -> torch.Tensor:
    """Model space -> uint8 pixels; like the reference, clamps and shifts `x` IN PLACE first (data/src/utils.py:21-25)."""
    x = torch.clamp(x, -0.5, 0.5)
    x += 0.5
    x = x * 255
    return torch.clamp(x, 0, 255).byte()
