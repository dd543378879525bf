"""Mirror of the reference's models/utils.py (split_feature, uniform_binning_correction, compute_same_pad)."""
from __future__ import annotations

import math

import torch


def compute_same_pad(kernel_size, stride):
    """'same' padding per side for odd kernels (reference: models/utils.py:5-23)."""
    if isinstance(kernel_size, int):
        kernel_size = [kernel_size]
    if isinstance(stride, int):
        stride = [stride]
    assert len(stride) == len(kernel_size), \
        "Pass kernel size and stride both as int, or both as equal length iterable"
    pads = []
    for k, s in zip(kernel_size, stride):
        side = ((k - 1) * s + 1) // 2
        pads += [side, side]
    return pads


def dequant_noise(x: torch.Tensor, n_bins: int) -> torch.Tensor:
    """U(0, 1/n_bins) noise with the shape of x. Kept as a separate hook so tests can substitute recorded noise
    (the RNG stays on the torch side, SURVEY.md §7 'RNG parity')."""
    return torch.zeros_like(x).uniform_(0, 1.0 / n_bins)


def uniform_binning_correction(x: torch.Tensor, n_bits: int = 8):
    """x <- x + U(0, 1/256) IN PLACE (the caller's batch is mutated, as in the reference) and the constant
    dequantisation log-det -ln(256) * C*H*W per sample (reference: models/utils.py:26-41)."""
    b, c, h, w = x.size()
    n_bins = 2 ** n_bits
    x += dequant_noise(x, n_bins)
    objective = torch.full((b,), -math.log(n_bins) * c * h * w, device=x.device, dtype=torch.float32)
    return x, objective


def can_fuse_dequant_squeeze(x: torch.Tensor) -> bool:
    return (x.is_cuda and x.dim() == 4 and x.shape[3] % 4 == 0 and x.shape[2] % 2 == 0 and x.is_contiguous()
            and x.dtype in (torch.float32, torch.uint8))


def dequantize_and_squeeze(x: torch.Tensor, n_bits: int = 8):
    """uniform_binning_correction (models/utils.py:26-41) + the first SqueezeLayer (models/layers.py:32-44) — and, when the
    batch arrives as raw uint8 pixels, preprocess (data/src/utils.py:7-18) — in ONE kernel (csrc/preproc.cu).
    Returns (x, objective, squeezed): for a float batch `x` is the caller's tensor with the noise added IN PLACE (the
    reference's semantics: a second call noises it again); for a uint8 batch it is a new fp32 tensor."""
    from .. import ops
    b, c, h, w = x.shape
    n_bins = 2 ** n_bits
    ref = x if x.dtype == torch.float32 else torch.empty(x.shape, device=x.device, dtype=torch.float32)
    noise = dequant_noise(ref, n_bins).contiguous()
    sq = torch.empty(b, 4 * c, h // 2, w // 2, device=x.device, dtype=torch.float32)
    ops.dequant_squeeze(x, noise, ref, sq, n_bits)
    objective = torch.full((b,), -math.log(n_bins) * c * h * w, device=x.device, dtype=torch.float32)
    return ref, objective, sq


def split_feature(tensor: torch.Tensor, type: str = "split"):
    """'split': first / second half of the channels; 'cross': even / odd channels (reference: models/utils.py:44-52)."""
    C = tensor.size(1)
    if type == "split":
        return tensor[:, : C // 2, ...], tensor[:, C // 2:, ...]
    elif type == "cross":
        return tensor[:, 0::2, ...], tensor[:, 1::2, ...]
