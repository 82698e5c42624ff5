"""CorrLookup with the reference's constructor and call signature (models/utils/corr_lookup.py:71-136)."""
from typing import Sequence

import torch
import torch.nn as nn
from torch import Tensor

from . import ops


def coords_grid(batch: int, xx: Tensor, yy: Tensor) -> Tensor:
    """(batch, 2, H, W) grid of (x, y) pixel coordinates (corr_lookup.py:11-28)."""
    ys, xs = torch.meshgrid(yy, xx, indexing='ij')
    return torch.stack([xs, ys], dim=0).float()[None].repeat(batch, 1, 1, 1)


class CorrLookup(nn.Module):
    """Correlation lookup operator: one CUDA gather kernel over all pyramid levels.

    Args mirror the reference. Only the configuration SCFlow uses is implemented on the GPU
    (mode='bilinear', padding_mode='zeros', align_corners=True - configs/refine_models/scflow.py:72); anything
    else raises instead of silently running different numerics.
    """

    def __init__(self, radius: int = 4, mode: str = 'bilinear', padding_mode: str = 'zeros',
                 align_corners: bool = True) -> None:
        super().__init__()
        if mode != 'bilinear' or padding_mode != 'zeros' or not align_corners:
            raise NotImplementedError('CorrLookup: only bilinear / zeros / align_corners=True is implemented')
        self.r = radius
        self.mode = mode
        self.padding_mode = padding_mode
        self.align_corners = align_corners

    def forward(self, corr_pyramid: Sequence[Tensor], flow: Tensor) -> Tensor:
        """corr_pyramid: list of [B*H*W, 1, Hl, Wl]; flow: [B, 2, H, W] at 1/8 res. Returns [B, L*(2r+1)^2, H, W]."""
        b, _, h, w = flow.shape
        flow8 = flow.permute(0, 2, 3, 1).contiguous()
        ch = len(corr_pyramid) * (2 * self.r + 1) ** 2
        out = ops.corr_lookup_nhwc([c.contiguous() for c in corr_pyramid], flow8, self.r)
        return ops.nhwc_to_nchw(out, channels=ch)
