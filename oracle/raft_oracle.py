"""CPU oracle pieces of the RAFT baseline decoders.  TEST INFRASTRUCTURE ONLY.
Parity pin: `convex_upsample` was checked against the reference's own RAFTDecoder._upsample
(models/decoder/raft_decoder.py:381-416) run through oracle/ref_shim.py (oracle/make_golden_raft.py; fixture
tests/golden/convex_upsample_b2_6x9.npz)."""
import torch
import torch.nn.functional as F


def convex_upsample(flow: torch.Tensor, mask: torch.Tensor, scale: int = 8, grid_side: int = 3) -> torch.Tensor:
    """raft_decoder.py:403-416 (num_levels = 4 => scale 8; radius = 4 => grid_size 9)."""
    n, _, h, w = flow.shape
    g = grid_side * grid_side
    m = torch.softmax(mask.view(n, 1, g, scale, scale, h, w), dim=2)
    up = F.unfold(scale * flow, [grid_side, grid_side], padding=1).view(n, 2, g, 1, 1, h, w)
    up = torch.sum(m * up, dim=2).permute(0, 1, 4, 2, 5, 3)
    return up.reshape(n, 2, scale * h, scale * w)


def make_upsample_case(seed: int, b: int, h: int, w: int):
    g = torch.Generator().manual_seed(seed)
    return 3. * torch.randn(b, 2, h, w, generator=g), 2. * torch.randn(b, 576, h, w, generator=g)
