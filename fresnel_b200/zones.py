"""Fresnel depth zones and the depth-edge detector: host-side mirrors of the two helper modules the decoder tail
and the loss read on this path (scripts/utils/fresnel_zones.py: ``FresnelZones`` :34-180, ``FresnelEdgeDetector``
:1084-1159).  Same constructor arguments, buffer / parameter names (the reference's ``state_dict`` loads) and
results; only the members the hot path touches exist.  The zone snap itself and the boundary mask run inside the
CUDA kernels (csrc/head.cu, csrc/loss.cu) from these modules' buffers; the PyTorch methods below are the CPU
restatement the tests compare with the reference and with the kernels.
"""

from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F


class FresnelZones(nn.Module):
    """``num_zones`` equal depth zones over ``depth_range``: boundaries = linspace(lo, hi, num_zones + 1), centres
    = midpoints (fresnel_zones.py:79-87)."""

    def __init__(self, num_zones: int = 8, depth_range: Tuple[float, float] = (0.0, 1.0),
                 boundary_threshold: float = 0.02, soft_boundaries: bool = True):
        super().__init__()
        self.num_zones = num_zones
        self.depth_range = depth_range
        self.boundary_threshold = boundary_threshold
        self.soft_boundaries = soft_boundaries
        boundaries = torch.linspace(depth_range[0], depth_range[1], num_zones + 1)
        self.register_buffer("zone_boundaries", boundaries)
        self.register_buffer("zone_centers", (boundaries[:-1] + boundaries[1:]) / 2)
        self.register_buffer("zone_width", torch.tensor((depth_range[1] - depth_range[0]) / num_zones))
        self.boundary_emphasis = nn.Parameter(torch.ones(num_zones + 1))      # present in the reference's state

    def quantize_depth(self, depth: torch.Tensor) -> torch.Tensor:
        """Zone index of every depth value (fresnel_zones.py:96-116)."""
        clamped = torch.clamp(depth, self.depth_range[0], self.depth_range[1])
        return torch.bucketize(clamped, self.zone_boundaries[1:-1])

    def get_zone_centers_for_depth(self, depth: torch.Tensor) -> torch.Tensor:
        """Depth snapped to the centre of its zone (fresnel_zones.py:118-139)."""
        return self.zone_centers[self.quantize_depth(depth)]

    def compute_boundary_mask(self, depth: torch.Tensor, threshold: Optional[float] = None) -> torch.Tensor:
        """Closeness to the nearest zone boundary (fresnel_zones.py:141-180): sigmoid(10 / thr * (thr - dist)) or,
        with hard boundaries, 1[dist < thr]."""
        thr = self.boundary_threshold if threshold is None else threshold
        dist = (depth.unsqueeze(-1) - self.zone_boundaries).abs().min(dim=-1).values
        if self.soft_boundaries:
            return torch.sigmoid((10.0 / thr) * (thr - dist))
        return (dist < thr).float()


class DepthEdgeDetector(nn.Module):
    """Three 3x3 convolutions over [depth, sobel_x(depth), sobel_y(depth)] -> edge strength in [0, 1]
    (FresnelEdgeDetector, fresnel_zones.py:1084-1159).  Plain cuDNN convolutions on a 37x37 grid: plumbing."""

    def __init__(self, in_channels: int = 1, hidden_channels: int = 16, use_depth_gradients: bool = True):
        super().__init__()
        self.use_depth_gradients = use_depth_gradients
        actual_in = in_channels + 2 if use_depth_gradients else in_channels
        self.conv1 = nn.Conv2d(actual_in, hidden_channels, kernel_size=3, padding=1)
        self.conv2 = nn.Conv2d(hidden_channels, hidden_channels, kernel_size=3, padding=1)
        self.conv3 = nn.Conv2d(hidden_channels, 1, kernel_size=3, padding=1)
        self.register_buffer("sobel_x", torch.tensor([[-1., 0., 1.], [-2., 0., 2.], [-1., 0., 1.]]).view(1, 1, 3, 3))
        self.register_buffer("sobel_y", torch.tensor([[-1., -2., -1.], [0., 0., 0.], [1., 2., 1.]]).view(1, 1, 3, 3))

    def forward(self, depth: torch.Tensor) -> torch.Tensor:
        if depth.dim() == 3:
            depth = depth.unsqueeze(1)
        x = depth
        if self.use_depth_gradients:
            x = torch.cat([depth, F.conv2d(depth, self.sobel_x, padding=1), F.conv2d(depth, self.sobel_y, padding=1)],
                          dim=1)
        x = F.relu(self.conv1(x))
        x = F.relu(self.conv2(x))
        return torch.sigmoid(self.conv3(x))
