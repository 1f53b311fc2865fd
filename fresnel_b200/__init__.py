"""fresnel_b200: B200-native (sm_100a) drop-in for the differentiable Gaussian-splatting
renderer of CalebisGross/fresnel (scripts/models/differentiable_renderer.py).

    from fresnel_b200 import TileBasedRenderer, Camera      # instead of models.differentiable_renderer

The compute path is the CUDA library ``csrc/libfresnel_b200.so`` behind ``include/fresnel_b200.h``;
importing this package does not need a GPU, rendering does.
"""

from .camera import Camera, camera_vector, create_camera_from_pose
from .renderer import DEFAULT_T_EPS, DifferentiableGaussianRenderer, TileBasedRenderer, build_bins, render_views
from .wave import ASMWaveFieldRenderer, WaveFieldRenderer
from .fourier import FourierGaussianRenderer
from .simple import SimplifiedRenderer
from .io import (load_gaussians_from_binary, load_gaussians_from_ply, save_gaussians_to_binary,
                 save_gaussians_to_ply)

__all__ = ["Camera", "camera_vector", "create_camera_from_pose", "TileBasedRenderer", "WaveFieldRenderer",
           "ASMWaveFieldRenderer", "DifferentiableGaussianRenderer", "FourierGaussianRenderer", "SimplifiedRenderer", "render_views",
           "build_bins", "DEFAULT_T_EPS", "load_gaussians_from_binary", "save_gaussians_to_binary",
           "load_gaussians_from_ply", "save_gaussians_to_ply"]
