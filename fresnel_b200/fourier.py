"""FourierGaussianRenderer drop-in over the C-ABI library.

Constructor and ``forward`` signature follow the reference module
(scripts/models/differentiable_renderer.py:1500-1774).  What the reference computes (DR:1693-1753) is an
order-free additive splat of isotropic Gaussians followed by a global-max normalisation; the splat runs
through the wave kernels with zero phases on FRB_MODE_FOURIER records, the epilogue in csrc/fourier.cu.
"""

from __future__ import annotations

from typing import Optional, Tuple

import numpy as np
import torch
import torch.nn as nn

from . import _lib
from .camera import camera_vector
from .renderer import RECORD_FLOATS, _call, _check_inputs, _ptr, _stream, empty_cloud_result
from .wave import WC_FLOATS, _prepare_wave_bins, _project_backward

MODE_FOURIER = 2


class _FourierRenderFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, positions, scales, rotations, colors, opacities, cfg):
        cam_vecs, n_views, width, height, max_radius, bg = cfg
        L = _lib.lib()
        dev = positions.device
        st = _stream()
        n = positions.shape[0]
        f32 = dict(dtype=torch.float32, device=dev)
        zero_phase = torch.zeros(n, **f32)
        bins, sorted_wc = _prepare_wave_bins(positions, scales, rotations, colors, opacities, zero_phase, 1, cfg,
                                             mode=MODE_FOURIER)
        accum = torch.empty(n_views, 8, height, width, **f32)
        rmax = torch.empty(n_views, dtype=torch.int32, device=dev)
        mx_key = torch.empty(n_views, dtype=torch.int32, device=dev)
        image = torch.empty(n_views, 3, height, width, **f32)
        _call("frb_wave_splat_fwd", L.frb_wave_splat_fwd, n_views, width, height, _ptr(bins.ranges),
              _ptr(bins.sorted_records), _ptr(sorted_wc), _ptr(accum), _ptr(rmax), st)
        _call("frb_fourier_finish_fwd", L.frb_fourier_finish_fwd, n_views, width, height, _ptr(accum),
              _ptr(mx_key), bg.ctypes.data, _ptr(image), st)
        ctx.cfg, ctx.n = cfg, n
        ctx.set_materialize_grads(False)
        ctx.save_for_backward(positions, scales, rotations, colors, zero_phase, bins.ranges, bins.sorted_records,
                              bins.sorted_gids, sorted_wc, accum, mx_key)
        return image

    @staticmethod
    def backward(ctx, g_image):
        cam_vecs, n_views, width, height, max_radius, bg = ctx.cfg
        (positions, scales, rotations, colors, zero_phase, ranges, sorted_records, sorted_gids, sorted_wc, accum,
         mx_key) = ctx.saved_tensors
        L = _lib.lib()
        dev = positions.device
        st = _stream()
        n = ctx.n
        f32 = dict(dtype=torch.float32, device=dev)
        if g_image is None:
            return (torch.zeros_like(positions), torch.zeros_like(scales), torch.zeros_like(rotations),
                    torch.zeros_like(colors), torch.zeros(n, **f32), None)
        g_image = g_image.contiguous().float()
        red = torch.empty(2 * n_views, **f32)
        gpix = torch.empty(n_views, 8, height, width, **f32)
        _call("frb_fourier_finish_bwd", L.frb_fourier_finish_bwd, n_views, width, height, _ptr(accum), _ptr(mx_key),
              bg.ctypes.data, _ptr(g_image), _ptr(red), _ptr(gpix), st)
        grad2d = torch.zeros(n, RECORD_FLOATS, **f32)
        gwc = torch.zeros(n, WC_FLOATS, **f32)
        _call("frb_wave_splat_bwd", L.frb_wave_splat_bwd, n_views, width, height, 0, _ptr(ranges),
              _ptr(sorted_records), _ptr(sorted_wc), _ptr(sorted_gids), None, _ptr(gpix), 1.0, _ptr(grad2d),
              _ptr(gwc), st)
        g_phase = torch.empty(n, **f32)          # d/dphase of the zero phases: discarded
        _call("frb_wave_chain_bwd", L.frb_wave_chain_bwd, n, _ptr(colors), _ptr(zero_phase), 1, _ptr(gwc),
              _ptr(grad2d), _ptr(g_phase), st)
        g = _project_backward((positions, scales, rotations), cam_vecs, n_views, grad2d, mode=MODE_FOURIER)
        return (*g, None)


class FourierGaussianRenderer(nn.Module):
    """Holographic Fourier Gaussian Splatting renderer - CUDA drop-in for the reference module of the same
    name (scripts/models/differentiable_renderer.py:1500-1774).

    As in the reference the wavelengths are (optionally learnable) parameters that the rendered image does
    not depend on (DR:1686-1691 computes phases that DR:1693-1738 never uses), ``rotations`` enter through the
    projected covariance only, and the depth map returned with ``return_depth=True`` is all zeros (DR:1759-1764).
    """

    def __init__(self, image_width: int, image_height: int,
                 background: Tuple[float, float, float] = (0.0, 0.0, 0.0), wavelength_r: float = 0.0635,
                 wavelength_g: float = 0.05, wavelength_b: float = 0.041, learnable_wavelengths: bool = True,
                 focal_depth: float = 0.5):
        super().__init__()
        self.width = image_width
        self.height = image_height
        self.focal_depth = focal_depth
        self.register_buffer("background", torch.tensor(background))
        self._background_host = np.asarray(background, np.float32)
        u = torch.fft.fftfreq(image_width)
        v = torch.fft.fftfreq(image_height)
        V, U = torch.meshgrid(v, u, indexing="ij")
        self.register_buffer("U", U)
        self.register_buffer("V", V)
        self.register_buffer("U2_V2", U ** 2 + V ** 2)
        wavelengths = torch.tensor([wavelength_r, wavelength_g, wavelength_b])
        if learnable_wavelengths:
            self.wavelengths = nn.Parameter(wavelengths)
        else:
            self.register_buffer("wavelengths", wavelengths)
        self.learnable_wavelengths = learnable_wavelengths
        self.wavelength_min = 0.01
        self.wavelength_max = 0.5

    def _get_constrained_wavelengths(self) -> torch.Tensor:
        return torch.clamp(torch.abs(self.wavelengths), self.wavelength_min, self.wavelength_max)

    def render_batch(self, positions, scales, rotations, colors, opacities, cameras):
        """(B, N, .) inputs, B cameras -> image (B, 3, H, W); the maximum is taken per view."""
        B, N = positions.shape[0], positions.shape[1]
        cams = list(cameras) if isinstance(cameras, (list, tuple)) else [cameras] * B
        t = _check_inputs(positions=positions.reshape(B * N, 3), scales=scales.reshape(B * N, 3),
                          rotations=rotations.reshape(B * N, 4), colors=colors.reshape(B * N, 3),
                          opacities=opacities.reshape(B * N))
        if N == 0:                                                                     # DR:1651-1657
            return empty_cloud_result(B, self.height, self.width, self._background_host, t["positions"], t["colors"],
                                      t["opacities"])[0]
        cam_vecs = np.ascontiguousarray(np.stack([camera_vector(c, self.width, self.height) for c in cams]),
                                        np.float32)
        cfg = (cam_vecs, B, int(self.width), int(self.height), 32000.0, self._background_host)
        with torch.cuda.device(t["positions"].device):
            return _FourierRenderFn.apply(t["positions"], t["scales"], t["rotations"], t["colors"],
                                          t["opacities"], cfg)

    def forward(self, positions: torch.Tensor, scales: torch.Tensor, rotations: torch.Tensor,
                colors: torch.Tensor, opacities: torch.Tensor, camera, return_depth: bool = False,
                phases: Optional[torch.Tensor] = None):
        image = self.render_batch(positions.unsqueeze(0), scales.unsqueeze(0), rotations.unsqueeze(0),
                                  colors.unsqueeze(0), opacities.reshape(1, -1), [camera]).squeeze(0)
        if return_depth:
            return image, torch.zeros(self.height, self.width, device=image.device)    # DR:1759-1764
        return image

    def extra_repr(self) -> str:
        lam = self._get_constrained_wavelengths()
        return (f"size=({self.height}, {self.width}), λ_rgb=[{lam[0]:.4f}, {lam[1]:.4f}, {lam[2]:.4f}], "
                f"learnable={self.learnable_wavelengths}")
