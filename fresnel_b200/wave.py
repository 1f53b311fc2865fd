"""WaveFieldRenderer and ASMWaveFieldRenderer drop-ins over the C-ABI library.

Constructor and ``forward`` signatures follow the reference modules
(scripts/models/differentiable_renderer.py:707-718 / 747-757 and :1082-1115 / 1150-1161).
"""

from __future__ import annotations

from typing import Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn as nn

from . import _lib
from .camera import camera_vector
from .renderer import RECORD_FLOATS, TILE, _call, _check_inputs, _ptr, _stream, build_bins, empty_cloud_result

WC_FLOATS = 8


def _phase_stride(phases: torch.Tensor, n: int) -> int:
    if phases.numel() == n:
        return 1
    if phases.numel() == 3 * n:
        return 3
    raise ValueError(f"phases must have shape (N,) or (N, 3); got {tuple(phases.shape)} for N={n}")


def _prepare_wave_bins(positions, scales, rotations, colors, opacities, phases, stride, cfg, low_word_fn=None,
                       mode: int = 0):
    """Projection, binning (no depth order needed) and the sorted 32-byte (colour cos/sin) side records."""
    cam_vecs, n_views, width, height, max_radius = cfg[:5]
    L = _lib.lib()
    dev = positions.device
    n = positions.shape[0]
    st = _stream()
    bins = build_bins(positions, scales, rotations, colors, opacities, cam_vecs, n_views, width, height,
                      max_radius, presort=low_word_fn is not None, low_word_fn=low_word_fn, mode=mode)
    wc = torch.empty(n, WC_FLOATS, dtype=torch.float32, device=dev)
    _call("frb_wave_prepare", L.frb_wave_prepare, n, _ptr(colors), _ptr(phases), stride, _ptr(wc), st)
    sorted_wc = torch.empty(max(bins.m_alloc, 1), WC_FLOATS, dtype=torch.float32, device=dev)[:max(bins.m, 1)]
    _call("frb_wave_gather", L.frb_wave_gather, bins.m, _ptr(bins.sorted_gids), _ptr(wc), _ptr(sorted_wc), st)
    return bins, sorted_wc


def _project_backward(ctx_inputs, cam_vecs, n_views, grad2d, mode: int = 0):
    positions, scales, rotations = ctx_inputs
    L = _lib.lib()
    n = positions.shape[0]
    f32 = dict(dtype=torch.float32, device=positions.device)
    g_pos, g_scl, g_rot = torch.empty(n, 3, **f32), torch.empty(n, 3, **f32), torch.empty(n, 4, **f32)
    g_col, g_opa = torch.empty(n, 3, **f32), torch.empty(n, **f32)
    cam = np.ascontiguousarray(cam_vecs, np.float32)
    _call("frb_project_bwd", L.frb_project_bwd_mode, n, n_views, _ptr(positions), _ptr(scales), _ptr(rotations),
          cam.ctypes.data, _ptr(grad2d), int(mode), _ptr(g_pos), _ptr(g_scl), _ptr(g_rot), _ptr(g_col), _ptr(g_opa),
          _stream())
    return g_pos, g_scl, g_rot, g_col, g_opa


class _WaveRenderFn(torch.autograd.Function):
    """WaveFieldRenderer.forward DR:747-926 for n_views views."""

    @staticmethod
    def forward(ctx, positions, scales, rotations, colors, opacities, phases, cfg):
        cam_vecs, n_views, width, height, max_radius, bg = cfg
        L = _lib.lib()
        dev = positions.device
        st = _stream()
        n = positions.shape[0]
        stride = _phase_stride(phases, n)
        bins, sorted_wc = _prepare_wave_bins(positions, scales, rotations, colors, opacities, phases, stride, cfg)
        f32 = dict(dtype=torch.float32, device=dev)
        accum = torch.empty(n_views, 8, height, width, **f32)
        rmax = torch.empty(n_views, dtype=torch.int32, device=dev)
        image = torch.empty(n_views, 3, height, width, **f32)
        depth = torch.empty(n_views, height, width, **f32)
        bg_host = np.asarray(bg, np.float32)
        _call("frb_wave_splat_fwd", L.frb_wave_splat_fwd, n_views, width, height, _ptr(bins.ranges),
              _ptr(bins.sorted_records), _ptr(sorted_wc), _ptr(accum), _ptr(rmax), st)
        _call("frb_wave_finish_fwd", L.frb_wave_finish_fwd, n_views, width, height, _ptr(accum), _ptr(rmax),
              bg_host.ctypes.data, _ptr(image), _ptr(depth), st)
        ctx.cfg, ctx.n, ctx.stride = cfg, n, stride
        ctx.phase_shape = phases.shape
        ctx.save_for_backward(positions, scales, rotations, colors, phases, bins.ranges, bins.sorted_records,
                              bins.sorted_gids, sorted_wc, accum, rmax)
        return image, depth

    @staticmethod
    def backward(ctx, g_image, g_depth):
        cam_vecs, n_views, width, height, max_radius, bg = ctx.cfg
        (positions, scales, rotations, colors, phases, ranges, sorted_records, sorted_gids, sorted_wc, accum,
         rmax) = ctx.saved_tensors
        L = _lib.lib()
        dev = positions.device
        st = _stream()
        n = ctx.n
        f32 = dict(dtype=torch.float32, device=dev)
        g_image = (torch.zeros(n_views, 3, height, width, **f32) if g_image is None
                   else g_image.contiguous().float())
        g_depth = None if g_depth is None else g_depth.contiguous().float()
        bg_host = np.asarray(bg, np.float32)
        red = torch.empty(2 * n_views, **f32)
        gpix = torch.empty(n_views, 8, height, width, **f32)
        _call("frb_wave_finish_bwd", L.frb_wave_finish_bwd, n_views, width, height, _ptr(accum), _ptr(rmax),
              bg_host.ctypes.data, _ptr(g_image), _ptr(g_depth), _ptr(red), _ptr(gpix), st)
        grad2d = torch.zeros(n, RECORD_FLOATS, **f32)
        gwc = torch.zeros(n, WC_FLOATS, **f32)
        _call("frb_wave_splat_bwd", L.frb_wave_splat_bwd, n_views, width, height, 0, _ptr(ranges),
              _ptr(sorted_records), _ptr(sorted_wc), _ptr(sorted_gids), None, _ptr(gpix), 1.0, _ptr(grad2d),
              _ptr(gwc), st)
        g_phases = torch.empty(ctx.phase_shape, **f32)
        _call("frb_wave_chain_bwd", L.frb_wave_chain_bwd, n, _ptr(colors), _ptr(phases), ctx.stride, _ptr(gwc),
              _ptr(grad2d), _ptr(g_phases), st)
        g = _project_backward((positions, scales, rotations), cam_vecs, n_views, grad2d)
        return (*g, g_phases, None)


# The per-plane complex fields (forward) and their gradients (backward) are the largest buffers of the ASM renderer
# (384 H W bytes per view: 403 MB at 1024^2) and live only inside one call.  They are kept between calls, one per
# (device, stream, size), instead of going through the caching allocator every step: a block of that size that another
# allocation has split costs a cudaMalloc of tens of milliseconds in the middle of a training loop.
_PLANE_SCRATCH: dict = {}


def _plane_scratch(numel: int, dev: torch.device) -> torch.Tensor:
    if torch.cuda.is_current_stream_capturing():         # a graph owns its memory: plain allocation
        return torch.empty(numel, dtype=torch.float32, device=dev)
    key = (dev.index, torch.cuda.current_stream(dev).cuda_stream, numel)
    buf = _PLANE_SCRATCH.get(key)
    if buf is None:
        if len(_PLANE_SCRATCH) >= 4:
            _PLANE_SCRATCH.clear()
        buf = _PLANE_SCRATCH[key] = torch.empty(numel, dtype=torch.float32, device=dev)
    return buf


class _AsmRenderFn(torch.autograd.Function):
    """ASMWaveFieldRenderer.forward DR:1150-1344 for n_views views (wavelengths are constants)."""

    @staticmethod
    def forward(ctx, positions, scales, rotations, colors, opacities, phases, cfg):
        (cam_vecs, n_views, width, height, max_radius, bg, planes, focal, pitch, wavelengths) = cfg
        L = _lib.lib()
        dev = positions.device
        st = _stream()
        n = positions.shape[0]
        n_planes = int(planes.shape[0])
        stride = _phase_stride(phases, n)

        def plane_words(b):
            idx = torch.empty(n, dtype=torch.int32, device=dev)
            _call("frb_asm_assign_planes", L.frb_asm_assign_planes, n, _ptr(b.records), n_planes,
                  planes.ctypes.data, _ptr(idx), st)
            return idx

        bins, sorted_wc = _prepare_wave_bins(positions, scales, rotations, colors, opacities, phases, stride, cfg,
                                             low_word_fn=plane_words)
        f32 = dict(dtype=torch.float32, device=dev)
        fields = _plane_scratch(n_views * n_planes * 3 * height * width * 2, dev)
        total = torch.empty(n_views, 3, height, width, 2, **f32)
        rmax = torch.empty(n_views, dtype=torch.int32, device=dev)
        image = torch.empty(n_views, 3, height, width, **f32)
        bg_host = np.asarray(bg, np.float32)
        _call("frb_asm_splat_fwd", L.frb_asm_splat_fwd, n_views, width, height, n_planes, _ptr(bins.ranges),
              _ptr(bins.sorted_records), _ptr(sorted_wc), _ptr(bins.keys), _ptr(fields), st)
        _call("frb_asm_propagate_fwd", L.frb_asm_propagate_fwd, n_views, width, height, n_planes,
              planes.ctypes.data, float(focal), float(pitch), wavelengths.ctypes.data, bg_host.ctypes.data,
              _ptr(fields), _ptr(total), _ptr(rmax), _ptr(image), st)
        del fields                     # its FFT is not needed by the backward pass
        ctx.cfg, ctx.n, ctx.stride = cfg, n, stride
        ctx.phase_shape = phases.shape
        ctx.save_for_backward(positions, scales, rotations, colors, phases, bins.ranges, bins.sorted_records,
                              bins.sorted_gids, sorted_wc, bins.keys, total, rmax)
        return image

    @staticmethod
    def backward(ctx, g_image):
        (cam_vecs, n_views, width, height, max_radius, bg, planes, focal, pitch, wavelengths) = ctx.cfg
        (positions, scales, rotations, colors, phases, ranges, sorted_records, sorted_gids, sorted_wc, keys, total,
         rmax) = ctx.saved_tensors
        L = _lib.lib()
        dev = positions.device
        st = _stream()
        n = ctx.n
        n_planes = int(planes.shape[0])
        f32 = dict(dtype=torch.float32, device=dev)
        g_image = g_image.contiguous().float()
        bg_host = np.asarray(bg, np.float32)
        red = torch.empty(2 * n_views, **f32)
        g_total = torch.empty(n_views, 3, height, width, 2, **f32)
        d_fields = _plane_scratch(n_views * n_planes * 3 * height * width * 2, dev)
        _call("frb_asm_propagate_bwd", L.frb_asm_propagate_bwd, n_views, width, height, n_planes,
              planes.ctypes.data, float(focal), float(pitch), wavelengths.ctypes.data, bg_host.ctypes.data,
              _ptr(total), _ptr(rmax), _ptr(g_image), _ptr(red), _ptr(g_total), _ptr(d_fields), st)
        grad2d = torch.zeros(n, RECORD_FLOATS, **f32)
        gwc = torch.zeros(n, WC_FLOATS, **f32)
        _call("frb_wave_splat_bwd", L.frb_wave_splat_bwd, n_views, width, height, n_planes, _ptr(ranges),
              _ptr(sorted_records), _ptr(sorted_wc), _ptr(sorted_gids), _ptr(keys), _ptr(d_fields),
              1.0 / float(width * height), _ptr(grad2d), _ptr(gwc), st)
        g_phases = torch.empty(ctx.phase_shape, **f32)
        _call("frb_wave_chain_bwd", L.frb_wave_chain_bwd, n, _ptr(colors), _ptr(phases), ctx.stride, _ptr(gwc),
              _ptr(grad2d), _ptr(g_phases), st)
        g = _project_backward((positions, scales, rotations), cam_vecs, n_views, grad2d)
        return (*g, g_phases, None)


def _flatten_views(positions, scales, rotations, colors, opacities, phases):
    B, N = positions.shape[0], positions.shape[1]
    ph = phases.reshape(B * N) if phases.numel() == B * N else phases.reshape(B * N, 3)
    return _check_inputs(positions=positions.reshape(B * N, 3), scales=scales.reshape(B * N, 3),
                         rotations=rotations.reshape(B * N, 4), colors=colors.reshape(B * N, 3),
                         opacities=opacities.reshape(B * N), phases=ph)


class WaveFieldRenderer(nn.Module):
    """True wave optics renderer with complex field accumulation - CUDA drop-in for the reference
    ``WaveFieldRenderer`` (scripts/models/differentiable_renderer.py:689-926)."""

    def __init__(self, image_width: int, image_height: int,
                 background: Tuple[float, float, float] = (0.0, 0.0, 0.0), max_radius: int = 64):
        super().__init__()
        self.width = image_width
        self.height = image_height
        self.background = torch.tensor(background)
        self.max_radius = max_radius

    def render_batch(self, positions, scales, rotations, colors, opacities, cameras, phases):
        """(B, N, .) inputs, B cameras -> image (B, 3, H, W), depth (B, H, W)."""
        if phases is None:
            raise ValueError("WaveFieldRenderer requires phases tensor. "
                             "Use PhysicsDirectPatchDecoder to generate phases.")      # DR:779-780
        B = positions.shape[0]
        cams = list(cameras) if isinstance(cameras, (list, tuple)) else [cameras] * B
        t = _flatten_views(positions, scales, rotations, colors, opacities, phases)
        if positions.shape[1] == 0:                                                    # DR:801-808
            img, dep, _ = empty_cloud_result(B, self.height, self.width, self.background.tolist(), t["positions"],
                                             t["colors"], t["opacities"])
            return img, dep
        cam_vecs = np.stack([camera_vector(c, self.width, self.height) for c in cams])
        cfg = (cam_vecs, B, int(self.width), int(self.height), float(self.max_radius),
               tuple(float(x) for x in self.background.tolist()))
        with torch.cuda.device(t["positions"].device):
            return _WaveRenderFn.apply(t["positions"], t["scales"], t["rotations"], t["colors"], t["opacities"],
                                       t["phases"], cfg)

    def forward(self, positions, scales, rotations, colors, opacities, camera, return_depth: bool = False,
                phases: Optional[torch.Tensor] = None):
        if phases is None:
            raise ValueError("WaveFieldRenderer requires phases tensor. "
                             "Use PhysicsDirectPatchDecoder to generate phases.")      # DR:779-780
        image, depth = self.render_batch(positions.unsqueeze(0), scales.unsqueeze(0), rotations.unsqueeze(0),
                                         colors.unsqueeze(0), opacities.reshape(1, -1), [camera],
                                         phases.unsqueeze(0))
        return (image.squeeze(0), depth.squeeze(0)) if return_depth else image.squeeze(0)


class ASMWaveFieldRenderer(nn.Module):
    """Wave field renderer with Angular Spectrum Method propagation - CUDA drop-in for the reference
    ``ASMWaveFieldRenderer`` (scripts/models/differentiable_renderer.py:1068-1344).

    The reference's scalar-wavelength default path raises (SURVEY.md note 4); here a missing
    ``wavelengths_rgb`` means the constructor's ``wavelength`` for all three channels.  Wavelengths
    are treated as constants (the reference's gradient with respect to them is NaN under band limiting).
    """

    def __init__(self, image_width: int, image_height: int,
                 background: Tuple[float, float, float] = (0.0, 0.0, 0.0), max_radius: int = 64,
                 num_depth_planes: int = 16, depth_range: Tuple[float, float] = (0.1, 2.0),
                 focal_depth: float = 0.5, pixel_pitch: float = 1.0 / 256.0, wavelength: float = 0.05):
        super().__init__()
        if not 1 <= num_depth_planes <= 64:
            raise ValueError("num_depth_planes must be in [1, 64]")
        self.width = image_width
        self.height = image_height
        self.max_radius = max_radius
        self.num_depth_planes = num_depth_planes
        self.depth_range = depth_range
        self.focal_depth = focal_depth
        self.pixel_pitch = pixel_pitch
        self.wavelength = wavelength
        self.register_buffer("background", torch.tensor(background))
        self.register_buffer("depth_planes", torch.linspace(depth_range[0], depth_range[1], num_depth_planes))
        # host copies for the C-ABI arguments (reading the buffers back would synchronise once per call)
        self._planes_host = np.ascontiguousarray(self.depth_planes.numpy().astype(np.float32))
        self._background_host = tuple(float(x) for x in background)

    def render_batch(self, positions, scales, rotations, colors, opacities, cameras, phases, wavelengths_rgb=None):
        """(B, N, .) inputs, B cameras -> image (B, 3, H, W)."""
        if phases is None:
            raise ValueError("ASMWaveFieldRenderer requires phases tensor.")            # DR:1187-1188
        B = positions.shape[0]
        cams = list(cameras) if isinstance(cameras, (list, tuple)) else [cameras] * B
        t = _flatten_views(positions, scales, rotations, colors, opacities, phases)
        if positions.shape[1] == 0:                                                    # DR:1208-1212
            return empty_cloud_result(B, self.height, self.width, self._background_host, t["positions"], t["colors"],
                                      t["opacities"])[0]
        cam_vecs = np.stack([camera_vector(c, self.width, self.height) for c in cams])
        if wavelengths_rgb is None:
            wl = np.full(3, self.wavelength, np.float32)
        else:
            wl = np.ascontiguousarray(torch.as_tensor(wavelengths_rgb).detach().float().cpu().numpy().reshape(3))
        planes = self._planes_host
        cfg = (cam_vecs, B, int(self.width), int(self.height), float(self.max_radius),
               self._background_host, planes, float(self.focal_depth),
               float(self.pixel_pitch), wl)
        with torch.cuda.device(t["positions"].device):
            return _AsmRenderFn.apply(t["positions"], t["scales"], t["rotations"], t["colors"], t["opacities"],
                                      t["phases"], cfg)

    def forward(self, positions, scales, rotations, colors, opacities, camera, return_depth: bool = False,
                phases: Optional[torch.Tensor] = None, wavelengths_rgb: Optional[torch.Tensor] = None):
        if phases is None:
            raise ValueError("ASMWaveFieldRenderer requires phases tensor.")            # DR:1187-1188
        image = self.render_batch(positions.unsqueeze(0), scales.unsqueeze(0), rotations.unsqueeze(0),
                                  colors.unsqueeze(0), opacities.reshape(1, -1), [camera], phases.unsqueeze(0),
                                  wavelengths_rgb).squeeze(0)
        if return_depth:
            return image, torch.zeros(self.height, self.width, device=image.device)     # DR:1339-1342
        return image
