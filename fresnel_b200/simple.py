"""SimplifiedRenderer drop-in over the C-ABI library.

Constructor and ``forward`` signature follow the reference module
(scripts/models/differentiable_renderer.py:1347-1458): isotropic point splats with an integer radius, blended
with "over".  The reference blends back to front; the same sum is evaluated front to back by the tile
compositor (csrc/composite.cu with alpha_max = 1 - 2^-24 for the reference's clamp(alpha, 0, 1)), the
projection / radius rule, the depth map and their backward live in csrc/simple.cu.

Order on exact depth ties: the reference's (stable) descending argsort puts the lower index further back; the
inputs are therefore processed in reversed index order, for which the kernels' stable ascending sort gives the
same sequence.  ``scales`` and ``rotations`` receive zero gradients (the reference returns None for them: the
radius goes through ``.item()`` and rotations are unused).
"""

from __future__ import annotations

from typing import Tuple

import numpy as np
import torch
import torch.nn as nn

from . import _lib
from .camera import camera_vector
from .renderer import RECORD_FLOATS, _call, _check_inputs, _ptr, _stream, build_bins, empty_cloud_result

ALPHA_MAX = 1.0 - 2.0 ** -24       # clamp(alpha, 0, 1) of DR:1430 with a backward that can divide by 1 - alpha


class _SimpleRenderFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, positions, scales, colors, opacities, cfg):
        cam_vecs, width, height, bg = cfg
        L = _lib.lib()
        dev = positions.device
        st = _stream()
        n = positions.shape[0]
        f32 = dict(dtype=torch.float32, device=dev)

        def project(b, cam):
            _call("frb_simple_project_fwd", L.frb_simple_project_fwd, n, 1, _ptr(positions), _ptr(scales), _ptr(colors),
                  _ptr(opacities), cam.ctypes.data, _ptr(b.records), _ptr(b.depth_bits), _ptr(b.touched), st)

        bins = build_bins(positions, scales, None, colors, opacities, cam_vecs, 1, width, height, 20.0,
                          project_fn=project)
        image = torch.empty(1, 3, height, width, **f32)
        depth_sum = torch.empty(1, height, width, **f32)           # the compositor's sum c * depth: unused here
        alpha = torch.empty(1, height, width, **f32)
        state_T = torch.empty(1, height, width, **f32)
        state_n = torch.empty(1, height, width, dtype=torch.int32, device=dev)
        _call("frb_composite_fwd", L.frb_composite_fwd_cap, 1, width, height, None, _ptr(bins.ranges),
              _ptr(bins.sorted_records), None, 0.0, bg.ctypes.data, 0.0, ALPHA_MAX, _ptr(image), _ptr(depth_sum),
              _ptr(alpha), _ptr(state_T), _ptr(state_n), None, st)
        depth = torch.empty(1, height, width, **f32)
        hit = torch.empty(1, height, width, dtype=torch.int32, device=dev)
        _call("frb_simple_depth_fwd", L.frb_simple_depth_fwd, 1, width, height, _ptr(bins.ranges),
              _ptr(bins.sorted_records), _ptr(bins.sorted_gids), _ptr(depth), _ptr(hit), st)
        ctx.cfg, ctx.n = cfg, n
        ctx.set_materialize_grads(False)
        ctx.save_for_backward(positions, bins.ranges, bins.sorted_records, bins.sorted_gids, state_T, state_n, hit)
        return image, depth

    @staticmethod
    def backward(ctx, g_image, g_depth):
        cam_vecs, width, height, bg = ctx.cfg
        positions, ranges, sorted_records, sorted_gids, state_T, state_n, hit = ctx.saved_tensors
        L = _lib.lib()
        dev = positions.device
        st = _stream()
        n = ctx.n
        f32 = dict(dtype=torch.float32, device=dev)
        g_image = torch.zeros(1, 3, height, width, **f32) if g_image is None else g_image.contiguous().float()
        grad2d = torch.zeros(n, RECORD_FLOATS, **f32)
        _call("frb_composite_bwd", L.frb_composite_bwd_cap, 1, width, height, None, _ptr(ranges), _ptr(sorted_records),
              _ptr(sorted_gids), None, 0.0, bg.ctypes.data, ALPHA_MAX, _ptr(state_T), _ptr(state_n), None,
              _ptr(g_image), None, None, _ptr(grad2d), None, st)
        if g_depth is not None:
            g_depth = g_depth.contiguous().float()
            _call("frb_simple_depth_bwd", L.frb_simple_depth_bwd, 1, width, height, _ptr(hit), _ptr(g_depth),
                  _ptr(grad2d), st)
        g_pos, g_col, g_opa = torch.empty(n, 3, **f32), torch.empty(n, 3, **f32), torch.empty(n, **f32)
        _call("frb_simple_project_bwd", L.frb_simple_project_bwd, n, 1, _ptr(positions), cam_vecs.ctypes.data,
              _ptr(grad2d), _ptr(g_pos), _ptr(g_col), _ptr(g_opa), st)
        return g_pos, None, g_col, g_opa, None


class SimplifiedRenderer(nn.Module):
    """Simplified renderer for faster training - CUDA drop-in for the reference module of the same name
    (scripts/models/differentiable_renderer.py:1347-1458)."""

    def __init__(self, image_width: int, image_height: int, splat_size: int = 3,
                 background: Tuple[float, float, float] = (0.0, 0.0, 0.0)):
        super().__init__()
        self.width = image_width
        self.height = image_height
        self.splat_size = splat_size                   # unused by the reference's forward as well
        self.background = torch.tensor(background)

    def forward(self, positions: torch.Tensor, scales: torch.Tensor, rotations: torch.Tensor,
                colors: torch.Tensor, opacities: torch.Tensor, camera, return_depth: bool = False):
        t = _check_inputs(positions=positions, scales=scales, colors=colors, opacities=opacities.reshape(-1))
        bg = np.asarray(self.background.tolist(), np.float32)
        if positions.shape[0] == 0:
            image, depth, _ = empty_cloud_result(1, self.height, self.width, bg, t["positions"], t["colors"],
                                                 t["opacities"])
        else:
            cam_vecs = np.ascontiguousarray(camera_vector(camera, self.width, self.height)[None], np.float32)
            cfg = (cam_vecs, int(self.width), int(self.height), bg)
            with torch.cuda.device(t["positions"].device):
                # reversed index order: see the module docstring (depth ties)
                image, depth = _SimpleRenderFn.apply(t["positions"].flip(0), t["scales"].detach().flip(0),
                                                     t["colors"].flip(0), t["opacities"].flip(0), cfg)
        if return_depth:
            return image.squeeze(0), depth.squeeze(0)
        return image.squeeze(0)
