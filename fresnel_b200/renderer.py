"""Drop-in renderer modules over the C-ABI library.

``TileBasedRenderer`` has the constructor and ``forward`` signature of the reference module
(scripts/models/differentiable_renderer.py:434-450, 489-511) and is what
scripts/training/train_gaussian_decoder.py:1212-1221 calls once per view.  The work is done by
the CUDA kernels in ``csrc/`` through ``include/fresnel_b200.h``; this file only allocates
device buffers, sequences the launches on the current stream and registers the backward pass
as a ``torch.autograd.Function``.  There is no CPU path: non-CUDA inputs raise.
"""

from __future__ import annotations

import ctypes
import math
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn as nn

from . import _lib
from .camera import camera_vector

TILE = 16
RECORD_FLOATS = 12
DEFAULT_T_EPS = 2.0 ** -20   # pixel stops once transmittance < t_eps; 0 = never (exact mode)
INSTANCE_BYTES = 2 * 12 + 48 + 32            # staged path: keys+ids (two copies), sorted record, side record
FUSED_INSTANCE_BYTES = 8                     # whole-pass path: depth rank + Gaussian id (records are gathered by TMA)
SYNC_FREE_BUDGET_BYTES = 4 << 30             # worst-case instance buffers above this use the host-sync path


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _stream():
    # raw cudaStream_t of torch's current stream (the stream every kernel of a call is launched on)
    return torch._C._cuda_getCurrentRawStream(torch.cuda.current_device())


class StageTimer:
    """Per-stage CUDA-event timing of the C-ABI calls (used by bench.py for the roofline line).

    with StageTimer() as t: ...render...; t.summary() -> {stage: [ms, ...]} after a synchronize.
    Events are recorded on the current stream, the one every kernel is launched on.
    """

    def __init__(self):
        self.events = []

    def __enter__(self):
        global _TIMER
        self._prev, _TIMER = _TIMER, self
        return self

    def __exit__(self, *exc):
        global _TIMER
        _TIMER = self._prev

    def summary(self):
        torch.cuda.synchronize()
        out = {}
        for name, a, b in self.events:
            out.setdefault(name, []).append(a.elapsed_time(b))
        return out


_TIMER: Optional[StageTimer] = None


class FusedStageTimer:
    """Per-stage times of the WHOLE-PASS entry points (frb_tile_render_fwd / _bwd), measured by CUDA events the
    library records between the stages it enqueues (csrc/pipeline.cu): the same code path as the throughput
    figure, unlike StageTimer, which forces the stage-by-stage path.  Eager calls only (not inside a capture).

        with FusedStageTimer() as t: ...render + backward...
        t.summary() -> {stage: [ms, ...]}
    """

    def __enter__(self):
        _lib.check(_lib.lib().frb_stage_timing_enable(1), "frb_stage_timing_enable")
        return self

    def __exit__(self, *exc):
        self._out = self._read()
        _lib.check(_lib.lib().frb_stage_timing_enable(0), "frb_stage_timing_enable")

    def _read(self):
        L = _lib.lib()
        out = {}
        name, ms = ctypes.c_char_p(), ctypes.c_float()
        for i in range(L.frb_stage_timing_count()):
            _lib.check(L.frb_stage_timing_get(i, ctypes.byref(name), ctypes.byref(ms)), "frb_stage_timing_get")
            out.setdefault(name.value.decode(), []).append(float(ms.value))
        return out

    def summary(self):
        return self._out


def _call(name, fn, *args):
    if _TIMER is None:
        _lib.check(fn(*args), name)
        return
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    _lib.check(fn(*args), name)
    b.record()
    _TIMER.events.append((name, a, b))


def _check_inputs(**tensors):
    out = {}
    dev = None
    for name, t in tensors.items():
        if t is None:
            out[name] = None
            continue
        if not isinstance(t, torch.Tensor):
            raise TypeError(f"{name} must be a torch.Tensor")
        if not t.is_cuda:
            raise TypeError(
                f"{name} is on {t.device}: fresnel_b200 renders on CUDA only (no CPU fallback)")
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise TypeError(f"{name} is on {t.device}, expected {dev}")
        if t.dtype != torch.float32:
            t = t.float()
        t = t.contiguous()
        if t.data_ptr() % 16:
            # the kernels read rotations (and write their gradient) as 16-byte vectors; a contiguous view into
            # a flat buffer (torch.split, flat[a:b].view(n, 4)) may start at any float.  The reference accepts
            # any tensor, so re-home it (a differentiable copy) instead of faulting on the device
            t = t.clone()
        out[name] = t
    return out


class TileBins:
    """Sorted tile instance lists of one batch of views (result of the binning stage)."""

    __slots__ = ("m", "m_alloc", "m_dev", "ranges", "sorted_records", "sorted_gids", "sorted_phases", "keys", "order",
                 "records", "rects", "depth_bits", "touched", "tile_order")


def build_bins(positions, scales, rotations, colors, opacities, cam_vecs: np.ndarray, n_views: int,
               width: int, height: int, max_radius: float, phases=None, keep_debug: bool = False,
               sort: bool = True, low_word_fn=None, presort: bool = True, sync: Optional[bool] = None,
               mode: int = 0, project_fn=None) -> TileBins:
    """Projection + binning: everything up to the per-tile sorted record lists.

    Sequences frb_project_fwd -> frb_depth_order -> frb_tile_offsets -> frb_bin_emit ->
    frb_radix_sort_pairs (tile bits only) -> frb_tile_ranges -> frb_gather_records.
    One device->host read (the instance count M) sizes the instance buffers.

    ``sync``: True reads the instance count M back (one 4-byte device->host read, buffers sized exactly);
    False never synchronises: buffers are sized for the worst case (every Gaussian touching
    ``max_tiles_per_gaussian`` tiles) and the kernels read M on the device; None (default) picks False
    when the worst case fits in ``SYNC_FREE_BUDGET_BYTES``.

    ``sort=False`` emits in index order and sorts all 64 key bits (test path).  ``low_word_fn(bins)``
    replaces the depth bits as the low key word (ASM: the depth-plane index).  ``presort=False`` skips
    the depth order altogether (order-free renderers): lists come out in ascending Gaussian index.
    ``mode``: projection mode of frb_project_fwd_mode (0 tile, 1 dense, 2 Fourier); modes 1 and 2 have
    image-sized rectangles, so they always take the exact-size (sync) path.  ``project_fn(bins, cam)`` replaces the
    projection kernel (it must fill ``bins.records``, ``bins.depth_bits`` and ``bins.touched``).
    """
    L = _lib.lib()
    dev = positions.device
    n = positions.shape[0]
    st = _stream()
    i32 = dict(dtype=torch.int32, device=dev)
    tiles_x, tiles_y = (width + TILE - 1) // TILE, (height + TILE - 1) // TILE
    n_tiles = n_views * tiles_x * tiles_y
    cam = np.ascontiguousarray(cam_vecs, np.float32)

    b = TileBins()
    b.tile_order = None
    b.records = torch.empty(n, RECORD_FLOATS, dtype=torch.float32, device=dev)
    b.depth_bits = torch.empty(n, **i32)
    b.touched = torch.empty(n, **i32)
    b.rects = torch.empty(n, 4, **i32) if keep_debug else None
    if project_fn is not None:
        project_fn(b, cam)
    else:
        _call("frb_project_fwd", L.frb_project_fwd_mode, n, n_views, _ptr(positions), _ptr(scales), _ptr(rotations),
              _ptr(colors), _ptr(opacities), cam.ctypes.data, float(max_radius), int(mode), _ptr(b.records),
              _ptr(b.rects), _ptr(b.depth_bits), _ptr(b.touched), None, st)

    if low_word_fn is not None:
        b.depth_bits = low_word_fn(b)
    b.order = None
    rank = None
    sort_err = None
    if sort and presort and n > 0:
        b.order = torch.empty(n, **i32)
        rank = torch.empty(n, **i32)
        ws = torch.empty(L.frb_depth_order_workspace_bytes(n), dtype=torch.uint8, device=dev)
        sort_err = (ws, L.frb_depth_order_error_word(n, _ptr(ws)))       # ws kept alive with its error word
        if low_word_fn is None and project_fn is None:
            # keys = depth bits over the cameras' [near, far] only: fewer radix passes for narrow slabs (csrc/sort.cu)
            cam2 = cam.reshape(-1, 20)
            _call("frb_depth_order", L.frb_depth_order_range, n, _ptr(b.depth_bits), float(cam2[:, 18].min()),
                  float(cam2[:, 19].max()), _ptr(b.order), _ptr(rank), _ptr(ws), st)
        else:
            _call("frb_depth_order", L.frb_depth_order_rank, n, _ptr(b.depth_bits), _ptr(b.order), _ptr(rank),
                  _ptr(ws), st)

    if (sort and presort and low_word_fn is None and project_fn is None and n > 0 and TILE_LISTS
            and n <= L.frb_tile_lists_max_gaussians() and n_tiles <= L.frb_tile_lists_max_tiles()):
        return _tile_lists(b, rank, sort_err, n, n_views, n_tiles, width, height, max_radius, phases, keep_debug, sync,
                           mode, dev, st)

    offsets = torch.empty(n + 1, **i32)
    ws = torch.empty(max(L.frb_scan_workspace_bytes(n), 4), dtype=torch.uint8, device=dev)
    _call("frb_tile_offsets", L.frb_tile_offsets, n, _ptr(b.touched), _ptr(b.order), _ptr(offsets), _ptr(ws), st)
    span = -(-(2 * int(math.ceil(max_radius)) + 2) // TILE) + 1       # tiles a rectangle can span per axis
    worst = n * min(tiles_x * tiles_y, span * span)
    if sync is None:
        sync = (keep_debug or not sort or phases is not None or low_word_fn is not None or not presort
                or mode != 0 or project_fn is not None or worst * INSTANCE_BYTES > SYNC_FREE_BUDGET_BYTES)
    if sync:
        m = int(offsets[n].item())      # the one host sync of the forward pass
        m_dev = None
    else:
        m = worst                       # capacity; the true count stays on the device (offsets[n])
        m_dev = offsets[n:]
    b.m = m
    # buffers are sized for a rounded-up instance count (views of length m are handed out): in an optimisation
    # loop m drifts a little every step, and every new size class is a cudaMalloc of tens of milliseconds
    cap = b.m_alloc = instance_capacity(m) if sync else m

    b.ranges = torch.empty(n_tiles, 2, **i32)
    b.keys = torch.empty(cap, dtype=torch.int64, device=dev)[:m]
    b.sorted_gids = torch.empty(cap, **i32)[:m]
    b.sorted_records = torch.empty(max(cap, 1), RECORD_FLOATS, dtype=torch.float32, device=dev)[:max(m, 1)]
    b.sorted_phases = None
    if m > 0:
        _call("frb_bin_emit", L.frb_bin_emit, n, n_views, width, height, _ptr(b.records), _ptr(b.depth_bits),
              _ptr(b.order), _ptr(offsets), _ptr(b.keys), _ptr(b.sorted_gids), st)
        keys_tmp = torch.empty(cap, dtype=torch.int64, device=dev)
        vals_tmp = torch.empty(cap, **i32)
        ws = torch.empty(max(L.frb_sort_workspace_bytes(m), L.frb_sort_workspace_bytes(cap)), dtype=torch.uint8,
                         device=dev)
        tile_bits = max(1, int(math.ceil(math.log2(max(n_tiles, 2)))))
        begin = 32 if sort else 0
        if sync:
            _call("frb_radix_sort_pairs", L.frb_radix_sort_pairs, m, _ptr(b.keys), _ptr(b.sorted_gids),
                  _ptr(keys_tmp), _ptr(vals_tmp), begin, 32 + tile_bits, _ptr(ws), st)
        else:
            _call("frb_radix_sort_pairs", L.frb_radix_sort_pairs_dev, m, _ptr(m_dev), _ptr(b.keys),
                  _ptr(b.sorted_gids), _ptr(keys_tmp), _ptr(vals_tmp), begin, 32 + tile_bits, _ptr(ws), st)
    if phases is not None:
        b.sorted_phases = torch.empty(max(cap, 1), dtype=torch.float32, device=dev)[:max(m, 1)]
    if sync:
        _call("frb_ranges_and_gather", L.frb_ranges_and_gather, m, _ptr(b.keys), _ptr(b.sorted_gids), n_tiles,
              _ptr(b.ranges), _ptr(b.records), _ptr(b.sorted_records), _ptr(phases), _ptr(b.sorted_phases), st)
    else:
        _call("frb_ranges_and_gather", L.frb_ranges_and_gather_dev, m, _ptr(m_dev), _ptr(b.keys),
              _ptr(b.sorted_gids), n_tiles, _ptr(b.ranges), _ptr(b.records), _ptr(b.sorted_records), _ptr(phases),
              _ptr(b.sorted_phases), st)
    b.m_dev = m_dev
    return b


TILE_LISTS = True       # tile lists by counting + bitmap ranking (csrc/tile_lists.cu); False = the 64-bit key sort


def _tile_lists(b: TileBins, rank, sort_err, n, n_views, n_tiles, width, height, max_radius, phases, keep_debug, sync,
                mode, dev, st):
    """Binning without a sort over the instances: frb_tile_count -> frb_tile_scan -> frb_tile_emit ->
    frb_tile_rank_gather (csrc/tile_lists.cu).  Same lists, bit for bit, as the key sort (``b.keys`` is produced
    only with ``keep_debug``)."""
    L = _lib.lib()
    i32 = dict(dtype=torch.int32, device=dev)
    ws = torch.empty(L.frb_tile_lists_workspace_bytes(n, n_tiles), dtype=torch.uint8, device=dev)
    worst = worst_case_instances(n, n_views, width, height, max_radius)
    if sync is None:
        sync = keep_debug or phases is not None or mode != 0 or worst * INSTANCE_BYTES > SYNC_FREE_BUDGET_BYTES
    b.ranges = torch.empty(n_tiles, 2, **i32)
    tile_order = torch.empty(n_tiles, **i32)
    m_out = torch.empty(2, **i32)       # [instance count, status]
    # sync: scan without a capacity, read the instance count back (the one host sync), size the buffers exactly;
    # otherwise the buffers hold the worst case and the count stays on the device
    cap = (2 ** 31 - 1) if sync else worst
    _call("frb_tile_count_scan", L.frb_tile_count_scan, n, n_views, width, height, _ptr(b.records), cap, _ptr(b.ranges),
          _ptr(tile_order), _ptr(m_out), sort_err[1] if sort_err is not None else None, _ptr(ws), st)
    if sync:
        m, status = m_out.tolist()           # the one host sync: count and status in one read
        if status & 2:
            raise _lib.FresnelB200Error("fresnel_b200: the depth sort's decoupled look-back gave up (SPIN_LIMIT); "
                                        "the tile lists of this call are not trustworthy")
        m_dev = None
        cap = b.m_alloc = instance_capacity(m)
    else:
        m, m_dev = worst, m_out[:1]
        b.m_alloc = worst
    b.m, b.m_dev = m, m_dev
    b.tile_order = tile_order
    b.sorted_gids = torch.empty(cap, **i32)[:m]
    b.sorted_records = torch.empty(max(cap, 1), RECORD_FLOATS, dtype=torch.float32, device=dev)[:max(m, 1)]
    b.sorted_phases = torch.empty(max(cap, 1), dtype=torch.float32, device=dev)[:max(m, 1)] if phases is not None else None
    b.keys = torch.empty(cap, dtype=torch.int64, device=dev)[:m] if keep_debug else None
    if m > 0:
        inst_rank = torch.empty(cap, **i32)
        _call("frb_tile_emit", L.frb_tile_emit, n, n_views, width, height, _ptr(b.records), _ptr(rank), cap,
              _ptr(ws), _ptr(inst_rank), st)
        _call("frb_tile_rank_gather", L.frb_tile_rank_gather, n, n_tiles, _ptr(tile_order), _ptr(b.ranges),
              _ptr(inst_rank), _ptr(b.order), _ptr(b.records), _ptr(b.depth_bits), _ptr(phases), _ptr(b.sorted_gids),
              _ptr(b.sorted_records), _ptr(b.sorted_phases), _ptr(b.keys), st)
    return b


def instance_capacity(m: int) -> int:
    """Allocation size for m tile instances: m rounded up to 1/16 of the next power of two (at least 65,536), i.e. at
    most ~12 % more, so that nearby instance counts share one size."""
    if m <= 0:
        return 0
    g = max(1 << 16, 1 << max(0, (m - 1).bit_length() - 4))
    return -(-m // g) * g


def worst_case_instances(n: int, n_views: int, width: int, height: int, max_radius: float) -> int:
    """Upper bound on the number of tile instances: a rectangle is at most 2*max_radius + 2 pixels wide."""
    tiles_x, tiles_y = (width + TILE - 1) // TILE, (height + TILE - 1) // TILE
    span = -(-(2 * int(math.ceil(max_radius)) + 2) // TILE) + 1
    return n * min(tiles_x * tiles_y, span * span)


def empty_cloud_result(n_views: int, height: int, width: int, background, positions, colors, opacities):
    """N = 0: the reference's "no visible Gaussians" branch (DR:545-552) - the background image and zero depth /
    alpha, tied to the inputs by the same zero-valued anchor so that ``backward()`` runs."""
    dev = positions.device
    anchor = (colors.sum() + opacities.sum() + positions.sum()) * 0.0
    bg = torch.as_tensor(np.asarray(background, np.float32), device=dev)
    image = bg.view(1, 3, 1, 1).expand(n_views, 3, height, width) + anchor
    zeros = torch.zeros(n_views, height, width, dtype=torch.float32, device=dev)
    return image, zeros + anchor, zeros.clone()


FUSED_CALLS = True      # one C call per pass (frb_tile_render_fwd / _bwd); False = stage by stage


_LAYOUT_CACHE = {}


def _tile_layout(n, n_views, width, height, max_radius):
    """(capacity, FrbTileLayout) of the whole-pass entry points, cached per problem shape."""
    key = (n, n_views, width, height, max_radius)
    hit = _LAYOUT_CACHE.get(key)
    if hit is None:
        cap = worst_case_instances(n, n_views, width, height, max_radius)
        lay = _lib.TileLayout()
        _lib.check(_lib.lib().frb_tile_layout(n, n_views, width, height, cap, ctypes.byref(lay)), "frb_tile_layout")
        if len(_LAYOUT_CACHE) > 256:
            _LAYOUT_CACHE.clear()
        hit = _LAYOUT_CACHE[key] = (cap, max(lay.persist_bytes, 256), max(lay.scratch_bytes, 256))
    return hit


class _TileRenderFusedFn(torch.autograd.Function):
    """TileBasedRenderer through the whole-pass C entry points: sync-free, two ctypes calls per frame."""

    @staticmethod
    def forward(ctx, positions, scales, rotations, colors, opacities, cfg):
        (cam, n_views, width, height, bg, max_radius, t_eps, phase_amp) = cfg[:8]
        L = _lib.lib()
        dev = positions.device
        n = positions.shape[0]
        cap, persist_bytes, scratch_bytes = _tile_layout(n, n_views, width, height, max_radius)
        persist = torch.empty(persist_bytes, dtype=torch.uint8, device=dev)
        scratch = torch.empty(scratch_bytes, dtype=torch.uint8, device=dev)
        # one buffer: [image 3 | depth 1 | alpha 1] x views x H x W (image and depth adjacent: one D2H copy)
        hw = height * width
        out = torch.empty(5 * n_views * hw, dtype=torch.float32, device=dev)
        image = out[:3 * n_views * hw].view(n_views, 3, height, width)
        depth = out[3 * n_views * hw:4 * n_views * hw].view(n_views, height, width)
        alpha = out[4 * n_views * hw:].view(n_views, height, width)
        base = out.data_ptr()
        _call("frb_tile_render_fwd", L.frb_tile_render_fwd, n, n_views, _ptr(positions), _ptr(scales),
              _ptr(rotations), _ptr(colors), _ptr(opacities), cam.ctypes.data, float(max_radius), width, height,
              bg.ctypes.data, float(t_eps), cap, _ptr(persist), _ptr(scratch), base, base + 12 * n_views * hw,
              base + 16 * n_views * hw, _stream())
        ctx.cfg, ctx.n, ctx.cap = cfg, n, cap
        ctx.set_materialize_grads(False)
        ctx.save_for_backward(positions, scales, rotations, persist)
        return image, depth, alpha

    @staticmethod
    def backward(ctx, g_image, g_depth, g_alpha):
        (cam, n_views, width, height, bg, max_radius, t_eps, phase_amp) = ctx.cfg[:8]
        positions, scales, rotations, persist = ctx.saved_tensors
        L = _lib.lib()
        dev = positions.device
        n = ctx.n
        f32 = dict(dtype=torch.float32, device=dev)
        g_image = (torch.zeros(n_views, 3, height, width, **f32) if g_image is None
                   else g_image.contiguous().float())
        g_depth = None if g_depth is None else g_depth.contiguous().float()
        g_alpha = None if g_alpha is None else g_alpha.contiguous().float()
        # one allocation: [grad2d 12 | rotations 4 | positions 3 | scales 3 | colors 3 | opacities 1] x n
        # (the float4-typed segments first: 16-byte aligned for every n; the 14 n gradient floats are contiguous)
        buf = torch.empty(26 * n, **f32)
        grad2d, g_rot = buf[:12 * n], buf[12 * n:16 * n].view(n, 4)
        g_pos, g_scl = buf[16 * n:19 * n].view(n, 3), buf[19 * n:22 * n].view(n, 3)
        g_col, g_opa = buf[22 * n:25 * n].view(n, 3), buf[25 * n:]
        _call("frb_tile_render_bwd", L.frb_tile_render_bwd, n, n_views, _ptr(positions), _ptr(scales),
              _ptr(rotations), cam.ctypes.data, width, height, bg.ctypes.data, ctx.cap, _ptr(persist),
              _ptr(g_image), _ptr(g_depth), _ptr(g_alpha), _ptr(grad2d), _ptr(g_pos), _ptr(g_scl), _ptr(g_rot),
              _ptr(g_col), _ptr(g_opa), _stream())
        return g_pos, g_scl, g_rot, g_col, g_opa, None


class _TileRenderFn(torch.autograd.Function):
    """Forward / backward of TileBasedRenderer for n_views views in one call."""

    @staticmethod
    def forward(ctx, positions, scales, rotations, colors, opacities, phases, cfg):
        (cam_vecs, n_views, width, height, bg, max_radius, t_eps, phase_amp) = cfg[:8]
        mode = cfg[8] if len(cfg) > 8 else 0
        L = _lib.lib()
        dev = positions.device
        st = _stream()
        n = positions.shape[0]
        bins = build_bins(positions, scales, rotations, colors, opacities, cam_vecs, n_views, width, height,
                          max_radius, phases=phases, mode=mode)
        f32 = dict(dtype=torch.float32, device=dev)
        image = torch.empty(n_views, 3, height, width, **f32)
        depth = torch.empty(n_views, height, width, **f32)
        alpha = torch.empty(n_views, height, width, **f32)
        state_T = torch.empty(n_views, height, width, **f32)
        state_n = torch.empty(n_views, height, width, dtype=torch.int32, device=dev)
        bg_host = bg
        ckpt = None
        if phases is not None:
            n_tiles = bins.ranges.shape[0]
            ckpt = torch.empty(max(L.frb_phase_ckpt_floats(bins.m_alloc, n_tiles), 1), **f32)
        tile_order = bins.tile_order
        if tile_order is None:
            tile_order = torch.empty(bins.ranges.shape[0], dtype=torch.int32, device=dev)
            _call("frb_tile_schedule", L.frb_tile_schedule, bins.ranges.shape[0], _ptr(bins.ranges), _ptr(tile_order), st)
        _call("frb_composite_fwd", L.frb_composite_fwd_sched, n_views, width, height, _ptr(tile_order),
              _ptr(bins.ranges), _ptr(bins.sorted_records),
                                       _ptr(bins.sorted_phases), float(phase_amp), bg_host.ctypes.data,
                                       float(t_eps), _ptr(image), _ptr(depth), _ptr(alpha), _ptr(state_T),
                                       _ptr(state_n), _ptr(ckpt), st)
        ctx.cfg = cfg
        ctx.n = n
        ctx.set_materialize_grads(False)
        ctx.has_phase = phases is not None
        ctx.save_for_backward(positions, scales, rotations, bins.ranges, bins.sorted_records, bins.sorted_gids,
                              state_T, state_n, tile_order,
                              bins.sorted_phases if phases is not None else positions.new_empty(0),
                              ckpt if ckpt is not None else positions.new_empty(0))
        return image, depth, alpha

    @staticmethod
    def backward(ctx, g_image, g_depth, g_alpha):
        (cam_vecs, n_views, width, height, bg, max_radius, t_eps, phase_amp) = ctx.cfg[:8]
        mode = ctx.cfg[8] if len(ctx.cfg) > 8 else 0
        (positions, scales, rotations, ranges, sorted_records, sorted_gids, state_T, state_n, tile_order,
         sorted_phases, ckpt) = ctx.saved_tensors
        L = _lib.lib()
        dev = positions.device
        st = _stream()
        n = ctx.n
        f32 = dict(dtype=torch.float32, device=dev)
        g_image = (torch.zeros(n_views, 3, height, width, **f32) if g_image is None
                   else g_image.contiguous().float())
        g_depth = None if g_depth is None else g_depth.contiguous().float()
        g_alpha = None if g_alpha is None else g_alpha.contiguous().float()
        grad2d = torch.zeros(n, RECORD_FLOATS, **f32)
        g_phases = torch.zeros(n, **f32) if ctx.has_phase else None
        bg_host = bg
        _call("frb_composite_bwd", L.frb_composite_bwd_sched, n_views, width, height, _ptr(tile_order),
              _ptr(ranges), _ptr(sorted_records),
                                       _ptr(sorted_gids), _ptr(sorted_phases) if ctx.has_phase else None,
                                       float(phase_amp), bg_host.ctypes.data, _ptr(state_T), _ptr(state_n),
                                       _ptr(ckpt) if ctx.has_phase else None, _ptr(g_image), _ptr(g_depth),
                                       _ptr(g_alpha), _ptr(grad2d), _ptr(g_phases), st)
        g_pos = torch.empty(n, 3, **f32)
        g_scl = torch.empty(n, 3, **f32)
        g_rot = torch.empty(n, 4, **f32)
        g_col = torch.empty(n, 3, **f32)
        g_opa = torch.empty(n, **f32)
        cam = cam_vecs
        _call("frb_project_bwd", L.frb_project_bwd_mode, n, n_views, _ptr(positions), _ptr(scales),
              _ptr(rotations), cam.ctypes.data, _ptr(grad2d), int(mode), _ptr(g_pos), _ptr(g_scl), _ptr(g_rot),
              _ptr(g_col), _ptr(g_opa), st)
        return g_pos, g_scl, g_rot, g_col, g_opa, g_phases, None


def render_views(positions, scales, rotations, colors, opacities, cameras: Sequence, width: int, height: int,
                 background=(0.0, 0.0, 0.0), max_radius: float = 64, t_eps: float = DEFAULT_T_EPS,
                 phases=None, phase_amplitude: float = 0.25, mode: int = 0
                 ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Render B views in one pass of the kernels (the loop at train_gaussian_decoder.py:1209-1223).

    positions (B, N, 3), scales (B, N, 3), rotations (B, N, 4), colors (B, N, 3), opacities (B, N),
    phases (B, N) or None, cameras: B camera objects.  Returns image (B, 3, H, W), depth (B, H, W),
    alpha (B, H, W); differentiable with respect to the five (six) inputs.
    """
    B, N = positions.shape[0], positions.shape[1]
    if len(cameras) != B:
        raise ValueError(f"{len(cameras)} cameras for {B} views")
    if B > 32:
        outs = [render_views(positions[i:i + 32], scales[i:i + 32], rotations[i:i + 32], colors[i:i + 32],
                             opacities[i:i + 32], cameras[i:i + 32], width, height, background, max_radius,
                             t_eps, None if phases is None else phases[i:i + 32], phase_amplitude, mode)
                for i in range(0, B, 32)]
        return tuple(torch.cat([o[k] for o in outs]) for k in range(3))
    t = _check_inputs(positions=positions.reshape(B * N, 3), scales=scales.reshape(B * N, 3),
                      rotations=rotations.reshape(B * N, 4), colors=colors.reshape(B * N, 3),
                      opacities=opacities.reshape(B * N),
                      phases=None if phases is None else phases.reshape(B * N))
    if N == 0:
        return empty_cloud_result(B, int(height), int(width), background, t["positions"], t["colors"], t["opacities"])
    cam_vecs = np.ascontiguousarray(np.stack([camera_vector(c, width, height) for c in cameras]), np.float32)
    cfg = (cam_vecs, B, int(width), int(height), np.asarray(background, np.float32), float(max_radius),
           float(t_eps), float(phase_amplitude), int(mode))
    n_total = B * N
    n_tiles = B * (-(-int(width) // TILE)) * (-(-int(height) // TILE))
    lists = (n_total <= _lib.lib().frb_tile_lists_max_gaussians() and n_tiles <= _lib.lib().frb_tile_lists_max_tiles())
    fused = (FUSED_CALLS and _TIMER is None and t["phases"] is None and n_total > 0 and mode == 0 and
             worst_case_instances(n_total, B, int(width), int(height), float(max_radius))
             * (FUSED_INSTANCE_BYTES if lists else INSTANCE_BYTES) <= SYNC_FREE_BUDGET_BYTES)
    with torch.cuda.device(t["positions"].device):
        if fused:
            return _TileRenderFusedFn.apply(t["positions"], t["scales"], t["rotations"], t["colors"],
                                            t["opacities"], cfg)
        return _TileRenderFn.apply(t["positions"], t["scales"], t["rotations"], t["colors"], t["opacities"],
                                   t["phases"], cfg)


class TileBasedRenderer(nn.Module):
    """Memory-efficient tile-based Gaussian renderer - CUDA drop-in for the reference module.

    Same constructor and call signature as the reference ``TileBasedRenderer``
    (scripts/models/differentiable_renderer.py:412-686).  Extra keyword-only knobs:
    ``t_eps`` (early termination threshold on the transmittance; 0 reproduces the reference's
    exhaustive loop) and ``return_alpha`` on ``forward``.
    """

    def __init__(self, image_width: int, image_height: int,
                 background: Tuple[float, float, float] = (0.0, 0.0, 0.0), max_radius: int = 64,
                 use_phase_blending: bool = False, phase_amplitude: float = 0.25, *,
                 t_eps: float = DEFAULT_T_EPS):
        super().__init__()
        self.width = image_width
        self.height = image_height
        self.background = torch.tensor(background)
        self.max_radius = max_radius
        self.use_phase_blending = use_phase_blending
        self.phase_amplitude = phase_amplitude
        self.t_eps = t_eps

    def forward(self, positions: torch.Tensor, scales: torch.Tensor, rotations: torch.Tensor,
                colors: torch.Tensor, opacities: torch.Tensor, camera, return_depth: bool = False,
                phases: Optional[torch.Tensor] = None, return_alpha: bool = False):
        use_phase = self.use_phase_blending and phases is not None          # DR:571-572
        bg = tuple(float(x) for x in self.background.tolist())
        image, depth, alpha = render_views(
            positions.unsqueeze(0), scales.unsqueeze(0), rotations.unsqueeze(0), colors.unsqueeze(0),
            opacities.reshape(1, -1), [camera], self.width, self.height, bg, self.max_radius, self.t_eps,
            phases.reshape(1, -1) if use_phase else None, self.phase_amplitude)
        # squeeze (a view: its backward launches nothing), not image[0] (select backward = fill + copy)
        out: List[torch.Tensor] = [image.squeeze(0)]
        if return_depth:
            out.append(depth.squeeze(0))
        if return_alpha:
            out.append(alpha.squeeze(0))
        return out[0] if len(out) == 1 else tuple(out)

    def render_batch(self, positions, scales, rotations, colors, opacities, cameras, phases=None):
        """All views of a training batch in one call: (B, N, .) inputs, B cameras (or one camera
        shared by all views) -> image (B, 3, H, W), depth (B, H, W), alpha (B, H, W)."""
        B = positions.shape[0]
        cams = list(cameras) if isinstance(cameras, (list, tuple)) else [cameras] * B
        bg = tuple(float(x) for x in self.background.tolist())
        use_phase = self.use_phase_blending and phases is not None
        return render_views(positions, scales, rotations, colors, opacities, cams, self.width, self.height,
                            bg, self.max_radius, self.t_eps, phases if use_phase else None,
                            self.phase_amplitude)


class DifferentiableGaussianRenderer(nn.Module):
    """Differentiable 2D Gaussian splatting renderer - CUDA drop-in for the reference module of the same name
    (scripts/models/differentiable_renderer.py:245-409): every visible Gaussian is evaluated at every pixel
    and composited front to back.

    The reference materialises (512, H, W) Gaussian images; here the dense evaluation runs through the tile
    compositor with per-Gaussian support rectangles of 7.5 standard deviations (csrc/frb_math.h,
    FRB_MODE_DENSE) - the dropped tail is below 1e-12 per Gaussian and pixel.  Visibility is the reference's
    (frustum and a 100-pixel margin on the projected centre, DR:315-318); ``t_eps`` as in TileBasedRenderer.
    """

    def __init__(self, image_width: int, image_height: int,
                 background: Tuple[float, float, float] = (0.0, 0.0, 0.0), *, t_eps: float = DEFAULT_T_EPS):
        super().__init__()
        self.width = image_width
        self.height = image_height
        self.background = torch.tensor(background)
        self.t_eps = t_eps

    def forward(self, positions: torch.Tensor, scales: torch.Tensor, rotations: torch.Tensor,
                colors: torch.Tensor, opacities: torch.Tensor, camera, return_depth: bool = False):
        bg = tuple(float(x) for x in self.background.tolist())
        image, depth, _ = render_views(
            positions.unsqueeze(0), scales.unsqueeze(0), rotations.unsqueeze(0), colors.unsqueeze(0),
            opacities.reshape(1, -1), [camera], self.width, self.height, bg, 32000.0, self.t_eps, None, 0.25,
            mode=1)
        if return_depth:
            return image.squeeze(0), depth.squeeze(0)
        return image.squeeze(0)
