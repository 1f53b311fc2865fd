"""ctypes binding of the C-ABI library (include/fresnel_b200.h).

There is no CPU fallback: if ``libfresnel_b200.so`` is missing or a call fails, this module
raises.  ``lib()`` loads the library built in-tree by ``fresnel_b200.build``.
"""

from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_size_t, c_void_p

from . import build as _build

_LIB = None

P = c_void_p
_SIGNATURES = {
    "frb_version": (c_int, []),
    "frb_error_string": (c_char_p, [c_int]),
    "frb_launch_count": (ctypes.c_ulonglong, []),
    "frb_project_fwd": (c_int, [c_int, c_int, P, P, P, P, P, P, c_float, P, P, P, P, P, P]),
    "frb_project_bwd": (c_int, [c_int, c_int, P, P, P, P, P, P, P, P, P, P, P]),
    "frb_project_fwd_mode": (c_int, [c_int, c_int, P, P, P, P, P, P, c_float, c_int, P, P, P, P, P, P]),
    "frb_project_bwd_mode": (c_int, [c_int, c_int, P, P, P, P, P, c_int, P, P, P, P, P, P]),
    "frb_sort_workspace_bytes": (c_size_t, [c_int]),
    "frb_radix_sort_pairs": (c_int, [c_int, P, P, P, P, c_int, c_int, P, P]),
    "frb_radix_sort_pairs_dev": (c_int, [c_int, P, P, P, P, P, c_int, c_int, P, P]),
    "frb_ranges_and_gather_dev": (c_int, [c_int, P, P, P, c_int, P, P, P, P, P, P]),
    "frb_depth_order_workspace_bytes": (c_size_t, [c_int]),
    "frb_depth_order": (c_int, [c_int, P, P, P, P]),
    "frb_depth_order_rank": (c_int, [c_int, P, P, P, P, P]),
    "frb_depth_order_range": (c_int, [c_int, P, c_float, c_float, P, P, P, P]),
    "frb_depth_sort_in_cluster": (c_int, [c_int]),
    "frb_scan_workspace_bytes": (c_size_t, [c_int]),
    "frb_tile_offsets": (c_int, [c_int, P, P, P, P, P]),
    "frb_bin_emit": (c_int, [c_int, c_int, c_int, c_int, P, P, P, P, P, P, P]),
    "frb_bin_sort_dev": (c_int, [c_int, c_int, c_int, c_int, P, P, P, P, c_int, P, P, P, P, P, c_int, P, P, P]),
    "frb_tile_ranges": (c_int, [c_int, P, c_int, P, P]),
    "frb_gather_records": (c_int, [c_int, P, P, P, P, P, P]),
    "frb_ranges_and_gather": (c_int, [c_int, P, P, c_int, P, P, P, P, P, P]),
    "frb_phase_ckpt_floats": (c_size_t, [c_int, c_int]),
    "frb_composite_fwd": (c_int, [c_int, c_int, c_int, P, P, P, c_float, P, c_float, P, P, P, P, P, P, P]),
    "frb_composite_bwd": (c_int, [c_int, c_int, c_int, P, P, P, P, c_float, P, P, P, P, P, P, P, P, P, P]),
    "frb_tile_layout": (c_int, [c_int, c_int, c_int, c_int, c_int, P]),
    "frb_tile_render_fwd": (c_int, [c_int, c_int, P, P, P, P, P, P, c_float, c_int, c_int, P, c_float, c_int, P, P,
                                    P, P, P, P]),
    "frb_tile_render_bwd": (c_int, [c_int, c_int, P, P, P, P, c_int, c_int, P, c_int, P, P, P, P, P, P, P, P, P, P,
                                    P]),
    "frb_tile_lists_max_gaussians": (c_int, []),
    "frb_tile_lists_max_tiles": (c_int, []),
    "frb_tile_lists_workspace_bytes": (c_size_t, [c_int, c_int]),
    "frb_tile_count": (c_int, [c_int, c_int, c_int, c_int, P, P, P]),
    "frb_tile_scan": (c_int, [c_int, c_int, c_int, P, P, P, P, P, P]),
    "frb_depth_order_error_word": (P, [c_int, P]),
    "frb_tile_count_scan": (c_int, [c_int, c_int, c_int, c_int, P, c_int, P, P, P, P, P, P]),
    "frb_tile_emit": (c_int, [c_int, c_int, c_int, c_int, P, P, c_int, P, P, P]),
    "frb_tile_rank_gather": (c_int, [c_int, c_int, P, P, P, P, P, P, P, P, P, P, P, P]),
    "frb_stage_timing_enable": (c_int, [c_int]),
    "frb_stage_timing_count": (c_int, []),
    "frb_stage_timing_get": (c_int, [c_int, P, P]),
    "frb_tile_schedule": (c_int, [c_int, P, P, P]),
    "frb_composite_fwd_sched": (c_int, [c_int, c_int, c_int, P, P, P, P, c_float, P, c_float, P, P, P, P, P, P, P]),
    "frb_composite_bwd_sched": (c_int, [c_int, c_int, c_int, P, P, P, P, P, c_float, P, P, P, P, P, P, P, P, P, P]),
    "frb_wave_prepare": (c_int, [c_int, P, P, c_int, P, P]),
    "frb_wave_gather": (c_int, [c_int, P, P, P, P]),
    "frb_wave_splat_fwd": (c_int, [c_int, c_int, c_int, P, P, P, P, P, P]),
    "frb_wave_finish_fwd": (c_int, [c_int, c_int, c_int, P, P, P, P, P, P]),
    "frb_wave_finish_bwd": (c_int, [c_int, c_int, c_int, P, P, P, P, P, P, P, P]),
    "frb_wave_splat_bwd": (c_int, [c_int, c_int, c_int, c_int, P, P, P, P, P, P, c_float, P, P, P]),
    "frb_wave_chain_bwd": (c_int, [c_int, P, P, c_int, P, P, P, P]),
    "frb_fourier_finish_fwd": (c_int, [c_int, c_int, c_int, P, P, P, P, P]),
    "frb_fourier_finish_bwd": (c_int, [c_int, c_int, c_int, P, P, P, P, P, P, P]),
    "frb_unpack_gaussians": (c_int, [c_int, c_int, P, P, P, P, P, P, P]),
    "frb_pack_gaussians": (c_int, [c_int, c_int, P, P, P, P, P, P, P]),
    "frb_decode_head_fwd": (c_int, [c_int, c_int, c_int, c_int, P, P, P, P, c_float, c_float, P, c_int, P, P, P, P, P, P]),
    "frb_decode_head_bwd": (c_int, [c_int, c_int, c_int, c_int, P, P, c_float, c_float, P, c_int, P, P, P, P, P, P,
                                    P, P]),
    "frb_decode_head_fwd_ex": (c_int, [c_int, c_int, c_int, c_int, P, P, P, P, P, c_int, P, P, P, P, P, P]),
    "frb_decode_head_bwd_ex": (c_int, [c_int, c_int, c_int, c_int, P, P, P, c_int, P, P, P, P, P, P, P, P, P]),
    "frb_recon_loss_workspace_bytes": (c_size_t, []),
    "frb_recon_loss_fwd": (c_int, [ctypes.c_longlong, ctypes.c_longlong, P, P, P, P, c_float, c_float, P, P, P]),
    "frb_recon_loss_bwd": (c_int, [ctypes.c_longlong, ctypes.c_longlong, P, P, P, P, c_float, c_float, P, P, P, P, P]),
    "frb_recon_loss_fwd_ex": (c_int, [ctypes.c_longlong, ctypes.c_longlong, ctypes.c_longlong, P, P, P, P, c_float,
                                      c_float, c_float, c_int, P, c_float, c_int, P, P, P]),
    "frb_recon_loss_bwd_ex": (c_int, [ctypes.c_longlong, ctypes.c_longlong, ctypes.c_longlong, P, P, P, P, c_float,
                                      c_float, c_float, c_int, P, c_float, c_int, P, P, P, P, P]),
    "frb_composite_fwd_cap": (c_int, [c_int, c_int, c_int, P, P, P, P, c_float, P, c_float, c_float, P, P, P, P, P, P, P]),
    "frb_composite_fwd_gather": (c_int, [c_int, c_int, c_int, P, P, P, c_int, P, P, c_float, c_float, P, P, P, P, P, P]),
    "frb_composite_bwd_gather": (c_int, [c_int, c_int, c_int, P, P, P, c_int, P, P, c_float, P, P, P, P, P, P, P]),
    "frb_composite_bwd_cap": (c_int, [c_int, c_int, c_int, P, P, P, P, P, c_float, P, c_float, P, P, P, P, P, P, P, P,
                                      P]),
    "frb_simple_project_fwd": (c_int, [c_int, c_int, P, P, P, P, P, P, P, P, P]),
    "frb_simple_project_bwd": (c_int, [c_int, c_int, P, P, P, P, P, P, P]),
    "frb_simple_depth_fwd": (c_int, [c_int, c_int, c_int, P, P, P, P, P, P]),
    "frb_simple_depth_bwd": (c_int, [c_int, c_int, c_int, P, P, P, P]),
    "frb_peer_shard_floats": (ctypes.c_longlong, [c_int, c_int, ctypes.c_longlong]),
    "frb_peer_adam_step": (c_int, [c_int, c_int, ctypes.c_longlong, P, P, P, P, P, P, ctypes.c_double, ctypes.c_double,
                                   ctypes.c_double, c_float, c_float, P]),
    "frb_asm_assign_planes": (c_int, [c_int, P, c_int, P, P, P]),
    "frb_asm_splat_fwd": (c_int, [c_int, c_int, c_int, c_int, P, P, P, P, P, P]),
    "frb_asm_propagate_fwd": (c_int, [c_int, c_int, c_int, c_int, P, c_float, c_float, P, P, P, P, P, P, P]),
    "frb_asm_propagate_bwd": (c_int, [c_int, c_int, c_int, c_int, P, c_float, c_float, P, P, P, P, P, P, P, P, P]),
}

_OPTIONAL = {}


class TileLayout(ctypes.Structure):
    """FrbTileLayout of include/fresnel_b200.h."""
    _fields_ = [(name, c_size_t) for name in (
        "ranges", "tile_order", "state_T", "state_n", "sorted_gids", "sorted_records", "records", "persist_bytes", "depth_bits",
        "touched", "order", "offsets", "depth_ws", "scan_ws", "keys", "keys_tmp", "vals_tmp", "sort_ws",
        "tile_ws", "inst_rank", "rank", "scratch_bytes")]


class HeadExtras(ctypes.Structure):
    """FrbHeadExtras of include/fresnel_b200.h."""
    _fields_ = [("edge", c_void_p), ("edge_scale_factor", c_float), ("edge_opacity_boost", c_float),
                ("num_zones", c_int), ("zone_boundaries_host", c_void_p), ("zone_centers_host", c_void_p),
                ("pose_trig", c_void_p)]


class FresnelB200Error(RuntimeError):
    pass


def library_path() -> str:
    return _build.LIB


def lib() -> ctypes.CDLL:
    """The loaded C-ABI library.  Raises if it has not been built (no fallback)."""
    global _LIB
    if _LIB is None:
        path = library_path()
        if not os.path.exists(path):
            raise FresnelB200Error(
                f"{path} is missing: build it with `python -m fresnel_b200.build` "
                "(fresnel_b200 has no CPU or PyTorch fallback path)")
        handle = ctypes.CDLL(path)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(handle, name)      # AttributeError if the symbol is not exported
            fn.restype, fn.argtypes = res, args
        for name, (res, args) in _OPTIONAL.items():
            if hasattr(handle, name):
                fn = getattr(handle, name)
                fn.restype, fn.argtypes = res, args
        _LIB = handle
    return _LIB


def check(code: int, what: str) -> None:
    if code != 0:
        msg = lib().frb_error_string(code)
        raise FresnelB200Error(f"{what} failed ({code}): {msg.decode() if msg else '?'}")


def exported_symbols():
    """Names declared in include/fresnel_b200.h (used by the CPU test tier)."""
    return sorted(_SIGNATURES)
