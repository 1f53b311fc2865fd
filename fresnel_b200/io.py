"""Gaussian cloud files -> device tensors in the renderer's layout (and back).

``load_gaussians_from_binary`` / ``save_gaussians_to_binary`` keep the names, arguments and dict keys of the
reference functions (scripts/models/differentiable_renderer.py:1461-1497); the ``.ply`` pair is the Python face
of ``GaussianCloud::load_ply`` / ``save_ply`` (src/core/renderer/renderer.cpp:649-793), which the reference only
has in C++.  With ``device=`` the file body goes to the GPU in ONE pinned H2D copy and is split (and, for .ply,
de-parameterised: exp scale, SH-DC colour, sigmoid opacity) by ``frb_unpack_gaussians``; no per-field host work.
Without ``device`` the .bin loader returns CPU tensors exactly like the reference (a reshape, no arithmetic);
the .ply transforms exist on the GPU only.
"""

from __future__ import annotations

from typing import Dict, Optional, Tuple

import numpy as np
import torch

from . import _lib

ROW = 14
KEYS = ("positions", "scales", "rotations", "colors", "opacities")
PLY_PROPERTIES = ("x", "y", "z", "scale_0", "scale_1", "scale_2", "rot_0", "rot_1", "rot_2", "rot_3",
                  "f_dc_0", "f_dc_1", "f_dc_2", "opacity")


def _stream(device):
    return torch._C._cuda_getCurrentRawStream(device.index if device.index is not None else torch.cuda.current_device())


def _rows_to_device(rows: np.ndarray, device, ply: bool) -> Dict[str, torch.Tensor]:
    device = torch.device(device)
    if device.type != "cuda":
        raise TypeError("fresnel_b200.io: device must be a CUDA device (the split / transform kernel has no CPU path)")
    n = rows.shape[0]
    f32 = dict(dtype=torch.float32, device=device)
    out = {"positions": torch.empty(n, 3, **f32), "scales": torch.empty(n, 3, **f32),
           "rotations": torch.empty(n, 4, **f32), "colors": torch.empty(n, 3, **f32),
           "opacities": torch.empty(n, **f32)}
    if n == 0:
        return out
    with torch.cuda.device(device):
        staged = torch.from_numpy(np.ascontiguousarray(rows, np.float32)).pin_memory()
        dev_rows = staged.to(device, non_blocking=True)
        _lib.check(_lib.lib().frb_unpack_gaussians(n, int(ply), dev_rows.data_ptr(), *(out[k].data_ptr() for k in KEYS),
                                                   _stream(device)), "frb_unpack_gaussians")
    return out


def _device_to_rows(gaussians: Dict[str, torch.Tensor], ply: bool) -> np.ndarray:
    pos = gaussians["positions"]
    if not pos.is_cuda:
        raise TypeError("fresnel_b200.io: tensors must be on a CUDA device for the packed path")
    n = pos.shape[0]
    t = {k: gaussians[k].detach().to(pos.device, torch.float32).contiguous() for k in KEYS}
    rows = torch.empty(n, ROW, dtype=torch.float32, device=pos.device)
    if n:
        with torch.cuda.device(pos.device):
            _lib.check(_lib.lib().frb_pack_gaussians(n, int(ply), *(t[k].data_ptr() for k in KEYS), rows.data_ptr(),
                                                     _stream(pos.device)), "frb_pack_gaussians")
    return rows.cpu().numpy()


def load_gaussians_from_binary(path: str, device=None) -> Dict[str, torch.Tensor]:
    """14 floats per Gaussian (DR:1461-1482).  ``device=None``: CPU tensors, as the reference returns them."""
    data = np.fromfile(path, dtype=np.float32)
    n = len(data) // ROW
    data = data[:n * ROW].reshape(n, ROW)
    if device is None:
        return {"positions": torch.from_numpy(data[:, 0:3].copy()), "scales": torch.from_numpy(data[:, 3:6].copy()),
                "rotations": torch.from_numpy(data[:, 6:10].copy()), "colors": torch.from_numpy(data[:, 10:13].copy()),
                "opacities": torch.from_numpy(data[:, 13].copy())}
    return _rows_to_device(data, device, ply=False)


def save_gaussians_to_binary(path: str, gaussians: Dict[str, torch.Tensor]) -> None:
    """DR:1485-1497.  CUDA tensors are interleaved on the device and leave in one D2H copy."""
    if gaussians["positions"].is_cuda:
        _device_to_rows(gaussians, ply=False).tofile(path)
        return
    n = gaussians["positions"].shape[0]
    data = np.zeros((n, ROW), dtype=np.float32)
    data[:, 0:3] = gaussians["positions"].detach().cpu().numpy()
    data[:, 3:6] = gaussians["scales"].detach().cpu().numpy()
    data[:, 6:10] = gaussians["rotations"].detach().cpu().numpy()
    data[:, 10:13] = gaussians["colors"].detach().cpu().numpy()
    data[:, 13] = gaussians["opacities"].detach().cpu().numpy()
    data.tofile(path)


def read_ply_rows(path: str) -> np.ndarray:
    """Header parse of GaussianCloud::load_ply (renderer.cpp:727-752): the vertex count comes from the
    ``element vertex`` line, the body is ``count`` rows of 14 little-endian floats after ``end_header``."""
    with open(path, "rb") as f:
        count, done = 0, False
        while True:
            line = f.readline()
            if not line:
                break
            text = line.decode("ascii", "replace").rstrip("\n").rstrip("\r")
            if "element vertex" in text:
                parts = text.split()
                count = int(parts[2]) if len(parts) > 2 else 0
            elif text == "end_header":
                done = True
                break
        if not done or count == 0:
            raise ValueError("Invalid PLY header or no vertices")          # renderer.cpp:748-751
        body = np.fromfile(f, dtype="<f4", count=count * ROW)
    if body.size != count * ROW:
        raise ValueError(f"Failed reading Gaussian {body.size // ROW}")    # renderer.cpp:758-761
    return body.reshape(count, ROW)


def load_gaussians_from_ply(path: str, device="cuda") -> Dict[str, torch.Tensor]:
    """3DGS-style .ply written by the reference viewer -> activated parameters on ``device``."""
    return _rows_to_device(read_ply_rows(path), device, ply=True)


def save_gaussians_to_ply(path: str, gaussians: Dict[str, torch.Tensor]) -> None:
    """GaussianCloud::save_ply (renderer.cpp:649-721): log scale, SH-DC colour, logit opacity."""
    rows = _device_to_rows(gaussians, ply=True)
    with open(path, "wb") as f:
        f.write(b"ply\nformat binary_little_endian 1.0\n")
        f.write(f"element vertex {rows.shape[0]}\n".encode())
        for p in PLY_PROPERTIES:
            f.write(f"property float {p}\n".encode())
        f.write(b"end_header\n")
        f.write(rows.astype("<f4").tobytes())
