"""Pinhole camera with the reference's interface.

Mirrors ``Camera`` of the reference (scripts/models/differentiable_renderer.py:24-95) and
``create_camera_from_pose`` (scripts/training/train_gaussian_decoder.py:684-757) so that
callers can swap ``from models.differentiable_renderer import Camera`` for this module.
Any object with the same attributes (fx, fy, cx, cy, near, far, view_matrix) is accepted by
the renderers, including the reference's own ``Camera``.
"""

from __future__ import annotations

import math
from typing import Tuple

import numpy as np
import torch


class Camera:
    """Simple pinhole camera model (OpenGL convention: the camera looks down -Z)."""

    def __init__(self, fx: float, fy: float, cx: float, cy: float, width: int, height: int,
                 near: float = 0.01, far: float = 100.0):
        self.fx = fx
        self.fy = fy
        self.cx = cx
        self.cy = cy
        self.width = width
        self.height = height
        self.near = near
        self.far = far
        self.view_matrix = torch.eye(4)

    def set_view(self, view_matrix: torch.Tensor):
        """Set view matrix (world-to-camera transform)."""
        self.view_matrix = view_matrix

    def project(self, points_3d: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """Project (N, 3) world-space points to pixel coordinates and depths (DR:54-85)."""
        ones = torch.ones(points_3d.shape[0], 1, device=points_3d.device)
        points_homo = torch.cat([points_3d, ones], dim=1)
        view = self.view_matrix.to(points_3d.device)
        points_cam = (view @ points_homo.T).T[:, :3]
        x, y, z = points_cam[:, 0], points_cam[:, 1], points_cam[:, 2]
        z = torch.clamp(z.abs(), min=self.near) * torch.sign(z + 1e-8)
        u = self.fx * x / (-z) + self.cx
        v = self.fy * (-y) / (-z) + self.cy
        return torch.stack([u, v], dim=1), -z

    def get_intrinsics(self) -> torch.Tensor:
        """3x3 intrinsic matrix (DR:87-95)."""
        return torch.tensor([[self.fx, 0, self.cx], [0, self.fy, self.cy], [0, 0, 1]],
                            dtype=torch.float32)


def create_camera_from_pose(elevation_rad: float, azimuth_rad: float, render_size: int,
                            focal_length_mult: float = 0.8, distance: float = 2.0) -> Camera:
    """Look-at camera on a sphere of radius ``distance`` around the origin.

    Same construction as train_gaussian_decoder.py:684-757 (angles in radians): position from
    (elevation, azimuth), forward towards the origin, rows of the rotation [right; up; -forward],
    translation -R @ position, built in float64 and stored as float32.
    """
    ce, se = math.cos(elevation_rad), math.sin(elevation_rad)
    pos = np.array([distance * ce * math.sin(azimuth_rad), distance * se,
                    distance * ce * math.cos(azimuth_rad)])
    fwd = -pos
    n = np.linalg.norm(fwd)
    fwd = np.array([0.0, 0.0, -1.0]) if n < 1e-6 else fwd / n
    right = np.cross(fwd, np.array([0.0, 1.0, 0.0]))
    n = np.linalg.norm(right)
    right = np.array([1.0, 0.0, 0.0]) if n < 1e-6 else right / n
    up = np.cross(right, fwd)
    rot = np.stack([right, up, -fwd])
    view = torch.eye(4)
    view[:3, :3] = torch.from_numpy(rot).float()
    view[:3, 3] = torch.from_numpy(-rot @ pos).float()
    cam = Camera(render_size * focal_length_mult, render_size * focal_length_mult,
                 render_size / 2, render_size / 2, render_size, render_size)
    cam.set_view(view)
    return cam


def camera_vector(camera, width: int, height: int) -> np.ndarray:
    """The 20-float C-ABI camera (include/fresnel_b200.h): view rows 0..2, fx, fy, cx, cy,
    width, height, near, far.  ``width``/``height`` are the renderer's, as in the reference
    (DR:541-543 cull against self.width / self.height, not the camera's).

    The reference moves ``camera.view_matrix`` to the device (DR:151); reading it back is a blocking
    device->host copy (and illegal inside a stream capture), so the host copy of the 12 view floats is cached on the
    camera object and reused while the matrix is the same tensor at the same version (in-place edits bump
    ``_version``; ``set_view`` installs a new tensor)."""
    vm = camera.view_matrix
    key = (id(vm), getattr(vm, "_version", None))
    hit = getattr(camera, "_frb_view_cache", None)
    if hit is not None and hit[0] == key:
        view12 = hit[1]
    else:
        view12 = vm.detach().to("cpu", torch.float32).numpy()[:3, :].reshape(-1).copy()
        try:
            camera._frb_view_cache = (key, view12, vm)      # vm kept alive: its id cannot be reused
        except AttributeError:          # objects with __slots__: no cache
            pass
    out = np.empty(20, np.float32)
    out[:12] = view12
    out[12:] = (camera.fx, camera.fy, camera.cx, camera.cy, width, height, camera.near, camera.far)
    return out
