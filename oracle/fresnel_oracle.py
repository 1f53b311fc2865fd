"""CPU restatement of the reference renderer hot path (torch CPU + numpy).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Every function cites the
reference lines it follows; paths are relative to ``/root/reference`` and
``DR`` abbreviates ``scripts/models/differentiable_renderer.py``.

Design notes
------------
* The reference evaluates the projection with small ``torch`` matmuls whose
  summation order belongs to the BLAS backend.  The oracle writes every
  contraction out as elementwise fp32 operations in a FIXED left-to-right
  order (no fused multiply-add).  IEEE add/mul/div/sqrt are correctly rounded
  on the CPU and on the GPU, so the CUDA kernels - which use the same order
  with ``__fmul_rn`` / ``__fadd_rn`` - reproduce these numbers bit for bit.
  That is what makes "bit-exact tile assignment and sort order" a testable
  statement.  Against the reference itself the projected floats agree to a
  few ulp (checked by ``oracle/make_golden.py``), and the integer pins agree
  exactly on the margin-filtered golden fixtures.
* Depth order is ``argsort(depth, stable=True)``: ties resolve by ascending
  input index.  ``DR:527`` uses the default (unstable) argsort, whose order on
  ties is implementation-defined (SURVEY.md note 6).
* The compositing loops keep the reference's structure (one Gaussian at a
  time, in-place slice accumulation) so that the oracle doubles as the CPU
  baseline "port" with the same cost profile as the reference.
"""

from __future__ import annotations

import math
from typing import Dict, Optional, Sequence, Tuple

import numpy as np
import torch

TILE = 16  # tile edge used by the derived integer pins (tile|depth keys)


# --------------------------------------------------------------------------
# Camera (DR:24-95) and look-at pose
# (scripts/training/train_gaussian_decoder.py:684-757)
# --------------------------------------------------------------------------
class Camera:
    """Pinhole camera, OpenGL convention (camera looks down -Z).  DR:24-52."""

    def __init__(self, fx, fy, cx, cy, width, height, near=0.01, far=100.0):
        self.fx, self.fy, self.cx, self.cy = fx, fy, cx, cy
        self.width, self.height = width, height
        self.near, self.far = near, far
        self.view_matrix = torch.eye(4)

    def set_view(self, view_matrix: torch.Tensor) -> None:
        self.view_matrix = view_matrix


def camera_from_pose(elevation_rad, azimuth_rad, render_size,
                     focal_length_mult=0.8, distance=2.0) -> Camera:
    """Look-at camera on a sphere around the origin.

    Follows train_gaussian_decoder.py:684-757: spherical position, forward =
    -position normalised, right = forward x world-up, rows of the rotation are
    [right; up; -forward], translation = -R @ position, all in float64 and
    cast to float32 at the end.
    """
    ce, se = math.cos(elevation_rad), math.sin(elevation_rad)
    pos = np.array([distance * ce * math.sin(azimuth_rad),
                    distance * se,
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


# --------------------------------------------------------------------------
# Projection (DR:98-195) in a fixed elementwise order
# --------------------------------------------------------------------------
def _f32(x) -> torch.Tensor:
    return torch.tensor(float(x), dtype=torch.float32)


class _ExactSqrt(torch.autograd.Function):
    """Correctly rounded fp32 sqrt.

    torch.sqrt on CPU goes through a vectorised routine that is NOT correctly
    rounded for about 1% of inputs (measured: 33 of 4096 differ from IEEE
    sqrtf by one ulp); numpy's float32 sqrt is the hardware instruction and is.
    The GPU kernels use __fsqrt_rn, so the oracle must be IEEE-exact here for
    the bit-exact pins to be meaningful.
    """

    @staticmethod
    def forward(ctx, x):
        y = torch.from_numpy(np.sqrt(x.detach().to(torch.float32).contiguous().numpy()))
        ctx.save_for_backward(y)
        return y

    @staticmethod
    def backward(ctx, g):
        (y,) = ctx.saved_tensors
        return g / (2 * y)


def _sqrt(x: torch.Tensor) -> torch.Tensor:
    with np.errstate(invalid="ignore"):
        return _ExactSqrt.apply(x)


def project(positions, scales, rotations, camera) -> Dict[str, torch.Tensor]:
    """EWA projection of N Gaussians.  Differentiable.

    Follows compute_2d_covariance (DR:123-195) and
    quaternion_to_rotation_matrix (DR:98-120), including the Jacobian sign
    quirk at DR:185 (J[1,2] = +fy*y/z^2) and the hard-coded 0.01 clamp at
    DR:175.  Returns a dict with u, v, depth, cov (a, b, c, d) and the
    camera-space point.
    """
    V = camera.view_matrix.to(torch.float32)
    x, y, z = positions[:, 0], positions[:, 1], positions[:, 2]

    def row(i):
        return ((V[i, 0] * x + V[i, 1] * y) + V[i, 2] * z) + V[i, 3]

    pcx, pcy, pcz = row(0), row(1), row(2)          # DR:149-152
    depth = -pcz                                    # DR:155

    # F.normalize(q, dim=-1): q / max(||q||, 1e-12)   DR:109
    qw, qx, qy, qz = (rotations[:, i] for i in range(4))
    n2 = ((qw * qw + qx * qx) + qy * qy) + qz * qz
    pos_n2 = n2 > 0        # norm has a zero sub-gradient at q = 0, as vector_norm does
    nrm = torch.where(pos_n2, _sqrt(torch.where(pos_n2, n2, torch.ones_like(n2))),
                      torch.zeros_like(n2))
    den = torch.clamp(nrm, min=1e-12)
    qw, qx, qy, qz = qw / den, qx / den, qy / den, qz / den

    # DR:114-118, same expression trees as the reference
    R = [[1 - 2 * qy * qy - 2 * qz * qz, 2 * qx * qy - 2 * qw * qz, 2 * qx * qz + 2 * qw * qy],
         [2 * qx * qy + 2 * qw * qz, 1 - 2 * qx * qx - 2 * qz * qz, 2 * qy * qz - 2 * qw * qx],
         [2 * qx * qz - 2 * qw * qy, 2 * qy * qz + 2 * qw * qx, 1 - 2 * qx * qx - 2 * qy * qy]]

    # M = V_rot @ R @ diag(s)   DR:162-165
    M = [[None] * 3 for _ in range(3)]
    for i in range(3):
        for j in range(3):
            rc = (V[i, 0] * R[0][j] + V[i, 1] * R[1][j]) + V[i, 2] * R[2][j]
            M[i][j] = rc * scales[:, j]
    # Sigma3 = M @ M^T   DR:166
    S3 = [[None] * 3 for _ in range(3)]
    for i in range(3):
        for j in range(i, 3):
            S3[i][j] = (M[i][0] * M[j][0] + M[i][1] * M[j][1]) + M[i][2] * M[j][2]
            S3[j][i] = S3[i][j]

    fx, fy = _f32(camera.fx), _f32(camera.fy)
    cx, cy = _f32(camera.cx), _f32(camera.cy)
    zs = torch.clamp(pcz.abs(), min=0.01) * torch.sign(pcz + 1e-8)    # DR:175
    z2 = zs * zs
    j00 = fx / (-zs)                # DR:182
    j02 = (fx * pcx) / z2           # DR:183
    j11 = fy / zs                   # DR:184
    j12 = (fy * pcy) / z2           # DR:185 (sign quirk kept)

    # T = J @ Sigma3 ; cov = T @ J^T   DR:188
    t00 = j00 * S3[0][0] + j02 * S3[2][0]
    t01 = j00 * S3[0][1] + j02 * S3[2][1]
    t02 = j00 * S3[0][2] + j02 * S3[2][2]
    t10 = j11 * S3[1][0] + j12 * S3[2][0]
    t11 = j11 * S3[1][1] + j12 * S3[2][1]
    t12 = j11 * S3[1][2] + j12 * S3[2][2]
    a = t00 * j00 + t02 * j02
    b = t01 * j11 + t02 * j12
    c = t10 * j00 + t12 * j02
    d = t11 * j11 + t12 * j12

    u = (fx * pcx) / (-zs) + cx     # DR:191
    v = (fy * (-pcy)) / (-zs) + cy  # DR:192
    return dict(u=u, v=v, depth=depth, a=a, b=b, c=c, d=d, pcx=pcx, pcy=pcy, pcz=pcz)


def compute_radius(a, b, c, d, max_radius) -> torch.Tensor:
    """3-sigma radius from the UN-regularised covariance.  DR:452-487."""
    trace = a + d
    det = torch.clamp(a * d - b * c, min=1e-6)
    disc = torch.clamp(trace * trace - 4 * det, min=0)
    lam = (trace + _sqrt(disc)) / 2
    r = 3.0 * _sqrt(torch.clamp(lam, min=1e-6))
    return torch.clamp(r, max=float(max_radius))


def visibility(u, v, depth, radius, camera, width, height) -> torch.Tensor:
    """Frustum + bounding-box visibility, strict inequalities.  DR:541-543."""
    vis = (depth > camera.near) & (depth < camera.far)
    vis &= (u + radius > 0) & (u - radius < width)
    vis &= (v + radius > 0) & (v - radius < height)
    return vis


def rects(u, v, radius, width, height) -> np.ndarray:
    """Integer pixel rectangles [x0, x1, y0, y1) per Gaussian.  DR:594-597.

    The reference subtracts two Python floats obtained with ``.item()``, i.e.
    the arithmetic is float64 on float32 values, then truncates with ``int``.
    """
    u64 = u.detach().double().numpy()
    v64 = v.detach().double().numpy()
    r64 = radius.detach().double().numpy()
    with np.errstate(invalid="ignore"):
        x0 = np.maximum(0, np.trunc(u64 - r64))
        x1 = np.minimum(width, np.trunc(u64 + r64) + 1)
        y0 = np.maximum(0, np.trunc(v64 - r64))
        y1 = np.minimum(height, np.trunc(v64 + r64) + 1)
    out = np.stack([x0, x1, y0, y1], axis=1)
    out = np.nan_to_num(out, nan=0.0, posinf=0.0, neginf=0.0)
    return np.clip(out, -2**31, 2**31 - 1).astype(np.int64)


def depth_bits(depth: torch.Tensor) -> np.ndarray:
    """IEEE-754 bit pattern of the fp32 depth (monotone for depth > 0)."""
    return depth.detach().to(torch.float32).contiguous().numpy().view(np.uint32).copy()


def tile_keys(vis: np.ndarray, rect: np.ndarray, dbits: np.ndarray, width: int, height: int
              ) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """The derived integer pins: sorted (tile|depth) keys, Gaussian ids, ranges.

    No reference counterpart; this is the
    definition the CUDA binning must reproduce bit for bit.  For every
    visible Gaussian with a non-empty rectangle (DR:599-600), one 64-bit key
    ``(tile_id << 32) | depth_bits`` per overlapped 16x16 tile, emitted in
    ascending Gaussian index, then stably sorted.  A stable sort on
    (tile, depth) is the reference's global stable depth order (DR:527)
    restricted to each tile's list.
    Returns (keys uint64 [M], gaussian ids int32 [M], ranges int32 [tiles, 2]).
    """
    tiles_x = (width + TILE - 1) // TILE
    tiles_y = (height + TILE - 1) // TILE
    keys, gids = [], []
    idx = np.nonzero(vis)[0]
    for i in idx:
        x0, x1, y0, y1 = rect[i]
        if x0 >= x1 or y0 >= y1:
            continue
        tx0, tx1 = x0 // TILE, (x1 - 1) // TILE
        ty0, ty1 = y0 // TILE, (y1 - 1) // TILE
        ty, tx = np.meshgrid(np.arange(ty0, ty1 + 1), np.arange(tx0, tx1 + 1), indexing="ij")
        t = (ty * tiles_x + tx).reshape(-1).astype(np.uint64)
        keys.append((t << np.uint64(32)) | np.uint64(dbits[i]))
        gids.append(np.full(t.shape[0], i, dtype=np.int32))
    if keys:
        keys = np.concatenate(keys)
        gids = np.concatenate(gids)
    else:
        keys = np.zeros(0, np.uint64)
        gids = np.zeros(0, np.int32)
    order = np.argsort(keys, kind="stable")
    keys, gids = keys[order], gids[order]
    tile_of = (keys >> np.uint64(32)).astype(np.int64)
    n_tiles = tiles_x * tiles_y
    starts = np.searchsorted(tile_of, np.arange(n_tiles), side="left")
    ends = np.searchsorted(tile_of, np.arange(n_tiles), side="right")
    return keys, gids, np.stack([starts, ends], axis=1).astype(np.int32)


def pins(positions, scales, rotations, camera, width, height, max_radius=64) -> Dict[str, np.ndarray]:
    """All integer pins for one view (no gradients)."""
    with torch.no_grad():
        p = project(positions, scales, rotations, camera)
        r = compute_radius(p["a"], p["b"], p["c"], p["d"], max_radius)
        vis = visibility(p["u"], p["v"], p["depth"], r, camera, width, height).numpy()
        rc = rects(p["u"], p["v"], r, width, height)
        db = depth_bits(p["depth"])
        order = torch.argsort(p["depth"], stable=True).numpy().astype(np.int32)
    keys, gids, ranges = tile_keys(vis, rc, db, width, height)
    return dict(visible=vis, rect=rc.astype(np.int32), depth_bits=db, order=order,
                keys=keys, gids=gids, ranges=ranges,
                u=p["u"].numpy(), v=p["v"].numpy(), depth=p["depth"].numpy(),
                radius=r.numpy(),
                cov=torch.stack([p["a"], p["b"], p["c"], p["d"]], 1).numpy())


# --------------------------------------------------------------------------
# TileBasedRenderer.forward (DR:489-686)
# --------------------------------------------------------------------------
def _prepare(positions, scales, rotations, camera, width, height, max_radius, sort: bool):
    p = project(positions, scales, rotations, camera)
    radius = compute_radius(p["a"], p["b"], p["c"], p["d"], max_radius)     # DR:524
    cov = torch.stack([torch.stack([p["a"], p["b"]], -1),
                       torch.stack([p["c"], p["d"]], -1)], -2)               # (N,2,2)
    vis = visibility(p["u"], p["v"], p["depth"], radius, camera, width, height)
    if sort:
        order = torch.argsort(p["depth"], stable=True)                       # DR:527 (stable)
        order = order[vis[order]]                                            # DR:541-562
    else:
        order = torch.nonzero(vis).squeeze(1)                                # DR:810-816
    return p, radius, cov, order


def _zero_anchor(colors, opacities, positions):
    return (colors.sum() + opacities.sum() + positions.sum()) * 0.0          # DR:548


def render_tile_based(positions, scales, rotations, colors, opacities, camera,
                      width, height, background=(0.0, 0.0, 0.0), max_radius=64,
                      use_phase_blending=False, phase_amplitude=0.25, phases=None
                      ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Front-to-back alpha compositing, one Gaussian at a time.  DR:489-686.

    Returns (image (3,H,W), depth (H,W), alpha (H,W)); alpha is the
    reference's ``accumulated_alpha`` (not returned by the reference).
    With phase blending the three slice reads at DR:632, 663, 665 are cloned so
    that autograd can run (the reference raises there; SURVEY.md note 3) -
    the forward values are unchanged.
    """
    H, W = height, width
    bg = torch.tensor(background, dtype=torch.float32)
    p, radius, cov, order = _prepare(positions, scales, rotations, camera, W, H, max_radius, True)

    if order.numel() == 0:                                                   # DR:545-552
        anchor = _zero_anchor(colors, opacities, positions)
        img = bg.view(3, 1, 1).expand(3, H, W) + anchor
        return img, torch.zeros(H, W) + anchor, torch.zeros(H, W)

    u, v = p["u"][order], p["v"][order]
    depth = p["depth"][order]
    col, opa = colors[order], opacities[order]
    rad = radius[order]
    phs = phases[order] if phases is not None else None
    rc = rects(u, v, rad, W, H)
    inv = torch.linalg.pinv(cov[order] + 1e-4 * torch.eye(2).unsqueeze(0))   # DR:578-579

    acc_c = torch.zeros(H, W, 3)
    acc_a = torch.zeros(H, W)
    acc_d = torch.zeros(H, W)
    blend = use_phase_blending and phs is not None
    acc_p = torch.zeros(H, W) if blend else None

    for i in range(order.numel()):                                           # DR:582
        x0, x1, y0, y1 = (int(t) for t in rc[i])
        if x0 >= x1 or y0 >= y1:                                             # DR:599
            continue
        ly, lx = torch.meshgrid(torch.arange(y0, y1, dtype=torch.float32),
                                torch.arange(x0, x1, dtype=torch.float32), indexing="ij")
        dx = lx - u[i]
        dy = ly - v[i]
        m = inv[i, 0, 0] * dx * dx + (inv[i, 0, 1] + inv[i, 1, 0]) * dx * dy + inv[i, 1, 1] * dy * dy
        alpha = torch.exp(-0.5 * m) * opa[i]                                 # DR:618-624
        if blend:                                                            # DR:629-645
            prev = acc_p[y0:y1, x0:x1].clone()
            diff = torch.abs(phs[i] - prev)
            diff = torch.min(diff, 1.0 - diff)
            alpha = alpha * ((1.0 - phase_amplitude) + phase_amplitude * torch.cos(diff * 2 * 3.14159))
        alpha = torch.clamp(alpha, 0, 0.99)                                  # DR:647
        contrib = alpha * (1.0 - acc_a[y0:y1, x0:x1])                        # DR:650-653
        acc_c[y0:y1, x0:x1] += contrib.unsqueeze(-1) * col[i].view(1, 1, 3)  # DR:656
        acc_d[y0:y1, x0:x1] += contrib * depth[i]                            # DR:657
        acc_a[y0:y1, x0:x1] += contrib                                       # DR:658
        if blend:                                                            # DR:661-667
            pc = contrib / acc_a[y0:y1, x0:x1].clone().clamp(min=1e-6)
            acc_p[y0:y1, x0:x1] = acc_p[y0:y1, x0:x1].clone() * (1 - pc) + phs[i] * pc

    acc_c = acc_c + (1.0 - acc_a).unsqueeze(-1) * bg.view(1, 1, 3)           # DR:670-671
    image = torch.clamp(acc_c.permute(2, 0, 1), 0, 1)                        # DR:674-675
    if not image.requires_grad and any(t.requires_grad for t in (colors, opacities, positions)):
        anchor = _zero_anchor(colors, opacities, positions)                  # DR:679-682
        image = image + anchor
        acc_d = acc_d + anchor
    return image, acc_d, acc_a


# --------------------------------------------------------------------------
# WaveFieldRenderer.forward (DR:747-926)
# --------------------------------------------------------------------------
def _splat_terms(i, u, v, inv, rc, opa):
    x0, x1, y0, y1 = (int(t) for t in rc[i])
    if x0 >= x1 or y0 >= y1:
        return None
    ly, lx = torch.meshgrid(torch.arange(y0, y1, dtype=torch.float32),
                            torch.arange(x0, x1, dtype=torch.float32), indexing="ij")
    dx = lx - u[i]
    dy = ly - v[i]
    m = inv[i, 0, 0] * dx * dx + (inv[i, 0, 1] + inv[i, 1, 0]) * dx * dy + inv[i, 1, 1] * dy * dy
    return (x0, x1, y0, y1), torch.exp(-0.5 * m) * opa[i]


def render_wave(positions, scales, rotations, colors, opacities, camera, width, height,
                phases, background=(0.0, 0.0, 0.0), max_radius=64
                ) -> Tuple[torch.Tensor, torch.Tensor]:
    """Unsorted complex-amplitude splat + intensity.  DR:747-926.

    ``phases`` is (N,) or (N,3) radians (DR:874-881).  Returns (image, depth).
    """
    if phases is None:
        raise ValueError("WaveFieldRenderer requires phases tensor.")        # DR:779-780
    H, W = height, width
    bg = torch.tensor(background, dtype=torch.float32)
    p, radius, cov, order = _prepare(positions, scales, rotations, camera, W, H, max_radius, False)
    if order.numel() == 0:                                                   # DR:801-808
        anchor = _zero_anchor(colors, opacities, positions)
        return bg.view(3, 1, 1).expand(3, H, W) + anchor, torch.zeros(H, W) + anchor
    u, v, depth = p["u"][order], p["v"][order], p["depth"][order]
    col, opa, phs = colors[order], opacities[order], phases[order]
    rc = rects(u, v, radius[order], W, H)
    inv = torch.linalg.pinv(cov[order] + 1e-4 * torch.eye(2).unsqueeze(0))   # DR:828-829

    re = torch.zeros(H, W, 3)
    im = torch.zeros(H, W, 3)
    acc_d = torch.zeros(H, W)
    wsum = torch.zeros(H, W)
    for i in range(order.numel()):                                           # DR:832
        t = _splat_terms(i, u, v, inv, rc, opa)
        if t is None:
            continue
        (x0, x1, y0, y1), amp = t
        cosp, sinp = torch.cos(phs[i]), torch.sin(phs[i])                    # DR:874-881
        re[y0:y1, x0:x1] += amp.unsqueeze(-1) * col[i].view(1, 1, 3) * cosp  # DR:886
        im[y0:y1, x0:x1] += amp.unsqueeze(-1) * col[i].view(1, 1, 3) * sinp  # DR:887
        acc_d[y0:y1, x0:x1] += amp * depth[i]                                # DR:890
        wsum[y0:y1, x0:x1] += amp                                            # DR:891

    inten = re ** 2 + im ** 2                                                # DR:894
    rend = torch.sqrt(inten + 1e-8)                                          # DR:898
    rend = rend / rend.max().clamp(min=1.0)                                  # DR:902-903
    rend = torch.clamp(rend, 0, 1)                                           # DR:905
    tot = torch.sqrt((re ** 2 + im ** 2).sum(dim=-1, keepdim=True) + 1e-8).clamp(0, 1)   # DR:908-909
    rend = rend + bg.view(1, 1, 3) * (1 - tot)                               # DR:910
    image = torch.clamp(rend.permute(2, 0, 1), 0, 1)                         # DR:913-914
    return image, acc_d / (wsum + 1e-8)                                      # DR:924


# --------------------------------------------------------------------------
# AngularSpectrumPropagator (DR:929-1065) and ASMWaveFieldRenderer (DR:1150-1344)
# --------------------------------------------------------------------------
def asm_transfer(height, width, pixel_pitch, z_distance, wavelength) -> torch.Tensor:
    """H = exp(i 2 pi z sqrt(max(1/lambda^2 - fx^2 - fy^2, 0))).  DR:958-1001.

    The reference builds FX, FY with ``meshgrid(fx, fy, indexing='xy')``
    (DR:959-961), i.e. arrays of shape (len(fy), len(fx)) = (H, W).
    """
    fx = torch.fft.fftfreq(width, d=pixel_pitch)
    fy = torch.fft.fftfreq(height, d=pixel_pitch)
    FX, FY = torch.meshgrid(fx, fy, indexing="xy")
    kz_sq = torch.clamp((1.0 / wavelength) ** 2 - FX ** 2 - FY ** 2, min=0)
    return torch.exp(1j * 2 * torch.pi * z_distance * torch.sqrt(kz_sq))


def asm_propagate(field, height, width, pixel_pitch, z_distance, wavelength) -> torch.Tensor:
    """ifft2(fft2(field) * H) for one (H, W) complex field.  DR:1041-1047."""
    return torch.fft.ifft2(torch.fft.fft2(field) * asm_transfer(height, width, pixel_pitch,
                                                                z_distance, wavelength))


def render_asm(positions, scales, rotations, colors, opacities, camera, width, height,
               phases, wavelengths_rgb, background=(0.0, 0.0, 0.0), max_radius=64,
               num_depth_planes=16, depth_range=(0.1, 2.0), focal_depth=0.5,
               pixel_pitch=1.0 / 256.0) -> Tuple[torch.Tensor, torch.Tensor]:
    """Per-plane complex splat, ASM propagation to the focal plane.  DR:1150-1344.

    ``wavelengths_rgb`` must be a (3,) tensor: the reference's scalar default
    crashes at DR:1036 (SURVEY.md note 4).  Wavelengths are constants here.
    Depth output is zeros as in the reference (DR:1339-1342).
    """
    if phases is None:
        raise ValueError("ASMWaveFieldRenderer requires phases tensor.")     # DR:1187-1188
    H, W = height, width
    bg = torch.tensor(background, dtype=torch.float32)
    p, radius, cov, order = _prepare(positions, scales, rotations, camera, W, H, max_radius, False)
    if order.numel() == 0:                                                   # DR:1207-1212
        anchor = _zero_anchor(colors, opacities, positions)
        return bg.view(3, 1, 1).expand(3, H, W) + anchor, torch.zeros(H, W) + anchor
    u, v, depth = p["u"][order], p["v"][order], p["depth"][order]
    col, opa, phs = colors[order], opacities[order], phases[order]
    rc = rects(u, v, radius[order], W, H)
    planes = torch.linspace(depth_range[0], depth_range[1], num_depth_planes)     # DR:1106
    plane_idx = (depth.detach().unsqueeze(1) - planes.unsqueeze(0)).abs().argmin(dim=1)  # DR:1147-1148
    inv = torch.linalg.pinv(cov[order] + 1e-4 * torch.eye(2).unsqueeze(0))   # DR:1229-1230

    fields = torch.zeros(num_depth_planes, H, W, 3, 2)                       # DR:1233-1235
    for i in range(order.numel()):                                           # DR:1238
        t = _splat_terms(i, u, v, inv, rc, opa)
        if t is None:
            continue
        (x0, x1, y0, y1), amp = t
        k = int(plane_idx[i])
        cosp, sinp = torch.cos(phs[i]), torch.sin(phs[i])
        fields[k, y0:y1, x0:x1, :, 0] += amp.unsqueeze(-1) * col[i].view(1, 1, 3) * cosp   # DR:1278
        fields[k, y0:y1, x0:x1, :, 1] += amp.unsqueeze(-1) * col[i].view(1, 1, 3) * sinp   # DR:1281

    total = torch.zeros(H, W, 3, dtype=torch.cfloat)
    focal = torch.tensor(focal_depth)
    for k in range(num_depth_planes):                                        # DR:1291
        fc = torch.complex(fields[k, :, :, :, 0], fields[k, :, :, :, 1])
        if fc.abs().max() < 1e-8:                                            # DR:1302
            continue
        z_prop = focal - planes[k]                                           # DR:1293
        chans = [asm_propagate(fc[..., c], H, W, pixel_pitch, z_prop, wavelengths_rgb[c])
                 for c in range(3)]                                          # DR:1306-1313
        total = total + torch.stack(chans, dim=-1)

    inten = total.real ** 2 + total.imag ** 2                                # DR:1316
    rend = torch.sqrt(inten + 1e-8)                                          # DR:1319
    rend = rend / rend.max().clamp(min=1.0)                                  # DR:1322-1323
    rend = torch.clamp(rend, 0, 1)
    tot = total.abs().sum(dim=-1, keepdim=True).clamp(0, 1)                  # DR:1327
    rend = rend + bg.view(1, 1, 3) * (1 - tot)                               # DR:1328
    image = torch.clamp(rend.permute(2, 0, 1), 0, 1)                         # DR:1331-1332
    return image, torch.zeros(H, W)                                          # DR:1339-1342


# --------------------------------------------------------------------------
# Synthetic workloads
# --------------------------------------------------------------------------
# DifferentiableGaussianRenderer.forward (DR:245-409)
# --------------------------------------------------------------------------
def render_dense(positions, scales, rotations, colors, opacities, camera, width, height,
                 background=(0.0, 0.0, 0.0)) -> Tuple[torch.Tensor, torch.Tensor]:
    """Every visible Gaussian evaluated at every pixel, composited front to back.  DR:274-409.

    Visibility is the frustum plus a 100-pixel margin on the projected centre (DR:315-318); the
    quadratic form uses pinv(cov + 1e-4 I) with both off-diagonal terms (gaussian_2d, DR:198-242).
    Returns (image (3,H,W), depth (H,W)).
    """
    H, W = height, width
    bg = torch.tensor(background, dtype=torch.float32)
    p = project(positions, scales, rotations, camera)
    order = torch.argsort(p["depth"], stable=True)                           # DR:306 (stable)
    vis = (p["depth"] > camera.near) & (p["depth"] < camera.far)             # DR:315
    vis &= (p["u"] > -100) & (p["u"] < W + 100)                              # DR:316
    vis &= (p["v"] > -100) & (p["v"] < H + 100)                              # DR:317
    order = order[vis[order]]
    if order.numel() == 0:                                                   # DR:319-325
        anchor = _zero_anchor(colors, opacities, positions)
        return bg.view(3, 1, 1).expand(3, H, W) + anchor, torch.zeros(H, W) + anchor
    u, v, depth = p["u"][order], p["v"][order], p["depth"][order]
    col, opa = colors[order], opacities[order]
    cov = torch.stack([torch.stack([p["a"], p["b"]], -1), torch.stack([p["c"], p["d"]], -1)], -2)[order]
    inv = torch.linalg.pinv(cov + 1e-4 * torch.eye(2).unsqueeze(0))          # DR:219-220
    ly, lx = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(W, dtype=torch.float32),
                            indexing="ij")
    acc_c = torch.zeros(H, W, 3)
    acc_a = torch.zeros(H, W)
    acc_d = torch.zeros(H, W)
    for i in range(order.numel()):                                           # DR:348-384
        dx = lx - u[i]
        dy = ly - v[i]
        m = inv[i, 0, 0] * dx * dx + (inv[i, 0, 1] + inv[i, 1, 0]) * dx * dy + inv[i, 1, 1] * dy * dy
        alpha = torch.clamp(torch.exp(-0.5 * m) * opa[i], 0, 0.99)           # DR:240, 362-365
        contrib = alpha * (1.0 - acc_a)                                      # DR:370-373
        acc_c = acc_c + contrib.unsqueeze(-1) * col[i].view(1, 1, 3)         # DR:376
        acc_d = acc_d + contrib * depth[i]                                   # DR:379
        acc_a = acc_a + contrib                                              # DR:382
    acc_c = acc_c + (1.0 - acc_a).unsqueeze(-1) * bg.view(1, 1, 3)           # DR:386-387
    return torch.clamp(acc_c.permute(2, 0, 1), 0, 1), acc_d                  # DR:390-393


# --------------------------------------------------------------------------
# FourierGaussianRenderer.forward (DR:1582-1766)
# --------------------------------------------------------------------------
def render_fourier(positions, scales, rotations, colors, opacities, camera, width, height,
                   background=(0.0, 0.0, 0.0)) -> torch.Tensor:
    """Isotropic additive splat over the whole image + global-max normalisation.  DR:1693-1753.

    sigma^2 = (a + d)/2 + 1e-8 from the un-regularised projected covariance (DR:1673-1677); visibility is
    the frustum plus one image size of margin on the centre (DR:1647-1649).  The wavelengths / phases of
    the reference do not enter the image (DR:1686-1691 is dead code).  Returns the image (3,H,W).
    """
    H, W = height, width
    bg = torch.tensor(background, dtype=torch.float32)
    p = project(positions, scales, rotations, camera)
    vis = (p["depth"] > camera.near) & (p["depth"] < camera.far)
    vis &= (p["u"] > -W) & (p["u"] < 2 * W) & (p["v"] > -H) & (p["v"] < 2 * H)
    idx = torch.nonzero(vis).squeeze(1)
    if idx.numel() == 0:                                                     # DR:1651-1657
        return bg.view(3, 1, 1).expand(3, H, W) + _zero_anchor(colors, opacities, positions)
    u, v = p["u"][idx], p["v"][idx]
    sigma = torch.sqrt((p["a"][idx] + p["d"][idx]) / 2 + 1e-8)               # DR:1677
    col, opa = colors[idx], opacities[idx]
    ly, lx = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(W, dtype=torch.float32),
                            indexing="ij")
    image = torch.zeros(3, H, W)
    for s0 in range(0, idx.numel(), 16):                                     # DR:1701-1738
        sl = slice(s0, min(s0 + 16, idx.numel()))
        dx = lx.unsqueeze(0) - u[sl].view(-1, 1, 1)
        dy = ly.unsqueeze(0) - v[sl].view(-1, 1, 1)
        g = torch.exp(-(dx ** 2 + dy ** 2) / (2 * sigma[sl].view(-1, 1, 1) ** 2 + 1e-8)) * opa[sl].view(-1, 1, 1)
        image = image + torch.stack([(col[sl, c].view(-1, 1, 1) * g).sum(0) for c in range(3)])
    mx = image.max()                                                         # DR:1741
    if mx > 1e-8:
        image = image / mx
    bgw = torch.clamp(1.0 - image.sum(dim=0, keepdim=True), 0, 1)            # DR:1746-1747
    return torch.clamp(image + bg.view(3, 1, 1) * bgw, 0, 1)                 # DR:1748-1751


# --------------------------------------------------------------------------
# SimplifiedRenderer.forward (DR:1347-1458) and Camera.project (DR:54-85)
# --------------------------------------------------------------------------
def camera_project(positions, camera) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Pixel coordinates and depth of N points, fixed elementwise fp32 order.  DR:54-85."""
    V = camera.view_matrix.to(torch.float32)
    x, y, z = positions[:, 0], positions[:, 1], positions[:, 2]

    def row(i):
        return ((V[i, 0] * x + V[i, 1] * y) + V[i, 2] * z) + V[i, 3]

    pcx, pcy, pcz = row(0), row(1), row(2)
    zs = torch.clamp(pcz.abs(), min=camera.near) * torch.sign(pcz + 1e-8)     # DR:78
    u = (_f32(camera.fx) * pcx) / (-zs) + _f32(camera.cx)                      # DR:81
    v = (_f32(camera.fy) * (-pcy)) / (-zs) + _f32(camera.cy)                   # DR:82
    return u, v, -zs                                                           # DR:85


def simplified_radius(scales_i: torch.Tensor, fx: float, d: torch.Tensor) -> int:
    """DR:1402-1403, the reference's own expression (Python float product, fp32 tensor division)."""
    return min(int(max(scales_i.mean().item() * fx / d, 1)), 20)


def render_simplified(positions, scales, colors, opacities, camera, width, height,
                      background=(0.0, 0.0, 0.0)) -> Tuple[torch.Tensor, torch.Tensor]:
    """Point splats with an integer radius, blended back to front with "over".  DR:1369-1458.
    Returns (image (3,H,W), depth (H,W) with inf replaced by 0 as DR:1453 does)."""
    H, W = height, width
    bg = torch.tensor(background, dtype=torch.float32)
    u, v, depth = camera_project(positions, camera)
    order = torch.argsort(depth, descending=True, stable=True)                # DR:1381 (stable)
    image = bg.view(3, 1, 1).expand(3, H, W).clone()
    depth_map = torch.full((H, W), float("inf"))
    for i in order.tolist():                                                  # DR:1394
        d = depth[i]
        if d <= 0:                                                            # DR:1398
            continue
        radius = simplified_radius(scales[i], camera.fx, d.detach())
        x_int, y_int = int(u[i].item()), int(v[i].item())                     # DR:1406
        x0, x1 = max(0, x_int - radius), min(W, x_int + radius + 1)
        y0, y1 = max(0, y_int - radius), min(H, y_int + radius + 1)
        if x0 >= x1 or y0 >= y1:                                              # DR:1412
            continue
        yy, xx = torch.meshgrid(torch.arange(y0, y1, dtype=torch.float32),
                                torch.arange(x0, x1, dtype=torch.float32), indexing="ij")
        dist_sq = (xx - u[i]) ** 2 + (yy - v[i]) ** 2                         # DR:1423
        weight = torch.exp(-dist_sq / (2 * max(radius / 2, 1) ** 2))          # DR:1424
        alpha = torch.clamp(weight * opacities[i], 0, 1)                      # DR:1427-1428
        new = alpha.unsqueeze(0) * colors[i].view(3, 1, 1) + (1 - alpha).unsqueeze(0) * image[:, y0:y1, x0:x1]
        image = torch.cat([image[:, :y0], torch.cat([image[:, y0:y1, :x0], new, image[:, y0:y1, x1:]], 2),
                           image[:, y1:]], 1)                                 # DR:1431-1436 without in-place writes
        dm = depth_map[y0:y1, x0:x1]
        dnew = torch.where(alpha > 0.1, torch.minimum(dm, d.expand(y1 - y0, x1 - x0)), dm)   # DR:1438-1442
        depth_map = torch.cat([depth_map[:y0], torch.cat([depth_map[y0:y1, :x0], dnew, depth_map[y0:y1, x1:]], 1),
                               depth_map[y1:]], 0)
    depth_map = torch.where(depth_map == float("inf"), torch.zeros_like(depth_map), depth_map)   # DR:1453
    return image, depth_map


# --------------------------------------------------------------------------
# On-disk formats: GaussianCloud::save_ply / load_ply (src/core/renderer/renderer.cpp:649-793)
# PARITY UNPINNED for the .ply transforms: the C++ reference needs GLM / Kompute (absent here) and has no
# test vectors; this is a restatement of the published 3DGS parameterisation as the reference writes it.
# The .bin format IS pinned: tests/golden/cloud_97.bin is written by the reference's own
# save_gaussians_to_binary (DR:1485-1497) and read back by its load_gaussians_from_binary (DR:1461-1482).
# --------------------------------------------------------------------------
SH_C0 = np.float32(0.28209479177387814)


def ply_decode_rows(rows: np.ndarray) -> Dict[str, np.ndarray]:
    """14-float .ply rows -> activated parameters.  renderer.cpp:754-785."""
    r = np.asarray(rows, np.float32)
    return dict(positions=r[:, 0:3].copy(),
                scales=np.exp(r[:, 3:6]).astype(np.float32),                                    # :766-768
                rotations=r[:, 6:10].copy(),
                colors=np.clip(r[:, 10:13] * SH_C0 + np.float32(0.5), 0.0, 1.0).astype(np.float32),   # :777-779
                opacities=(1.0 / (1.0 + np.exp(-r[:, 13]))).astype(np.float32))                 # :782


def ply_encode_rows(g: Dict[str, np.ndarray]) -> np.ndarray:
    """Activated parameters -> 14-float .ply rows.  renderer.cpp:679-717."""
    n = g["positions"].shape[0]
    r = np.zeros((n, 14), np.float32)
    r[:, 0:3] = g["positions"]
    r[:, 3:6] = np.log(np.maximum(g["scales"], np.float32(1e-7)))                               # :687-689
    r[:, 6:10] = g["rotations"]
    r[:, 10:13] = (g["colors"] - np.float32(0.5)) / SH_C0                                       # :708-710
    op = g["opacities"].astype(np.float32)
    r[:, 13] = np.log(op / np.maximum(np.float32(1.0) - op, np.float32(1e-7)))                  # :716
    return r


# --------------------------------------------------------------------------
def synthetic_cloud(n: int, seed: int = 0, s_lo: float = 0.005, s_hi: float = 0.03,
                    phase_hi: float = 1.0) -> Dict[str, torch.Tensor]:
    """The seeded synthetic Gaussian cloud every config uses (fp32, CPU)."""
    g = torch.Generator().manual_seed(seed)
    pos = torch.randn(n, 3, generator=g) * 0.5
    pos[:, 2] -= 2.0
    return dict(
        positions=pos,
        scales=torch.rand(n, 3, generator=g) * (s_hi - s_lo) + s_lo,
        rotations=torch.randn(n, 4, generator=g),
        colors=torch.rand(n, 3, generator=g),
        opacities=torch.rand(n, generator=g) * 0.8 + 0.1,
        phases=torch.rand(n, generator=g) * phase_hi,
    )


def default_camera(res_w: int, res_h: Optional[int] = None) -> Camera:
    """fx = fy = 0.8*res, principal point at the centre, identity view
    (train_gaussian_decoder.py:1910-1917)."""
    res_h = res_w if res_h is None else res_h
    return Camera(0.8 * res_w, 0.8 * res_w, res_w / 2, res_h / 2, res_w, res_h)
