"""Shared helpers of the parity tests (oracle side runs on the CPU)."""
import numpy as np
import torch

from oracle import fresnel_oracle as fo

GRAD_NAMES = ("positions", "scales", "rotations", "colors", "opacities")


def oracle_camera(cam_vec, W, H):
    c = np.asarray(cam_vec, np.float64)
    cam = fo.Camera(float(c[12]), float(c[13]), float(c[14]), float(c[15]), W, H, float(c[18]), float(c[19]))
    view = torch.eye(4)
    view[:3, :] = torch.from_numpy(c[:12].reshape(3, 4)).float()
    cam.set_view(view)
    return cam


def golden_inputs(z, device="cpu", requires_grad=False, with_phases=False):
    names = GRAD_NAMES + (("phases",) if with_phases else ())
    out = {}
    for k in names:
        t = torch.from_numpy(z["in_" + k]).to(device)
        out[k] = t.requires_grad_(True) if requires_grad else t
    return out


def rel(a, b):
    """max |a-b| / max(max|b|, 1e-3): the per-tensor relative error of SURVEY.md section 8c."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-3))


def synthetic(n, seed=0, **kw):
    return fo.synthetic_cloud(n, seed=seed, **kw)
