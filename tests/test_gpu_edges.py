"""GPU tier: edge cases of the drop-in boundary - empty and single-Gaussian inputs, image sides that are not
multiples of the 16-pixel tile, everything culled, non-contiguous / non-fp32 inputs, inputs that do not require
gradients, larger-than-32 view batches - for every renderer module."""
import math

import numpy as np
import pytest
import torch

from oracle import fresnel_oracle as fo
import fresnel_b200
from helpers import GRAD_NAMES, rel

pytestmark = pytest.mark.gpu
IMG_TOL, GRAD_TOL = 1e-5, 1e-4


def dev():
    assert torch.cuda.is_available(), "GPU tier needs a CUDA device"
    return torch.device("cuda:0")


def renderers(W, H, bg):
    return {
        "tile": (fresnel_b200.TileBasedRenderer(W, H, background=bg), False),
        "phase": (fresnel_b200.TileBasedRenderer(W, H, background=bg, use_phase_blending=True), True),
        "wave": (fresnel_b200.WaveFieldRenderer(W, H, background=bg), True),
        "asm": (fresnel_b200.ASMWaveFieldRenderer(W, H, background=bg).to(dev()), True),
        "dense": (fresnel_b200.DifferentiableGaussianRenderer(W, H, background=bg), False),
        "fourier": (fresnel_b200.FourierGaussianRenderer(W, H, background=bg).to(dev()), False),
        "simple": (fresnel_b200.SimplifiedRenderer(W, H, background=bg), False),
    }


def call(ren, needs_phase, L, cam):
    kw = {}
    if needs_phase:
        kw["phases"] = L["phases"]
    if isinstance(ren, fresnel_b200.ASMWaveFieldRenderer):
        kw["wavelengths_rgb"] = torch.tensor([0.0635, 0.05, 0.041])
    return ren(L["positions"], L["scales"], L["rotations"], L["colors"], L["opacities"], cam, **kw)


@pytest.mark.parametrize("kind", ["tile", "phase", "wave", "asm", "dense", "fourier", "simple"])
def test_empty_cloud_renders_the_background(kind):
    """N = 0 (the reference's 'no visible Gaussians' branch, DR:545-552): background image, backward runs."""
    W, H, bg = 40, 24, (0.25, 0.5, 0.75)
    ren, needs_phase = renderers(W, H, bg)[kind]
    cam = fresnel_b200.Camera(0.8 * W, 0.8 * W, W / 2, H / 2, W, H)
    L = {k: torch.zeros((0, c) if c else (0,), device=dev(), requires_grad=True)
         for k, c in (("positions", 3), ("scales", 3), ("rotations", 4), ("colors", 3), ("opacities", 0), ("phases", 0))}
    img = call(ren, needs_phase, L, cam)
    assert img.shape == (3, H, W)
    want = torch.tensor(bg).view(3, 1, 1).expand(3, H, W)
    assert torch.allclose(img.detach().cpu(), want, atol=1e-6)
    img.sum().backward()
    assert L["positions"].grad is None or L["positions"].grad.shape == (0, 3)


@pytest.mark.parametrize("kind", ["tile", "dense"])
@pytest.mark.parametrize("size", [(1, 1), (17, 5), (33, 47), (130, 16)])
def test_single_gaussian_and_ragged_image_sizes(kind, size):
    """One Gaussian, image sides that are not multiples of the tile: against the oracle, image and gradients."""
    W, H = size
    bg = (0.1, 0.3, 0.2)
    inp = {"positions": torch.tensor([[0.05, -0.02, -1.5]]), "scales": torch.tensor([[0.08, 0.03, 0.05]]),
           "rotations": torch.tensor([[0.9, 0.1, -0.3, 0.2]]), "colors": torch.tensor([[0.9, 0.2, 0.4]]),
           "opacities": torch.tensor([0.8])}
    cam_o = fo.default_camera(W, H)
    cam = fresnel_b200.Camera(cam_o.fx, cam_o.fy, cam_o.cx, cam_o.cy, W, H)
    g = torch.Generator().manual_seed(W * 100 + H)
    gi, gd = torch.rand(3, H, W, generator=g) * 2 - 1, torch.rand(H, W, generator=g) * 2 - 1
    Lo = {k: v.clone().requires_grad_(True) for k, v in inp.items()}
    if kind == "tile":
        io, do, _ = fo.render_tile_based(Lo["positions"], Lo["scales"], Lo["rotations"], Lo["colors"], Lo["opacities"],
                                         cam_o, W, H, background=bg)
        ren = fresnel_b200.TileBasedRenderer(W, H, background=bg, t_eps=0.0)
    else:
        io, do = fo.render_dense(Lo["positions"], Lo["scales"], Lo["rotations"], Lo["colors"], Lo["opacities"], cam_o,
                                 W, H, background=bg)
        ren = fresnel_b200.DifferentiableGaussianRenderer(W, H, background=bg, t_eps=0.0)
    torch.autograd.backward((io, do), (gi, gd))
    L = {k: v.to(dev()).requires_grad_(True) for k, v in inp.items()}
    img, dep = ren(L["positions"], L["scales"], L["rotations"], L["colors"], L["opacities"], cam, return_depth=True)
    torch.autograd.backward((img, dep), (gi.to(dev()), gd.to(dev())))
    assert rel(img.detach().cpu(), io.detach()) < IMG_TOL and rel(dep.detach().cpu(), do.detach()) < IMG_TOL
    for k in GRAD_NAMES:
        want = Lo[k].grad if Lo[k].grad is not None else torch.zeros_like(Lo[k])
        assert rel(L[k].grad.cpu(), want) < GRAD_TOL, k


def test_inputs_need_not_be_contiguous_fp32_or_require_grad():
    """Strided views, float64 inputs and inputs without requires_grad are accepted like the reference accepts
    them; gradients flow only to the tensors that ask for them."""
    W = H = 64
    inp = fo.synthetic_cloud(500, seed=8, s_lo=0.01, s_hi=0.06)
    cam = fresnel_b200.Camera(0.8 * W, 0.8 * W, W / 2, H / 2, W, H)
    ren = fresnel_b200.TileBasedRenderer(W, H)
    base = {k: v.to(dev()) for k, v in inp.items()}
    ref = ren(base["positions"], base["scales"], base["rotations"], base["colors"], base["opacities"], cam)
    wide = torch.zeros(500, 6, device=dev())
    wide[:, ::2] = base["positions"]
    pos_strided = wide[:, ::2]                                   # non-contiguous view
    col64 = base["colors"].double().requires_grad_(True)
    img = ren(pos_strided, base["scales"], base["rotations"], col64, base["opacities"], cam)
    assert torch.equal(img, ref)
    img.sum().backward()
    assert col64.grad is not None and col64.grad.dtype == torch.float64 and bool(torch.isfinite(col64.grad).all())
    with torch.no_grad():
        img2 = ren(base["positions"], base["scales"], base["rotations"], base["colors"], base["opacities"], cam)
    assert torch.equal(img2, ref) and not img2.requires_grad


def test_more_than_32_views_in_one_batch():
    """render_batch splits batches above the kernels' 32-view limit; the result equals per-view calls."""
    W = H = 32
    B, N = 37, 60
    clouds = [fo.synthetic_cloud(N, seed=200 + b, s_lo=0.02, s_hi=0.08) for b in range(B)]
    stack = {k: torch.stack([c[k] for c in clouds]).to(dev()) for k in GRAD_NAMES}
    cam = fresnel_b200.Camera(0.8 * W, 0.8 * W, W / 2, H / 2, W, H)
    ren = fresnel_b200.TileBasedRenderer(W, H, background=(0.2, 0.2, 0.2))
    img, dep, alpha = ren.render_batch(stack["positions"], stack["scales"], stack["rotations"], stack["colors"],
                                       stack["opacities"], cam)
    assert img.shape == (B, 3, H, W) and dep.shape == (B, H, W)
    for b in (0, 31, 32, 36):
        one, d1 = ren(stack["positions"][b], stack["scales"][b], stack["rotations"][b], stack["colors"][b],
                      stack["opacities"][b], cam, return_depth=True)
        assert torch.equal(one, img[b]) and torch.equal(d1, dep[b])


def test_cpu_tensors_are_refused_by_every_renderer():
    W = H = 16
    inp = fo.synthetic_cloud(10, seed=1)
    cam = fresnel_b200.Camera(12.8, 12.8, 8, 8, W, H)
    for kind, (ren, needs_phase) in renderers(W, H, (0, 0, 0)).items():
        with pytest.raises(TypeError):
            call(ren, needs_phase, inp, cam)


def test_missing_phases_raise_value_error_like_the_reference():
    W = H = 16
    L = {k: v.to(dev()) for k, v in fo.synthetic_cloud(10, seed=1).items()}
    cam = fresnel_b200.Camera(12.8, 12.8, 8, 8, W, H)
    for ren in (fresnel_b200.WaveFieldRenderer(W, H), fresnel_b200.ASMWaveFieldRenderer(W, H).to(dev())):
        with pytest.raises(ValueError):                          # DR:779-780, DR:1187-1188
            ren(L["positions"], L["scales"], L["rotations"], L["colors"], L["opacities"], cam)
