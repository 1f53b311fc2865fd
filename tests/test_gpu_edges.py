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


def test_parameter_views_at_any_float_offset_are_accepted():
    """The reference accepts any tensor; the kernels read rotations (and write their gradient) as 16-byte vectors.
    Contiguous (N, 4) views into a flat buffer that start at a float offset which is not a multiple of four
    (torch.split of a flat parameter vector, flat[a:b].view(n, 4) with odd n) must render like a fresh tensor instead
    of faulting on the device - and the C ABI refuses a misaligned pointer with FRB_E_INVALID."""
    from fresnel_b200 import _lib
    from fresnel_b200.renderer import _ptr, _stream
    from fresnel_b200.camera import camera_vector
    d = dev()
    W, H, n = 64, 48, 333                                  # odd n: 3n is odd, so the rotations start misaligned
    inp = fo.synthetic_cloud(n, seed=77, s_lo=0.01, s_hi=0.08)
    cam = fo.default_camera(W, H)
    flat = torch.cat([inp["positions"].reshape(-1), inp["rotations"].reshape(-1), inp["scales"].reshape(-1),
                      inp["colors"].reshape(-1), inp["opacities"]]).to(d).requires_grad_(True)
    pos, rot, scl, col, opa = torch.split(flat, [3 * n, 4 * n, 3 * n, 3 * n, n])
    rot_v = rot.view(n, 4)
    assert rot_v.data_ptr() % 16 != 0
    ren = fresnel_b200.TileBasedRenderer(W, H, background=(0.1, 0.2, 0.3))
    img, dep = ren(pos.view(n, 3), scl.view(n, 3), rot_v, col.view(n, 3), opa, cam, return_depth=True)
    (img.sum() + dep.sum()).backward()
    L = {k: inp[k].to(d).requires_grad_(True) for k in GRAD_NAMES}
    img2, dep2 = ren(L["positions"], L["scales"], L["rotations"], L["colors"], L["opacities"], cam, return_depth=True)
    (img2.sum() + dep2.sum()).backward()
    assert torch.equal(img, img2) and torch.equal(dep, dep2)
    want = torch.cat([L["positions"].grad.reshape(-1), L["rotations"].grad.reshape(-1), L["scales"].grad.reshape(-1),
                      L["colors"].grad.reshape(-1), L["opacities"].grad])
    assert rel(flat.grad.cpu(), want.cpu()) < 1e-6
    # C ABI: the same misaligned pointer is an argument error, not a device fault
    lib = _lib.lib()
    rec = torch.empty(n, 12, device=d)
    db = torch.empty(n, dtype=torch.int32, device=d)
    tt = torch.empty(n, dtype=torch.int32, device=d)
    camv = camera_vector(cam, W, H)
    rc = lib.frb_project_fwd(n, 1, _ptr(L["positions"]), _ptr(L["scales"]), rot_v.data_ptr(), _ptr(L["colors"]),
                             _ptr(L["opacities"]), camv.ctypes.data, 64.0, _ptr(rec), None, _ptr(db), _ptr(tt), None,
                             _stream())
    assert rc != 0 and b"invalid" in lib.frb_error_string(rc)
    torch.cuda.synchronize()                               # no sticky error


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs in one process")
def test_forward_and_backward_on_two_devices_of_one_process():
    """The opt-in for more than 48 KB of dynamic shared memory is a per-DEVICE attribute: one process that renders on
    cuda:0 and then on cuda:1 must be able to launch the backward kernels (105 KB) on both, for the plain and the
    phase-blending compositor, and get the same result on both."""
    W, H, n = 96, 80, 1500
    inp = fo.synthetic_cloud(n, seed=79, s_lo=0.01, s_hi=0.06)
    cam = fo.default_camera(W, H)
    outs = []
    for idx in (0, 1, 0):
        d = torch.device("cuda", idx)
        for phase in (False, True):
            ren = fresnel_b200.TileBasedRenderer(W, H, use_phase_blending=phase)
            L = {k: inp[k].to(d).requires_grad_(True) for k in GRAD_NAMES + ("phases",)}
            img, dep = ren(L["positions"], L["scales"], L["rotations"], L["colors"], L["opacities"], cam,
                           return_depth=True, phases=L["phases"] if phase else None)
            (img.sum() + dep.sum()).backward()
            torch.cuda.synchronize(d)
            outs.append((idx, phase, img.detach().cpu(), L["positions"].grad.cpu()))
    for idx, phase, img, g in outs[2:]:
        ref = outs[1 if phase else 0]
        assert torch.equal(img, ref[2]), (idx, phase)
        assert rel(g, ref[3]) < 1e-5, (idx, phase)


@pytest.mark.parametrize("switches", [{"FRB_FWD_BY_RECORD": "1", "FRB_BWD_HALVES": "1"}, {"FRB_GATHER": "0"},
                                      {"FRB_CLUSTER_SORT": "1", "FRB_PDL": "0"}])
def test_alternate_kernel_paths_pass_the_golden_tests(switches):
    """The library keeps the slower form of four choices behind environment switches (the
    record-by-record forward, the one-CTA-per-tile backward, the sorted record copy instead of the TMA gather, the
    cluster-resident depth sort, plain launches).  They are read once per process, so the golden parity tests of the
    tile renderer run again in a child process with the switches set."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, **switches)
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(root, "tests", "test_gpu_parity.py"), "-x", "-q",
                        "-k", "matches_reference_golden or tile_keys_bit_exact or batched"], cwd=root, env=env,
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-1000:]
    assert " passed" in r.stdout
