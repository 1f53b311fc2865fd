#!/usr/bin/env python3
"""Generate tests/golden/*.npz by executing the UNMODIFIED reference module.

Run in the build container only (needs /root/reference, CPU):

    python oracle/make_golden.py [--only NAME]

For each fixture this script
  1. draws seeded synthetic inputs and re-draws "fragile" Gaussians - those
     whose integer rectangle, visibility bit or depth rank could flip under a
     few-ulp change of the projected floats - so that the integer pins are a
     property of the input and not of one BLAS build (SURVEY.md section 8c:
     "fixtures must be tie-free"); fixtures tagged ``ties`` keep exact depth
     ties on purpose (identity view, so depth = -z is exact everywhere);
  2. runs the reference renderer forward + backward with ``torch.argsort``
     forced to ``stable=True`` (the only patch; DR:527 is otherwise
     implementation-defined on ties);
  3. derives the integer pins (visible, rect, stable depth order) from the
     REFERENCE's own intermediates (compute_2d_covariance / _compute_radius);
  4. cross-checks the oracle restatement against all of it and refuses to write
     the fixture if the oracle disagrees;
  5. writes inputs, outputs, gradients and pins to tests/golden/NAME.npz.

Phase-blending gradients cannot come from the reference (it raises inside
autograd, SURVEY.md note 3); they come from the oracle's ``.clone()``
restatement and the fixture says so (``grad_source = 'oracle_clone'``).
"""

from __future__ import annotations

import argparse
import math
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference/scripts")

from oracle import fresnel_oracle as fo  # noqa: E402

_orig_argsort = torch.argsort


def _stable_argsort(x, *a, **k):
    k["stable"] = True
    return _orig_argsort(x, *a, **k)


torch.argsort = _stable_argsort
import models.differentiable_renderer as dr  # noqa: E402  (the reference)

GOLD = os.path.join(ROOT, "tests", "golden")


def ref_camera(cam: fo.Camera) -> "dr.Camera":
    c = dr.Camera(cam.fx, cam.fy, cam.cx, cam.cy, cam.width, cam.height, cam.near, cam.far)
    c.set_view(cam.view_matrix.clone())
    return c


def cam_vec(cam: fo.Camera) -> np.ndarray:
    v = cam.view_matrix.to(torch.float32).numpy()[:3, :].reshape(-1)
    return np.concatenate([v, np.array([cam.fx, cam.fy, cam.cx, cam.cy, cam.width, cam.height,
                                        cam.near, cam.far])]).astype(np.float64)


def fragile_mask(inp, cam, W, H, max_radius, allow_ties) -> np.ndarray:
    """Gaussians whose integer pins sit within a safety margin of flipping."""
    pn = fo.pins(inp["positions"], inp["scales"], inp["rotations"], cam, W, H, max_radius)
    u, v, r = (pn[k].astype(np.float64) for k in ("u", "v", "radius"))
    d = pn["depth"].astype(np.float64)
    eps = 2e-3
    frag = np.zeros(u.shape[0], bool)
    for e in (u - r, u + r, v - r, v + r):
        frag |= np.abs(e - np.round(e)) < eps
    frag |= (np.abs(d - cam.near) < 1e-4) | (np.abs(d - cam.far) < 1e-2)
    frag |= ~np.isfinite(u) | ~np.isfinite(v) | ~np.isfinite(r)
    if not allow_ties:
        bits = pn["depth_bits"].astype(np.int64)
        o = np.argsort(bits, kind="stable")
        gap = np.diff(bits[o])
        close = np.zeros_like(frag)
        near = np.nonzero(gap < 16)[0]
        close[o[near]] = True
        close[o[near + 1]] = True
        frag |= close
    return frag


def settle(inp, cam, W, H, max_radius, seed, redraw, allow_ties=False):
    """Re-draw fragile Gaussians until none is left."""
    g = torch.Generator().manual_seed(seed + 1000)
    for it in range(50):
        frag = fragile_mask(inp, cam, W, H, max_radius, allow_ties)
        n = int(frag.sum())
        if n == 0:
            return inp
        idx = torch.from_numpy(np.nonzero(frag)[0])
        redraw(inp, idx, g)
    raise RuntimeError("could not settle fixture")


def upstream(H, W, seed=1):
    g = torch.Generator().manual_seed(seed)
    return (torch.rand(3, H, W, generator=g) * 2 - 1, torch.rand(H, W, generator=g) * 2 - 1)


def leafs(inp, names):
    return {k: (v.clone().requires_grad_(True) if k in names else v.clone()) for k, v in inp.items()}


GRAD_NAMES = ("positions", "scales", "rotations", "colors", "opacities")


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-3))


def run_tile(name, inp, cam, W, H, bg=(0.0, 0.0, 0.0), max_radius=64, phase=False, amp=0.25,
             note="", forward_only=False):
    """forward_only: no gradients in the fixture (full-size phase-blending fixture: the clone restatement's tape
    would need tens of GB); image / depth from the reference, oracle cross-checked under no_grad."""
    t0 = time.time()
    gi, gd = upstream(H, W)
    rc = ref_camera(cam)
    ren = dr.TileBasedRenderer(W, H, background=bg, max_radius=max_radius,
                               use_phase_blending=phase, phase_amplitude=amp)
    # reference forward (+ backward when it can)
    L = leafs(inp, GRAD_NAMES + (("phases",) if phase else ()))
    ph = L["phases"] if phase else None
    if phase:
        with torch.no_grad():
            img_r, dep_r = ren(L["positions"], L["scales"], L["rotations"], L["colors"],
                               L["opacities"], rc, return_depth=True, phases=ph)
        grads_r = None
    else:
        img_r, dep_r = ren(L["positions"], L["scales"], L["rotations"], L["colors"],
                           L["opacities"], rc, return_depth=True, phases=None)
        ((img_r * gi).sum() + (dep_r * gd).sum()).backward()
        grads_r = {k: (L[k].grad if L[k].grad is not None else torch.zeros_like(L[k])).numpy()
                   for k in GRAD_NAMES}
    # reference-derived pins
    with torch.no_grad():
        cov, m2, dep = dr.compute_2d_covariance(inp["positions"], inp["scales"], inp["rotations"], rc)
        rad = ren._compute_radius(cov)
        vis = fo.visibility(m2[:, 0], m2[:, 1], dep, rad, cam, W, H).numpy()
        rect = fo.rects(m2[:, 0], m2[:, 1], rad, W, H).astype(np.int32)
        order = torch.argsort(dep).numpy().astype(np.int32)
    # oracle
    Lo = leafs(inp, GRAD_NAMES + (("phases",) if phase else ()))
    with torch.set_grad_enabled(not forward_only):
        img_o, dep_o, alpha_o = fo.render_tile_based(
            Lo["positions"], Lo["scales"], Lo["rotations"], Lo["colors"], Lo["opacities"], cam, W, H,
            background=bg, max_radius=max_radius, use_phase_blending=phase, phase_amplitude=amp,
            phases=Lo["phases"] if phase else None)
    grads_o = {}
    if not forward_only:
        ((img_o * gi).sum() + (dep_o * gd).sum()).backward()
        grads_o = {k: (Lo[k].grad if Lo[k].grad is not None else torch.zeros_like(Lo[k])).numpy()
                   for k in GRAD_NAMES + (("phases",) if phase else ())}
    pn = fo.pins(inp["positions"], inp["scales"], inp["rotations"], cam, W, H, max_radius)

    # cross-checks: oracle vs reference
    e_img, e_dep = rel(img_o.detach(), img_r.detach()), rel(dep_o.detach(), dep_r.detach())
    assert e_img < 1e-5 and e_dep < 1e-5, (name, e_img, e_dep)
    assert np.array_equal(pn["visible"], vis), name
    vi = np.nonzero(vis)[0]
    assert np.array_equal(pn["rect"][vi], rect[vi]), name
    assert np.array_equal(pn["order"], order), name
    worst = 0.0
    if grads_r is not None:
        for k in GRAD_NAMES:
            e = rel(grads_o[k], grads_r[k])
            assert e < 1e-4, (name, k, e)
            worst = max(worst, e)
        grads, src = grads_r, "reference"
    else:
        grads, src = grads_o, ("none" if forward_only else "oracle_clone")

    out = dict(cam=cam_vec(cam), W=W, H=H, bg=np.array(bg, np.float32), max_radius=max_radius,
               phase_blending=int(phase), phase_amplitude=amp,
               image=img_r.detach().numpy(), depth=dep_r.detach().numpy(),
               alpha=alpha_o.detach().numpy(),
               gimage=gi.numpy(), gdepth=gd.numpy(),
               visible=vis, rect=rect, order=order, grad_source=src, note=note)
    for k in GRAD_NAMES + ("phases",):
        out["in_" + k] = inp[k].numpy()
    for k, g in grads.items():
        out["grad_" + k] = g
    os.makedirs(GOLD, exist_ok=True)
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), **out)
    print(f"{name}: N={inp['positions'].shape[0]} vis={int(vis.sum())} M={pn['keys'].shape[0]} "
          f"oracle-vs-ref img {e_img:.2e} depth {e_dep:.2e} grads {worst:.2e} [{src}] "
          f"{time.time() - t0:.1f}s")


def add_f64_companion(name, kind):
    """Appends image64 / depth64 to tests/golden/NAME.npz: the SAME unmodified reference module run in float64
    (torch default dtype switched for the call).  It tells which side of a 1e-5 disagreement is the noisy one: at
    ~1000 overlaps per pixel the reference's own fp32 run is ~1e-5 away from its fp64 run."""
    path = os.path.join(GOLD, name + ".npz")
    z = dict(np.load(path))
    W, H = int(z["W"]), int(z["H"])
    c = z["cam"]
    torch.set_default_dtype(torch.float64)
    try:
        rc = dr.Camera(float(c[12]), float(c[13]), float(c[14]), float(c[15]), W, H, float(c[18]), float(c[19]))
        view = torch.eye(4)
        view[:3, :] = torch.from_numpy(np.asarray(c[:12], np.float32).reshape(3, 4)).double()
        rc.set_view(view)
        names = GRAD_NAMES + (("phases",) if kind == "wave" else ())
        inp = {k: torch.from_numpy(z["in_" + k]).double() for k in names}
        bg = tuple(float(x) for x in z["bg"])
        with torch.no_grad():
            if kind == "wave":
                ren = dr.WaveFieldRenderer(W, H, background=bg)
                img, dep = ren(inp["positions"], inp["scales"], inp["rotations"], inp["colors"], inp["opacities"], rc,
                               return_depth=True, phases=inp["phases"])
            else:
                ren = dr.TileBasedRenderer(W, H, background=bg, max_radius=int(z["max_radius"]))
                img, dep = ren(inp["positions"], inp["scales"], inp["rotations"], inp["colors"], inp["opacities"], rc,
                               return_depth=True)
    finally:
        torch.set_default_dtype(torch.float32)
    z["image64"], z["depth64"] = img.numpy(), dep.numpy()
    np.savez_compressed(path, **z)
    print(f"{name}: fp64 run of the reference added; fp32 reference vs fp64: image {rel(z['image'], z['image64']):.2e} "
          f"depth {rel(z['depth'], z['depth64']):.2e}")


def run_wave(name, inp, cam, W, H, bg, per_channel, note=""):
    gi, gd = upstream(H, W)
    rc = ref_camera(cam)
    names = GRAD_NAMES + ("phases",)
    ren = dr.WaveFieldRenderer(W, H, background=bg)
    L = leafs(inp, names)
    img_r, dep_r = ren(L["positions"], L["scales"], L["rotations"], L["colors"], L["opacities"],
                       rc, return_depth=True, phases=L["phases"])
    ((img_r * gi).sum() + (dep_r * gd).sum()).backward()
    Lo = leafs(inp, names)
    img_o, dep_o = fo.render_wave(Lo["positions"], Lo["scales"], Lo["rotations"], Lo["colors"],
                                  Lo["opacities"], cam, W, H, Lo["phases"], background=bg)
    ((img_o * gi).sum() + (dep_o * gd).sum()).backward()
    assert rel(img_o.detach(), img_r.detach()) < 5e-6 and rel(dep_o.detach(), dep_r.detach()) < 2e-5   # summation order
    out = dict(cam=cam_vec(cam), W=W, H=H, bg=np.array(bg, np.float32), max_radius=64,
               image=img_r.detach().numpy(), depth=dep_r.detach().numpy(),
               gimage=gi.numpy(), gdepth=gd.numpy(), grad_source="reference", note=note)
    for k in names:
        out["in_" + k] = inp[k].numpy()
        gr, go = L[k].grad.numpy(), Lo[k].grad.numpy()
        assert rel(go, gr) < 5e-5, (name, k, rel(go, gr))
        out["grad_" + k] = gr
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), **out)
    print(f"{name}: wave ok, img max {float(img_r.max()):.3f}")


def run_asm(name, inp, cam, W, H, bg, wl, depth_range, note="", num_depth_planes=16, focal_depth=0.5,
            pixel_pitch=1.0 / 256):
    gi, _ = upstream(H, W)
    rc = ref_camera(cam)
    names = GRAD_NAMES + ("phases",)
    ren = dr.ASMWaveFieldRenderer(W, H, background=bg, depth_range=depth_range, num_depth_planes=num_depth_planes,
                                  focal_depth=focal_depth, pixel_pitch=pixel_pitch)
    L = leafs(inp, names)
    img_r = ren(L["positions"], L["scales"], L["rotations"], L["colors"], L["opacities"], rc,
                phases=L["phases"], wavelengths_rgb=wl)
    (img_r * gi).sum().backward()
    Lo = leafs(inp, names)
    img_o, _ = fo.render_asm(Lo["positions"], Lo["scales"], Lo["rotations"], Lo["colors"],
                             Lo["opacities"], cam, W, H, Lo["phases"], wl, background=bg,
                             depth_range=depth_range, num_depth_planes=num_depth_planes, focal_depth=focal_depth,
                             pixel_pitch=pixel_pitch)
    (img_o * gi).sum().backward()
    assert rel(img_o.detach(), img_r.detach()) < 5e-6
    out = dict(cam=cam_vec(cam), W=W, H=H, bg=np.array(bg, np.float32), max_radius=64,
               wavelengths=wl.numpy(), depth_range=np.array(depth_range), num_depth_planes=num_depth_planes,
               focal_depth=focal_depth, pixel_pitch=pixel_pitch,
               image=img_r.detach().numpy(), gimage=gi.numpy(), grad_source="reference", note=note)
    for k in names:
        out["in_" + k] = inp[k].numpy()
        gr, go = L[k].grad.numpy(), Lo[k].grad.numpy()
        assert rel(go, gr) < 1e-4, (name, k, rel(go, gr))
        out["grad_" + k] = gr
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), **out)
    print(f"{name}: asm ok, img max {float(img_r.max()):.3f}")


def run_dense(name, inp, cam, W, H, bg, note=""):
    """DifferentiableGaussianRenderer (DR:245-409): reference forward + backward, oracle cross-check."""
    gi, gd = upstream(H, W)
    rc = ref_camera(cam)
    ren = dr.DifferentiableGaussianRenderer(W, H, background=bg)
    L = leafs(inp, GRAD_NAMES)
    img_r, dep_r = ren(L["positions"], L["scales"], L["rotations"], L["colors"], L["opacities"], rc,
                       return_depth=True)
    ((img_r * gi).sum() + (dep_r * gd).sum()).backward()
    Lo = leafs(inp, GRAD_NAMES)
    img_o, dep_o = fo.render_dense(Lo["positions"], Lo["scales"], Lo["rotations"], Lo["colors"],
                                   Lo["opacities"], cam, W, H, background=bg)
    ((img_o * gi).sum() + (dep_o * gd).sum()).backward()
    e_img, e_dep = rel(img_o.detach(), img_r.detach()), rel(dep_o.detach(), dep_r.detach())
    assert e_img < 1e-5 and e_dep < 1e-5, (name, e_img, e_dep)
    out = dict(cam=cam_vec(cam), W=W, H=H, bg=np.array(bg, np.float32),
               image=img_r.detach().numpy(), depth=dep_r.detach().numpy(),
               gimage=gi.numpy(), gdepth=gd.numpy(), grad_source="reference", note=note)
    worst = 0.0
    for k in GRAD_NAMES:
        out["in_" + k] = inp[k].numpy()
        gr, go = L[k].grad.numpy(), Lo[k].grad.numpy()
        worst = max(worst, rel(go, gr))
        assert rel(go, gr) < 1e-4, (name, k, rel(go, gr))
        out["grad_" + k] = gr
    out["in_phases"] = inp["phases"].numpy()
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), **out)
    print(f"{name}: dense ok, oracle-vs-ref img {e_img:.2e} depth {e_dep:.2e} grads {worst:.2e}")


def run_fourier(name, inp, cam, W, H, bg, note=""):
    """FourierGaussianRenderer (DR:1500-1774): reference forward + backward, oracle cross-check."""
    gi, _ = upstream(H, W)
    rc = ref_camera(cam)
    ren = dr.FourierGaussianRenderer(W, H, background=bg)
    L = leafs(inp, GRAD_NAMES)
    img_r = ren(L["positions"], L["scales"], L["rotations"], L["colors"], L["opacities"], rc)
    (img_r * gi).sum().backward()
    Lo = leafs(inp, GRAD_NAMES)
    img_o = fo.render_fourier(Lo["positions"], Lo["scales"], Lo["rotations"], Lo["colors"], Lo["opacities"],
                              cam, W, H, background=bg)
    (img_o * gi).sum().backward()
    e_img = rel(img_o.detach(), img_r.detach())
    assert e_img < 1e-5, (name, e_img)
    out = dict(cam=cam_vec(cam), W=W, H=H, bg=np.array(bg, np.float32), image=img_r.detach().numpy(),
               gimage=gi.numpy(), grad_source="reference", note=note)
    worst = 0.0
    for k in GRAD_NAMES:
        out["in_" + k] = inp[k].numpy()
        gr = (L[k].grad if L[k].grad is not None else torch.zeros_like(L[k])).numpy()
        go = (Lo[k].grad if Lo[k].grad is not None else torch.zeros_like(Lo[k])).numpy()
        worst = max(worst, rel(go, gr))
        assert rel(go, gr) < 1e-4, (name, k, rel(go, gr))
        out["grad_" + k] = gr
    out["in_phases"] = inp["phases"].numpy()
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), **out)
    print(f"{name}: fourier ok, oracle-vs-ref img {e_img:.2e} grads {worst:.2e}, img max {float(img_r.max()):.3f}")


def run_simplified(name, inp, cam, W, H, bg, note=""):
    """SimplifiedRenderer (DR:1347-1458): reference forward + backward, oracle cross-check."""
    gi, gd = upstream(H, W)
    rc = ref_camera(cam)
    ren = dr.SimplifiedRenderer(W, H, background=bg)
    names = ("positions", "colors", "opacities")
    # the reference cannot backpropagate through this renderer (in-place slice writes on tensors saved for the
    # backward: "modified by an inplace operation", as on the phase-blending path): forward from the reference,
    # gradients from the oracle's functional restatement of the same expressions
    with torch.no_grad():
        img_r, dep_r = ren(inp["positions"], inp["scales"], inp["rotations"], inp["colors"], inp["opacities"], rc,
                           return_depth=True)
    Lo = leafs(inp, names)
    img_o, dep_o = fo.render_simplified(Lo["positions"], Lo["scales"], Lo["colors"], Lo["opacities"], cam, W, H,
                                        background=bg)
    ((img_o * gi).sum() + (dep_o * gd).sum()).backward()
    e_img, e_dep = rel(img_o.detach(), img_r.detach()), rel(dep_o.detach(), dep_r.detach())
    assert e_img < 1e-5 and e_dep < 1e-6, (name, e_img, e_dep)
    out = dict(cam=cam_vec(cam), W=W, H=H, bg=np.array(bg, np.float32), image=img_r.detach().numpy(),
               depth=dep_r.detach().numpy(), gimage=gi.numpy(), gdepth=gd.numpy(), grad_source="oracle_functional",
               note=note)
    for k in GRAD_NAMES + ("phases",):
        out["in_" + k] = inp[k].numpy()
    for k in names:
        out["grad_" + k] = Lo[k].grad.numpy()
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), **out)
    print(f"{name}: simplified ok, oracle-vs-ref img {e_img:.2e} depth {e_dep:.2e} [gradients: oracle], "
          f"depth pixels hit {int((dep_r > 0).sum())}")


# ---------------------------------------------------------------- fixtures
def redraw_std(inp, idx, g):
    n = idx.numel()
    p = torch.randn(n, 3, generator=g) * 0.5
    p[:, 2] -= 2.0
    inp["positions"][idx] = p


def fx_c1():
    """Config 1 (BASELINE.json configs[0]): 16,384 Gaussians, 256x256, one view."""
    W = H = 256
    cam = fo.default_camera(W)
    inp = fo.synthetic_cloud(16384, seed=0)
    inp = settle(inp, cam, W, H, 64, 0, redraw_std)
    run_tile("c1_tile_16k_256", inp, cam, W, H, note="config 1, identity view, black background")


def fx_rot():
    """Rotated look-at camera, non-black background, non-square non-multiple-of-16 image."""
    W, H = 144, 120
    cam = fo.camera_from_pose(math.radians(20.0), math.radians(35.0), 128)
    cam.width, cam.height, cam.cx, cam.cy = W, H, W / 2, H / 2
    inp = fo.synthetic_cloud(2048, seed=3, s_lo=0.01, s_hi=0.06)
    inp["positions"][:, 2] += 2.0            # cloud around the origin; camera orbits at distance 2

    def redraw(inp, idx, g):
        inp["positions"][idx] = torch.randn(idx.numel(), 3, generator=g) * 0.5

    inp = settle(inp, cam, W, H, 64, 3, redraw)
    run_tile("tile_rotcam_2k_144x120", inp, cam, W, H, bg=(0.2, 0.3, 0.4),
             note="create_camera_from_pose(el=20deg, az=35deg), W=144 H=120, coloured background")


def fx_edge():
    """Near-plane / behind-camera / border / radius-cap / depth-tie / saturation edge cases."""
    W, H = 96, 80
    cam = fo.default_camera(W, H)
    n = 1024
    g = torch.Generator().manual_seed(7)
    inp = fo.synthetic_cloud(n, seed=7, s_lo=0.01, s_hi=0.08)
    pos = inp["positions"]
    pos[:, :2] *= 2.0                                   # many centres outside the image
    pos[:128, 2] = torch.rand(128, generator=g) * 0.3 - 0.15     # around and behind the camera plane
    pos[128:256, 2] = -(torch.rand(128, generator=g) * 0.2 + 0.05)   # very close: radius cap
    inp["scales"][256:320] *= 8.0                       # big splats: 64-px cap, whole-image rects
    zq = torch.round(pos[320:704, 2] * 8) / 8           # exact depth ties (Fresnel-zone style)
    pos[320:704, 2] = zq
    inp["opacities"][704:832] = 1.5                     # alpha clamp at 0.99 active
    inp["opacities"][832:840] = -0.2                    # alpha clamp at 0 active
    inp["rotations"][840:848] = 0.0                     # zero quaternion: normalise eps path
    inp["colors"][848:912] *= 3.0                       # image clamp at 1 active

    def redraw(inp, idx, g):
        m = idx[(idx >= 320) & (idx < 704)]
        o = idx[(idx < 320) | (idx >= 704)]
        # keep the special structure: nudge x,y only (depth ties and near-plane z stay)
        inp["positions"][idx, 0] += (torch.rand(idx.numel(), generator=g) - 0.5) * 0.05
        inp["positions"][idx, 1] += (torch.rand(idx.numel(), generator=g) - 0.5) * 0.05
        inp["scales"][idx] *= 1.0 + 0.01 * torch.rand(idx.numel(), 3, generator=g)
        del m, o

    inp = settle(inp, cam, W, H, 64, 7, redraw, allow_ties=True)
    run_tile("tile_edge_1k_96x80", inp, cam, W, H, bg=(0.1, 0.0, 0.3),
             note="ties; near plane, behind camera, border, radius cap, alpha and image clamps, zero quat")


def fx_culled():
    W = H = 64
    cam = fo.default_camera(W)
    inp = fo.synthetic_cloud(64, seed=9)
    inp["positions"][:, 2] = inp["positions"][:, 2].abs() + 1.0     # all behind the camera
    run_tile("tile_allculled_64", inp, cam, W, H, bg=(0.25, 0.5, 0.75), note="nothing visible (DR:545-552)")


def fx_params():
    """Non-default renderer / camera parameters: max_radius 24 (the cap is active for most Gaussians), near 1.4 and
    far 2.6 (both planes cut through the cloud), fx != fy and an off-centre principal point."""
    W, H = 96, 64
    cam = fo.Camera(0.9 * W, 0.7 * W, W / 2 + 5.0, H / 2 - 3.0, W, H, 1.4, 2.6)
    inp = fo.synthetic_cloud(1500, seed=37, s_lo=0.02, s_hi=0.15)
    inp = settle(inp, cam, W, H, 24, 37, redraw_std)
    run_tile("tile_params_1500_96x64", inp, cam, W, H, bg=(0.05, 0.1, 0.2), max_radius=24,
             note="max_radius=24, near=1.4, far=2.6, fx=0.9W, fy=0.7W, principal point (W/2+5, H/2-3)")


def fx_phase():
    W = H = 128
    cam = fo.default_camera(W)
    inp = fo.synthetic_cloud(2048, seed=11, s_lo=0.01, s_hi=0.05)
    inp = settle(inp, cam, W, H, 64, 11, redraw_std)
    run_tile("tile_phase_2k_128", inp, cam, W, H, bg=(0.0, 0.0, 0.0), phase=True, amp=0.25,
             note="use_phase_blending=True; forward from the reference, gradients from the oracle clone restatement")


def fx_phase_rot():
    """Phase blending on a non-square image (sides not multiples of the tile), rotated look-at camera, coloured
    background, amplitude 0.4."""
    W, H = 112, 80
    cam = fo.camera_from_pose(math.radians(10.0), math.radians(120.0), 128)
    cam.width, cam.height, cam.cx, cam.cy = W, H, W / 2, H / 2
    inp = fo.synthetic_cloud(1500, seed=29, s_lo=0.01, s_hi=0.06)
    inp["positions"][:, 2] += 2.0

    def redraw(inp, idx, g):
        inp["positions"][idx] = torch.randn(idx.numel(), 3, generator=g) * 0.5

    inp = settle(inp, cam, W, H, 64, 29, redraw)
    run_tile("tile_phase_rot_1500_112x80", inp, cam, W, H, bg=(0.3, 0.1, 0.2), phase=True, amp=0.4,
             note="use_phase_blending=True, amplitude 0.4, look-at camera el 10 az 120, W=112 H=80; forward from the "
                  "reference, gradients from the oracle clone restatement")


def fx_wave():
    W = H = 128
    cam = fo.default_camera(W)
    inp = fo.synthetic_cloud(2048, seed=13, s_lo=0.01, s_hi=0.05, phase_hi=2 * math.pi)
    inp = settle(inp, cam, W, H, 64, 13, redraw_std, allow_ties=True)
    run_wave("wave_scalar_2k_128", inp, cam, W, H, (0.1, 0.2, 0.3), False, note="(N,) phases")
    g = torch.Generator().manual_seed(14)
    inp3 = dict(inp)
    inp3["phases"] = torch.rand(2048, 3, generator=g) * 2 * math.pi
    run_wave("wave_rgb_2k_128", inp3, cam, W, H, (0.1, 0.2, 0.3), True, note="(N,3) phases")
    add_f64_companion("wave_scalar_2k_128", "wave")
    add_f64_companion("wave_rgb_2k_128", "wave")


def fx_wave_rot():
    """WaveFieldRenderer on a non-square image (sides not multiples of the tile) seen by a rotated look-at camera,
    per-channel phases."""
    W, H = 112, 80
    cam = fo.camera_from_pose(math.radians(25.0), math.radians(-50.0), 128)
    cam.width, cam.height, cam.cx, cam.cy = W, H, W / 2, H / 2
    inp = fo.synthetic_cloud(1500, seed=27, s_lo=0.01, s_hi=0.05, phase_hi=2 * math.pi)
    inp["positions"][:, 2] += 2.0

    def redraw(inp, idx, g):
        inp["positions"][idx] = torch.randn(idx.numel(), 3, generator=g) * 0.5

    inp = settle(inp, cam, W, H, 64, 27, redraw, allow_ties=True)
    g = torch.Generator().manual_seed(28)
    inp["phases"] = torch.rand(1500, 3, generator=g) * 2 * math.pi
    run_wave("wave_rot_1500_112x80", inp, cam, W, H, (0.05, 0.2, 0.1), True,
             note="look-at camera el 25 az -50, W=112 H=80, (N,3) phases")
    add_f64_companion("wave_rot_1500_112x80", "wave")


def fx_asm():
    W = H = 64
    cam = fo.default_camera(W)
    inp = fo.synthetic_cloud(1024, seed=15, s_lo=0.01, s_hi=0.05, phase_hi=2 * math.pi)
    inp = settle(inp, cam, W, H, 64, 15, redraw_std, allow_ties=True)
    wl = torch.tensor([0.0635, 0.05, 0.041])
    run_asm("asm_1k_64", inp, cam, W, H, (0.05, 0.1, 0.15), wl, (0.1, 4.0),
            note="wavelengths_rgb=(0.0635,0.05,0.041), depth_range=(0.1,4.0), 16 planes")


def fx_asm_rot():
    """ASM on a non-square image whose sides are not multiples of the tile, rotated look-at camera: the frequency grids
    of the propagator (fftfreq per axis, meshgrid 'xy', DR:1001-1008) and the ragged tile edges both differ from the
    square default-camera fixture."""
    W, H = 112, 80
    cam = fo.camera_from_pose(math.radians(-15.0), math.radians(40.0), 128)
    cam.width, cam.height, cam.cx, cam.cy = W, H, W / 2, H / 2
    inp = fo.synthetic_cloud(1500, seed=23, s_lo=0.01, s_hi=0.05, phase_hi=2 * math.pi)
    inp["positions"][:, 2] += 2.0            # cloud around the origin; the camera orbits at distance 2

    def redraw(inp, idx, g):
        inp["positions"][idx] = torch.randn(idx.numel(), 3, generator=g) * 0.5

    inp = settle(inp, cam, W, H, 64, 23, redraw, allow_ties=True)
    wl = torch.tensor([0.0635, 0.05, 0.041])
    run_asm("asm_rot_1500_112x80", inp, cam, W, H, (0.15, 0.05, 0.1), wl, (0.1, 4.0),
            note="look-at camera el -15 az 40, W=112 H=80, wavelengths_rgb=(0.0635,0.05,0.041), depth_range=(0.1,4.0)")


def fx_asm_params():
    """Non-default propagator parameters: 8 depth planes, focal depth 1.2, pixel pitch 1/128, other wavelengths."""
    W, H = 80, 64
    cam = fo.default_camera(W, H)
    inp = fo.synthetic_cloud(1200, seed=41, s_lo=0.01, s_hi=0.05, phase_hi=2 * math.pi)
    inp = settle(inp, cam, W, H, 64, 41, redraw_std, allow_ties=True)
    wl = torch.tensor([0.07, 0.045, 0.03])
    run_asm("asm_params_1200_80x64", inp, cam, W, H, (0.0, 0.1, 0.05), wl, (0.5, 3.0),
            note="num_depth_planes=8, focal_depth=1.2, pixel_pitch=1/128, wavelengths (0.07,0.045,0.03), depth_range (0.5,3.0)",
            num_depth_planes=8, focal_depth=1.2, pixel_pitch=1.0 / 128)


def fx_dense():
    """DifferentiableGaussianRenderer: rotated look-at camera, some centres outside the image (100-px margin)."""
    W, H = 96, 80
    cam = fo.camera_from_pose(math.radians(20.0), math.radians(35.0), W)
    cam.cx, cam.cy, cam.width, cam.height = W / 2, H / 2, W, H
    inp = fo.synthetic_cloud(700, seed=17, s_lo=0.01, s_hi=0.08)
    inp["positions"][:, 2] += 2.0                      # the look-at camera orbits the origin
    inp["positions"][:60, 0] *= 4.0                    # far off-centre: in the margin or beyond it
    run_dense("dense_700_96x80", inp, cam, W, H, (0.15, 0.05, 0.3),
              note="DifferentiableGaussianRenderer, look-at camera el 20 az 35, non-zero background")


def fx_fourier():
    """FourierGaussianRenderer: identity view, non-zero background."""
    W, H = 96, 80
    cam = fo.default_camera(W, H)
    inp = fo.synthetic_cloud(1500, seed=19, s_lo=0.005, s_hi=0.04)
    run_fourier("fourier_1500_96x80", inp, cam, W, H, (0.2, 0.1, 0.05), note="FourierGaussianRenderer")


def fx_bin():
    """.bin fixture written and read back by the reference's own functions (DR:1461-1497)."""
    inp = fo.synthetic_cloud(97, seed=23)
    path = os.path.join(GOLD, "cloud_97.bin")
    dr.save_gaussians_to_binary(path, {k: inp[k] for k in GRAD_NAMES})
    back = dr.load_gaussians_from_binary(path)
    np.savez_compressed(os.path.join(GOLD, "cloud_97_loaded.npz"), **{k: v.numpy() for k, v in back.items()})
    for k in GRAD_NAMES:
        assert np.array_equal(back[k].numpy(), inp[k].numpy()), k
    print(f"cloud_97.bin: {os.path.getsize(path)} bytes, reference save -> load round trip exact")


def fx_ply():
    """3DGS .ply fixture written and read back by the REFERENCE's own C++ functions (GaussianCloud::save_ply / load_ply,
    src/core/renderer/renderer.cpp:649-793), compiled by oracle/build_ref.sh into oracle/_ref/libref_cloud_io.so.
    Rows include the transforms' edge cases: scales below the 1e-7 floor, opacity 0 / 1 (logit of 0 and of 1 / 1e-7),
    colours outside [0, 1] (clamped on load)."""
    import ctypes
    import subprocess
    subprocess.check_call(["bash", os.path.join(ROOT, "oracle", "build_ref.sh")])
    lib = ctypes.CDLL(os.path.join(ROOT, "oracle", "_ref", "libref_cloud_io.so"))
    inp = fo.synthetic_cloud(97, seed=31)
    rows = np.concatenate([inp[k].numpy().reshape(97, -1) for k in GRAD_NAMES], axis=1).astype(np.float32)
    rows[0, 3:6] = (0.0, 1e-9, 3e-8)            # below the 1e-7 floor of save_ply
    rows[1, 13], rows[2, 13], rows[3, 13] = 0.0, 1.0, 0.9999999
    rows[4, 10:13] = (-0.3, 1.7, 0.5)           # outside [0, 1]: clamped by load_ply
    rows[5, 3:6] = (40.0, 1e-3, 7.0)
    rows = np.ascontiguousarray(rows)
    ply = os.path.join(GOLD, "cloud_97.ply")
    fp = ctypes.POINTER(ctypes.c_float)
    assert lib.ref_save_ply(rows.ctypes.data_as(fp), 97, ply.encode()) == 0
    back = np.zeros((97, 14), np.float32)
    assert lib.ref_load_ply(ply.encode(), back.ctypes.data_as(fp), 97) == 97
    # the reference's .bin functions on the same cloud: C++ writer against the Python reader / writer of DR:1461-1497
    binp = os.path.join("/tmp", "cloud_97_cpp.bin")
    assert lib.ref_save_binary(rows.ctypes.data_as(fp), 97, binp.encode()) == 0
    py = dr.load_gaussians_from_binary(binp)
    assert np.array_equal(np.concatenate([py[k].numpy().reshape(97, -1) for k in GRAD_NAMES], axis=1), rows)
    body = np.fromfile(ply, dtype="<f4", offset=os.path.getsize(ply) - 97 * 14 * 4).reshape(97, 14)
    # the oracle's restatement against both directions
    enc = fo.ply_encode_rows({k: torch.from_numpy(rows[:, a:b].reshape(97, -1) if b - a > 1 else rows[:, a]).numpy()
                              for k, (a, b) in zip(GRAD_NAMES, ((0, 3), (3, 6), (6, 10), (10, 13), (13, 14)))})
    with np.errstate(all="ignore"):
        ok = np.isclose(enc, body, rtol=2e-6, atol=2e-6) | (np.isinf(enc) & np.isinf(body) & (np.sign(enc) == np.sign(body)))
    assert ok.all(), np.argwhere(~ok)
    dec = fo.ply_decode_rows(body)
    dec_rows = np.concatenate([np.asarray(dec[k]).reshape(97, -1) for k in GRAD_NAMES], axis=1)
    assert np.allclose(dec_rows, back, rtol=2e-6, atol=1e-7), np.abs(dec_rows - back).max()
    np.savez_compressed(os.path.join(GOLD, "cloud_97_ply.npz"), rows_in=rows, file_rows=body, rows_loaded=back)
    print(f"cloud_97.ply: {os.path.getsize(ply)} bytes written by the reference's save_ply, read back by its load_ply; "
          f"oracle encode / decode agree")


def fx_simplified():
    """SimplifiedRenderer: rotated look-at camera; Gaussians whose integer radius or integer pixel centre could
    flip under a one-ulp change are re-drawn (the fixture pins the arithmetic, not one rounding)."""
    W, H = 96, 80
    cam = fo.camera_from_pose(math.radians(15.0), math.radians(-30.0), W)
    cam.cx, cam.cy, cam.width, cam.height = W / 2, H / 2, W, H
    inp = fo.synthetic_cloud(900, seed=29, s_lo=0.02, s_hi=0.25)
    inp["positions"][:, 2] += 2.0
    inp["positions"][:40] *= 3.0                       # some behind the camera / outside the image
    g = torch.Generator().manual_seed(30)
    for _ in range(20):
        u, v, d = fo.camera_project(inp["positions"], cam)
        q = inp["scales"].mean(dim=1) * cam.fx / d
        def near_int(t, tol):
            return (t - t.round()).abs() < tol
        fragile = (d > 0) & (near_int(q, 2e-3) | near_int(u, 2e-3) | near_int(v, 2e-3) | (d.abs() < 0.02))
        idx = torch.nonzero(fragile).squeeze(1)
        if idx.numel() == 0:
            break
        p = torch.randn(idx.numel(), 3, generator=g) * 0.5
        inp["positions"][idx] = p
        inp["scales"][idx] = torch.rand(idx.numel(), 3, generator=g) * 0.23 + 0.02
    else:
        raise RuntimeError("could not settle the simplified fixture")
    run_simplified("simplified_900_96x80", inp, cam, W, H, (0.1, 0.15, 0.2),
                   note="SimplifiedRenderer, look-at camera el 15 az -30; gradients for positions, colours, opacities")


def fx_overlap():
    """>= 700 overlapping Gaussians per pixel ("re-check product-form transmittance at ~700 overlaps/px"):
    20,000 big Gaussians on a 128x128 image.  The reference keeps the SUM form of the accumulated alpha
    (DR:647-658: contribution = alpha * (1 - accumulated_alpha)); the kernels carry the transmittance as a product.
    Two opacity regimes: the standard U(0.1, 0.9) (the transmittance dies after a few dozen entries, so the tail
    probes the sum form's quantisation of 1 - A near 1) and a faint one (x 0.04: hundreds of entries contribute)."""
    W = H = 128
    cam = fo.default_camera(W)
    for name, seed, k in (("tile_overlap_20k_128", 43, 1.0), ("tile_overlap_faint_20k_128", 47, 0.04)):
        inp = fo.synthetic_cloud(20000, seed=seed, s_lo=0.05, s_hi=0.15)
        inp["opacities"] = inp["opacities"] * k
        inp = settle(inp, cam, W, H, 64, seed, redraw_std)
        run_tile(name, inp, cam, W, H, bg=(0.1, 0.2, 0.3),
                 note=f"20k Gaussians, scales U(0.05,0.15), opacities U(0.1,0.9)*{k}: ~1000 rectangle overlaps per pixel")
        add_f64_companion(name, "tile")


def zone_snap_cloud(n, seed, num_zones=8, s_lo=0.005, s_hi=0.03):
    """BASELINE configs[3] inputs: the synthetic cloud with the decoder-side Fresnel treatment applied -
    depth snapped to the centre of its zone by the REFERENCE's FresnelZones.get_zone_centers_for_depth
    (gaussian_decoder_models.py:833-841: base_z = offset + zone_centre * (-2)), i.e. only `num_zones` distinct depths
    (massive exact ties), and the edge-aware modulation of gaussian_decoder_models.py:881-895 (scales shrunk by up to
    50 %, opacity boosted by up to 0.2 and clamped) with a seeded per-Gaussian edge strength."""
    from utils.fresnel_zones import FresnelZones
    inp = fo.synthetic_cloud(n, seed=seed, s_lo=s_lo, s_hi=s_hi)
    zones = FresnelZones(num_zones=num_zones, depth_range=(0.0, 1.0))
    d01 = ((-inp["positions"][:, 2]) - 1.0) / 2.0                    # camera depth 1..3 -> 0..1
    zc = zones.get_zone_centers_for_depth(d01)
    inp["positions"][:, 2] = -1.0 + zc * (-2.0)
    g = torch.Generator().manual_seed(seed + 500)
    edge = torch.rand(n, generator=g) ** 2
    inp["scales"] = inp["scales"] * (1.0 - 0.5 * edge).unsqueeze(-1)
    inp["opacities"] = torch.clamp(inp["opacities"] + 0.2 * edge, 0, 1)
    return inp


def redraw_xy(inp, idx, g):
    inp["positions"][idx, :2] = torch.randn(idx.numel(), 2, generator=g) * 0.5


def fx_c4_small():
    """configs[3] at a size whose gradients the oracle clone restatement finishes quickly: 20k @ 256^2, 8 zones."""
    W = H = 256
    cam = fo.default_camera(W)
    inp = zone_snap_cloud(20000, 53)
    inp = settle(inp, cam, W, H, 64, 53, redraw_xy, allow_ties=True)
    assert np.unique(inp["positions"][:, 2].numpy()).size <= 8
    run_tile("c4_zones_phase_20k_256", inp, cam, W, H, phase=True, amp=0.25,
             note="configs[3] inputs (8 Fresnel depth zones => exact depth ties, edge-aware scales / opacities), "
                  "use_phase_blending=True; forward from the reference, gradients from the oracle clone restatement")
    run_tile("c4_zones_tile_20k_256", inp, cam, W, H, phase=False,
             note="same inputs without phase blending: forward and gradients from the reference")
    add_f64_companion("c4_zones_tile_20k_256", "tile")


def fx_c2():
    """BASELINE configs[1] at full size: 100,000 Gaussians, 512x512, fwd + bwd of the unmodified reference."""
    W = H = 512
    cam = fo.default_camera(W)
    inp = fo.synthetic_cloud(100000, seed=0)
    inp = settle(inp, cam, W, H, 64, 0, redraw_std)
    run_tile("c2_tile_100k_512", inp, cam, W, H, note="config 2 (BASELINE configs[1]) at full size")


def fx_c4():
    """BASELINE configs[3] at full size: 200,000 Gaussians, 512x512, 8 zones, edge-aware inputs, phase blending."""
    W = H = 512
    cam = fo.default_camera(W)
    inp = zone_snap_cloud(200000, 59)
    inp = settle(inp, cam, W, H, 64, 59, redraw_xy, allow_ties=True)
    run_tile("c4_zones_phase_200k_512", inp, cam, W, H, phase=True, amp=0.25,
             note="configs[3] at full size; forward from the reference (no gradients in this fixture: they are "
                  "pinned by c4_zones_phase_20k_256)", forward_only=True)


FIXTURES = dict(simplified=fx_simplified, bin=fx_bin, ply=fx_ply, dense=fx_dense, fourier=fx_fourier, culled=fx_culled, edge=fx_edge, rot=fx_rot, params=fx_params, phase=fx_phase, phase_rot=fx_phase_rot, wave=fx_wave,
                wave_rot=fx_wave_rot, asm=fx_asm, asm_rot=fx_asm_rot, asm_params=fx_asm_params, c1=fx_c1, overlap=fx_overlap,
                c4_small=fx_c4_small)
# full-size fixtures: tens of minutes of CPU each, only with --only
BIG = dict(c2=fx_c2, c4=fx_c4)

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default=None)
    ap.add_argument("--threads", type=int, default=os.cpu_count() or 1)
    a = ap.parse_args()
    torch.set_num_threads(a.threads)
    if a.only in BIG:
        BIG[a.only]()
    for k, f in FIXTURES.items():
        if a.only in (None, k):
            f()
