"""CPU tier: fresnel_b200/csrc/frb_math.h (the arithmetic the CUDA kernels inline), compiled with
g++ by tests/host_shim.cpp, against the oracle: bit-exact projection / radius / visibility /
rectangles, and the hand-derived projection backward against the oracle's autograd."""
import ctypes
import math

import numpy as np
import pytest
import torch

from oracle import fresnel_oracle as fo
from fresnel_b200.camera import camera_vector
from helpers import golden_inputs, oracle_camera, rel


def P(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def shim_forward(shim, inp, camv, max_radius=64.0):
    n = inp["positions"].shape[0]
    p, s, q = (np.ascontiguousarray(inp[k].detach().numpy()) for k in ("positions", "scales", "rotations"))
    of = np.zeros((n, 11), np.float32)
    oi = np.zeros((n, 5), np.int32)
    shim.shim_project_fwd(n, P(p), P(s), P(q), P(camv), ctypes.c_float(max_radius), P(of), P(oi))
    return of, oi


def cases(golden):
    out = []
    W = H = 256
    out.append((fo.synthetic_cloud(4096, seed=0), fo.default_camera(W), W, H))
    cam = fo.camera_from_pose(math.radians(20.0), math.radians(35.0), 128)
    inp = fo.synthetic_cloud(2048, seed=3, s_lo=0.01, s_hi=0.06)
    inp["positions"][:, 2] += 2.0
    out.append((inp, cam, 128, 128))
    z = golden("tile_edge_1k_96x80")
    out.append((golden_inputs(z), oracle_camera(z["cam"], 96, 80), 96, 80))
    return out


def bits_equal(a, b):
    a = np.ascontiguousarray(a, np.float32)
    b = np.ascontiguousarray(b, np.float32)
    return bool(np.all((a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b))))


def test_projection_is_bit_exact(host_shim, golden):
    for inp, cam, W, H in cases(golden):
        camv = camera_vector(cam, W, H)
        of, oi = shim_forward(host_shim, inp, camv)
        pn = fo.pins(inp["positions"], inp["scales"], inp["rotations"], cam, W, H, 64)
        assert bits_equal(of[:, 0], pn["u"]) and bits_equal(of[:, 1], pn["v"])
        assert bits_equal(of[:, 2], pn["depth"])
        assert bits_equal(of[:, 3:7], pn["cov"])
        assert bits_equal(of[:, 7], pn["radius"])
        assert np.array_equal(oi[:, 0].astype(bool), pn["visible"])
        vi = pn["visible"]
        r = pn["rect"][vi].copy()
        r[(r[:, 0] >= r[:, 1]) | (r[:, 2] >= r[:, 3])] = 0
        assert np.array_equal(oi[vi, 1:], r)


def test_conic_matches_pinv(host_shim, golden):
    for inp, cam, W, H in cases(golden):
        camv = camera_vector(cam, W, H)
        of, oi = shim_forward(host_shim, inp, camv)
        pn = fo.pins(inp["positions"], inp["scales"], inp["rotations"], cam, W, H, 64)
        vi = pn["visible"]
        cov = torch.from_numpy(pn["cov"][vi]).reshape(-1, 2, 2)
        inv = torch.linalg.pinv(cov + 1e-4 * torch.eye(2).unsqueeze(0)).numpy()
        k = -0.5 * 1.4426950408889634
        ref = np.stack([inv[:, 0, 0], inv[:, 0, 1] + inv[:, 1, 0], inv[:, 1, 1]], 1) * k
        got = of[vi, 8:11]
        err = np.abs(got - ref) / (np.abs(ref).max(axis=1, keepdims=True) + 1e-30)
        assert err.max() < 2e-4, err.max()      # fp32 pinv (SVD) vs closed form on ill-conditioned cases
        assert np.median(err) < 1e-6


def test_projection_backward_matches_autograd(host_shim, golden):
    for inp, cam, W, H in cases(golden):
        n = inp["positions"].shape[0]
        camv = camera_vector(cam, W, H)
        L = {k: inp[k].clone().requires_grad_(True) for k in ("positions", "scales", "rotations")}
        pr = fo.project(L["positions"], L["scales"], L["rotations"], cam)
        rad = fo.compute_radius(pr["a"], pr["b"], pr["c"], pr["d"], 64)
        vis = fo.visibility(pr["u"], pr["v"], pr["depth"], rad, cam, W, H)
        cov = torch.stack([torch.stack([pr["a"], pr["b"]], -1), torch.stack([pr["c"], pr["d"]], -1)], -2)
        inv = torch.linalg.pinv(cov + 1e-4 * torch.eye(2).unsqueeze(0))
        k = -0.5 * 1.4426950408889634
        A, B, C = inv[:, 0, 0] * k, (inv[:, 0, 1] + inv[:, 1, 0]) * k, inv[:, 1, 1] * k
        g2d = torch.randn(n, 6, generator=torch.Generator().manual_seed(5)) * vis[:, None]
        loss = pr["u"] * g2d[:, 0] + pr["v"] * g2d[:, 1] + A * g2d[:, 2] + B * g2d[:, 3] + C * g2d[:, 4] \
            + pr["depth"] * g2d[:, 5]
        torch.where(vis, loss, torch.zeros_like(loss)).sum().backward()
        p, s, q = (np.ascontiguousarray(inp[k].numpy()) for k in ("positions", "scales", "rotations"))
        gp, gs, gq = np.zeros((n, 3), np.float32), np.zeros((n, 3), np.float32), np.zeros((n, 4), np.float32)
        g2 = np.ascontiguousarray(g2d.numpy())
        host_shim.shim_project_bwd(n, P(p), P(s), P(q), P(camv), P(g2), P(gp), P(gs), P(gq))
        m = vis.numpy()
        for name, mine in (("positions", gp), ("scales", gs), ("rotations", gq)):
            assert rel(mine[m], L[name].grad.numpy()[m]) < 1e-4, name


def test_fourier_mode_conic_and_backward(host_shim, golden):
    """FRB_MODE_FOURIER: A = C = -log2(e) / (2 sigma^2 + 1e-8), sigma^2 = (a + d)/2 + 1e-8 (DR:1677, 1725),
    the reference's visibility rule (DR:1647-1649), and the chain back to the 3D parameters."""
    for inp, cam, W, H in cases(golden):
        n = inp["positions"].shape[0]
        camv = camera_vector(cam, W, H)
        p, s, q = (np.ascontiguousarray(inp[k].numpy()) for k in ("positions", "scales", "rotations"))
        of, oi = np.zeros((n, 11), np.float32), np.zeros((n, 5), np.int32)
        host_shim.shim_project_fwd_mode(n, P(p), P(s), P(q), P(camv), ctypes.c_float(32000.0), 2, P(of), P(oi))
        L = {k: inp[k].clone().requires_grad_(True) for k in ("positions", "scales", "rotations")}
        pr = fo.project(L["positions"], L["scales"], L["rotations"], cam)
        vis = (pr["depth"] > cam.near) & (pr["depth"] < cam.far) & (pr["u"] > -W) & (pr["u"] < 2 * W) \
            & (pr["v"] > -H) & (pr["v"] < 2 * H)
        assert np.array_equal(oi[:, 0].astype(bool), vis.numpy())
        sigma = torch.sqrt((pr["a"] + pr["d"]) / 2 + 1e-8)
        A = -1.4426950408889634 / (2 * sigma ** 2 + 1e-8)
        m = vis.numpy()
        assert rel(of[m, 8], A.detach().numpy()[m]) < 1e-6 and np.all(of[:, 9] == 0) and np.array_equal(of[:, 8], of[:, 10])
        # support rectangle: 7.5 sigma + 1 around the centre, clipped to the image
        r = 7.5 * sigma.detach().numpy() + 1.0
        u, v = pr["u"].detach().numpy(), pr["v"].detach().numpy()
        inside = m & (u - r < W) & (u + r > 0) & (v - r < H) & (v + r > 0) & (oi[:, 2] > oi[:, 1])
        assert np.all(oi[inside, 1] <= np.maximum(0, np.floor(u[inside] - r[inside]) + 1))
        assert np.all(oi[inside, 2] >= np.minimum(W, np.floor(u[inside] + r[inside])))
        g2d = torch.randn(n, 6, generator=torch.Generator().manual_seed(7)) * vis[:, None]
        loss = pr["u"] * g2d[:, 0] + pr["v"] * g2d[:, 1] + A * g2d[:, 2] + A * g2d[:, 4] + pr["depth"] * g2d[:, 5]
        torch.where(vis, loss, torch.zeros_like(loss)).sum().backward()
        gp, gs, gq = np.zeros((n, 3), np.float32), np.zeros((n, 3), np.float32), np.zeros((n, 4), np.float32)
        g2 = np.ascontiguousarray(g2d.numpy())
        host_shim.shim_project_bwd_mode(n, P(p), P(s), P(q), P(camv), P(g2), 2, P(gp), P(gs), P(gq))
        for name, mine in (("positions", gp), ("scales", gs), ("rotations", gq)):
            assert rel(mine[m], L[name].grad.numpy()[m]) < 1e-4, name


def test_dense_mode_visibility_and_support(host_shim, golden):
    """FRB_MODE_DENSE: frustum + 100-pixel margin on the centre (DR:315-318); the rectangle contains every
    pixel where exp(-0.5 m) >= 2^-40, and the conic is the tile renderer's."""
    for inp, cam, W, H in cases(golden):
        n = inp["positions"].shape[0]
        camv = camera_vector(cam, W, H)
        p, s, q = (np.ascontiguousarray(inp[k].numpy()) for k in ("positions", "scales", "rotations"))
        of, oi = np.zeros((n, 11), np.float32), np.zeros((n, 5), np.int32)
        host_shim.shim_project_fwd_mode(n, P(p), P(s), P(q), P(camv), ctypes.c_float(32000.0), 1, P(of), P(oi))
        of0, _ = shim_forward(host_shim, inp, camv)
        assert np.array_equal(of[:, 8:11], of0[:, 8:11])
        pr = fo.project(inp["positions"], inp["scales"], inp["rotations"], cam)
        vis = (pr["depth"] > cam.near) & (pr["depth"] < cam.far) & (pr["u"] > -100) & (pr["u"] < W + 100) \
            & (pr["v"] > -100) & (pr["v"] < H + 100)
        assert np.array_equal(oi[:, 0].astype(bool), vis.numpy())
        # brute force on a subset: power (log2 units) outside the rectangle is below -40
        ys, xs = np.mgrid[0:H, 0:W].astype(np.float32)
        for i in np.nonzero(vis.numpy())[0][:64]:
            u, v, A, B, C = of[i, 0], of[i, 1], of[i, 8], of[i, 9], of[i, 10]
            power = A * (xs - u) ** 2 + B * (xs - u) * (ys - v) + C * (ys - v) ** 2
            x0, x1, y0, y1 = oi[i, 1:]
            outside = np.ones((H, W), bool)
            outside[y0:y1, x0:x1] = False
            assert not np.any(power[outside] > -40.0)
