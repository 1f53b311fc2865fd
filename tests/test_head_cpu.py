"""CPU tier: fresnel_b200/csrc/frb_head.h (the decoder output head the CUDA kernels inline), compiled with g++ by
tests/host_shim.cpp, against the PyTorch restatement of the reference head (fresnel_b200/training.py, itself
checked against the reference module in test_training_cpu.py) and its autograd."""
import ctypes

import numpy as np
import torch
import torch.nn.functional as F

from fresnel_b200.training import rotation_6d_to_quaternion
from helpers import rel


def P(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def torch_head(raw, base, edge, esf, eob):
    """DirectPatchDecoder.forward tail (gaussian_decoder_models.py:844-895) on flat (n, 16) raw outputs."""
    pos = torch.stack([base[:, 0] + raw[:, 0] * 0.25, base[:, 1] + raw[:, 1] * 0.25, base[:, 2]], -1)
    scl = torch.clamp(F.softplus(torch.clamp(raw[:, 3:6], min=-10, max=20) + 1.0) * 0.15, min=1e-6, max=2.0)
    scl = scl * (1.0 - esf * edge)[:, None]
    rot = rotation_6d_to_quaternion(raw[:, 6:12])
    col = torch.sigmoid(raw[:, 12:15])
    opa = torch.clamp(torch.sigmoid(raw[:, 15]) + eob * edge, 0, 1)
    return torch.cat([pos, scl, rot, col, opa[:, None]], -1)


def cases():
    g = torch.Generator().manual_seed(11)
    n = 6000
    raw = torch.randn(n, 16, generator=g) * 2.0
    raw[:200, 3:6] = torch.randn(200, 3, generator=g) * 30.0          # scale clamps and the softplus threshold
    raw[200:260, 6:12] *= 1e-7                                         # tiny 6D vectors: the eps branches
    raw[260:300, 9:12] = raw[260:300, 6:9] * 1.5 + 1e-3 * torch.randn(40, 3, generator=g)   # a2 almost parallel to a1
    raw[300:340, 15] = torch.randn(40, generator=g) * 12.0             # saturated opacity
    base = torch.randn(n, 3, generator=g)
    edge0 = torch.zeros(n)
    edge1 = torch.rand(n, generator=g)
    return [(raw, base, edge0, 0.0, 0.0), (raw, base, edge1, 0.5, 0.9)]


def test_head_forward_matches_torch(host_shim):
    for raw, base, edge, esf, eob in cases():
        n = raw.shape[0]
        out = np.zeros((n, 14), np.float32)
        host_shim.shim_head_fwd(n, P(raw.numpy()), P(base.numpy()), P(edge.numpy()), ctypes.c_float(esf),
                                ctypes.c_float(eob), P(out))
        want = torch_head(raw, base, edge, esf, eob).numpy()
        hard = np.zeros(n, bool)
        hard[200:300] = True          # tiny / almost parallel 6D vectors amplify rounding by ~1e3
        assert np.allclose(out[~hard], want[~hard], rtol=2e-5, atol=2e-6), float(np.abs(out - want)[~hard].max())
        assert np.allclose(out[hard], want[hard], rtol=0, atol=2e-3), float(np.abs(out - want)[hard].max())


def test_head_degenerate_rotation_inputs_stay_finite(host_shim):
    """a2 exactly parallel to a1 (b2 is rounding noise + 1e-8, GM:206-209) and all-zero 6D vectors: the result is
    ill-conditioned by construction, so only finiteness and unit length are required - forward and backward."""
    g = torch.Generator().manual_seed(2)
    n = 64
    raw = torch.randn(n, 16, generator=g)
    raw[:32, 9:12] = raw[:32, 6:9] * 1.5
    raw[32:, 6:12] = 0.0
    base, edge = torch.zeros(n, 3), torch.zeros(n)
    out = np.zeros((n, 14), np.float32)
    host_shim.shim_head_fwd(n, P(raw.numpy()), P(base.numpy()), P(edge.numpy()), ctypes.c_float(0.0),
                            ctypes.c_float(0.0), P(out))
    assert np.isfinite(out).all()
    assert np.allclose(np.linalg.norm(out[:, 6:10], axis=1), 1.0, atol=1e-5)
    g_raw, g_z = np.zeros((n, 16), np.float32), np.zeros(n, np.float32)
    g_out = np.ones((n, 14), np.float32)
    host_shim.shim_head_bwd(n, P(raw.numpy()), P(edge.numpy()), ctypes.c_float(0.0), ctypes.c_float(0.0), P(g_out),
                            P(g_raw), P(g_z))
    assert np.isfinite(g_raw).all()


def test_head_backward_matches_autograd(host_shim):
    for raw, base, edge, esf, eob in cases():
        n = raw.shape[0]
        r = raw.clone().requires_grad_(True)
        b = base.clone().requires_grad_(True)
        out = torch_head(r, b, edge, esf, eob)
        g_out = torch.randn(n, 14, generator=torch.Generator().manual_seed(3))
        (out * g_out).sum().backward()
        g_raw = np.zeros((n, 16), np.float32)
        g_z = np.zeros(n, np.float32)
        host_shim.shim_head_bwd(n, P(raw.numpy()), P(edge.numpy()), ctypes.c_float(esf), ctypes.c_float(eob),
                                P(np.ascontiguousarray(g_out.numpy())), P(g_raw), P(g_z))
        want = r.grad.numpy()
        # rows where autograd itself is ill-conditioned (near-zero 6D vectors: gradients ~1e6) are compared relatively
        err = np.abs(g_raw - want) / np.maximum(np.abs(want).max(axis=1, keepdims=True), 1e-3)
        assert err.max() < 2e-3, (float(err.max()), int(err.max(axis=1).argmax()))
        assert np.median(err.max(axis=1)) < 1e-5
        assert rel(g_z, b.grad.numpy()[:, 2]) < 1e-6


def test_full_head_zone_snap_pose_rotation_and_edge_gradient(host_shim):
    """The head as head.cu sequences it (frb_zone_center -> frb_head_fwd_one -> frb_pose_rotate, and the backward with
    frb_pose_rotate_bwd and the edge-strength gradient) against the PyTorch restatement of
    gaussian_decoder_models.py:833-838, :51-104 / :860 and :881-895 and its autograd."""
    from fresnel_b200.training import rotate_positions_for_pose
    from fresnel_b200.zones import FresnelZones
    g = torch.Generator().manual_seed(21)
    n = 4000
    raw = torch.randn(n, 16, generator=g) * 1.5
    base_xy = torch.rand(n, 2, generator=g) * 2 - 1
    depth = torch.rand(n, generator=g) * 1.3 - 0.15            # some outside the zone range: clamped
    zones = FresnelZones(8, (0.0, 1.0))
    depth[:9] = zones.zone_boundaries                          # exactly on the boundaries (bucketize, right=False)
    edge = torch.rand(n, generator=g)
    el, az = torch.rand(n, generator=g) - 0.5, torch.rand(n, generator=g) * 6.28
    trig = torch.stack([torch.cos(az), torch.sin(az), torch.cos(el), torch.sin(el)], -1).contiguous()
    off, esf, eob = -2.0, 0.5, 0.9
    zb, zc = zones.zone_boundaries.numpy().copy(), zones.zone_centers.numpy().copy()

    def restatement(raw_t, edge_t, off_t):
        z = off_t + zones.get_zone_centers_for_depth(depth) * (-2)
        base = torch.cat([base_xy, z[:, None]], -1)
        out = torch_head(raw_t, base, edge_t, esf, eob)
        pos = rotate_positions_for_pose(out[:, None, :3], el, az)[:, 0]      # one "view" per Gaussian
        return torch.cat([pos, out[:, 3:]], -1)

    out = np.zeros((n, 14), np.float32)
    host_shim.shim_head_full_fwd(n, P(raw.numpy()), P(base_xy.numpy()), P(depth.numpy()), ctypes.c_float(off),
                                 P(edge.numpy()), ctypes.c_float(esf), ctypes.c_float(eob), 8, P(zb), P(zc),
                                 P(trig.numpy()), P(out))
    want = restatement(raw, edge, torch.tensor(off)).numpy()
    assert np.allclose(out, want, rtol=2e-5, atol=3e-6), float(np.abs(out - want).max())

    r, e, o = raw.clone().requires_grad_(True), edge.clone().requires_grad_(True), torch.tensor(off, requires_grad=True)
    g_out = torch.randn(n, 14, generator=torch.Generator().manual_seed(5))
    (restatement(r, e, o) * g_out).sum().backward()
    g_raw, g_z, g_e = np.zeros((n, 16), np.float32), np.zeros(n, np.float32), np.zeros(n, np.float32)
    host_shim.shim_head_full_bwd(n, P(raw.numpy()), P(edge.numpy()), ctypes.c_float(esf), ctypes.c_float(eob),
                                 P(trig.numpy()), P(np.ascontiguousarray(g_out.numpy())), P(g_raw), P(g_z), P(g_e))
    err = np.abs(g_raw - r.grad.numpy()) / np.maximum(np.abs(r.grad.numpy()).max(axis=1, keepdims=True), 1e-3)
    assert err.max() < 2e-3 and np.median(err.max(axis=1)) < 1e-5, float(err.max())
    assert rel(g_e, e.grad.numpy()) < 1e-5
    assert abs(float(g_z.astype(np.float64).sum()) - float(o.grad)) < 1e-3 * max(1.0, abs(float(o.grad)))
