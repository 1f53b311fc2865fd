"""CPU tier: the oracle restatement against the golden vectors generated from the reference
(oracle/make_golden.py), and the derived integer pins."""
import numpy as np
import pytest
import torch

from oracle import fresnel_oracle as fo
from helpers import GRAD_NAMES, golden_inputs, oracle_camera, rel

TILE_FIXTURES = ["tile_allculled_64", "tile_edge_1k_96x80", "tile_rotcam_2k_144x120", "tile_params_1500_96x64"]


@pytest.mark.parametrize("name", TILE_FIXTURES)
def test_oracle_tile_matches_reference_golden(golden, name):
    z = golden(name)
    W, H = int(z["W"]), int(z["H"])
    cam = oracle_camera(z["cam"], W, H)
    L = golden_inputs(z, requires_grad=True)
    img, dep, alpha = fo.render_tile_based(L["positions"], L["scales"], L["rotations"], L["colors"],
                                           L["opacities"], cam, W, H, background=tuple(z["bg"]),
                                           max_radius=int(z["max_radius"]))
    assert rel(img.detach(), z["image"]) < 1e-5
    assert rel(dep.detach(), z["depth"]) < 1e-5
    (img * torch.from_numpy(z["gimage"])).sum().add((dep * torch.from_numpy(z["gdepth"])).sum()).backward()
    for k in GRAD_NAMES:
        g = L[k].grad if L[k].grad is not None else torch.zeros_like(L[k])
        assert rel(g, z["grad_" + k]) < 1e-4, k


@pytest.mark.parametrize("name", TILE_FIXTURES + ["c1_tile_16k_256", "tile_phase_2k_128", "tile_phase_rot_1500_112x80"])
def test_oracle_pins_match_reference_pins(golden, name):
    """visible / rect / stable depth order derived from the REFERENCE's intermediates."""
    z = golden(name)
    W, H = int(z["W"]), int(z["H"])
    cam = oracle_camera(z["cam"], W, H)
    L = golden_inputs(z)
    pn = fo.pins(L["positions"], L["scales"], L["rotations"], cam, W, H, int(z["max_radius"]))
    assert np.array_equal(pn["visible"], z["visible"])
    vi = np.nonzero(z["visible"])[0]
    assert np.array_equal(pn["rect"][vi], z["rect"][vi])
    assert np.array_equal(pn["order"], z["order"])
    # keys are sorted, ranges partition them, every tile list is in stable depth order
    keys, gids, ranges = pn["keys"], pn["gids"], pn["ranges"]
    assert np.all(np.diff(keys.astype(np.uint64)) >= 0) if keys.size else True
    rank = np.empty(L["positions"].shape[0], np.int64)
    rank[pn["order"]] = np.arange(rank.shape[0])
    for s, e in ranges:
        assert np.all(np.diff(rank[gids[s:e]]) > 0)
    assert int(ranges[-1, 1]) == keys.shape[0] or keys.shape[0] == 0 or ranges[:, 1].max() == keys.shape[0]


@pytest.mark.parametrize("fixture", ["tile_phase_2k_128", "tile_phase_rot_1500_112x80"])
def test_oracle_phase_forward_matches_reference(golden, fixture):
    z = golden(fixture)
    W, H = int(z["W"]), int(z["H"])
    cam = oracle_camera(z["cam"], W, H)
    L = golden_inputs(z, with_phases=True)
    with torch.no_grad():
        img, dep, _ = fo.render_tile_based(L["positions"], L["scales"], L["rotations"], L["colors"],
                                           L["opacities"], cam, W, H, background=tuple(z["bg"]),
                                           use_phase_blending=True,
                                           phase_amplitude=float(z["phase_amplitude"]), phases=L["phases"])
    assert rel(img, z["image"]) < 1e-5
    assert rel(dep, z["depth"]) < 1e-5


@pytest.mark.parametrize("name", ["wave_scalar_2k_128", "wave_rgb_2k_128", "wave_rot_1500_112x80"])
def test_oracle_wave_matches_reference(golden, name):
    z = golden(name)
    W, H = int(z["W"]), int(z["H"])
    cam = oracle_camera(z["cam"], W, H)
    L = golden_inputs(z, with_phases=True)
    with torch.no_grad():
        img, dep = fo.render_wave(L["positions"], L["scales"], L["rotations"], L["colors"], L["opacities"],
                                  cam, W, H, L["phases"], background=tuple(z["bg"]))
    assert rel(img, z["image"]) < 1e-5
    assert rel(dep, z["depth"]) < 5e-5


@pytest.mark.parametrize("fixture", ["asm_1k_64", "asm_rot_1500_112x80", "asm_params_1200_80x64"])
def test_oracle_asm_matches_reference(golden, fixture):
    z = golden(fixture)
    W, H = int(z["W"]), int(z["H"])
    cam = oracle_camera(z["cam"], W, H)
    L = golden_inputs(z, with_phases=True)
    extra = {}
    if "num_depth_planes" in z:
        extra = dict(num_depth_planes=int(z["num_depth_planes"]), focal_depth=float(z["focal_depth"]),
                     pixel_pitch=float(z["pixel_pitch"]))
    with torch.no_grad():
        img, _ = fo.render_asm(L["positions"], L["scales"], L["rotations"], L["colors"], L["opacities"], cam,
                               W, H, L["phases"], torch.from_numpy(z["wavelengths"]),
                               background=tuple(z["bg"]), depth_range=tuple(z["depth_range"]), **extra)
    assert rel(img, z["image"]) < 1e-5


def test_oracle_dense_matches_reference(golden):
    """DifferentiableGaussianRenderer restatement against the reference's own output and gradients."""
    z = golden("dense_700_96x80")
    W, H = int(z["W"]), int(z["H"])
    cam = oracle_camera(z["cam"], W, H)
    L = golden_inputs(z, requires_grad=True)
    img, dep = fo.render_dense(L["positions"], L["scales"], L["rotations"], L["colors"], L["opacities"], cam, W, H,
                               background=tuple(z["bg"]))
    assert rel(img.detach(), z["image"]) < 1e-5
    assert rel(dep.detach(), z["depth"]) < 1e-5
    (img * torch.from_numpy(z["gimage"])).sum().add((dep * torch.from_numpy(z["gdepth"])).sum()).backward()
    for k in GRAD_NAMES:
        assert rel(L[k].grad, z["grad_" + k]) < 1e-4, k


def test_oracle_fourier_matches_reference(golden):
    """FourierGaussianRenderer restatement against the reference's own output and gradients."""
    z = golden("fourier_1500_96x80")
    W, H = int(z["W"]), int(z["H"])
    cam = oracle_camera(z["cam"], W, H)
    L = golden_inputs(z, requires_grad=True)
    img = fo.render_fourier(L["positions"], L["scales"], L["rotations"], L["colors"], L["opacities"], cam, W, H,
                            background=tuple(z["bg"]))
    assert rel(img.detach(), z["image"]) < 1e-5
    (img * torch.from_numpy(z["gimage"])).sum().backward()
    for k in GRAD_NAMES:
        g = L[k].grad if L[k].grad is not None else torch.zeros_like(L[k])
        assert rel(g, z["grad_" + k]) < 1e-4, k


def test_oracle_simplified_matches_reference(golden):
    """SimplifiedRenderer restatement against the reference's own forward output (the reference cannot
    backpropagate through this renderer: in-place writes on saved tensors)."""
    z = golden("simplified_900_96x80")
    W, H = int(z["W"]), int(z["H"])
    cam = oracle_camera(z["cam"], W, H)
    L = golden_inputs(z)
    with torch.no_grad():
        img, dep = fo.render_simplified(L["positions"], L["scales"], L["colors"], L["opacities"], cam, W, H,
                                        background=tuple(z["bg"]))
    assert rel(img, z["image"]) < 1e-5
    assert rel(dep, z["depth"]) < 1e-6
