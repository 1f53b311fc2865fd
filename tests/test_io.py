"""On-disk formats: the reference's 14-float .bin and the viewer's 3DGS .ply."""
import os

import numpy as np
import pytest
import torch

from oracle import fresnel_oracle as fo
import fresnel_b200
from fresnel_b200 import io as fio
from helpers import GRAD_NAMES

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_bin_loader_matches_reference_fixture():
    """cloud_97.bin was written by the reference's save_gaussians_to_binary; cloud_97_loaded.npz is what the
    reference's load_gaussians_from_binary returned for it."""
    got = fresnel_b200.load_gaussians_from_binary(os.path.join(GOLD, "cloud_97.bin"))
    want = np.load(os.path.join(GOLD, "cloud_97_loaded.npz"))
    assert set(got) == set(want.files)
    for k in want.files:
        assert got[k].dtype == torch.float32 and np.array_equal(got[k].numpy(), want[k]), k


def test_bin_save_cpu_reproduces_reference_bytes(tmp_path):
    src = os.path.join(GOLD, "cloud_97.bin")
    g = fresnel_b200.load_gaussians_from_binary(src)
    out = tmp_path / "again.bin"
    fresnel_b200.save_gaussians_to_binary(str(out), g)
    assert out.read_bytes() == open(src, "rb").read()


def test_bin_ragged_and_empty(tmp_path):
    """Trailing floats that do not fill a row are ignored (DR:1475-1476); an empty file gives N = 0."""
    raw = np.arange(14 * 3 + 5, dtype=np.float32)
    p = tmp_path / "ragged.bin"
    raw.tofile(p)
    g = fresnel_b200.load_gaussians_from_binary(str(p))
    assert g["positions"].shape == (3, 3) and float(g["opacities"][2]) == 14 * 2 + 13
    e = tmp_path / "empty.bin"
    e.write_bytes(b"")
    assert fresnel_b200.load_gaussians_from_binary(str(e))["positions"].shape == (0, 3)


def test_ply_header_parse_and_errors(tmp_path):
    rows = np.random.default_rng(0).standard_normal((5, 14)).astype(np.float32)
    p = tmp_path / "a.ply"
    hdr = "ply\r\nformat binary_little_endian 1.0\r\ncomment x\r\nelement vertex 5\r\n" + \
          "".join(f"property float {n}\r\n" for n in fio.PLY_PROPERTIES) + "end_header\n"
    p.write_bytes(hdr.encode() + rows.tobytes())
    assert np.array_equal(fio.read_ply_rows(str(p)), rows)           # CRLF tolerated (renderer.cpp:735-737)
    bad = tmp_path / "bad.ply"
    bad.write_bytes(b"ply\nelement vertex 0\nend_header\n")
    with pytest.raises(ValueError):
        fio.read_ply_rows(str(bad))
    short = tmp_path / "short.ply"
    short.write_bytes(hdr.encode() + rows.tobytes()[:-8])
    with pytest.raises(ValueError):
        fio.read_ply_rows(str(short))


def test_loaders_refuse_cpu_device(tmp_path):
    with pytest.raises(TypeError):
        fresnel_b200.load_gaussians_from_binary(os.path.join(GOLD, "cloud_97.bin"), device="cpu")


def test_oracle_ply_round_trip():
    inp = fo.synthetic_cloud(300, seed=4)
    g = {k: inp[k].numpy() for k in GRAD_NAMES}
    back = fo.ply_decode_rows(fo.ply_encode_rows(g))
    for k in GRAD_NAMES:
        assert np.allclose(back[k], g[k], rtol=2e-6, atol=2e-7), k


@pytest.mark.gpu
def test_bin_to_device_and_back(tmp_path):
    dev = torch.device("cuda:0")
    src = os.path.join(GOLD, "cloud_97.bin")
    want = np.load(os.path.join(GOLD, "cloud_97_loaded.npz"))
    g = fresnel_b200.load_gaussians_from_binary(src, device=dev)
    for k in want.files:
        assert g[k].is_cuda and np.array_equal(g[k].cpu().numpy(), want[k]), k
    out = tmp_path / "dev.bin"
    fresnel_b200.save_gaussians_to_binary(str(out), g)
    assert out.read_bytes() == open(src, "rb").read()
    img = fresnel_b200.TileBasedRenderer(64, 64)(g["positions"], g["scales"], g["rotations"], g["colors"],
                                                 g["opacities"], fresnel_b200.Camera(51.2, 51.2, 32, 32, 64, 64))
    assert img.shape == (3, 64, 64) and bool(torch.isfinite(img).all())


@pytest.mark.gpu
def test_ply_device_transforms_match_oracle(tmp_path):
    dev = torch.device("cuda:0")
    n = 100_003                                  # odd size, several grid-stride rounds
    inp = fo.synthetic_cloud(n, seed=6)
    g = {k: inp[k].to(dev) for k in GRAD_NAMES}
    p = tmp_path / "cloud.ply"
    fresnel_b200.save_gaussians_to_ply(str(p), g)
    rows = fio.read_ply_rows(str(p))
    want_rows = fo.ply_encode_rows({k: inp[k].numpy() for k in GRAD_NAMES})
    assert np.allclose(rows, want_rows, rtol=2e-6, atol=2e-6)
    back = fresnel_b200.load_gaussians_from_ply(str(p), device=dev)
    want = fo.ply_decode_rows(rows)
    for k in GRAD_NAMES:
        assert np.allclose(back[k].cpu().numpy(), want[k], rtol=2e-6, atol=1e-7), k
        assert np.allclose(back[k].cpu().numpy(), inp[k].numpy(), rtol=1e-5, atol=1e-6), k   # save -> load round trip


def _ply_fixture():
    z = np.load(os.path.join(GOLD, "cloud_97_ply.npz"))
    return z["rows_in"], z["file_rows"], z["rows_loaded"]


def _split(rows):
    return {"positions": rows[:, 0:3], "scales": rows[:, 3:6], "rotations": rows[:, 6:10], "colors": rows[:, 10:13],
            "opacities": rows[:, 13]}


def _close_or_same_inf(a, b, rtol, atol):
    with np.errstate(all="ignore"):
        return np.isclose(a, b, rtol=rtol, atol=atol) | (np.isinf(a) & np.isinf(b) & (np.sign(a) == np.sign(b)))


def test_oracle_ply_matches_reference_cpp_fixture():
    """The .ply pin: tests/golden/cloud_97.ply was written by the reference's own GaussianCloud::save_ply and read back
    by its load_ply (renderer.cpp:649-793, compiled by oracle/build_ref.sh).  The oracle's encode reproduces the file
    body, its decode the loaded values, including the 1e-7 scale floor, logit(0) = -inf, logit(1) and the colour clamp."""
    rows_in, file_rows, rows_loaded = _ply_fixture()
    body = fio.read_ply_rows(os.path.join(GOLD, "cloud_97.ply"))
    assert np.array_equal(body.view(np.uint32), file_rows.view(np.uint32))
    enc = fo.ply_encode_rows(_split(rows_in))
    assert _close_or_same_inf(enc, file_rows, 2e-6, 2e-6).all()
    dec = fo.ply_decode_rows(file_rows)
    for k, v in _split(rows_loaded).items():
        assert np.allclose(np.asarray(dec[k]).reshape(v.shape), v, rtol=2e-6, atol=1e-7), k


@pytest.mark.gpu
def test_ply_device_loader_and_saver_match_reference_cpp_fixture(tmp_path):
    """frb_unpack_gaussians / frb_pack_gaussians (ply mode) against the file the reference's C++ wrote and the values
    its C++ loaded."""
    dev = torch.device("cuda:0")
    rows_in, file_rows, rows_loaded = _ply_fixture()
    g = fresnel_b200.load_gaussians_from_ply(os.path.join(GOLD, "cloud_97.ply"), device=dev)
    for k, v in _split(rows_loaded).items():
        assert np.allclose(g[k].cpu().numpy().reshape(v.shape), v, rtol=2e-6, atol=1e-7), k
    out = tmp_path / "mine.ply"
    fresnel_b200.save_gaussians_to_ply(str(out), {k: torch.from_numpy(np.ascontiguousarray(v)).to(dev)
                                                  for k, v in _split(rows_in).items()})
    mine = fio.read_ply_rows(str(out))
    assert _close_or_same_inf(mine, file_rows, 2e-6, 2e-6).all(), np.argwhere(~_close_or_same_inf(mine, file_rows, 2e-6, 2e-6))
    # same header as the reference's writer, byte for byte
    ref_bytes = open(os.path.join(GOLD, "cloud_97.ply"), "rb").read()
    assert out.read_bytes()[:-97 * 14 * 4] == ref_bytes[:-97 * 14 * 4]
