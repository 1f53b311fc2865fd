"""CPU tier: the C-ABI library loads and exports every symbol include/fresnel_b200.h declares,
argument validation works without a GPU, and the Python boundary mirrors the reference's."""
import ctypes
import inspect
import os
import re

import numpy as np
import pytest
import torch

import fresnel_b200
from fresnel_b200 import _lib
from oracle import fresnel_oracle as fo

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "fresnel_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(frb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(cuda_lib):
    syms = header_symbols()
    assert len(syms) >= 15
    for s in syms:
        assert hasattr(cuda_lib, s), f"{s} declared in include/fresnel_b200.h but not exported"
    assert set(_lib.exported_symbols()) <= set(syms)
    assert cuda_lib.frb_version() >= 100


def test_argument_validation_needs_no_gpu(cuda_lib):
    cam = np.zeros(20, np.float32)
    # n not divisible by n_views, too many views, null pointers -> FRB_E_INVALID, never a crash
    assert cuda_lib.frb_project_fwd(3, 2, None, None, None, None, None, cam.ctypes.data, 64.0, None, None, None,
                                    None, None, None) == -1
    assert cuda_lib.frb_project_fwd(4, 64, None, None, None, None, None, cam.ctypes.data, 64.0, None, None, None,
                                    None, None, None) == -1
    cam[16] = cam[17] = 40000.0
    assert cuda_lib.frb_project_fwd(4, 1, None, None, None, None, None, cam.ctypes.data, 64.0, None, None, None,
                                    None, None, None) == -2
    assert cuda_lib.frb_radix_sort_pairs(-1, None, None, None, None, 0, 64, None, None) == -1
    assert cuda_lib.frb_composite_fwd(1, 0, 16, None, None, None, 0.0, None, 0.0, None, None, None, None, None,
                                      None, None) == -1
    assert b"invalid" in cuda_lib.frb_error_string(-1)
    assert cuda_lib.frb_sort_workspace_bytes(1 << 20) >= 256 * 4 * (1 << 20) // 2048


def test_depth_sort_mode_switch_is_a_host_side_setting(cuda_lib):
    """frb_depth_sort_in_cluster: -1 (environment) is the initial state, the call returns the previous mode, other
    values are refused and leave the mode alone."""
    L = cuda_lib
    assert L.frb_depth_sort_in_cluster(1) == -1
    assert L.frb_depth_sort_in_cluster(0) == 1
    assert L.frb_depth_sort_in_cluster(7) == -1          # FRB_E_INVALID
    assert L.frb_depth_sort_in_cluster(-2) == -1
    assert L.frb_depth_sort_in_cluster(-1) == 0          # unchanged by the refused calls; back to the initial state
    assert L.frb_depth_sort_in_cluster(-1) == -1


def test_no_cpu_fallback():
    r = fresnel_b200.TileBasedRenderer(32, 32)
    z = torch.zeros
    with pytest.raises(TypeError, match="CUDA only"):
        r(z(4, 3), z(4, 3), z(4, 4), z(4, 3), z(4), fresnel_b200.Camera(1, 1, 1, 1, 32, 32))


def test_signature_mirrors_reference():
    sig = inspect.signature(fresnel_b200.TileBasedRenderer.__init__)
    assert list(sig.parameters)[:7] == ["self", "image_width", "image_height", "background", "max_radius",
                                        "use_phase_blending", "phase_amplitude"]
    fwd = inspect.signature(fresnel_b200.TileBasedRenderer.forward)
    assert list(fwd.parameters)[:9] == ["self", "positions", "scales", "rotations", "colors", "opacities",
                                        "camera", "return_depth", "phases"]
    r = fresnel_b200.TileBasedRenderer(48, 32, background=(0.1, 0.2, 0.3))
    assert len(list(r.parameters())) == 0 and (r.width, r.height, r.max_radius) == (48, 32, 64)


def test_camera_mirror_matches_oracle_pose():
    import math
    a = fresnel_b200.create_camera_from_pose(math.radians(20), math.radians(35), 128)
    b = fo.camera_from_pose(math.radians(20), math.radians(35), 128)
    assert torch.equal(a.view_matrix, b.view_matrix)
    assert (a.fx, a.fy, a.cx, a.cy) == (b.fx, b.fy, b.cx, b.cy)
    v = fresnel_b200.camera_vector(a, 128, 128)
    assert v.shape == (20,) and v.dtype == np.float32 and v[16] == 128


@pytest.mark.parametrize("world", [1, 2, 3, 8, 16])
def test_peer_shard_partition_covers_the_buffer(cuda_lib, world):
    """frb_peer_shard_floats (host function of csrc/exchange.cu): the shards of the fused exchange + Adam kernel
    are float4-aligned, disjoint, in rank order, and cover every float (the < 4 float tail goes to the last rank);
    invalid arguments are refused, as is a launch with null pointers (no GPU needed for either)."""
    for n in (0, 1, 3, 4, 5, 1003, 4096, 15_000_000, 15_000_003):
        shards = [cuda_lib.frb_peer_shard_floats(world, r, n) for r in range(world)]
        assert sum(shards) == n, (world, n, shards)
        assert all(s >= 0 for s in shards)
        assert all(s % 4 == 0 for s in shards[:-1])
        per = ((n // 4) + world - 1) // world * 4
        assert all(s <= per + 3 for s in shards)
    assert cuda_lib.frb_peer_shard_floats(0, 0, 8) == -1
    assert cuda_lib.frb_peer_shard_floats(2, 2, 8) == -1
    assert cuda_lib.frb_peer_adam_step(2, 0, 16, None, None, None, None, None, None, 1e-3, 0.9, 0.999, 1e-8, 1.0,
                                       None) == -1
    assert cuda_lib.frb_peer_adam_step(17, 0, 16, None, None, None, None, None, None, 1e-3, 0.9, 0.999, 1e-8, 1.0,
                                       None) == -1


def build_c_example(out_path):
    """gcc build of examples/render_c_abi.c against the header and the library (no torch, no C++)."""
    import subprocess
    cmd = ["gcc", "-std=c99", "-O2", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"),
           "-I/usr/local/cuda/include", os.path.join(ROOT, "examples", "render_c_abi.c"), "-o", out_path,
           _lib.library_path(), "-L/usr/local/cuda/lib64", "-lcudart", "-lm",
           "-Wl,-rpath," + os.path.dirname(_lib.library_path()), "-Wl,-rpath,/usr/local/cuda/lib64"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return out_path


def test_header_is_plain_c_and_the_c_example_links(cuda_lib, tmp_path):
    """include/fresnel_b200.h is valid C99 on its own, and a plain C program using the whole-pass entry points
    compiles and links against the library (it is run by the GPU tier)."""
    import subprocess
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-fsyntax-only", "-x", "c",
                        os.path.join(ROOT, "include", "fresnel_b200.h")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    exe = build_c_example(str(tmp_path / "render_c_abi"))
    assert os.path.exists(exe)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 2 and "usage" in r.stderr            # argument check runs without a GPU


def test_host_sessions_refuse_cpu_devices():
    from fresnel_b200.host import HostRenderPipeline, HostRenderSession
    r = fresnel_b200.TileBasedRenderer(32, 32)
    with pytest.raises(TypeError, match="CUDA"):
        HostRenderSession(r, 16, torch.device("cpu"))
    with pytest.raises(TypeError, match="CUDA"):
        HostRenderPipeline(r, 16, torch.device("cpu"))
    with pytest.raises(ValueError):
        HostRenderPipeline(r, 16, torch.device("cpu"), depth=0)


def test_instance_capacity_rounding():
    """renderer.instance_capacity: never below m, at most ~12.5 % (or 65,536 entries) above it, monotone, and constant
    over the small drifts of an optimisation loop."""
    from fresnel_b200.renderer import instance_capacity
    assert instance_capacity(0) == 0 and instance_capacity(-5) == 0
    prev = 0
    for m in list(range(1, 2000, 37)) + [65535, 65536, 65537, 770_054, 5_820_000, 5_820_321, 10_000_000, 2**27 + 3]:
        c = instance_capacity(m)
        assert c >= m and c - m < max(65536, m // 8 + 1), (m, c)
        assert c >= prev or m < 2000
        prev = c
    assert instance_capacity(5_820_000) == instance_capacity(5_820_321) == instance_capacity(5_900_000)


def test_reference_arm_prints_the_contract_line():
    """``bench.py --impl reference`` (the oracle port timed on the host cores; no GPU involved): one JSON line with the
    keys the driver reads, the reference-arm extras, and a value that is a rate.  One bounded sample step, no full
    frame (--ref-full-budget 0), so the test stays under a minute."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "0", "--ref-full-budget", "0"], cwd=root, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "e2e", "cpu_baseline"):
        assert key in line, key
    assert line["impl"] == "reference" and line["unit"] == "frames/s" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["ms_per_step"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["config"]["full_frame_measured"] is False and "workload" in line["config"]
