#!/usr/bin/env python3
"""Time the BASELINE.json configurations that are not the bench headline (one GPU, CUDA events).

    python tools/measure_configs.py [--out profiles/rN_configs.json]

C1 16k @256^2, C2 100k @512^2, C4 200k @512^2 phase blending, C5 1M @1024^2 ASM (per view), plus
WaveFieldRenderer at C2 size.  Synthetic clouds of SURVEY.md section 8d.  Reports ms for forward,
backward and the tile-instance count.
"""
import argparse
import json
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fresnel_b200  # noqa: E402
from fresnel_b200.renderer import StageTimer  # noqa: E402


def cloud(n, seed, s_lo, s_hi, phase_hi, dev):
    g = torch.Generator().manual_seed(seed)
    pos = torch.randn(n, 3, generator=g) * 0.5
    pos[:, 2] -= 2.0
    d = dict(positions=pos, scales=torch.rand(n, 3, generator=g) * (s_hi - s_lo) + s_lo,
             rotations=torch.randn(n, 4, generator=g), colors=torch.rand(n, 3, generator=g),
             opacities=torch.rand(n, generator=g) * 0.8 + 0.1, phases=torch.rand(n, generator=g) * phase_hi)
    return {k: v.to(dev).requires_grad_(True) for k, v in d.items()}


def time_it(fn_fwd, iters=5, warm=2):
    fwd, bwd = [], []
    for i in range(warm + iters):
        a, b, c = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        a.record()
        outs, grads = fn_fwd()
        b.record()
        torch.autograd.backward(outs, grads)
        c.record()
        torch.cuda.synchronize()
        if i >= warm:
            fwd.append(a.elapsed_time(b)); bwd.append(b.elapsed_time(c))
    return sum(fwd) / len(fwd), sum(bwd) / len(bwd)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--skip-c5", action="store_true")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    rows = []

    def run(name, res, n, s_lo, s_hi, kind, phase_hi=1.0, **kw):
        L = cloud(n, 0, s_lo, s_hi, phase_hi, dev)
        cam = fresnel_b200.Camera(0.8 * res, 0.8 * res, res / 2, res / 2, res, res)
        g = torch.Generator().manual_seed(1)
        gi = (torch.rand(3, res, res, generator=g) * 2 - 1).to(dev)
        gd = (torch.rand(res, res, generator=g) * 2 - 1).to(dev)
        args5 = (L["positions"], L["scales"], L["rotations"], L["colors"], L["opacities"])
        if kind == "tile":
            ren = fresnel_b200.TileBasedRenderer(res, res)
            f = lambda: (ren(*args5, cam, return_depth=True), (gi, gd))
        elif kind == "phase":
            ren = fresnel_b200.TileBasedRenderer(res, res, use_phase_blending=True, phase_amplitude=0.25)
            f = lambda: (ren(*args5, cam, return_depth=True, phases=L["phases"]), (gi, gd))
        elif kind == "wave":
            ren = fresnel_b200.WaveFieldRenderer(res, res)
            f = lambda: (ren(*args5, cam, return_depth=True, phases=L["phases"]), (gi, gd))
        else:
            ren = fresnel_b200.ASMWaveFieldRenderer(res, res, depth_range=(0.1, 4.0)).to(dev)
            wl = torch.tensor([0.0635, 0.05, 0.041])
            f = lambda: ((ren(*args5, cam, phases=L["phases"], wavelengths_rgb=wl),), (gi,))
        fwd, bwd = time_it(f)
        with StageTimer() as st:
            outs, grads = f()
            torch.autograd.backward(outs, grads)
        stages = {k: round(sum(v) / len(v), 4) for k, v in st.summary().items()}
        row = dict(config=name, renderer=kind, gaussians=n, resolution=res, fwd_ms=round(fwd, 4), bwd_ms=round(bwd, 4),
                   frames_per_s=round(1e3 / (fwd + bwd), 2), peak_mem_gb=round(torch.cuda.max_memory_allocated() / 2**30, 2),
                   stage_ms=stages)
        print(json.dumps(row), flush=True)
        rows.append(row)
        torch.cuda.reset_peak_memory_stats()

    run("C1", 256, 16384, 0.005, 0.03, "tile")
    run("C2", 512, 100000, 0.005, 0.03, "tile")
    run("C2-wave", 512, 100000, 0.005, 0.03, "wave", phase_hi=2 * math.pi)
    run("C4", 512, 200000, 0.005, 0.03, "phase")
    if not args.skip_c5:
        run("C5-view", 1024, 1000000, 0.002, 0.012, "asm", phase_hi=2 * math.pi)
        run("C5-tile", 1024, 1000000, 0.002, 0.012, "tile")
    if args.out:
        json.dump(rows, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
