"""Measured gradient margins of the tile renderer on every golden fixture that carries the reference's gradients:
max |a - b| / max(max |b|, 1e-3) per tensor (the quantity the parity tests bound by 1e-4), both t_eps.
    python tools/diag_gradients.py > gpurun_out/diag_gradients.json
"""
import glob
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import fresnel_b200  # noqa: E402
from test_gpu_parity import GRAD_NAMES, golden_inputs, oracle_camera, rel, render_gpu  # noqa: E402


def main():
    out = {}
    for path in sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "*.npz"))):
        z = np.load(path, allow_pickle=True)
        name = os.path.basename(path)[:-4]
        if "grad_positions" not in z.files or "gimage" not in z.files or not name.startswith(("tile_", "c2_", "c4_zones_tile")):
            continue
        if name.startswith("tile_phase"):
            continue
        W, H = int(z["W"]), int(z["H"])
        cam = oracle_camera(z["cam"], W, H)
        row = {"bg": [float(x) for x in z["bg"]]}
        for t_eps in (0.0, fresnel_b200.DEFAULT_T_EPS):
            try:
                img, dep, alpha, grads = render_gpu(golden_inputs(z), cam, W, H, tuple(float(x) for x in z["bg"]), t_eps,
                                                    int(z["max_radius"]), torch.from_numpy(z["gimage"]),
                                                    torch.from_numpy(z["gdepth"]))
            except Exception as e:  # a fixture of another renderer
                row[str(t_eps)] = f"skipped: {e}"
                continue
            row[str(t_eps)] = {k: rel(grads[k], z["grad_" + k]) for k in GRAD_NAMES}
            row[str(t_eps)]["image"] = rel(img, z["image"])
        out[name] = row
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
