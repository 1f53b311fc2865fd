"""Rendering with HOST-resident inputs and outputs: the copies overlap the kernels.

The reference's callers hold their tensors on the device (train_gaussian_decoder.py:1209-1223), but the
evaluation / data-generation scripts start from host data (``load_gaussians_from_binary`` DR:1461-1482,
generate_cvs_bootstrap_data.py:346-354).  ``HostRenderSession`` is that call with the transfers
scheduled on copy streams inside one step:

    step i:   [H2D parameters] -> forward ------------> backward ---------> [D2H gradients]
                                  [H2D upstream grads]  [D2H image, depth]

The forward only waits for the parameters, the upstream gradients arrive while it runs, and image / depth
leave while the backward runs.  Nothing of step i+1 starts before step i has ended (no cross-step
prefetch), so a per-step timing bracket on the compute stream contains every byte moved for that step.
The renderer is called through its ``nn.Module`` / autograd interface, i.e. the drop-in path.
"""

from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch

GRAD_NAMES = ("positions", "scales", "rotations", "colors", "opacities")
_SHAPES = {"positions": 3, "scales": 3, "rotations": 4, "colors": 3, "opacities": 0}


class HostRenderSession:
    """Forward + backward of one view per ``step`` with pinned host buffers on both sides.

    renderer: a fresnel_b200 renderer module with the reference call signature
    (positions, scales, rotations, colors, opacities, camera, return_depth=True).
    """

    def __init__(self, renderer, n_gaussians: int, device: torch.device):
        if device.type != "cuda":
            raise TypeError("HostRenderSession needs a CUDA device (fresnel_b200 has no CPU path)")
        self.renderer, self.n, self.device = renderer, int(n_gaussians), device
        h, w = renderer.height, renderer.width
        f32 = dict(dtype=torch.float32, device=device)
        self.dev_in = {k: torch.empty((self.n, c) if c else (self.n,), **f32) for k, c in _SHAPES.items()}
        self.dev_gimg = torch.empty(3, h, w, **f32)
        self.dev_gdep = torch.empty(h, w, **f32)
        self.out_image = torch.empty(3, h, w).pin_memory()
        self.out_depth = torch.empty(h, w).pin_memory()
        self.out_grads = {k: torch.empty((self.n, c) if c else (self.n,)).pin_memory() for k, c in _SHAPES.items()}
        self.s_in = torch.cuda.Stream(device)
        self.s_out = torch.cuda.Stream(device)
        mk = lambda: torch.cuda.Event()
        self.e_start, self.e_params, self.e_grads, self.e_fwd, self.e_bwd, self.e_out = (mk() for _ in range(6))
        self.h2d_bytes = 4 * (sum(t.numel() for t in self.dev_in.values()) + 4 * h * w)
        self.d2h_bytes = 4 * (sum(t.numel() for t in self.out_grads.values()) + 4 * h * w)

    def step(self, host_inputs: Dict[str, torch.Tensor], camera, g_image: torch.Tensor,
             g_depth: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, Dict[str, torch.Tensor]]:
        """host_inputs / g_image / g_depth: pinned host tensors.  Returns the pinned (image, depth, grads);
        they are complete once the current stream has caught up (``torch.cuda.current_stream().synchronize()``)."""
        main = torch.cuda.current_stream(self.device)
        self.e_start.record(main)
        self.s_in.wait_event(self.e_start)            # the previous step has finished with the input buffers
        with torch.cuda.stream(self.s_in):
            for k in GRAD_NAMES:
                self.dev_in[k].copy_(host_inputs[k], non_blocking=True)
            self.e_params.record(self.s_in)
            self.dev_gimg.copy_(g_image, non_blocking=True)
            self.dev_gdep.copy_(g_depth, non_blocking=True)
            self.e_grads.record(self.s_in)
        main.wait_event(self.e_params)
        leaves = {k: self.dev_in[k].detach().requires_grad_(True) for k in GRAD_NAMES}
        image, depth = self.renderer(leaves["positions"], leaves["scales"], leaves["rotations"], leaves["colors"],
                                     leaves["opacities"], camera, return_depth=True)
        self.e_fwd.record(main)
        self.s_out.wait_event(self.e_fwd)
        with torch.cuda.stream(self.s_out):
            self.out_image.copy_(image.detach(), non_blocking=True)
            self.out_depth.copy_(depth.detach(), non_blocking=True)
        image.record_stream(self.s_out)
        depth.record_stream(self.s_out)
        main.wait_event(self.e_grads)
        torch.autograd.backward((image, depth), (self.dev_gimg, self.dev_gdep))
        self.e_bwd.record(main)
        self.s_out.wait_event(self.e_bwd)
        with torch.cuda.stream(self.s_out):
            for k in GRAD_NAMES:
                g = leaves[k].grad
                self.out_grads[k].copy_(g, non_blocking=True)
                g.record_stream(self.s_out)
            self.e_out.record(self.s_out)
        main.wait_event(self.e_out)                   # the step ends when its last byte is on the host
        return self.out_image, self.out_depth, self.out_grads


class BatchPrefetcher:
    """Double-buffered host -> device staging of training batches (what a DataLoader with ``pin_memory`` and
    ``non_blocking`` copies does): ``next()`` returns the device copy of the batch submitted before and starts
    copying the following one on a side stream while the caller computes.  ``fence()`` makes the current stream
    wait for the copy in flight, so that a timing bracket closed after it contains the transfer."""

    def __init__(self, device: torch.device):
        self.device = device
        self.stream = torch.cuda.Stream(device)
        self.ready = torch.cuda.Event()
        self.staged: Optional[Tuple[torch.Tensor, ...]] = None

    def submit(self, host_batch) -> None:
        main = torch.cuda.current_stream(self.device)
        start = torch.cuda.Event()
        start.record(main)
        self.stream.wait_event(start)
        with torch.cuda.stream(self.stream):
            self.staged = tuple(t.to(self.device, non_blocking=True) for t in host_batch)
            self.ready.record(self.stream)

    def fence(self) -> None:
        torch.cuda.current_stream(self.device).wait_event(self.ready)

    def take(self):
        self.fence()
        batch, self.staged = self.staged, None
        main = torch.cuda.current_stream(self.device)
        for t in batch:
            t.record_stream(main)
        return batch
