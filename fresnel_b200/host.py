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


def _carve(block: torch.Tensor, n: int):
    """Views of the five parameter tensors inside one flat fp32 block; every segment starts 256-byte aligned
    (rotations are read as float4)."""
    out, off = {}, 0
    for k, c in _SHAPES.items():
        cnt = n * max(c, 1)
        out[k] = block[off:off + cnt].view((n, c) if c else (n,))
        off += (cnt + 63) // 64 * 64
    return out


def _block_floats(n: int) -> int:
    return sum((n * max(c, 1) + 63) // 64 * 64 for c in _SHAPES.values())


class HostRenderSession:
    """Forward + backward of one view per ``step`` with pinned host buffers on both sides.

    renderer: a fresnel_b200 renderer module with the reference call signature
    (positions, scales, rotations, colors, opacities, camera, return_depth=True).

    The caller writes the parameters into ``host_inputs[name]`` and the upstream gradients into
    ``host_g_image`` / ``host_g_depth`` (pinned staging buffers owned by the session: one block per
    direction, so a step moves its data in two H2D and three D2H copies), calls ``step(camera)`` and reads
    ``out_image``, ``out_depth`` and ``out_grads[name]`` after synchronising the current stream.
    """

    def __init__(self, renderer, n_gaussians: int, device: torch.device, outputs=("image", "depth", "grads")):
        """``outputs``: which results are copied back to the host every step (a training loop that keeps the image on
        the device asks for ("grads",): 5.6 MB instead of 9.8 MB per step at 100k Gaussians / 512x512)."""
        if device.type != "cuda":
            raise TypeError("HostRenderSession needs a CUDA device (fresnel_b200 has no CPU path)")
        bad = set(outputs) - {"image", "depth", "grads"}
        if bad:
            raise ValueError(f"unknown outputs {sorted(bad)}")
        self.outputs = tuple(outputs)
        self.renderer, self.n, self.device = renderer, int(n_gaussians), device
        n, h, w = self.n, renderer.height, renderer.width
        f32 = dict(dtype=torch.float32, device=device)
        blk = _block_floats(n)
        self._host_in_block = torch.zeros(blk).pin_memory()
        self._dev_in_block = torch.zeros(blk, **f32)
        self.host_inputs = _carve(self._host_in_block, n)
        self.dev_in = _carve(self._dev_in_block, n)
        self._host_g_block = torch.zeros(4 * h * w).pin_memory()
        self._dev_g_block = torch.zeros(4 * h * w, **f32)
        self.host_g_image = self._host_g_block[:3 * h * w].view(3, h, w)
        self.host_g_depth = self._host_g_block[3 * h * w:].view(h, w)
        self.dev_gimg = self._dev_g_block[:3 * h * w].view(3, h, w)
        self.dev_gdep = self._dev_g_block[3 * h * w:].view(h, w)
        self.out_image = torch.empty(3, h, w).pin_memory()
        self.out_depth = torch.empty(h, w).pin_memory()
        # gradients in the renderer's own order (rotations first), so that one copy can take all five
        self._out_g_block = torch.empty(14 * n).pin_memory()
        self.out_grads = {"rotations": self._out_g_block[:4 * n].view(n, 4),
                          "positions": self._out_g_block[4 * n:7 * n].view(n, 3),
                          "scales": self._out_g_block[7 * n:10 * n].view(n, 3),
                          "colors": self._out_g_block[10 * n:13 * n].view(n, 3),
                          "opacities": self._out_g_block[13 * n:]}
        self.s_in = torch.cuda.Stream(device)
        self.s_out = torch.cuda.Stream(device)
        mk = lambda: torch.cuda.Event()
        self.e_start, self.e_params, self.e_grads, self.e_fwd, self.e_bwd, self.e_out = (mk() for _ in range(6))
        self.h2d_bytes = 4 * (blk + 4 * h * w)
        self.d2h_bytes = 4 * ((14 * n if "grads" in self.outputs else 0) + (3 * h * w if "image" in self.outputs else 0)
                              + (h * w if "depth" in self.outputs else 0))

    def load(self, inputs: Dict[str, torch.Tensor], g_image: torch.Tensor, g_depth: torch.Tensor) -> None:
        """Host-side copy of caller tensors into the staging buffers (not needed if the caller writes there)."""
        for k in GRAD_NAMES:
            self.host_inputs[k].copy_(inputs[k])
        self.host_g_image.copy_(g_image)
        self.host_g_depth.copy_(g_depth)

    def step(self, camera) -> Tuple[torch.Tensor, torch.Tensor, Dict[str, torch.Tensor]]:
        """One forward + backward of what the staging buffers hold.  Returns the pinned (image, depth, grads);
        they are complete once the current stream has caught up (``torch.cuda.current_stream().synchronize()``)."""
        main = torch.cuda.current_stream(self.device)
        s_in, s_out, n = self.s_in, self.s_out, self.n
        self.e_start.record(main)
        s_in.wait_event(self.e_start)                 # the previous step has finished with the input buffers
        with torch.cuda.stream(s_in):
            self._dev_in_block.copy_(self._host_in_block, non_blocking=True)
            self.e_params.record(s_in)
            self._dev_g_block.copy_(self._host_g_block, non_blocking=True)
            self.e_grads.record(s_in)
        main.wait_event(self.e_params)
        leaves = {k: self.dev_in[k].detach().requires_grad_(True) for k in GRAD_NAMES}
        image, depth = self.renderer(leaves["positions"], leaves["scales"], leaves["rotations"], leaves["colors"],
                                     leaves["opacities"], camera, return_depth=True)
        self.e_fwd.record(main)
        s_out.wait_event(self.e_fwd)
        with torch.cuda.stream(s_out):
            if "image" in self.outputs:
                self.out_image.copy_(image.detach(), non_blocking=True)
            if "depth" in self.outputs:
                self.out_depth.copy_(depth.detach(), non_blocking=True)
        # no record_stream: the step ends with the compute stream waiting for the last copy (e_out), so memory
        # freed after the step cannot be reused before the copies have read it
        main.wait_event(self.e_grads)
        torch.autograd.backward((image, depth), (self.dev_gimg, self.dev_gdep))
        self.e_bwd.record(main)
        s_out.wait_event(self.e_bwd)
        grads = {k: leaves[k].grad for k in GRAD_NAMES}
        g_rot = grads["rotations"]
        # the tile renderer returns its five gradients as consecutive segments of one buffer
        # (renderer._TileRenderFusedFn.backward: rotations | positions | scales | colors | opacities)
        p0, packed = g_rot.data_ptr(), True
        for k, off in (("positions", 4 * n), ("scales", 7 * n), ("colors", 10 * n), ("opacities", 13 * n)):
            packed = packed and grads[k].data_ptr() == p0 + 4 * off
        with torch.cuda.stream(s_out):
            if "grads" not in self.outputs:
                pass
            elif packed:
                flat = torch.as_strided(g_rot, (14 * n,), (1,), g_rot.storage_offset())
                self._out_g_block.copy_(flat, non_blocking=True)
            else:
                for k in GRAD_NAMES:
                    self.out_grads[k].copy_(grads[k], non_blocking=True)
            self.e_out.record(s_out)
        main.wait_event(self.e_out)                   # the step ends when its last byte is on the host
        return self.out_image, self.out_depth, self.out_grads


class HostRenderPipeline:
    """``depth`` HostRenderSession steps in flight, each replayed from ONE CUDA graph.

    A step of HostRenderSession is a fixed DAG (two H2D copies, the forward and backward kernels, three D2H
    copies over three streams) on static buffers, so it captures into a graph as it is: a step costs the host one
    graph launch instead of ~0.4 ms of Python.  Slot k has its own pinned staging buffers, device buffers and
    stream; consecutive steps go to alternating slots, so the parameters of step i+1 cross PCIe while the
    kernels of step i run and the gradients of step i leave while step i+1 computes.  Every byte of every step
    still moves inside the bracket that times the steps; only the latency of one step is no longer the period.

    Protocol: ``slot = pipe.acquire()`` (blocks until the step that last used the slot has ended), write the
    inputs into ``pipe.slots[slot].host_inputs / host_g_image / host_g_depth``, ``pipe.submit(camera, slot)``,
    later ``pipe.wait(slot)`` and read ``pipe.slots[slot].out_image / out_depth / out_grads``.
    Graphs are cached per (slot, camera); a new camera costs one capture.
    """

    def __init__(self, renderer, n_gaussians: int, device: torch.device, depth: int = 2, cuda_graph: bool = True,
                 outputs=("image", "depth", "grads"), cluster_sort=None):
        if depth < 1:
            raise ValueError("depth must be >= 1")
        self.device, self.depth, self.cuda_graph = device, int(depth), bool(cuda_graph)
        # the depth order as ONE kernel on a 16-CTA cluster (frb_depth_sort_in_cluster): slower than the multi-kernel
        # chain for a frame alone, but it leaves 132 SMs to the other frames in flight - on when three or more are
        self.cluster_sort = (self.depth >= 3) if cluster_sort is None else bool(cluster_sort)
        self.slots = [HostRenderSession(renderer, n_gaussians, device, outputs=outputs) for _ in range(self.depth)]
        self.streams = [torch.cuda.Stream(device) for _ in range(self.depth)]
        self.done = [torch.cuda.Event() for _ in range(self.depth)]
        self._busy = [False] * self.depth
        self._graphs = [dict() for _ in range(self.depth)]
        self._next = 0
        self.h2d_bytes, self.d2h_bytes = self.slots[0].h2d_bytes, self.slots[0].d2h_bytes
        self.kernels_per_step = 0

    def _camera_key(self, camera):
        from .camera import camera_vector
        r = self.slots[0].renderer
        return camera_vector(camera, r.width, r.height).tobytes()

    def _graph(self, slot: int, camera):
        key = self._camera_key(camera)
        g = self._graphs[slot].get(key)
        if g is None:
            from . import _lib
            st, sess = self.streams[slot], self.slots[slot]
            st.wait_stream(torch.cuda.current_stream(self.device))
            previous = _lib.lib().frb_depth_sort_in_cluster(1) if self.cluster_sort else None
            try:
                with torch.cuda.stream(st):             # warm-up on the slot's stream: allocator, lazy attributes
                    for _ in range(2):
                        sess.step(camera)
                st.synchronize()
                g = torch.cuda.CUDAGraph()
                before = _lib.lib().frb_launch_count()
                with torch.cuda.graph(g, stream=st):
                    sess.step(camera)
                self.kernels_per_step = int(_lib.lib().frb_launch_count() - before)
            finally:
                if previous is not None:
                    _lib.lib().frb_depth_sort_in_cluster(previous)
            if len(self._graphs[slot]) >= 8:
                self._graphs[slot].clear()
            self._graphs[slot][key] = g
        return g

    def acquire(self) -> int:
        """Next slot in round-robin order, free for the caller to fill."""
        slot = self._next
        self._next = (slot + 1) % self.depth
        self.wait(slot)
        return slot

    def submit(self, camera, slot: int) -> None:
        st = self.streams[slot]
        if self.cuda_graph:
            g = self._graph(slot, camera)
            with torch.cuda.stream(st):
                g.replay()
                self.done[slot].record(st)
        else:
            with torch.cuda.stream(st):
                self.slots[slot].step(camera)
                self.done[slot].record(st)
        self._busy[slot] = True

    def wait(self, slot: int) -> None:
        if self._busy[slot]:
            self.done[slot].synchronize()
            self._busy[slot] = False

    def drain(self) -> None:
        for k in range(self.depth):
            self.wait(k)


class BatchPrefetcher:
    """Double-buffered host -> device staging of training batches (what a DataLoader with ``pin_memory`` and
    ``non_blocking`` copies does): ``take()`` returns the device copy of the batch submitted before, ``submit()``
    starts copying the following one on a side stream while the caller computes.  ``fence()`` makes the current
    stream wait for the copy in flight, so that a timing bracket closed after it contains the transfer.
    Two persistent sets of device buffers: no allocator traffic per step."""

    def __init__(self, device: torch.device):
        self.device = device
        self.stream = torch.cuda.Stream(device)
        self.ready = torch.cuda.Event()
        self.start = torch.cuda.Event()
        self.buffers = [None, None]
        self.slot = 0                   # slot the next submit writes
        self.staged: Optional[Tuple[torch.Tensor, ...]] = None

    def submit(self, host_batch) -> None:
        """The slot written here was handed out two submits ago; its reader must have been enqueued on the
        current stream before this call (it has: take() precedes the step that consumes the batch)."""
        if self.buffers[self.slot] is None:
            self.buffers[self.slot] = tuple(torch.empty(t.shape, dtype=t.dtype, device=self.device)
                                            for t in host_batch)
        dst = self.buffers[self.slot]
        self.start.record(torch.cuda.current_stream(self.device))
        self.stream.wait_event(self.start)
        with torch.cuda.stream(self.stream):
            for d, h in zip(dst, host_batch):
                d.copy_(h, non_blocking=True)
            self.ready.record(self.stream)
        self.staged = dst
        self.slot ^= 1

    def fence(self) -> None:
        torch.cuda.current_stream(self.device).wait_event(self.ready)

    def take(self):
        self.fence()
        batch, self.staged = self.staged, None
        return batch
