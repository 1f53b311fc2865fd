#!/usr/bin/env python3
"""Benchmark of the hot path: differentiable Gaussian-splat render, forward + backward.

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched by torchrun)
    python bench.py --impl reference ...                     (CPU arm: the oracle port)

Workload (BASELINE.json configs[1]): 100,000 synthetic Gaussians, 512x512, one view per rank,
forward + backward with upstream gradients for image and depth.  A "step" is one frame.
At N > 1 every rank renders its own view of its own cloud (the path shards by view, no data-path
collective) -> weak scaling, value = frames of all ranks / max-over-ranks device time.

Prints ONE JSON line (rank 0).  See DESIGN.md section "Measurement" for the definitions of
value / e2e / roofline / cpu_baseline.
"""

from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "render fwd+bwd frames/sec at 512x512 (100k Gaussians)"
UNIT = "frames/s"
N_GAUSS = 100_000
RES = 512
GRAD_NAMES = ("positions", "scales", "rotations", "colors", "opacities")


def synthetic_cloud(n, seed):
    """SURVEY.md section 8d inputs (same generator as oracle.fresnel_oracle.synthetic_cloud)."""
    g = torch.Generator().manual_seed(seed)
    pos = torch.randn(n, 3, generator=g) * 0.5
    pos[:, 2] -= 2.0
    return dict(positions=pos,
                scales=torch.rand(n, 3, generator=g) * (0.03 - 0.005) + 0.005,
                rotations=torch.randn(n, 4, generator=g),
                colors=torch.rand(n, 3, generator=g),
                opacities=torch.rand(n, generator=g) * 0.8 + 0.1)


def upstream(seed=1):
    g = torch.Generator().manual_seed(seed)
    return torch.rand(3, RES, RES, generator=g) * 2 - 1, torch.rand(RES, RES, generator=g) * 2 - 1


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines, self.t_mark = index, None, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(self.index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def wait_first(self, timeout_s=8.0):
        """nvidia-smi needs up to seconds to start on an 8-GPU box: block until it has delivered one sample, so
        that short timed regions are still covered by the samples taken around them under the same load."""
        t0 = time.time()
        while self.proc is not None and not self.lines and time.time() - t0 < timeout_s:
            time.sleep(0.02)
        self.t_mark = time.time()            # samples from here on are "during the timed region"

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        inside = [ln for t, ln in self.lines if self.t_mark is not None and t >= self.t_mark]
        for ln in (inside or [ln for _, ln in self.lines]):
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, val in zip(names, f[4:8]):
                if val.lower() == "active":
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# --------------------------------------------------------------------------------------
# CPU arm: the oracle port (the reference is a Python module and cannot travel to the GPU box)
# --------------------------------------------------------------------------------------
def cpu_frame_time(n_sample, seed=0):
    """Seconds for one fwd+bwd frame of the first n_sample Gaussians of the workload, oracle port."""
    from oracle import fresnel_oracle as fo
    inp = synthetic_cloud(N_GAUSS, seed)
    L = {k: inp[k][:n_sample].clone().requires_grad_(True) for k in GRAD_NAMES}
    gi, gd = upstream()
    cam = fo.default_camera(RES)
    t0 = time.perf_counter()
    img, dep, _ = fo.render_tile_based(L["positions"], L["scales"], L["rotations"], L["colors"], L["opacities"],
                                       cam, RES, RES)
    torch.autograd.backward((img, dep), (gi, gd))
    return time.perf_counter() - t0


def calibrate_sample(budget_s):
    """Number of Gaussians whose fwd+bwd takes about budget_s on this host (second probe: the first pays
    one-off import / allocator costs; cost grows a little faster than linearly, hence the 0.7)."""
    cpu_frame_time(100)
    t_probe = cpu_frame_time(400)
    return int(max(400, min(N_GAUSS, 400 * budget_s / max(t_probe, 1e-3) * 0.7)))


def cpu_baseline(budget_s=20.0):
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    n_s = calibrate_sample(budget_s)
    t = cpu_frame_time(n_s)
    est = t * (N_GAUSS / n_s)
    return {"value": 1.0 / est, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
            "sample": (f"oracle port (reference is Python; restated in oracle/fresnel_oracle.py), first {n_s} of "
                       f"{N_GAUSS} Gaussians at {RES}x{RES}, fwd+bwd {t:.2f} s, scaled linearly in N "
                       "(favours the CPU: its backward cost grows faster than N)")}


def run_reference(args, rank, world):
    """Reference arm: the oracle port of TileBasedRenderer on the host cores (the reference is a Python module and
    does not exist on the GPU box).  The headline figure is a MEASUREMENT of the stated config: one full
    100,000-Gaussian 512x512 forward + backward, timed once in a child process (a few minutes; killed by exact PID
    if it exceeds --ref-full-budget seconds).  The K bounded-sample steps the contract describes still run and are
    reported in ``cpu_baseline.sample`` with their linear extrapolation, for comparison."""
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    total = args.steps + args.warmup
    per_step_budget = max(1.0, min(10.0, 60.0 / max(total, 1)))
    n_s = calibrate_sample(per_step_budget)
    for _ in range(args.warmup):
        cpu_frame_time(n_s)
    times = [cpu_frame_time(n_s) for _ in range(args.steps)]
    t = sum(times) / len(times)
    est = t * (N_GAUSS / n_s)
    sampled = (f"{args.steps} sampled steps of the first {n_s} of {N_GAUSS} Gaussians at {RES}x{RES}: fwd+bwd "
               f"{t:.2f} s/step, scaled linearly in N -> {est:.1f} s/frame ({1.0 / est:.5f} frames/s, extrapolated)")
    full_s, full_note = None, "not attempted (--ref-full-budget 0)"
    if args.ref_full_budget > 0:
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--ref-full-child"],
                               capture_output=True, text=True, timeout=args.ref_full_budget)
            for ln in r.stdout.splitlines():
                if ln.startswith("FULL_FRAME_SECONDS "):
                    full_s = float(ln.split()[1])
            full_note = "ok" if full_s is not None else f"child exited {r.returncode}: {r.stderr.strip()[-300:]}"
        except subprocess.TimeoutExpired:
            full_s, full_note = None, f"child killed after {args.ref_full_budget:.0f} s"
    if full_s is not None:
        value, ms_per_step, steps, warmup = 1.0 / full_s, full_s * 1e3, 1, 0
        how = (f"oracle port of TileBasedRenderer (reference is Python and absent from the GPU box): ONE full frame, "
               f"{N_GAUSS} Gaussians at {RES}x{RES}, fwd+bwd measured {full_s:.1f} s on {cores} host threads; " + sampled)
        measured = True
    else:
        value, ms_per_step, steps, warmup = 1.0 / est, t * 1e3, args.steps, args.warmup
        how = ("oracle port of TileBasedRenderer; the full frame did not finish inside the budget, so `value` is the "
               "EXTRAPOLATION and ms_per_step the measured sample step: " + sampled)
        measured = False
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "steps_requested": args.steps, "warmup_requested": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"single-view render fwd+bwd, {N_GAUSS} Gaussians, {RES}x{RES} (BASELINE configs[1])",
                   "full_frame_measured": measured, "full_frame_child": full_note},
        "extrapolated_from_sample": {"value": 1.0 / est, "unit": UNIT, "sample_gaussians": n_s,
                                     "sample_ms_per_step": t * 1e3},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                         "sample": how},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def run_reference_full_child():
    """Child of run_reference: one full frame of the port, prints its wall time."""
    import resource
    import threading
    torch.set_num_threads(os.cpu_count() or 1)
    n = int(os.environ.get("FRB_REF_FULL_N", N_GAUSS))      # (test hook: a smaller cloud)
    # The tape of the full frame is a chain of ~10^5 in-place slice updates; tearing it down recurses about that deep in
    # libtorch and overflows the default 8 MB stack (the child died with SIGSEGV on the GPU box): run the frame on a
    # thread with a 1 GiB stack, with the process limit lifted where the hard limit allows.
    try:
        soft, hard = resource.getrlimit(resource.RLIMIT_STACK)
        resource.setrlimit(resource.RLIMIT_STACK, (hard, hard))
    except (ValueError, OSError):
        pass
    out = {}

    def work():
        cpu_frame_time(200)                   # import / allocator warm-up
        out["t"] = cpu_frame_time(n)

    threading.stack_size(1 << 30)
    th = threading.Thread(target=work)
    th.start()
    th.join()
    print(f"FULL_FRAME_SECONDS {out['t']:.3f}", flush=True)
    os._exit(0)                               # skip interpreter teardown of whatever is left of the tape


# --------------------------------------------------------------------------------------
# Second workload (BASELINE.json configs[2]): data-parallel decoder training step, views/s
#   python bench.py --workload train [--gpus N via torchrun]      (fast_mode: 64x64, K=256)
#   python bench.py --workload train_full                         (256x256, all 5,476 Gaussians)
# --------------------------------------------------------------------------------------
TRAIN_B = 16


def train_batch(rank, seed=0):
    g = torch.Generator().manual_seed(seed + 1000 * rank)
    return (torch.randn(TRAIN_B, 384, 37, 37, generator=g), torch.rand(TRAIN_B, 1, 256, 256, generator=g),
            torch.rand(TRAIN_B, 3, 256, 256, generator=g))


def run_train_reference(args, rank, full):
    """CPU arm of the training step: same decoder on the CPU, oracle renderer called once per view
    (the reference's loop, train_gaussian_decoder.py:1209-1223)."""
    if rank != 0:
        return
    from oracle import fresnel_oracle as fo
    from fresnel_b200.training import PatchGaussianDecoder, reconstruction_losses, subsample_by_opacity
    import torch.nn.functional as F
    torch.set_num_threads(os.cpu_count() or 1)
    res, k = (256, None) if full else (64, 256)
    views = 2 if full else TRAIN_B                       # bounded sample: views per step
    torch.manual_seed(0)
    model = PatchGaussianDecoder(384, 4)
    opt = torch.optim.AdamW(model.parameters(), lr=1e-4)
    feats, depth, images = (t[:views] for t in train_batch(0))
    cam = fo.default_camera(res)

    def step():
        opt.zero_grad()
        g = subsample_by_opacity(model(feats, depth), k)
        imgs, deps = [], []
        for b in range(views):
            i, d, _ = fo.render_tile_based(g["positions"][b], g["scales"][b], g["rotations"][b], g["colors"][b],
                                           g["opacities"][b], cam, res, res)
            imgs.append(i); deps.append(d)
        tgt = F.interpolate(images, size=(res, res), mode="bilinear", align_corners=False)
        tdep = F.interpolate(depth, size=(res, res), mode="bilinear", align_corners=False).squeeze(1)
        reconstruction_losses(torch.stack(imgs), tgt, torch.stack(deps), tdep).backward()
        torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)
        opt.step()

    steps = max(1, min(args.steps, 3 if full else 10))
    for _ in range(min(args.warmup, 1)):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    value = views / dt
    sample = (f"oracle port, {views} views per step at {res}x{res}, "
              f"{'all 5476' if full else 256} Gaussians per view, {steps} timed steps, {dt:.2f} s/step")
    print(json.dumps({
        "impl": "reference", "metric": TRAIN_METRIC, "value": value, "unit": "views/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": min(args.warmup, 1), "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": train_workload_name(full)},
        "cpu_baseline": {"value": value, "unit": "views/s", "cores": torch.get_num_threads(), "kind": "port",
                         "sample": sample},
        "e2e": {"value": value, "unit": "views/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


TRAIN_METRIC = "decoder training views/sec (experiment 2, B=16 views per GPU)"


def train_workload_name(full):
    return ("Gaussian decoder training step, experiment 2, DINOv2-small-shaped synthetic features, batch of 16 views "
            "per GPU, " + ("256x256, all 5476 Gaussians per view" if full else
                           "fast_mode: 64x64 render, 256 stochastic Gaussians per view") +
            " (BASELINE configs[2])")


def init_distributed(local, world):
    """One NCCL process group per bench process (torchrun supplies the rendezvous variables)."""
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product has no CPU path; use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    return dev


def _timed_steps(fn, steps, flush):
    """Per-step CUDA events on the launching (current) stream; L2 flushed (256 MiB write) before every step."""
    evs = []
    for _ in range(steps):
        flush.fill_(1.0)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        evs.append((a, b))
    torch.cuda.synchronize()
    return [a.elapsed_time(b) for a, b in evs]


def _barrier(world):
    import torch.distributed as dist
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def _max_over_ranks(values, dev, world):
    import torch.distributed as dist
    t = torch.tensor(values, dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.tolist()


def measure_train(args, rank, world, dev, full, steps, flush, sampler=None):
    """BASELINE configs[2]: data-parallel decoder training step (train_gaussian_decoder.py:1209-1266 with ONE batched
    render call), 16 views per GPU, decoder gradients averaged by one flat NCCL all-reduce per step.  Returns the
    result dictionary (identical on every rank: times are the max over ranks)."""
    from fresnel_b200 import _lib
    from fresnel_b200.host import BatchPrefetcher
    from fresnel_b200.training import DecoderTrainer, PatchGaussianDecoder
    L = _lib.lib()
    torch.manual_seed(0)
    model = PatchGaussianDecoder(384, 4).to(dev)
    res, k = (256, None) if full else (64, 256)
    trainer = DecoderTrainer(model, res, stochastic_k=k, seed=rank, cuda_graph=not args.no_cuda_graph)
    trainer.broadcast_parameters()
    host = [t.pin_memory() for t in train_batch(rank)]
    resident = [t.to(dev) for t in host]
    loss_host = torch.empty((), dtype=torch.float32).pin_memory()

    def step_resident():
        trainer.step(*resident)

    prefetch = BatchPrefetcher(dev)
    prefetch.submit(host)

    def step_e2e():
        # double-buffered input staging: this step's batch was copied while the previous step computed, the
        # next batch is copied (side stream) while this one computes; the fence puts that copy inside this
        # step's timing bracket, so K timed steps contain K batch copies.
        f, d, i = prefetch.take()
        prefetch.submit(host)
        loss_host.copy_(trainer.step(f, d, i), non_blocking=True)
        prefetch.fence()

    def exchange_only():
        trainer.exchange()

    warm = max(args.warmup, 3)
    for _ in range(warm):
        step_resident()
    step_e2e()
    _barrier(world)
    if sampler is not None:
        sampler.wait_first()
    l0 = L.frb_launch_count()
    _barrier(world)
    ms = _timed_steps(step_resident, steps, flush)
    _barrier(world)
    launches = L.frb_launch_count() - l0
    if trainer.cuda_graph:          # replayed kernels do not pass through the library's launch counter
        launches = trainer.kernels_per_replay * steps
    ms_e2e = _timed_steps(step_e2e, steps, flush)
    _barrier(world)
    ms_x = _timed_steps(exchange_only, max(5, min(steps, 20)), flush)      # the collective alone (0 at one rank)
    _barrier(world)
    tot_ms, tot_e2e, x_ms = _max_over_ranks([sum(ms), sum(ms_e2e), statistics.median(ms_x)], dev, world)
    h2d = sum(t.numel() * 4 for t in host)
    n_par = sum(p.numel() for p in model.parameters())
    step_ms = tot_ms / steps
    return {
        "metric": TRAIN_METRIC, "value": world * TRAIN_B * steps / (tot_ms * 1e-3), "unit": "views/s",
        "n_gpus": world, "steps": steps, "warmup": warm, "ms_per_step": step_ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": train_workload_name(full), "views_per_gpu": TRAIN_B,
                   "decoder_parameters": n_par, "cuda_graph": not args.no_cuda_graph,
                   "l2": "flushed between steps (256 MiB fill), per-step CUDA events summed",
                   "parallelism": f"dp{world}: view batch sharded by rank, one flat NCCL all-reduce of the "
                                  "decoder gradients per step"},
        "e2e": {"value": world * TRAIN_B * steps / (tot_e2e * 1e-3), "unit": "views/s",
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4, "ms_per_step": tot_e2e / steps},
        "exchange": {"kind": "one NCCL all-reduce (SUM) of the flat decoder gradient buffer" if world > 1
                             else "none (one rank)",
                     "bytes": 4 * n_par, "ms": x_ms, "share_of_step": x_ms / step_ms,
                     "note": "the collective timed alone (the pack / unpack kernels are inside the two captured "
                             "halves of the step), median, max over ranks"},
        "gpu_launches": int(launches)}


def run_train(args, rank, world, local, full):
    import torch.distributed as dist
    dev = init_distributed(local, world)
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    line = measure_train(args, rank, world, dev, full, args.steps, flush, sampler if rank == 0 else None)
    if rank == 0:
        line["clocks"] = sampler.stop()
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


# --------------------------------------------------------------------------------------
# Third workload (BASELINE.json configs[4]): multi-pose optimisation of ONE cloud, ASM renderer, one view per rank
#   python bench.py --workload multiview [--gpus N via torchrun] [--mv-gaussians 1000000 --mv-res 1024]
# --------------------------------------------------------------------------------------
MV_METRIC = "multi-pose novel-view training views/sec (1M Gaussians, 1024x1024, ASMWaveFieldRenderer)"
MV_WAVELENGTHS = (0.0635, 0.05, 0.041)


def mv_cloud(n, seed=0):
    import math
    g = torch.Generator().manual_seed(seed)
    return dict(positions=torch.randn(n, 3, generator=g) * 0.5,             # centred: the cameras orbit the origin
                scales=torch.rand(n, 3, generator=g) * (0.012 - 0.002) + 0.002,
                rotations=torch.randn(n, 4, generator=g), colors=torch.rand(n, 3, generator=g),
                opacities=torch.rand(n, generator=g) * 0.8 + 0.1,
                phases=torch.rand(n, generator=g) * 2 * math.pi)


def mv_workload_name(n, res):
    return (f"multi-pose novel-view optimisation of one cloud: {n} Gaussians at {res}x{res}, ASMWaveFieldRenderer "
            "(16 planes, depth_range (0.1, 4.0), RGB wavelengths), look-at poses az = 45 deg * rank, one view per "
            "GPU per step, per-Gaussian gradients summed over ranks (BASELINE configs[4])")


def measure_multiview(args, rank, world, dev, exchange, steps, flush, sampler=None, stages=False):
    """BASELINE configs[4]: multi-pose optimisation of one replicated cloud, ASM renderer, one view per rank per step;
    the per-Gaussian gradients are summed over the ranks by the fused peer-memory exchange + Adam kernel
    (``exchange='peer'``) or by NCCL all-reduce + replicated Adam (``'nccl'``)."""
    import math
    import fresnel_b200
    from fresnel_b200 import _lib
    from fresnel_b200.host import BatchPrefetcher
    from fresnel_b200.renderer import StageTimer
    from fresnel_b200.training import MultiViewTrainer, allreduce_flat
    L = _lib.lib()
    n, res = args.mv_gaussians, args.mv_res
    ren = fresnel_b200.ASMWaveFieldRenderer(res, res, depth_range=(0.1, 4.0)).to(dev)
    wl = torch.tensor(MV_WAVELENGTHS)
    trainer = MultiViewTrainer(ren, mv_cloud(n), dev, lr=1e-4, with_phases=True, render_kwargs=dict(wavelengths_rgb=wl),
                               exchange=exchange)
    cam = fresnel_b200.create_camera_from_pose(0.0, math.radians(45.0 * rank), res)
    g = torch.Generator().manual_seed(100 + rank)
    target_host = torch.rand(3, res, res, generator=g).pin_memory()
    target = target_host.to(dev)
    loss_host = torch.empty((), dtype=torch.float32).pin_memory()

    def step_resident():
        trainer.step(cam, target)

    prefetch = BatchPrefetcher(dev)
    prefetch.submit((target_host,))

    def step_e2e():
        # double-buffered target staging, as in the decoder-training e2e: this step's target was copied while the
        # previous step computed, the next one is copied while this step computes; the fence puts that copy inside
        # this step's bracket, so K timed steps contain K target copies and K loss read-backs.
        (tgt,) = prefetch.take()
        prefetch.submit((target_host,))
        loss_host.copy_(trainer.step(cam, tgt), non_blocking=True)
        prefetch.fence()

    def exchange_only():
        # the exchange (and the update it is fused with) alone, on whatever the gradient buffer holds
        if exchange == "peer":
            trainer.optimizer.step()
        else:
            allreduce_flat(trainer.params.flat.grad)
            trainer.optimizer.step()

    warm = max(args.warmup, 3)
    for _ in range(warm):
        step_resident()
    step_e2e()
    _barrier(world)
    if sampler is not None:
        sampler.wait_first()
    l0 = L.frb_launch_count()
    _barrier(world)
    mem0 = torch.cuda.memory_stats(dev)
    ms = _timed_steps(step_resident, steps, flush)
    mem1 = torch.cuda.memory_stats(dev)
    _barrier(world)
    launches = L.frb_launch_count() - l0
    ms_e2e = _timed_steps(step_e2e, steps, flush)
    _barrier(world)
    ms_x = _timed_steps(exchange_only, max(5, min(steps, 20)), flush)
    _barrier(world)
    stage_ms = None
    if stages:
        with StageTimer() as st:
            _timed_steps(step_resident, max(3, min(steps, 10)), flush)
        stage_ms = {k_: sum(v) / len(v) for k_, v in st.summary().items()}
        _barrier(world)
    tot_ms, tot_e2e, x_ms = _max_over_ranks([sum(ms), sum(ms_e2e), statistics.median(ms_x)], dev, world)
    n_f = trainer.params.flat.numel()
    step_ms = tot_ms / steps
    line = {
        "allocator": {k_: [mem0.get(k_, 0), mem1.get(k_, 0)] for k_ in
                      ("segment.all.allocated", "segment.all.freed", "num_alloc_retries", "num_device_alloc",
                       "num_device_free", "reserved_bytes.all.current")},
        "metric": MV_METRIC, "value": world * steps / (tot_ms * 1e-3), "unit": "views/s", "n_gpus": world,
        "steps": steps, "warmup": warm, "ms_per_step": step_ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": mv_workload_name(n, res), "gradient_floats_exchanged": n_f,
                   "l2": "flushed between steps (256 MiB fill), per-step CUDA events summed",
                   "exchange": exchange,
                   "parallelism": (f"dp{world}: one view per rank, cloud replicated, per-Gaussian gradients "
                                   "summed and Adam applied by ONE kernel per rank over NVLink peer memory "
                                   "(reduce-scatter by peer loads, sharded Adam, all-gather by peer stores; "
                                   "csrc/exchange.cu)" if exchange == "peer" else
                                   f"dp{world}: one view per rank, cloud replicated, one flat NCCL all-reduce "
                                   "(SUM) of the per-Gaussian gradients per step, replicated fused Adam")},
        "e2e": {"value": world * steps / (tot_e2e * 1e-3), "unit": "views/s",
                "h2d_bytes_per_step": target_host.numel() * 4, "d2h_bytes_per_step": 4,
                "ms_per_step": tot_e2e / steps},
        "exchange": {"kind": ("frb_peer_adam_step: reduce-scatter by peer loads + sharded Adam + all-gather by peer "
                              "stores, one kernel" if exchange == "peer" else
                              "NCCL all-reduce (SUM) + replicated fused torch Adam"),
                     "bytes": 4 * n_f, "ms": x_ms, "share_of_step": x_ms / step_ms,
                     "link_gbs_per_direction": (4 * n_f * (world - 1) / world / (x_ms * 1e-3) / 1e9) if world > 1 else 0.0,
                     "note": "exchange + optimiser update timed alone, median, max over ranks; link figure = bytes a "
                             "rank pulls (reduce-scatter) = bytes it pushes (all-gather) over that time"},
        "step_ms": {"min": min(ms), "median": statistics.median(ms), "max": max(ms),
                    **({"all": [round(x, 3) for x in ms]} if os.environ.get("FRB_BENCH_ALL_STEPS") else {})},
        "gpu_launches": int(launches)}
    if stage_ms:
        peak, peak_src = peaks()
        hw = res * res
        m_est = None
        alg = {   # algorithmic bytes per launch (FFT stage = fields in + out, single-iFFT form)
            "frb_asm_propagate_fwd": 16 * 3 * hw * 8 * 2 + 16 * 3 * hw * 8 + 3 * hw * 8 * 3 + 12 * hw,
            "frb_asm_propagate_bwd": 16 * 3 * hw * 8 * 2 + 16 * 3 * hw * 8 + 3 * hw * 8 * 3 + 12 * hw,
        }
        top = max(stage_ms, key=stage_ms.get)
        fft = "frb_asm_propagate_fwd"
        line["roofline"] = {
            "bound": "hbm", "kernel": fft, "achieved": alg[fft] / (stage_ms[fft] * 1e-3) / 1e9, "peak": peak,
            "unit": "GB/s", "frac": alg[fft] / (stage_ms[fft] * 1e-3) / 1e9 / peak, "traffic": None,
            "peak_source": peak_src, "algorithmic_bytes_per_launch": alg[fft], "kernel_ms": stage_ms[fft],
            "slowest_stage": top,
            "note": "the bandwidth-shaped stage of this workload (batched cuFFT over 16 planes x 3 channels, fused "
                    "transfer-function multiply + plane sum, one inverse FFT per channel); the splat kernels are "
                    "issue-bound (DESIGN.md section 4); stage times from an instrumented eager pass of the same step",
            "stage_ms": {k_: round(v, 4) for k_, v in sorted(stage_ms.items(), key=lambda kv: -kv[1])},
            "stage_gbs": {k_: round(alg[k_] / (v * 1e-3) / 1e9, 1) for k_, v in stage_ms.items() if k_ in alg}}
        del m_est
    del trainer
    torch.cuda.empty_cache()
    return line


def run_multiview(args, rank, world, local):
    import torch.distributed as dist
    dev = init_distributed(local, world)
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    line = measure_multiview(args, rank, world, dev, args.exchange, args.steps, flush,
                             sampler if rank == 0 else None, stages=True)
    if rank == 0:
        line["clocks"] = sampler.stop()
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def run_multiview_reference(args, rank):
    """CPU arm: the oracle ASM renderer, forward + backward of one view, at two bounded sample sizes; the cost is
    fitted as a + b * n (the FFT part does not depend on n) and extrapolated to the full cloud."""
    if rank != 0:
        return
    import math
    from oracle import fresnel_oracle as fo
    torch.set_num_threads(os.cpu_count() or 1)
    n, res = args.mv_gaussians, args.mv_res
    cloud = mv_cloud(n)
    cam = fo.camera_from_pose(0.0, 0.0, res)
    wl = torch.tensor(MV_WAVELENGTHS)
    target = torch.rand(3, res, res, generator=torch.Generator().manual_seed(100))

    def one(n_s):
        Lc = {k: cloud[k][:n_s].clone().requires_grad_(True) for k in cloud}
        t0 = time.perf_counter()
        img, _ = fo.render_asm(Lc["positions"], Lc["scales"], Lc["rotations"], Lc["colors"], Lc["opacities"], cam,
                               res, res, Lc["phases"], wl, depth_range=(0.1, 4.0))
        torch.nn.functional.l1_loss(img, target).backward()
        return time.perf_counter() - t0

    one(50)
    n1, n2 = 200, 800
    t1, t2 = one(n1), one(n2)
    b = max((t2 - t1) / (n2 - n1), 0.0)
    a = max(t1 - b * n1, 0.0)
    est = a + b * n
    value = 1.0 / est
    sample = (f"oracle port of ASMWaveFieldRenderer, one view at {res}x{res}: {n1} Gaussians {t1:.2f} s, {n2} Gaussians "
              f"{t2:.2f} s, fitted {a:.2f} s + {b * 1e3:.3f} ms per Gaussian, extrapolated to {n}")
    print(json.dumps({
        "impl": "reference", "metric": MV_METRIC, "value": value, "unit": "views/s", "n_gpus": args.gpus, "steps": 1,
        "warmup": 1, "ms_per_step": est * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": {"workload": mv_workload_name(n, res)},
        "cpu_baseline": {"value": value, "unit": "views/s", "cores": torch.get_num_threads(), "kind": "port",
                         "sample": sample},
        "e2e": {"value": value, "unit": "views/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


# --------------------------------------------------------------------------------------
# Fourth workload (BASELINE.json configs[3]): Fresnel-zone + edge-aware + phase-blending render, fwd + bwd
# --------------------------------------------------------------------------------------
C4_METRIC = "Fresnel-zone + edge-aware + phase-blending render fwd+bwd frames/sec (8 zones, 200k Gaussians, 512x512)"
C4_N = 200_000


def c4_cloud(n, seed):
    """configs[3] inputs (same recipe as oracle/make_golden.py zone_snap_cloud): the synthetic cloud with the
    decoder-side Fresnel treatment applied - depth snapped to the centre of one of 8 zones
    (gaussian_decoder_models.py:833-841: only eight distinct depths, i.e. massive exact ties for the stable order),
    scales shrunk / opacities boosted by a seeded edge strength (:881-895) - plus a phase per Gaussian."""
    from fresnel_b200.zones import FresnelZones
    c = synthetic_cloud(n, seed)
    zones = FresnelZones(num_zones=8, depth_range=(0.0, 1.0))
    d01 = ((-c["positions"][:, 2]) - 1.0) / 2.0
    c["positions"][:, 2] = -1.0 + zones.get_zone_centers_for_depth(d01) * (-2.0)
    g = torch.Generator().manual_seed(seed + 500)
    edge = torch.rand(n, generator=g) ** 2
    c["scales"] = c["scales"] * (1.0 - 0.5 * edge).unsqueeze(-1)
    c["opacities"] = torch.clamp(c["opacities"] + 0.2 * edge, 0, 1)
    c["phases"] = torch.rand(n, generator=g)
    return c


def measure_phase(args, rank, world, dev, steps, flush):
    """configs[3]: TileBasedRenderer(use_phase_blending=True) on the zone-snapped, edge-modulated cloud; one view per
    rank, forward + backward with upstream gradients for image and depth (six gradients: the phases too)."""
    import fresnel_b200
    from fresnel_b200 import _lib
    from fresnel_b200.renderer import StageTimer
    L = _lib.lib()
    ren = fresnel_b200.TileBasedRenderer(RES, RES, use_phase_blending=True, phase_amplitude=0.25)
    cam = fresnel_b200.Camera(0.8 * RES, 0.8 * RES, RES / 2, RES / 2, RES, RES)
    cloud = {k: v.to(dev).requires_grad_(True) for k, v in c4_cloud(C4_N, seed=rank).items()}
    gi_h, gd_h = upstream(1 + rank)
    gi, gd = gi_h.to(dev), gd_h.to(dev)

    def step():
        for v in cloud.values():
            v.grad = None
        img, dep = ren(cloud["positions"], cloud["scales"], cloud["rotations"], cloud["colors"], cloud["opacities"],
                       cam, return_depth=True, phases=cloud["phases"])
        torch.autograd.backward((img, dep), (gi, gd))

    warm = max(args.warmup, 3)
    for _ in range(warm):
        step()
    _barrier(world)
    l0 = L.frb_launch_count()
    ms = _timed_steps(step, steps, flush)
    launches = L.frb_launch_count() - l0
    _barrier(world)
    with StageTimer() as st:
        _timed_steps(step, max(3, min(steps, 10)), flush)
    stage_ms = {k: sum(v) / len(v) for k, v in st.summary().items()}
    (tot_ms,) = _max_over_ranks([sum(ms)], dev, world)
    with torch.no_grad():
        depths = torch.unique(cloud["positions"][:, 2]).numel()
    return {"metric": C4_METRIC, "value": world * steps / (tot_ms * 1e-3), "unit": "frames/s", "n_gpus": world,
            "steps": steps, "warmup": warm, "ms_per_step": tot_ms / steps, "higher_is_better": True, "scaling": "weak",
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"phase-blending render fwd+bwd, {C4_N} Gaussians, {RES}x{RES}, 8 Fresnel depth "
                                   "zones + edge-aware scales / opacities applied to the inputs (BASELINE configs[3]); "
                                   "one view per rank, eager calls (one 4-byte device->host read per frame sizes the "
                                   "instance buffers)",
                       "distinct_depths": int(depths),
                       "l2": "flushed between steps (256 MiB fill), per-step CUDA events summed"},
            "stage_ms": {k: round(v, 4) for k, v in sorted(stage_ms.items(), key=lambda kv: -kv[1])},
            "gpu_launches": int(launches)}


RENDER_BATCH_VIEWS = 4


def measure_render_batch(args, rank, world, dev, steps, flush):
    """The headline render (100k Gaussians @ 512x512, forward + backward) as ``render_batch`` makes it: four clouds of
    that size with one camera each in ONE pass of the kernels (the call a training step makes,
    train_gaussian_decoder.py:1209-1223).  The latency-bound stages of the chain - depth order, tile counting, scan -
    are paid once for the four frames."""
    import fresnel_b200
    from fresnel_b200 import _lib
    L = _lib.lib()
    B = RENDER_BATCH_VIEWS
    ren = fresnel_b200.TileBasedRenderer(RES, RES)
    cam = fresnel_b200.Camera(0.8 * RES, 0.8 * RES, RES / 2, RES / 2, RES, RES)
    clouds = [synthetic_cloud(N_GAUSS, seed=rank * B + i) for i in range(B)]
    batch = {k: torch.stack([c[k] for c in clouds]).to(dev).requires_grad_(True) for k in clouds[0]}
    ups = [upstream(1 + rank * B + i) for i in range(B)]
    gi = torch.stack([u[0] for u in ups]).to(dev)
    gd = torch.stack([u[1] for u in ups]).to(dev)

    def step():
        for v in batch.values():
            v.grad = None
        img, dep, _ = ren.render_batch(batch["positions"], batch["scales"], batch["rotations"], batch["colors"],
                                       batch["opacities"], cam)
        torch.autograd.backward((img, dep), (gi, gd))

    warm = max(args.warmup, 3)
    for _ in range(warm):
        step()
    _barrier(world)
    l0 = L.frb_launch_count()
    ms = _timed_steps(step, steps, flush)
    launches = L.frb_launch_count() - l0
    (tot_ms,) = _max_over_ranks([sum(ms)], dev, world)
    return {"metric": "render fwd+bwd frames/sec (100k Gaussians, 512x512), four frames per call",
            "value": world * B * steps / (tot_ms * 1e-3), "unit": "frames/s", "n_gpus": world, "steps": steps,
            "warmup": warm, "ms_per_step": tot_ms / steps, "frames_per_step": B, "higher_is_better": True,
            "scaling": "weak", "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"TileBasedRenderer.render_batch: {B} clouds of {N_GAUSS} Gaussians, {RES}x{RES}, "
                                   "one pass of the kernels, forward + backward, inputs resident",
                       "l2": "flushed between steps (256 MiB fill), per-step CUDA events summed"},
            "gpu_launches": int(launches)}


# --------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--t-eps", type=float, default=None)
    ap.add_argument("--workload", default="render", choices=["render", "train", "train_full", "multiview", "phase"])
    ap.add_argument("--mv-gaussians", type=int, default=1_000_000)
    ap.add_argument("--mv-res", type=int, default=1024)
    ap.add_argument("--pipeline-depth", type=int, default=6, help="render e2e: host-to-host steps in flight")
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"],
                    help="multiview: fused peer-memory exchange + Adam kernel, or NCCL all-reduce + torch Adam")
    ap.add_argument("--no-cuda-graph", action="store_true", help="train workloads: run the step eagerly")
    ap.add_argument("--ref-full-budget", type=float, default=540.0,
                    help="reference arm: seconds allowed for the one full-size frame (0 = sampled steps only)")
    ap.add_argument("--ref-full-child", action="store_true", help=argparse.SUPPRESS)
    ap.add_argument("--no-workloads", action="store_true",
                    help="render: skip the `workloads` block (decoder training and multi-view at the same N)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.ref_full_child:
        run_reference_full_child()
        return
    if args.workload == "multiview":
        if args.impl == "reference":
            run_multiview_reference(args, rank)
        else:
            run_multiview(args, rank, world, local)
        return
    if args.workload == "phase":
        if args.impl == "reference":
            raise SystemExit("bench.py: --workload phase has no CPU arm (use the render / train / multiview arms)")
        import torch.distributed as dist
        dev = init_distributed(local, world)
        flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
        line = measure_phase(args, rank, world, dev, args.steps, flush)
        if rank == 0:
            print(json.dumps(line))
        if world > 1:
            dist.destroy_process_group()
        return
    if args.workload != "render":
        full = args.workload == "train_full"
        if args.impl == "reference":
            run_train_reference(args, rank, full)
        else:
            run_train(args, rank, world, local, full)
        return
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch.distributed as dist
    import fresnel_b200
    from fresnel_b200 import _lib
    from fresnel_b200.renderer import FusedStageTimer

    dev = init_distributed(local, world)
    L = _lib.lib()
    t_eps = fresnel_b200.DEFAULT_T_EPS if args.t_eps is None else args.t_eps
    ren = fresnel_b200.TileBasedRenderer(RES, RES, t_eps=t_eps)
    cam = fresnel_b200.Camera(0.8 * RES, 0.8 * RES, RES / 2, RES / 2, RES, RES)

    host = synthetic_cloud(N_GAUSS, seed=rank)            # each rank: its own view / cloud
    gi_h, gd_h = upstream(1 + rank)
    resident = {k: v.to(dev).requires_grad_(True) for k, v in host.items()}
    gi, gd = gi_h.to(dev), gd_h.to(dev)
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)   # > 126 MB L2

    def step_resident():
        for v in resident.values():
            v.grad = None
        img, dep = ren(resident["positions"], resident["scales"], resident["rotations"], resident["colors"],
                       resident["opacities"], cam, return_depth=True)
        torch.autograd.backward((img, dep), (gi, gd))

    from fresnel_b200.host import HostRenderSession
    session = HostRenderSession(ren, N_GAUSS, dev)
    session.load(host, gi_h, gd_h)          # the caller's data sits in the session's pinned staging buffers

    def step_e2e():
        # the module called on pinned host buffers: H2D of the five parameter tensors and both upstream
        # gradients, D2H of image, depth and the five gradients, all inside the step (copy streams overlap the
        # kernels within the step; nothing is prefetched across steps)
        session.step(cam)

    h2d, d2h = session.h2d_bytes, session.d2h_bytes

    # Throughput form of the same call: pipeline_depth HostRenderSession steps in flight, each replayed from one CUDA
    # graph (fresnel_b200/host.py HostRenderPipeline).  Every step still copies its inputs up and its results down
    # inside the bracket; step i+1's H2D and step i's D2H overlap the other step's kernels.
    from fresnel_b200.host import HostRenderPipeline
    pipe = pipe_full = HostRenderPipeline(ren, N_GAUSS, dev, depth=args.pipeline_depth)
    for s_ in pipe.slots:
        s_.load(host, gi_h, gd_h)

    # the same pipeline returning the gradients only (image and depth stay on the device: what a training loop needs)
    pipe_lean = HostRenderPipeline(ren, N_GAUSS, dev, depth=args.pipeline_depth, outputs=("grads",))
    for s_ in pipe_lean.slots:
        s_.load(host, gi_h, gd_h)

    def timed_pipeline(steps, with_flush=True, pipe=None):
        pipe = pipe_full if pipe is None else pipe
        return _timed_pipeline(pipe, steps, with_flush)

    def _timed_pipeline(pipe, steps, with_flush=True):
        """ONE bracket around ``steps`` pipelined steps (they overlap, so per-step brackets would double count).
        The 256 MiB L2 flush is enqueued on the slot's stream before every step and is INSIDE the bracket."""
        main = torch.cuda.current_stream()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(main)
        for st_ in pipe.streams:
            st_.wait_event(a)
        for _ in range(steps):
            slot = pipe.acquire()
            if with_flush:
                with torch.cuda.stream(pipe.streams[slot]):
                    flush.fill_(1.0)
            pipe.submit(cam, slot)
        for st_ in pipe.streams:
            main.wait_stream(st_)
        b.record(main)
        torch.cuda.synchronize()
        pipe.drain()
        return a.elapsed_time(b)

    enqueue = {}

    def timed(fn, steps):
        """Per-step CUDA events on the launching stream; L2 flushed (256 MiB write) between steps."""
        evs = []
        t0 = time.perf_counter()
        for _ in range(steps):
            flush.fill_(1.0)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            evs.append((a, b))
        enqueue[fn.__name__] = (time.perf_counter() - t0) / steps * 1e3   # host time to enqueue one step
        torch.cuda.synchronize()
        return [a.elapsed_time(b) for a, b in evs]

    def barrier():
        _barrier(world)

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()                      # started before the warm-up: see ClockSampler.wait_first
    for _ in range(args.warmup):
        step_resident()
    for _ in range(max(2, args.warmup // 2)):
        step_e2e()
    barrier()

    # The resident step is replayed from a CUDA graph (one launch per frame instead of ~20 kernel launches and
    # ~0.4 ms of Python): the forward is sync-free and allocation-static, so module call + autograd backward
    # capture as they are.  --no-cuda-graph times the eager calls.
    kernels_per_step = None
    step_timed = step_resident
    if not args.no_cuda_graph:
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):
                step_resident()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        k0 = L.frb_launch_count()
        with torch.cuda.graph(graph):
            step_resident()
        kernels_per_step = int(L.frb_launch_count() - k0)

        def step_graph():
            graph.replay()
        step_timed = step_graph
        for _ in range(3):
            step_timed()
        torch.cuda.synchronize()

    if rank == 0:
        sampler.wait_first()
    launches0 = L.frb_launch_count()
    barrier()
    ms = timed(step_timed, args.steps)
    barrier()
    launches = L.frb_launch_count() - launches0
    if kernels_per_step is not None:         # replayed kernels do not pass through the library's launch counter
        launches = kernels_per_step * args.steps
    for _ in range(3):                       # the caching allocator re-settles after the graph took its pool
        step_e2e()
    barrier()
    ms_e2e = timed(step_e2e, args.steps)
    barrier()
    timed_pipeline(2 * pipe.depth)           # captures the slot graphs
    barrier()
    pipe_ms = timed_pipeline(args.steps)
    barrier()
    pipe_noflush_ms = timed_pipeline(args.steps, with_flush=False)
    barrier()
    timed_pipeline(2 * pipe_lean.depth, pipe=pipe_lean)
    barrier()
    pipe_lean_ms = timed_pipeline(args.steps, pipe=pipe_lean)
    barrier()

    # Host-link ceiling of the e2e figure, measured here with every rank copying at once (the same pinned buffers,
    # H2D and D2H concurrently on two streams, no kernels): frames/s the PCIe / host-memory side could deliver if
    # the GPU cost nothing.  On 8-GPU VMs this, not the GPU, bounds e2e.
    def link_probe(reps=8):
        s_up, s_dn = torch.cuda.Stream(), torch.cuda.Stream()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sess = pipe.slots[0]
        outs = (sess._out_g_block, sess.out_image, sess.out_depth)
        srcs = (torch.empty(14 * N_GAUSS, device=dev), torch.empty(3, RES, RES, device=dev),
                torch.empty(RES, RES, device=dev))
        torch.cuda.synchronize()
        barrier()
        a.record()
        s_up.wait_event(a); s_dn.wait_event(a)
        for _ in range(reps):
            with torch.cuda.stream(s_up):
                sess._dev_in_block.copy_(sess._host_in_block, non_blocking=True)
                sess._dev_g_block.copy_(sess._host_g_block, non_blocking=True)
            with torch.cuda.stream(s_dn):
                for o, src in zip(outs, srcs):
                    o.copy_(src, non_blocking=True)
        torch.cuda.current_stream().wait_stream(s_up)
        torch.cuda.current_stream().wait_stream(s_dn)
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps

    link_probe(2)
    link_ms = min(link_probe(16) for _ in range(3))      # best of three: a ceiling, not an average
    barrier()
    clocks = sampler.stop() if rank == 0 else None

    # per-stage timing for the roofline line: the SAME fused entry points as `value` (events recorded inside
    # frb_tile_render_fwd / _bwd between the stages), eager calls, L2 flushed between steps
    with FusedStageTimer() as st:
        timed(step_resident, args.steps)
    stages = {k: sum(v) / len(v) for k, v in st.summary().items()}

    tot = _max_over_ranks([sum(ms), sum(ms_e2e), pipe_ms, pipe_noflush_ms, link_ms, pipe_lean_ms], dev, world)
    tot_ms, tot_e2e_ms, tot_pipe_ms, tot_pipe_nf_ms, link_ms, tot_lean_ms = tot
    lean_d2h = pipe_lean.d2h_bytes

    # the workloads north_star asks scaling numbers for, measured in this same invocation at this same N
    workloads = {}
    pipe_cluster_sort = pipe_full.cluster_sort
    if not args.no_workloads:
        del pipe, pipe_full, pipe_lean, session
        torch.cuda.empty_cache()
        w_steps = max(5, min(args.steps, 20))
        workloads["train"] = measure_train(args, rank, world, dev, False, args.steps, flush)
        torch.cuda.empty_cache()
        workloads["phase"] = measure_phase(args, rank, world, dev, w_steps, flush)
        torch.cuda.empty_cache()
        workloads["render_batch"] = measure_render_batch(args, rank, world, dev, w_steps, flush)
        torch.cuda.empty_cache()
        workloads["multiview"] = measure_multiview(args, rank, world, dev, "peer", w_steps, flush, stages=True)
        if world > 1:
            nccl = measure_multiview(args, rank, world, dev, "nccl", w_steps, flush)
            workloads["multiview_nccl"] = {k: nccl[k] for k in ("value", "unit", "ms_per_step", "e2e", "exchange",
                                                                  "n_gpus", "steps")}

    if rank == 0:
        # workload counts for the algorithmic bytes: M = tile instances
        with torch.no_grad():
            from fresnel_b200.renderer import build_bins
            from fresnel_b200.camera import camera_vector
            b = build_bins(resident["positions"].detach(), resident["scales"].detach(),
                           resident["rotations"].detach(), resident["colors"].detach(),
                           resident["opacities"].detach(), camera_vector(cam, RES, RES)[None], 1, RES, RES, 64.0,
                           sync=True)      # sync=True: b.m is the true instance count, not the capacity
            M = b.m
        peak, peak_src = peaks()
        HW, N = RES * RES, N_GAUSS
        alg = stage_algorithmic_bytes(N, HW, M)
        top = max(stages, key=stages.get)
        traffic = None       # DRAM bytes per launch of the dominant kernel from the committed ncu capture
        issue = None         # the bound that actually holds: warp instructions issued per second against the SMs' peak
        prof, prof_name = load_traffic_profile()
        if prof:
            traffic = prof.get(top)
            n_inst = prof.get("warp_instructions", {}).get(top)
            if n_inst:
                sms = torch.cuda.get_device_properties(dev).multi_processor_count
                mhz = (clocks or {}).get("sm_mhz") or 1965.0
                peak_inst = sms * 4 * mhz * 1e6                      # one warp instruction per scheduler per clock
                ach = n_inst / (stages[top] * 1e-3)
                issue = {"warp_instructions_per_launch": n_inst, "achieved_ginst_per_s": ach / 1e9,
                         "peak_ginst_per_s": peak_inst / 1e9, "frac": ach / peak_inst,
                         "note": f"instruction count from the committed ncu capture (profiles/{prof_name}, taken at "
                                 f"commit {prof.get('commit', '?')}), duration measured live; peak = SMs x 4 "
                                 "schedulers x SM clock under load"}
        achieved = alg[top] / (stages[top] * 1e-3) / 1e9
        frame_bytes = 168 * N + 36 * HW + 96 * M
        value = world * args.steps / (tot_ms * 1e-3)
        e2e_serial = world * args.steps / (tot_e2e_ms * 1e-3)
        e2e = world * args.steps / (tot_pipe_ms * 1e-3)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": tot_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"single-view render fwd+bwd, {N} Gaussians, {RES}x{RES} (BASELINE configs[1]); "
                                   "one view per rank",
                       "tile_instances": M, "t_eps": t_eps, "cuda_graph": not args.no_cuda_graph,
                       "l2": "flushed between steps (256 MiB fill), per-step CUDA events summed",
                       "e2e": "HostRenderPipeline: HostRenderSession steps (pinned host buffers both ways, H2D of the "
                              "parameters and upstream gradients, D2H of image, depth and all gradients) replayed from "
                              "CUDA graphs, pipeline_depth steps in flight on as many streams; one bracket over all steps with the "
                              "256 MiB L2 flush of every step INSIDE it; e2e.serial = one step at a time, eager calls, "
                              "per-step brackets, flush outside",
                       "parallelism": f"views sharded over {world} rank(s), no data-path collective"},
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": tot_pipe_ms / args.steps, "pipeline_depth": args.pipeline_depth,
                    "depth_sort": ("one kernel on a 16-CTA cluster, keys in distributed shared memory "
                                   "(frb_depth_sort_in_cluster; leaves 132 SMs to the other frames in flight)"
                                   if pipe_cluster_sort else "one-sweep chain"),
                    "value_no_flush": world * args.steps / (tot_pipe_nf_ms * 1e-3),
                    "serial": {"value": e2e_serial, "ms_per_step": tot_e2e_ms / args.steps},
                    "gradients_only": {"value": world * args.steps / (tot_lean_ms * 1e-3), "unit": UNIT,
                                       "ms_per_step": tot_lean_ms / args.steps, "h2d_bytes_per_step": h2d,
                                       "d2h_bytes_per_step": lean_d2h,
                                       "note": "same pipeline with outputs=('grads',): image and depth stay on the "
                                               "device (HostRenderSession(outputs=...)), the five gradients return"},
                    "host_link_ceiling": {
                        "value": world / (link_ms * 1e-3), "unit": UNIT, "ms_per_step": link_ms,
                        "h2d_gbs_per_gpu": h2d / (link_ms * 1e-3) / 1e9, "d2h_gbs_per_gpu": d2h / (link_ms * 1e-3) / 1e9,
                        "note": "this step's H2D and D2H copies alone (same pinned buffers, both directions at once, "
                                "all ranks at once, no kernels; best of three 16-step trials), max over ranks: the "
                                "host-to-host figure is bounded by about this"}},
            "step_ms": {"min": min(ms), "median": statistics.median(ms), "max": max(ms),
                        "e2e_min": min(ms_e2e), "e2e_median": statistics.median(ms_e2e), "e2e_max": max(ms_e2e),
                        "host_enqueue": enqueue.get(step_timed.__name__)},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": top, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg[top], "kernel_ms": stages[top],
                         "issue_roofline": issue,
                         "frame_algorithmic_bytes": frame_bytes,
                         "frame_frac": frame_bytes / (tot_ms / args.steps * 1e-3) / 1e9 / peak,
                         "stage_source": "CUDA events recorded inside frb_tile_render_fwd / _bwd (the fused entry "
                                         "points `value` replays), eager calls, L2 flushed between steps",
                         "stage_ms": {k: round(v, 4) for k, v in sorted(stages.items(), key=lambda kv: -kv[1])},
                         "stage_gbs": {k: round(alg[k] / (v * 1e-3) / 1e9, 1) for k, v in stages.items() if k in alg}},
        }
        if workloads:
            line["workloads"] = workloads
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline()
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def stage_algorithmic_bytes(N, HW, M):
    """Algorithmic bytes per launch of each stage of the fused tile render."""
    return {
        "frb_project_fwd": 60 * N + 56 * N,
        "frb_depth_order": 4 * (2 * 8 * N) + 8 * N,
        "frb_tile_offsets": 3 * 4 * N,
        "frb_bin_emit": 20 * N + 12 * M,
        "frb_bin_sort_dev": 20 * N + 12 * M + 2 * (2 * 12 * M),
        "frb_tile_count_scan": 8 * N + 8 * (HW // 256) * 98,
        "frb_tile_emit": 12 * N + 4 * M,
        # whole-pass calls with the gather compositors: depth rank in, Gaussian id out, no record copy
        # (the staged path with a sorted record copy moves 96 M more)
        "frb_tile_rank_gather": 4 * M + 4 * M + 4 * N,
        "frb_radix_sort_pairs": 2 * (2 * 12 * M),
        "frb_tile_ranges": 8 * M,
        "frb_gather_records": 4 * M + 96 * M,
        "frb_ranges_and_gather": 8 * M + 4 * M + 96 * M,
        "frb_tile_sort_gather": 8 * M + 4 * M + 96 * M,
        "frb_tile_schedule": 12 * (HW // 256),
        "frb_composite_fwd": 48 * M + 28 * HW,
        "frb_composite_bwd": 52 * M + 24 * HW + 48 * N,
        "frb_project_bwd": 56 * N + 48 * N + 56 * N,
    }


def load_traffic_profile():
    """DRAM bytes and warp instructions per launch from the newest committed ncu summary (profiles/rN_traffic.json)."""
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_traffic.json")))
    if not files:
        return None, None
    return json.load(open(files[-1])), os.path.basename(files[-1])


if __name__ == "__main__":
    main()
