"""Host <-> device copy bandwidth per rank, alone and with every rank copying at once, before and after binding
the process to the CPUs NVML reports as local to its GPU (pinned buffers re-allocated after binding).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/probe_pcie.py
"""
import json
import os

import torch
import torch.distributed as dist


def local_cpus(index):
    import pynvml
    pynvml.nvmlInit()
    h = pynvml.nvmlDeviceGetHandleByIndex(index)
    words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
    cpus = [64 * w + b for w, word in enumerate(words) for b in range(64) if (word >> b) & 1]
    return cpus


def bw(dev, mb=64, iters=10):
    host = torch.empty(mb * 1024 * 1024 // 4).pin_memory()
    host.fill_(1.0)
    d = torch.empty_like(host, device=dev)
    out = {}
    for name, (dst, src) in {"h2d": (d, host), "d2h": (host, d)}.items():
        for _ in range(2):
            dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        dist.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(iters):
            dst.copy_(src, non_blocking=True)
        b.record()
        torch.cuda.synchronize()
        out[name] = round(mb * iters / 1024 / (a.elapsed_time(b) * 1e-3), 1)      # GiB/s
        dist.barrier()
    return out


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    res = {"rank": rank, "affinity_before": len(os.sched_getaffinity(0)), "cpu_count": os.cpu_count()}
    res["concurrent_before"] = bw(dev)
    try:
        cpus = local_cpus(local)
        res["nvml_local_cpus"] = [min(cpus), max(cpus), len(cpus)] if cpus else []
        allowed = sorted(set(cpus) & os.sched_getaffinity(0))
        if allowed:
            os.sched_setaffinity(0, allowed)
        res["affinity_after"] = len(os.sched_getaffinity(0))
    except Exception as e:      # noqa: BLE001
        res["nvml_error"] = repr(e)
    res["concurrent_after_bind"] = bw(dev)
    gathered = [None] * world
    dist.all_gather_object(gathered, res)
    if rank == 0:
        for g in gathered:
            print(json.dumps(g), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
