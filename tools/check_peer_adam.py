"""Multi-GPU check of the fused exchange + Adam kernel (csrc/exchange.cu) against NCCL all-reduce + torch Adam.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
        tools/check_peer_adam.py [--time]

Every rank feeds different gradients for several steps; the parameters must agree with the library form to
fp32 rounding, and be BIT-IDENTICAL across ranks (the owner of a shard computes it once and stores it to all).
--time also measures both forms at the size of BASELINE configs[4] (15 floats x 1M Gaussians).
Prints one JSON line on rank 0; exit code 1 on mismatch.
"""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fresnel_b200.training import PeerShardedAdam  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    out = {"world": world}
    ok = True
    for n in (1003, 4096, 262147):
        opt = PeerShardedAdam(n, dev, lr=1e-2)
        g0 = torch.Generator(device="cpu").manual_seed(5)
        init = torch.randn(n, generator=g0)
        opt.param.copy_(init)
        ref = init.to(dev).clone().requires_grad_(True)
        ref_opt = torch.optim.Adam([ref], lr=1e-2, fused=True)
        torch.cuda.synchronize()
        dist.barrier()
        for step in range(6):
            g = torch.Generator(device="cpu").manual_seed(100 * step + rank)
            grad = torch.randn(n, generator=g).to(dev) * (10.0 ** (step - 3))
            opt.grad.copy_(grad)
            opt.step()
            tot = grad.clone()
            dist.all_reduce(tot)
            ref.grad = tot
            ref_opt.step()
        torch.cuda.synchronize()
        err = float((opt.param - ref.detach()).abs().max() / ref.detach().abs().max())
        gathered = [torch.empty_like(opt.param) for _ in range(world)]
        dist.all_gather(gathered, opt.param.clone())
        same = all(torch.equal(gathered[0], t) for t in gathered)
        out[f"n{n}"] = {"rel_err_vs_nccl_adam": err, "bit_identical_across_ranks": same,
                        "steps_on_device": int(opt.state[1])}
        ok = ok and err < 1e-5 and same and int(opt.state[1]) == 6
        dist.barrier()
    if "--time" in sys.argv:
        n = 15_000_000
        opt = PeerShardedAdam(n, dev, lr=1e-3)
        opt.grad.normal_()
        flat = torch.zeros(n, device=dev, requires_grad=True)
        flat.grad = torch.randn(n, device=dev)
        ref_opt = torch.optim.Adam([flat], lr=1e-3, fused=True)

        def timed(fn, iters=20):
            for _ in range(3):
                fn()
            torch.cuda.synchronize(); dist.barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(iters):
                fn()
            b.record()
            torch.cuda.synchronize()
            t = torch.tensor([a.elapsed_time(b) / iters], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t)

        def lib_form():
            dist.all_reduce(flat.grad)
            ref_opt.step()

        out["time_ms"] = {"floats": n, "peer_fused": timed(opt.step), "nccl_allreduce_plus_adam": timed(lib_form)}
    if rank == 0:
        out["ok"] = ok
        print(json.dumps(out), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
