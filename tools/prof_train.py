"""Kernel-time breakdown of the decoder training step (eager, torch.profiler) - where the 4 ms go."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from fresnel_b200.training import DecoderTrainer, PatchGaussianDecoder
full = len(sys.argv) > 1 and sys.argv[1] == "full"
dev = torch.device("cuda:0")
torch.manual_seed(0)
model = PatchGaussianDecoder(384, 4).to(dev)
res, k = (256, None) if full else (64, 256)
tr = DecoderTrainer(model, res, stochastic_k=k, seed=0, cuda_graph=False)
batch = [t.to(dev) for t in bench.train_batch(0)]
for _ in range(5): tr.step(*batch)
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(10): tr.step(*batch)
    torch.cuda.synchronize()
ev = [e for e in prof.key_averages() if e.device_time_total > 0]
ev.sort(key=lambda e: -e.device_time_total)
tot = sum(e.device_time_total for e in ev if e.device_type.name == "CUDA")
print("total CUDA kernel time per step (us):", tot / 10)
for e in ev[:45]:
    if e.device_type.name != "CUDA": continue
    print(f"{e.device_time_total/10:9.1f} us  x{e.count/10:5.1f}  {e.key[:110]}")
