import sys, cProfile, pstats, torch, time
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, fresnel_b200
dev = torch.device('cuda:0')
ren = fresnel_b200.TileBasedRenderer(512, 512)
cam = fresnel_b200.Camera(0.8*512, 0.8*512, 256, 256, 512, 512)
host = bench.synthetic_cloud(100000, 0)
res = {k: v.to(dev).requires_grad_(True) for k, v in host.items()}
gi, gd = [t.to(dev) for t in bench.upstream(1)]
def step():
    for v in res.values(): v.grad = None
    img, dep = ren(res['positions'], res['scales'], res['rotations'], res['colors'], res['opacities'], cam, return_depth=True)
    torch.autograd.backward((img, dep), (gi, gd))
for _ in range(10): step()
torch.cuda.synchronize()
t0=time.perf_counter()
for _ in range(100): step()
t1=time.perf_counter(); torch.cuda.synchronize(); t2=time.perf_counter()
print('enqueue ms/step', (t1-t0)*10, 'total ms/step', (t2-t0)*10)
pr = cProfile.Profile(); pr.enable()
for _ in range(100): step()
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats('cumulative').print_stats(22)
from fresnel_b200.host import HostRenderSession
sess = HostRenderSession(ren, 100000, dev)
sess.load(host, *bench.upstream(1))
for _ in range(10): sess.step(cam)
torch.cuda.synchronize()
t0=time.perf_counter()
for _ in range(100): sess.step(cam)
t1=time.perf_counter(); torch.cuda.synchronize(); t2=time.perf_counter()
print('session enqueue ms/step', (t1-t0)*10, 'total ms/step', (t2-t0)*10)
pr = cProfile.Profile(); pr.enable()
for _ in range(100): sess.step(cam)
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats('tottime').print_stats(25)
