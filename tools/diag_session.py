import sys, os, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, fresnel_b200
from fresnel_b200.host import HostRenderSession
dev = torch.device('cuda:0')
ren = fresnel_b200.TileBasedRenderer(512, 512)
cam = fresnel_b200.Camera(0.8*512, 0.8*512, 256, 256, 512, 512)
host = bench.synthetic_cloud(100000, 0)
sess = HostRenderSession(ren, 100000, dev)
sess.load(host, *bench.upstream(1))
flush = torch.empty(64*1024*1024, dtype=torch.float32, device=dev)
for rep in range(3):
    evs=[]; hosts=[]
    st0 = torch.cuda.memory_stats()
    for i in range(60):
        flush.fill_(1.0)
        a,b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0=time.perf_counter(); a.record(); sess.step(cam); b.record(); hosts.append((time.perf_counter()-t0)*1e3)
        evs.append((a,b))
    torch.cuda.synchronize()
    ms=[a.elapsed_time(b) for a,b in evs]
    st1 = torch.cuda.memory_stats()
    print('rep',rep,'median',sorted(ms)[30],'max',max(ms),'argmax',ms.index(max(ms)),'host median',sorted(hosts)[30],'host max',max(hosts), hosts.index(max(hosts)),
          'segments+',st1['num_device_alloc']-st0['num_device_alloc'],'frees+',st1['num_device_free']-st0['num_device_free'],'retries',st1['num_alloc_retries'])
    print([round(x,2) for x in ms[:12]])
