import sys, time, json
sys.path.insert(0,'/root/repo')
import numpy as np, torch, viamd_b200 as vb
import bench as B
cfg=B.CONFIGS['bench']; dev=0
vb.bind_host_to_device(dev)
FPS=4736; steps=6
wl=B.Workload(vb,cfg,dev,0,steps*FPS)
plan=wl.plan(steps*FPS, batch_frames=148, num_streams=0, rdf_variant=0, ingest_mode=0, ingest_threads=0)
na,fstride,cell=wl.na,wl.fstride,wl.cell
hp=[vb.host_alloc_pinned(FPS*fstride*4) for _ in range(2)]
for b,h in enumerate(hp): vb.memcpy_d2h(dev,h,wl.d_frames+b*FPS*fstride*4,FPS*fstride*4)
plan.set_initial_frame(*wl.f0[0],cell)
res=[]
for i in range(steps):
    t0=time.perf_counter(); plan.eval_host_ptr(hp[i%2],fstride,na,cell,i*FPS,FPS); t1=time.perf_counter()
    plan.sync(); t2=time.perf_counter()
    s=0.0
    for name in cfg['results']: s+=float(plan.property_data(name).values[0])
    t3=time.perf_counter(); res.append(((t1-t0)*1e3,(t2-t1)*1e3,(t3-t2)*1e3))
for r in res: print('eval_host %.2f ms  sync(drain+fold) %.2f ms  property_data %.2f ms'%r)
