import sys, ctypes as C, numpy as np
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
from cairo_b200 import gpu, synth
L = gpu.lib()
L.evxgpu_debug_profile.argtypes=[C.c_void_p, C.c_int, C.c_void_p]
L.evxgpu_set_wave_grid.argtypes=[C.c_void_p, C.c_int]
W,H=1920,1080
frames=[synth.frame(W,H,t,0,'moving') for t in range(4)]
for grid in (48, 96, 148, 296):
    p = gpu.Pipeline(W,H,2,0,1)
    p.enable_timing(True)
    L.evxgpu_set_wave_grid(p.h, grid)
    for t in range(3): p.encode(frames[t], 0 if t==0 else 1, t, 16)
    L.evxgpu_debug_profile(p.h, 1, None)
    p.encode(frames[3],1,3,16)
    tm = p.timing()
    prof = np.zeros((p.nblocks,12),dtype=np.int64)
    L.evxgpu_debug_profile(p.h, 1, prof.ctypes.data_as(C.c_void_p))
    d = np.diff(prof[:,:9],axis=1)
    names=['wait','window','search5','subpel','classify+pred','fdct+quant','recon','release']
    print(f"grid={grid} wavefront={tm['wavefront']:.3f} ms  inter={tm['inter_search']:.3f}")
    print("  mean cycles per phase:", {n:int(v) for n,v in zip(names,d.mean(axis=0))}, "total(excl wait)", int(d[:,1:].sum(axis=1).mean()))
    # critical path: release(globaltimer) of mb (bx-1,by) -> deps satisfied of mb (bx,by)
    mbw=120
    rel = prof[:,11]; dep = prof[:,10]
    lat = []
    for by in range(10,60):
        for bx in range(10,110):
            mb=by*mbw+bx
            prev=max(rel[mb-1], rel[(by-1)*mbw+min(bx+2,mbw-1)])
            lat.append(dep[mb]-prev)
    lat=np.array(lat); print("  release->acquire ns: median",np.median(lat),"mean",lat.mean(), "p90", np.percentile(lat,90))
    # per-MB active ns (deps satisfied -> release)
    act = (prof[:,11]-prof[:,10]); print("  active ns per MB: median", np.median(act), "mean", act.mean())
    p.close()
