"""How often a round of the intra search keeps its centre (the state stands), per round: needs a build
with EVX_EXTRA_NVCC=-DEVX_K3_STATS.  python profiles/stats_k3.py"""
import sys, ctypes as C, numpy as np
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
from cairo_b200 import gpu, synth
L = gpu.lib()
L.evxgpu_debug_profile.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
W, H = 1920, 1080
for kind in ['moving', 'noise', 'static']:
    try:
        frames = [synth.frame(W, H, t, 0, kind) for t in range(4)]
    except Exception as ex:
        print(kind, 'unavailable', ex); continue
    p = gpu.Pipeline(W, H, 2, 0, 1)
    for t in range(3):
        p.encode(frames[t], 0 if t == 0 else 1, t, 16)
    L.evxgpu_debug_profile(p.h, 1, None)
    tbl, rec = p.encode(frames[3], 1, 3, 16)
    raw = np.zeros(((p.ah // 16) * 10 + p.nblocks * 4) * 8, dtype=np.int64)      # (eight pipeline frames deep; the stand-alone launch uses the first)
    L.evxgpu_debug_profile(p.h, 1, raw.ctypes.data_as(C.c_void_p))
    prof = raw[:(p.ah // 16) * 10].reshape(-1, 10)
    holds = prof[:, 5:10].sum(axis=0) / p.nblocks
    types, counts = np.unique(tbl['block_type'], return_counts=True)
    print(kind, 'fraction of macroblocks that waited for column n+2 of the row above:', round(float(holds[0]), 3), '| whose round 1..4 kept its centre:', np.round(holds[1:], 3), 'block types', dict(zip(types.tolist(), counts.tolist())))
    p.close()
