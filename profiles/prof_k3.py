"""Phase breakdown of the encoder wavefront kernel (evx_wavefront): per-row cycle sums of the
compute warps, via evxgpu_debug_profile.  Run on the GPU box: python profiles/prof_k3.py"""
import sys, ctypes as C, numpy as np
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
from cairo_b200 import gpu, synth
L = gpu.lib()
L.evxgpu_debug_profile.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
W, H = 1920, 1080
frames = [synth.frame(W, H, t, 0, 'moving') for t in range(4)]
p = gpu.Pipeline(W, H, 2, 0, 1)
p.enable_timing(True)
for t in range(3):
    p.encode(frames[t], 0 if t == 0 else 1, t, 16)
L.evxgpu_debug_profile(p.h, 1, None)
tbl, rec = p.encode(frames[3], 1, 3, 16)
tm = p.timing()
raw = np.zeros(((p.ah // 16) * 10 + p.nblocks * 4) * 8, dtype=np.int64)      # (the buffer is eight pipeline frames deep; the stand-alone launch uses the first)
L.evxgpu_debug_profile(p.h, 1, raw.ctypes.data_as(C.c_void_p))
prof = raw[:(p.ah // 16) * 10].reshape(-1, 10)
names = ['wait_columns', 'search5', 'subpel', 'classify+pred', 'transform+recon', 'far_col_waits', 'hold1', 'hold2', 'hold3', 'wait_block_loader(stats build)']
mbw = p.aw // 16
print(f"wavefront {tm['wavefront']:.3f} ms, inter {tm['inter_search']:.3f} ms, non-copy share {rec.shape[0] / p.nblocks:.2f}")
print("mean cycles per macroblock (rows 10..60):", {n: int(v) for n, v in zip(names, prof[10:60].mean(axis=0) / mbw)})
print("row 0:", {n: int(v) for n, v in zip(names, prof[0] / mbw)}, " row 67:", {n: int(v) for n, v in zip(names, prof[-1] / mbw)})
print("sum per MB (rows 10..60):", int(prof[10:60, :5].sum(axis=1).mean() / mbw))
