"""Phase breakdown of the frame kernel in the pipeline (bin-string output): per-row cycle sums of the compute warps of one
frame encoded alone.  Build with EVX_EXTRA_NVCC=-DEVX_K3_STATS to split the wait for the block loader from the wait for
the columns of the row above.  python profiles/prof_pipe.py [ref_count]"""
import sys, ctypes as C, numpy as np, time
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
from cairo_b200 import gpu, synth
L = gpu.lib()
L.evxgpu_debug_profile.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
W, H = 1920, 1080
R = int(sys.argv[1]) if len(sys.argv) > 1 else 2
frames = [synth.frame(W, H, t, 0, 'moving') for t in range(6)]
p = gpu.Pipeline(W, H, R, 0, 1)
p.set_output(1)
for t in range(5):
    p.encode_submit(frames[t], 0 if t == 0 else 1, t, 16); p.encode_collect_bins()
L.evxgpu_debug_profile(p.h, 1, None)
t0 = time.perf_counter()
p.encode_submit(frames[5], 1, 5, 16); p.encode_collect_bins()
dt = time.perf_counter() - t0
stride = (p.ah // 16) * 10 + p.nblocks * 4
raw = np.zeros(stride * 8, dtype=np.int64)          # eight frames deep; this frame is the sixth submitted (frame_seq % 8 = 6)
L.evxgpu_debug_profile(p.h, 1, raw.ctypes.data_as(C.c_void_p))
prof = raw[6 * stride:6 * stride + (p.ah // 16) * 10].reshape(-1, 10)
names = ['wait_columns', 'search5', 'subpel', 'classify+pred', 'transform+recon', 'far_col_waits', 'hold1', 'hold2', 'hold3', 'wait_block_loader(stats build)']
mbw = p.aw // 16
print(f"one frame alone, submit -> bins on the host: {dt * 1e3:.3f} ms")
print("mean cycles per macroblock (rows 10..60):", {n: int(v) for n, v in zip(names, prof[10:60].mean(axis=0) / mbw)})
print("row 0:", {n: int(v) for n, v in zip(names, prof[0] / mbw)})
print("row 67:", {n: int(v) for n, v in zip(names, prof[-1] / mbw)})
