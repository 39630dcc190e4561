"""Row timeline of the frame kernel in the pipeline (build with EVX_EXTRA_NVCC=-DEVX_K3_TIMELINE): for eight consecutive frames
in flight, when each wavefront row was claimed, started, finished its first macroblock and ended (globaltimer, us relative to the
first frame's first claim).  python profiles/timeline_pipe.py"""
import os, sys, ctypes as C, numpy as np
os.environ.setdefault('CUDA_DEVICE_MAX_CONNECTIONS', '32')
sys.path.insert(0, '/root/repo')
import torch
from cairo_b200 import gpu, synth
L = gpu.lib()
L.evxgpu_debug_profile.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
W, H, NF = 1920, 1080, 40
host = torch.empty((NF, H, W, 3), dtype=torch.uint8)
for t in range(NF):
    host.numpy()[t] = synth.frame(W, H, t, 0, 'moving')
dev = host.cuda()
p = gpu.Pipeline(W, H, 2, 0, 1)
p.set_output(1)
L.evxgpu_debug_profile(p.h, 1, None)
cap = p.encode_capacity()
inflight = 0
for t in range(NF):
    p.encode_submit(int(dev[t].data_ptr()), 0 if t == 0 else 1, t, 16); inflight += 1
    if inflight >= cap:
        p.encode_collect_bins(); inflight -= 1
while inflight:
    p.encode_collect_bins(); inflight -= 1
mbh, nmb = p.ah // 16, p.nblocks
stride = mbh * 10 + nmb * 4
raw = np.zeros(stride * 8, dtype=np.int64)
L.evxgpu_debug_profile(p.h, 1, raw.ctypes.data_as(C.c_void_p))
frames = []
for k in range(8):
    tl = raw[k * stride + mbh * 10: k * stride + mbh * 10 + mbh * 4].reshape(mbh, 4)
    frames.append(tl)
# frame_seq of the last 8 frames: NF-7 .. NF ; order them by time
order = sorted(range(8), key=lambda k: frames[k][0, 0])
t0 = frames[order[0]][:, 0].min()
print("slots", cap)
for k in order:
    tl = (frames[k] - t0) / 1e3
    print(f"frame@{k}: row0 claim {tl[0,0]:8.1f} start {tl[0,1]:8.1f} first-mb {tl[0,3]:8.1f} end {tl[0,2]:8.1f} | row33 claim {tl[33,0]:8.1f} start {tl[33,1]:8.1f} end {tl[33,2]:8.1f} | row67 claim {tl[67,0]:8.1f} start {tl[67,1]:8.1f} end {tl[67,2]:8.1f}")
k = order[3]
tl = (frames[k] - t0) / 1e3
print("one frame, every 4th row: (claim, start, first-mb, end)")
for r in range(0, mbh, 4):
    print(f"  row {r:2d}: {tl[r,0]:8.1f} {tl[r,1]:8.1f} {tl[r,3]:8.1f} {tl[r,2]:8.1f}   row time {tl[r,2]-tl[r,1]:6.1f}")
