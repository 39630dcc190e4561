"""Where a frame's time goes inside the frame pipeline (build with EVX_EXTRA_NVCC="-DEVX_K3_TIMELINE -DEVX_K3_STATS"):
per-phase cycles per macroblock of the compute warps and the row timeline of the last eight frames of a saturated
pipeline, next to the same frame encoded alone.  python profiles/pipe_phases.py [slots] [frames]"""
import os, sys, ctypes as C, numpy as np
os.environ.setdefault('CUDA_DEVICE_MAX_CONNECTIONS', '32')
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cairo_b200 import gpu, synth
L = gpu.lib()
L.evxgpu_debug_profile.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
SLOTS = int(sys.argv[1]) if len(sys.argv) > 1 else 0
NF = int(sys.argv[2]) if len(sys.argv) > 2 else 48
W, H = 1920, 1080
host = torch.empty((24, H, W, 3), dtype=torch.uint8)
for t in range(24):
    host.numpy()[t] = synth.frame(W, H, t, 0, 'moving')
dev = host.cuda()
fidx = lambda t: t if t < 24 else 1 + (t - 1) % 23
names = ['wait_cols', 'search5', 'subpel', 'class+pred', 'xform+recon', 'far_waits', 'hold1', 'hold2', 'hold3', 'wait_loader']
p = gpu.Pipeline(W, H, 2, 0, 1, frame_slots=SLOTS)
p.set_output(1)
L.evxgpu_debug_profile(p.h, 1, None)
cap = p.encode_capacity()
inflight = 0
for t in range(NF):
    p.encode_submit(int(dev[fidx(t)].data_ptr()), 0 if t == 0 else 1, t, 16); inflight += 1
    if inflight >= cap:
        p.encode_collect_bins(); inflight -= 1
while inflight:
    p.encode_collect_bins(); inflight -= 1
mbh, mbw, nmb = p.ah // 16, p.aw // 16, p.nblocks
stride = mbh * 10 + nmb * 4
raw = np.zeros(stride * 8, dtype=np.int64)
L.evxgpu_debug_profile(p.h, 1, raw.ctypes.data_as(C.c_void_p))
fr = []
for k in range(8):
    prof = raw[k * stride:k * stride + mbh * 10].reshape(mbh, 10)
    tl = raw[k * stride + mbh * 10:k * stride + mbh * 10 + mbh * 4].reshape(mbh, 4)
    fr.append((prof, tl))
order = sorted(range(8), key=lambda k: fr[k][1][0, 0])
if cap <= 2:
    order = order[-1:]
t0 = fr[order[0]][1][0, 0]
print(f"slots {cap}, {NF} frames")
prev = None
for k in order:
    prof, tl = fr[k]
    tl = (tl - t0) / 1e3
    mid = prof[10:60].mean(axis=0) / mbw
    line = f"frame@{k}: row0 start {tl[0,1]:8.1f} row33 start {tl[33,1]:8.1f} row67 end {tl[67,2]:8.1f} residence {tl[67,2]-tl[0,1]:7.1f} us"
    if prev is not None:
        line += f" | period(row33 start) {tl[33,1]-prev:6.1f}"
    prev = tl[33, 1]
    print(line)
    print("     cycles/MB rows 10..60:", {n: int(v) for n, v in zip(names, mid) if not n.startswith('hold') and n != 'far_waits'}, "row time us (median):", round(float(np.median(tl[10:60, 2] - tl[10:60, 1])), 1),
          "claim->start us:", round(float(np.median(tl[10:60, 1] - tl[10:60, 0])), 1),
          "| loader (build with -DEVX_K3_LOADER_STATS): own searches per row", round(float(prof[10:60, 6].mean()), 1), "cycles/MB in them", int(prof[10:60, 8].mean() / mbw),
          "cycles/MB waiting for stamps", int(prof[10:60, 7].mean() / mbw))
