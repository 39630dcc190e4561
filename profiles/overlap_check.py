"""Frame overlap (EVXGPU_FRAME_OVERLAP=1): the bin strings of a 1080p sequence with two frames in flight against the
one-frame-at-a-time strings, and the device period.  python profiles/overlap_check.py [frames]"""
import os, sys, time
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np, torch
from cairo_b200 import gpu, synth
W, H, Q = 1920, 1080, 16
N = int(sys.argv[1]) if len(sys.argv) > 1 else 12
R = int(sys.argv[2]) if len(sys.argv) > 2 else 2
frames = torch.empty((N, H, W, 3), dtype=torch.uint8)
for t in range(N): frames.numpy()[t] = synth.frame(W, H, t, 0, 'moving')
dev = frames.cuda()

def run(overlap):
    os.environ['EVXGPU_FRAME_OVERLAP'] = '1' if overlap else '0'
    p = gpu.Pipeline(W, H, R, 0, 1)
    p.set_output(1)
    out = []
    t0 = time.perf_counter()
    p.encode_submit(int(dev[0].data_ptr()), 0, 0, Q)
    for t in range(1, N):
        p.encode_submit(int(dev[t].data_ptr()), 1, t, Q)
        out.append(p.encode_collect_bins())
        if t == 3: torch.cuda.synchronize(); t0 = time.perf_counter(); n0 = t
    out.append(p.encode_collect_bins())
    dt = time.perf_counter() - t0
    rec = [a.copy() for a in p.planes(2, (N - 1) % R)]
    p.close()
    return out, 1e3 * dt / (N - n0), rec

a, ta, ra = run(False)
print(f"serial : {ta:.3f} ms/frame", flush=True)
b, tb, rb = run(True)
print(f"overlap: {tb:.3f} ms/frame", flush=True)
bad = [t for t in range(N) if not (a[t][1] == b[t][1] and a[t][2] == b[t][2] and (a[t][0] == b[t][0]).all())]
print("mismatching frames:", bad, "| final reconstruction equal:", all((x == y).all() for x, y in zip(ra, rb)))
