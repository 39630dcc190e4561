"""Timing-perturbation stress of the frame pipeline: the same 1080p sequence encoded with two to six frames in
flight on the device (a) alone, (b) while other streams hammer the GPU, against the frame-after-frame strings.  A race in the row
counters / gating would show up as a mismatching bin string.  python profiles/stress_overlap.py"""
import os, sys, threading
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
os.environ.setdefault('CUDA_DEVICE_MAX_CONNECTIONS', '32')
import numpy as np, torch
from cairo_b200 import gpu, synth
W, H, NF, R = 1920, 1080, 24, 2
frames = torch.empty((NF, H, W, 3), dtype=torch.uint8)
for t in range(NF): frames.numpy()[t] = synth.frame(W, H, t, 0, 'moving')
dev = frames.cuda()

def same(a, b):
    if a[1] != b[1] or a[2] != b[2]: return False
    n = a[1]; full, rest = n // 64, n % 64
    if not (a[0][:full] == b[0][:full]).all(): return False
    return rest == 0 or bool((a[0][full] & np.uint64((1 << rest) - 1)) == (b[0][full] & np.uint64((1 << rest) - 1)))

import time

def run(pipelined, full_depth=False):
    """pipelined: the frame pipeline (six frame slots, frames following each other on the device) against frame_slots = 1"""
    p = gpu.Pipeline(W, H, R, 0, 1, frame_slots=0 if pipelined else 1); p.set_output(1)
    out, inflight = [], 0
    depth = p.encode_capacity() if full_depth else 2
    for t in range(NF):
        p.encode_submit(int(dev[t].data_ptr()), 0 if t in (0, 11) else 1, t, 16); inflight += 1
        if inflight >= depth:
            out.append(p.encode_collect_bins()); inflight -= 1
    while inflight:
        out.append(p.encode_collect_bins()); inflight -= 1
    p.close()
    return out

base = run(False)
alone = run(True)
print('overlap alone  mismatching frames:', [t for t in range(NF) if not same(base[t], alone[t])], flush=True)
stop = False
def noise(seed):
    q = gpu.Pipeline(W, H, 2, 0, 1, frame_slots=1)
    t = 0
    while not stop:
        q.encode(int(dev[t % NF].data_ptr()), 0 if t == 0 else 1, t, 8 + (seed % 16)); t += 1
    q.close()
ths = [threading.Thread(target=noise, args=(i,)) for i in range(5)]
for x in ths: x.start()
# the first loaded run starts BEFORE the foreign encoders are up (they appear while frames are in flight), the others after
bad = 0
for rep in range(6):
    got = run(True, full_depth=rep % 2 == 0)
    m = [t for t in range(NF) if not same(base[t], got[t])]
    bad += len(m)
    print('overlap loaded', rep, 'mismatching frames:', m, flush=True)
stop = True
for x in ths: x.join()
print('stress done, mismatching frames:', bad)
