"""Timing-perturbation stress of the frame overlap: the same 1080p sequence encoded with two overlapped frames in
flight (a) alone, (b) while other streams hammer the GPU, against the frame-after-frame strings.  A race in the row
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

import ctypes as C, faulthandler
current = {}
def watchdog():
    """A collect that takes more than 10 s is a hang: print the row counters of the stuck handle and leave."""
    import time
    while True:
        time.sleep(1.0)
        t0 = current.get('t0')
        if t0 and time.time() - t0 > 10.0:
            st = (C.c_uint * 10)()
            rc = gpu.lib().evxgpu_debug_overlap_state(current['h'], st)
            v = list(st)
            print('HANG at frame', current.get('t'), 'rc', rc, 'slot0 rows/final/k2', [x & 4095 for x in v[0:3]], 'epoch', v[0] >> 12, v[1] >> 12, v[2] >> 12,
                  '| slot1', [x & 4095 for x in v[4:7]], 'epoch', v[4] >> 12, v[5] >> 12, v[6] >> 12, '| slot epochs', v[8] >> 12, v[9] >> 12, flush=True)
            faulthandler.dump_traceback()
            os._exit(3)
threading.Thread(target=watchdog, daemon=True).start()

def run(overlap):
    import time
    os.environ['EVXGPU_FRAME_OVERLAP'] = '1' if overlap else '0'
    p = gpu.Pipeline(W, H, R, 0, 1); p.set_output(1)
    gpu.lib().evxgpu_debug_overlap_state.argtypes = [C.c_void_p, C.c_void_p]
    current['h'] = p.h
    out = []
    p.encode_submit(int(dev[0].data_ptr()), 0, 0, 16)
    for t in range(1, NF):
        current['t'] = t; current['t0'] = time.time() if overlap else None
        p.encode_submit(int(dev[t].data_ptr()), 0 if t == 11 else 1, t, 16)
        out.append(p.encode_collect_bins())
    out.append(p.encode_collect_bins())
    current['t0'] = None
    p.close()
    return out

base = run(False)
alone = run(True)
print('overlap alone  mismatching frames:', [t for t in range(NF) if not same(base[t], alone[t])], flush=True)
stop = False
def noise(seed):
    os.environ['EVXGPU_FRAME_OVERLAP'] = '0'
    q = gpu.Pipeline(W, H, 2, 0, 1)
    t = 0
    while not stop:
        q.encode(int(dev[t % NF].data_ptr()), 0 if t == 0 else 1, t, 8 + (seed % 16)); t += 1
    q.close()
ths = [threading.Thread(target=noise, args=(i,)) for i in range(5)]
for x in ths: x.start()
import time; time.sleep(0.5)
bad = 0
for rep in range(6):
    got = run(True)
    m = [t for t in range(NF) if not same(base[t], got[t])]
    bad += len(m)
    print('overlap loaded', rep, 'mismatching frames:', m, flush=True)
stop = True
for x in ths: x.join()
print('stress done, mismatching frames:', bad)
