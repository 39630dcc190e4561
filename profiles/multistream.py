"""configs[4] on one GPU: S independent 1080p streams, one host thread + one handle + one CUDA
stream each.  Aggregate P-frame throughput of (a) the pixel pipeline with device-resident frames,
(b) the public API end to end (host frames -> bitstreams).  python profiles/multistream.py [S ...]"""
import os, sys, threading, time
os.environ.setdefault('CUDA_DEVICE_MAX_CONNECTIONS', '32')
os.environ.setdefault('EVXGPU_FRAME_OVERLAP', '0')      # many streams: frame after frame within each (the library would let two of them overlap)
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np
import torch
from cairo_b200 import api, gpu, synth

W, H, Q, R, NF = 1920, 1080, 16, 2, 24
counts = [int(a) for a in sys.argv[1:]] or [1, 2, 4, 8]
print("host cores:", os.cpu_count())
base = torch.empty((NF, H, W, 3), dtype=torch.uint8).pin_memory()
for t in range(NF):
    base.numpy()[t] = synth.frame(W, H, t, 0, 'moving')
dev = base.cuda()

for S in counts:
    # (a) pixel pipeline, frames resident in HBM
    pipes = [gpu.Pipeline(W, H, R, 0, 1) for _ in range(S)]
    bar = threading.Barrier(S + 1)
    def work(i):
        p = pipes[i]
        for t in range(4):
            p.encode(int(dev[t].data_ptr()), 0 if t == 0 else 1, t, Q)
        bar.wait()
        for t in range(4, NF):
            p.encode(int(dev[t].data_ptr()), 1, t, Q)
        bar.wait()
    th = [threading.Thread(target=work, args=(i,)) for i in range(S)]
    for x in th: x.start()
    bar.wait(); t0 = time.perf_counter(); bar.wait(); dt = time.perf_counter() - t0
    for x in th: x.join()
    for p in pipes: p.close()
    fps_gpu = S * (NF - 4) / dt
    # (b) public API, host frames
    encs = [api.evx1_encoder(ref_count=R) for _ in range(S)]
    for e in encs: e.set_quality(Q)
    bar = threading.Barrier(S + 1)
    def work2(i):
        e = encs[i]
        for t in range(4):
            e.encode((int(base[t].data_ptr()), W, H))
        bar.wait()
        for t in range(4, NF):
            e.encode((int(base[t].data_ptr()), W, H))
        bar.wait()
    th = [threading.Thread(target=work2, args=(i,)) for i in range(S)]
    for x in th: x.start()
    bar.wait(); t0 = time.perf_counter(); bar.wait(); dt2 = time.perf_counter() - t0
    for x in th: x.join()
    del encs
    # (c) public API, the two halves of encode with one frame in flight per stream
    encs = [api.evx1_encoder(ref_count=R) for _ in range(S)]
    for e in encs: e.set_quality(Q)
    bar = threading.Barrier(S + 1)
    def work3(i):
        e = encs[i]
        for t in range(4):
            e.encode((int(base[t].data_ptr()), W, H))
        bar.wait()
        e.submit((int(base[4].data_ptr()), W, H))
        for t in range(5, NF):
            e.submit((int(base[t].data_ptr()), W, H))
            e.collect()
        e.collect()
        bar.wait()
    th = [threading.Thread(target=work3, args=(i,)) for i in range(S)]
    for x in th: x.start()
    bar.wait(); t0 = time.perf_counter(); bar.wait(); dt3 = time.perf_counter() - t0
    for x in th: x.join()
    del encs
    print(f"S={S}: pixel pipeline {fps_gpu:8.1f} frames/s aggregate ({fps_gpu / S:6.1f}/stream) | end-to-end encode() {S * (NF - 4) / dt2:8.1f} frames/s "
          f"({(NF - 4) / dt2:6.1f}/stream) | submit/collect {S * (NF - 4) / dt3:8.1f} frames/s ({(NF - 4) / dt3:6.1f}/stream)")
