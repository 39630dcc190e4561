"""The frame pipeline alone: frames resident in HBM -> submit / collect_bins through the C-ABI, as many frames in flight as
the handle takes.  Prints frames/s for the environment it is run in (EVXGPU_FRAME_SLOTS, EVXGPU_K2_CTAS, EVXGPU_K3_REGS ...).

    python profiles/pipe_bench.py [frames] [ref_count] [width height] [streams]
"""
import os
import sys
import threading
import time
os.environ.setdefault('CUDA_DEVICE_MAX_CONNECTIONS', '32')
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from cairo_b200 import gpu, synth

NF = int(sys.argv[1]) if len(sys.argv) > 1 else 200
R = int(sys.argv[2]) if len(sys.argv) > 2 else 2
W, H = (int(sys.argv[3]), int(sys.argv[4])) if len(sys.argv) > 4 else (1920, 1080)
S = int(sys.argv[5]) if len(sys.argv) > 5 else 1
Q, UNIQ, WARM = 16, 40, 6

host = torch.empty((UNIQ, H, W, 3), dtype=torch.uint8)
for t in range(UNIQ):
    host.numpy()[t] = synth.frame(W, H, t, 0, 'moving')
dev = host.cuda()
fidx = lambda t: t if t < UNIQ else 1 + (t - 1) % (UNIQ - 1)

pipes = [gpu.Pipeline(W, H, R, 0, 1) for _ in range(S)]
for p in pipes:
    p.set_output(1)
bar = threading.Barrier(S + 1)
bits = [0] * S


def work(i):
    p = pipes[i]
    for t in range(WARM):
        p.encode_submit(int(dev[fidx(t)].data_ptr()), 0 if t == 0 else 1, t, Q)
        p.encode_collect_bins()
    cap = p.encode_capacity()
    bar.wait()
    ahead = min(cap - 1, NF - 1)
    for t in range(WARM, WARM + ahead):
        p.encode_submit(int(dev[fidx(t)].data_ptr()), 1, t, Q)
    for t in range(WARM + ahead, WARM + NF):
        p.encode_submit(int(dev[fidx(t)].data_ptr()), 1, t, Q)
        bits[i] += p.encode_collect_bins()[1]
    for _ in range(ahead):
        bits[i] += p.encode_collect_bins()[1]
    bar.wait()


th = [threading.Thread(target=work, args=(i,)) for i in range(S)]
for x in th:
    x.start()
bar.wait()
torch.cuda.synchronize()
t0 = time.perf_counter()
bar.wait()
torch.cuda.synchronize()
dt = time.perf_counter() - t0
for x in th:
    x.join()
env = {k: v for k, v in os.environ.items() if k.startswith('EVXGPU_')}
print(f"streams {S} slots {pipes[0].encode_capacity()} R {R} {W}x{H}: {S * NF / dt:8.1f} frames/s  ({1e3 * dt / NF:.3f} ms/frame/stream, {sum(bits)} bins)  {env}")
