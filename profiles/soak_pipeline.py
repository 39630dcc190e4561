"""Long run of the frame pipeline against the frame-after-frame kernels: N frames (cycling through 24 distinct pictures, an
intra frame every 97), every frame's bin string compared.  python profiles/soak_pipeline.py [frames] [ref_count] [w h]"""
import os, sys
os.environ.setdefault('CUDA_DEVICE_MAX_CONNECTIONS', '32')
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from cairo_b200 import gpu, synth
N = int(sys.argv[1]) if len(sys.argv) > 1 else 600
R = int(sys.argv[2]) if len(sys.argv) > 2 else 2
W, H = (int(sys.argv[3]), int(sys.argv[4])) if len(sys.argv) > 4 else (1920, 1080)
UNIQ = 24
host = torch.empty((UNIQ, H, W, 3), dtype=torch.uint8)
for t in range(UNIQ):
    host.numpy()[t] = synth.frame(W, H, t, 0, 'moving')
dev = host.cuda()
fidx = lambda t: t if t < UNIQ else 1 + (t - 1) % (UNIQ - 1)
ftype = lambda t: 0 if t % 97 == 0 else 1
q = lambda t: 16 if (t // 50) % 2 == 0 else 9          # the quality changes every 50 frames


def digest(r):
    words, nbins, ncoded = r
    full, rest = nbins // 64, nbins % 64
    h = int(np.bitwise_xor.reduce(words[:full] * np.arange(1, full + 1, dtype=np.uint64))) if full else 0
    tail = int(words[full]) & ((1 << rest) - 1) if rest else 0
    return (nbins, ncoded, h, tail)


one = gpu.Pipeline(W, H, R, 0, 1, frame_slots=1); one.set_output(1)
want = []
for t in range(N):
    one.encode_submit(int(dev[fidx(t)].data_ptr()), ftype(t), t, q(t)); want.append(digest(one.encode_collect_bins()))
one.close()
p = gpu.Pipeline(W, H, R, 0, 1); p.set_output(1)
cap, got, inflight = p.encode_capacity(), [], 0
for t in range(N):
    p.encode_submit(int(dev[fidx(t)].data_ptr()), ftype(t), t, q(t)); inflight += 1
    if inflight >= cap:
        got.append(digest(p.encode_collect_bins())); inflight -= 1
while inflight:
    got.append(digest(p.encode_collect_bins())); inflight -= 1
bad = [t for t in range(N) if got[t] != want[t]]
print(f"{W}x{H} R={R} slots {cap}: {N} frames, mismatching frames: {bad[:10]} ({len(bad)})")
