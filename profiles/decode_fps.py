"""Decoder throughput through the public API: N 1080p frames encoded once, then decoded with evx1_decoder::decode (one
frame at a time) and submit/collect (frames in flight).  python profiles/decode_fps.py [frames] [ahead]"""
import os, sys, time
os.environ.setdefault('CUDA_DEVICE_MAX_CONNECTIONS', '32')
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cairo_b200 import api, synth
N = int(sys.argv[1]) if len(sys.argv) > 1 else 240
AHEAD = int(sys.argv[2]) if len(sys.argv) > 2 else 7
W, H, UNIQ = 1920, 1080, 24
host = torch.empty((UNIQ, H, W, 3), dtype=torch.uint8).pin_memory()
for t in range(UNIQ):
    host.numpy()[t] = synth.frame(W, H, t, 0, 'moving')
fidx = lambda t: t if t < UNIQ else 1 + (t - 1) % (UNIQ - 1)
enc = api.evx1_encoder(ref_count=2); enc.set_quality(16)
coded = []
for t in range(N):
    d, b = enc.encode((int(host[fidx(t)].data_ptr()), W, H)); coded.append((d.copy(), b))
del enc
out = torch.empty((H, W, 3), dtype=torch.uint8).pin_memory().numpy()
ref = []
dec = api.evx1_decoder()
t0 = time.perf_counter()
for k, (d, b) in enumerate(coded):
    dec.decode(d, b, W, H, out=out)
    if k % 40 == 7: ref.append((k, out.copy()))
sync = N / (time.perf_counter() - t0)
del dec
dec = api.evx1_decoder()
check = dict(ref); bad = 0; got = 0
torch.cuda.synchronize(); t0 = time.perf_counter()
for d, b in coded[:AHEAD]:
    dec.submit(d, b)
for d, b in coded[AHEAD:]:
    dec.submit(d, b); dec.collect(W, H, out=out)
    if got in check and not (out == check[got]).all(): bad += 1
    got += 1
for _ in range(AHEAD):
    dec.collect(W, H, out=out)
    if got in check and not (out == check[got]).all(): bad += 1
    got += 1
pipe = N / (time.perf_counter() - t0)
print(f"decode() {sync:.0f} frames/s, submit/collect ({AHEAD} ahead) {pipe:.0f} frames/s, pictures differing from decode(): {bad} of {len(ref)} checked")
