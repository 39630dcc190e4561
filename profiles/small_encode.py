"""Two QCIF frames through the encoder and decoder pixel pipelines (the command compute-sanitizer wraps)."""
import sys
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
from cairo_b200 import gpu, synth
W, H = 176, 144
p = gpu.Pipeline(W, H, 4, 0, 1)
d = gpu.Pipeline(W, H, 4, 0, 1)
for t in range(2):
    tbl, rec = p.encode(synth.frame(W, H, t, 0, 'dark'), 0 if t == 0 else 1, t, 16)
    d.decode(tbl, rec, 0 if t == 0 else 1, t)
print("ok", rec.shape)
