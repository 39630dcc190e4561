"""Four 1080p frames (1 intra + 3 P) through the pixel pipeline with both outputs on (table + records
and the device-built bin string); the short command ncu wraps."""
import sys
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
from cairo_b200 import gpu, synth
W, H = 1920, 1080
p = gpu.Pipeline(W, H, 2, 0, 1)
p.set_output(2)
for t in range(4):
    p.encode_submit(synth.frame(W, H, t, 0, 'moving'), 0 if t == 0 else 1, t, 16)
    words, nbins, ncoded = p.encode_collect_bins()
    tbl, rec = p.encode_collect()
print("ok", rec.shape, nbins)
