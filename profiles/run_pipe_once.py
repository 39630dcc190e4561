"""Eight 1080p frames (1 intra + 7 P) through the frame pipeline (bin-string output, frames following each other on the
device): the short command ncu wraps for the follower kernels.  Under ncu the kernels are serialised, so a follower's
duration there is its work alone (what it would wait for is complete when it starts)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cairo_b200 import gpu, synth
W, H = 1920, 1080
R = int(sys.argv[1]) if len(sys.argv) > 1 else 2
p = gpu.Pipeline(W, H, R, 0, 1)
p.set_output(1)
n = 0
for t in range(8):
    p.encode_submit(synth.frame(W, H, t, 0, 'moving'), 0 if t == 0 else 1, t, 16)
    n += p.encode_collect_bins()[1]
print("ok", n)
