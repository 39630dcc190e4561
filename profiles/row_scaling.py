"""Wavefront kernel time vs number of macroblock rows (same width): separates the per-macroblock
cost of a row from the coupling between rows."""
import sys
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np
from cairo_b200 import gpu, synth
W = 1920
full = [synth.frame(W, 1080, t, 0, 'moving') for t in range(4)]
for H in (16, 32, 48, 64, 128, 256, 512, 1080):
    p = gpu.Pipeline(W, H, 2, 0, 1); p.enable_timing(True)
    ts = []
    for t in range(4):
        p.encode(np.ascontiguousarray(full[t][:H]), 0 if t == 0 else 1, t, 16)
        ts.append(p.timing()['wavefront'])
    rows = (H + 15) // 16
    print(f"rows={rows:3d}: wavefront {ts[-1]*1e3:8.1f} us  (per step of {120 + 3 * (rows - 1)}: {ts[-1]*1e3/(120 + 3*(rows-1)):.2f} us)")
    p.close()
