import os, sys
sys.path.insert(0, '/root/repo')
import numpy as np
from cairo_b200 import gpu, synth
w, h, q, NF = 1920, 1080, 16, 8
for seed in (0, 1):
    frames = [synth.frame(w, h, t, seed, "moving") for t in range(NF)]
    outs = {}
    for name, env in (("regs1", {"EVXGPU_FRAME_SLOTS": "1"}), ("regs2", {"EVXGPU_FRAME_SLOTS": "1", "EVXGPU_K3_REGS": "2"}), ("pipe2", {}), ("pipe1", {"EVXGPU_K3_REGS": "1"})):
        for k in ("EVXGPU_FRAME_SLOTS", "EVXGPU_K3_REGS"):
            os.environ.pop(k, None)
        os.environ.update(env)
        p = gpu.Pipeline(w, h, 2, 0, 1); p.set_output(1)
        out = []
        for t in range(NF):
            p.encode_submit(frames[t], 0 if t == 0 else 1, t, q); out.append(p.encode_collect_bins())
        outs[name] = out; p.close()
    for name in ("regs2", "pipe2", "pipe1"):
        diffs = [t for t in range(NF) if outs[name][t][1] != outs["regs1"][t][1] or outs[name][t][2] != outs["regs1"][t][2]]
        print("seed", seed, name, "vs regs1: frames with different bin count / coded blocks:", diffs, [(outs[name][t][1], outs["regs1"][t][1]) for t in diffs[:3]])
