"""Concurrent pipelined streams against their own synchronous runs (the parity stress of tests/test_gpu_configs.py, longer):
python profiles/stress_streams.py [streams] [frames] [rounds].  Prints every (round, stream, frame) whose bytes differ."""
import os, sys, threading
sys.path.insert(0, '/root/repo')
import numpy as np
from cairo_b200 import api, synth
S = int(sys.argv[1]) if len(sys.argv) > 1 else 6
NF = int(sys.argv[2]) if len(sys.argv) > 2 else 8
ROUNDS = int(sys.argv[3]) if len(sys.argv) > 3 else 3
LOOK = int(sys.argv[4]) if len(sys.argv) > 4 else 0          # frames of lookahead of every stream (0: 1 + stream % 4)
w, h, q = 1920, 1080, 16
frames = [[synth.frame(w, h, t, s, "moving") for t in range(NF)] for s in range(2)]
want = []
for s in range(2):
    enc = api.evx1_encoder(ref_count=2); enc.set_quality(q)
    out = []
    for t in range(NF):
        d, b = enc.encode(frames[s][t]); out.append((d.copy(), b))
    want.append(out)
    del enc
bad = 0
for rnd in range(ROUNDS):
    results = [None] * S
    def work(i):
        enc = api.evx1_encoder(ref_count=2); enc.set_quality(q)
        fr = frames[i % 2]; out = []
        look = LOOK if LOOK else 1 + (i % 4)
        for t in range(NF):
            enc.submit(fr[t])
            if t >= look:
                d, b = enc.collect(); out.append((d.copy(), b))
        while len(out) < NF:
            d, b = enc.collect(); out.append((d.copy(), b))
        results[i] = out
    th = [threading.Thread(target=work, args=(i,)) for i in range(S)]
    for x in th: x.start()
    for x in th: x.join()
    for i in range(S):
        for t in range(NF):
            d, b = results[i][t]; wd, wb = want[i % 2][t]
            if b != wb or not (d == wd).all():
                bad += 1
                print(f"MISMATCH round {rnd} stream {i} frame {t}: {b} vs {wb} bits")
print("env", {k: v for k, v in os.environ.items() if k.startswith('EVXGPU_')}, "mismatches:", bad)
