"""Which macroblocks differ under concurrency: streams at the table + records seam (one frame at a time per stream) against the
same content encoded alone.  python profiles/stress_tables.py [streams] [frames] [rounds] [w h]"""
import os, sys, threading
sys.path.insert(0, '/root/repo')
import numpy as np
from cairo_b200 import gpu, synth
S = int(sys.argv[1]) if len(sys.argv) > 1 else 16
NF = int(sys.argv[2]) if len(sys.argv) > 2 else 10
ROUNDS = int(sys.argv[3]) if len(sys.argv) > 3 else 3
w, h = (int(sys.argv[4]), int(sys.argv[5])) if len(sys.argv) > 5 else (352, 288)
q = 16
frames = [[synth.frame(w, h, t, s, "moving") for t in range(NF)] for s in range(2)]
want = []
for s in range(2):
    p = gpu.Pipeline(w, h, 2, 0, 1)
    out = []
    for t in range(NF):
        tbl, rec = p.encode(frames[s][t], 0 if t == 0 else 1, t, q)
        inter = None
        out.append((tbl.copy(), rec.copy(), [a.copy() for a in p.planes(2, t % 2)]))
    want.append(out); p.close()
mbw = (w + 15) // 16
for rnd in range(ROUNDS):
    results = [None] * S
    def work(i):
        p = gpu.Pipeline(w, h, 2, 0, 1)
        fr = frames[i % 2]; out = []
        for t in range(NF):
            tbl, rec = p.encode(fr[t], 0 if t == 0 else 1, t, q)
            out.append((tbl.copy(), rec.copy(), [a.copy() for a in p.planes(2, t % 2)]))
        results[i] = out; p.close()
    th = [threading.Thread(target=work, args=(i,)) for i in range(S)]
    for x in th: x.start()
    for x in th: x.join()
    shown = 0
    for i in range(S):
        for t in range(NF):
            a, b = results[i][t], want[i % 2][t]
            names = ["block_type", "prediction_target", "motion_x", "motion_y", "sp_pred", "sp_amount", "sp_index", "q_index", "variance"]
            diff = [m for m in range(a[0].shape[0]) if any(a[0][n][m] != b[0][n][m] for n in names)]
            if diff or a[1].shape != b[1].shape or not (a[1] == b[1]).all():
                planes_equal = all((x == y).all() for x, y in zip(a[2], b[2]))
                if shown < 4:
                    shown += 1
                    print(f"round {rnd} stream {i} frame {t}: {len(diff)} table entries differ, first at (x,y)={[(m % mbw, m // mbw) for m in diff[:6]]}; recon equal: {planes_equal}")
                    for m in diff[:3]:
                        print("    got ", {n: int(a[0][n][m]) for n in names}); print("    want", {n: int(b[0][n][m]) for n in names})
                break
print("done")
