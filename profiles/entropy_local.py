"""Host entropy stage alone, on CPU: slices of the 1080p bench sequence (tables + records from the
oracle) through evx1c_slice_writer_serialize; prints ms per P-frame.  Needs no GPU.
    python profiles/entropy_local.py [--gen]     # --gen (re)creates /tmp/evx_entropy_in.npz with the oracle
"""
import sys, time, os
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np
F = '/tmp/evx_entropy_in.npz'
W, H, NF = 1920, 1080, 5
if '--gen' in sys.argv or not os.path.exists(F):
    import oracleharness as O
    from cairo_b200 import synth, gpu
    o = O.Oracle(W, H, 2, 0, 1)
    d = {}
    for t in range(NF):
        o.convert_in(synth.frame(W, H, t, 0, 'moving')); o.encode_slice(0 if t == 0 else 1, t, 16)
        tbl = o.block_table().copy()
        rec = gpu.planes_to_records(tbl, o.planes(1), o.aw)
        d['t%d' % t] = np.frombuffer(tbl.tobytes(), np.uint8); d['r%d' % t] = rec
        o.deblock(t)
    np.savez(F, **d)
from cairo_b200 import api, gpu
z = np.load(F)
w = api.SliceWriter(120, 68, 2)
tabs = [np.frombuffer(z['t%d' % t].tobytes(), gpu.BLOCK_DESC_DTYPE) for t in range(NF)]
recs = [z['r%d' % t] for t in range(NF)]
best = {}
for rep in range(7):
    w2 = api.SliceWriter(120, 68, 2)
    for t in range(NF):
        t0 = time.perf_counter(); bits = w2.serialize(tabs[t], recs[t]); dt = (time.perf_counter() - t0) * 1e3
        best[t] = min(best.get(t, 1e9), dt)
for t in range(NF): print("frame", t, "noncopy", len(recs[t]), "ms %.3f" % best[t])
