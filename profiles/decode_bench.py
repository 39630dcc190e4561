import sys, time
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np
from cairo_b200 import api, gpu, synth
W, H = 1920, 1080
enc = api.evx1_encoder(ref_count=2); enc.set_quality(16)
dec = api.evx1_decoder()
streams = []
for t in range(12):
    d, b = enc.encode(synth.frame(W, H, t, 0, 'moving')); streams.append((d.copy(), b))
import ctypes as C
L = gpu.lib()
ptr = L.evxgpu_host_alloc(W * H * 3)
out = np.ctypeslib.as_array((C.c_uint8 * (W * H * 3)).from_address(ptr)).reshape(H, W, 3)
for t in range(4): dec.decode(*streams[t], W, H, out=out)
t0 = time.perf_counter(); ent = g = 0.0
for t in range(4, 12):
    dec.decode(*streams[t], W, H, out=out); st = dec.stats(); ent += st["entropy_ms"]; g += st["gpu_ms"]
dt = (time.perf_counter() - t0) / 8
print(f"decode e2e (pinned output) {dt*1e3:.2f} ms/frame = {1/dt:.1f} fps; unserialize {ent/8:.2f} ms, submit->collect {g/8:.2f} ms")
# kernel-level timing through a raw pipeline
p = gpu.Pipeline(W, H, 2, 0, 1); p.enable_timing(True)
q = gpu.Pipeline(W, H, 2, 0, 1)
for t in range(6):
    tbl, rec = q.encode(synth.frame(W, H, t, 0, 'moving'), 0 if t == 0 else 1, t, 16)
    p.decode(tbl, rec, 0 if t == 0 else 1, t)
    print(t, {k: round(v, 3) for k, v in p.timing().items() if v > 0})
