"""Where the host thread's time goes per frame in the overlapped encode loop (submit / collect split)."""
import sys, time
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np, torch
from cairo_b200 import api, synth
W, H, N = 1920, 1080, 40
host = torch.empty((N, H, W, 3), dtype=torch.uint8).pin_memory()
for t in range(N): host.numpy()[t] = synth.frame(W, H, t, 0, 'moving')
enc = api.evx1_encoder(ref_count=2); enc.set_quality(16)
for t in range(4): enc.encode((int(host[t].data_ptr()), W, H))
ts = tc = ent = wt = 0.0
t0 = time.perf_counter()
enc.submit((int(host[4].data_ptr()), W, H))
enc.submit((int(host[5].data_ptr()), W, H))
for t in range(6, N):
    a = time.perf_counter(); enc.submit((int(host[t].data_ptr()), W, H)); b = time.perf_counter()
    enc.collect(); c = time.perf_counter()
    st = enc.stats(); ts += b - a; tc += c - b; ent += st["entropy_ms"]; wt += st["wait_ms"]
enc.collect(); enc.collect()
dt = time.perf_counter() - t0
n = N - 6
print(f"period {1e3*dt/(N-4):.3f} ms; submit {1e3*ts/n:.3f} ms, collect {1e3*tc/n:.3f} ms (coder+write {ent/n:.3f} ms, waiting for the device {wt/n:.3f} ms)")
