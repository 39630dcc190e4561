"""Per-kernel times of the stand-alone kernels (frame after frame): python profiles/kernel_times.py [ref_count] [frames]"""
import os, sys
os.environ.setdefault('CUDA_DEVICE_MAX_CONNECTIONS', '32')
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cairo_b200 import gpu, synth
R = int(sys.argv[1]) if len(sys.argv) > 1 else 2
NF = int(sys.argv[2]) if len(sys.argv) > 2 else 24
W, H = 1920, 1080
host = torch.empty((NF + 4, H, W, 3), dtype=torch.uint8)
for t in range(NF + 4):
    host.numpy()[t] = synth.frame(W, H, t, 0, 'moving')
dev = host.cuda()
p = gpu.Pipeline(W, H, R, 0, 1, frame_slots=1)
p.enable_timing(True); p.set_output(1)
for t in range(4):
    p.encode_submit(int(dev[t].data_ptr()), 0 if t == 0 else 1, t, 16); p.encode_collect_bins()
p.timing_sum(reset=True); p.counters(reset=True)
for t in range(4, 4 + NF):
    p.encode_submit(int(dev[t].data_ptr()), 1, t, 16); p.encode_collect_bins()
ks = {k: round(1e3 * v / NF, 1) for k, v in p.timing_sum().items() if v}
c = [x / NF for x in p.counters_split()]
ops = c[0] * 1024 + c[1] * 2560
peak = gpu.lib().evxgpu_measure_int_peak(0, 1)
print(f"R={R} us per frame: {ks} | K2 {ops / (ks['inter_search'] * 1e-6) / 1e12:.2f} Tiop/s = {ops / (ks['inter_search'] * 1e-6) / 1e12 / peak:.3f} of {peak:.2f}")
