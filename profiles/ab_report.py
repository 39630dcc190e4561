"""Prints the table of an A/B run made by profiles/k3_ab.sh (gpurun_out/ab_*.json)."""
import json, os, sys
d0 = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "gpurun_out")
for line in open(os.path.join(d0, "ab_tags.txt")):
    tag, v = line.strip().split(":", 1)
    try:
        d = json.loads(open(os.path.join(d0, f"ab_{tag}.json")).read().strip().splitlines()[-1])
        k = d["kernel_ms_per_step"]
        t = open(os.path.join(d0, f"ab_test_{tag}.log")).read().strip().splitlines()[-1]
        print(f"{tag} [{v.strip() or 'base'}] k2 {k['inter_search']*1e3:.1f} us  k3 {k['wavefront']:.4f} ms  value {d['value']:.1f}  e2e {d['e2e']['value']:.1f}  frac {d['roofline']['frac']:.3f}  | {t}")
    except Exception as e:
        print(tag, v, "ERR", e)
