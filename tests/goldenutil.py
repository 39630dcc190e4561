"""Loader for tests/golden/*.npz (made by tests/golden/make_golden.py from the reference)."""
import glob
import os

import numpy as np

from cairo_b200 import synth

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def names():
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


class Golden:
    def __init__(self, name):
        self.z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
        (self.w, self.h, self.frames, self.R, self.q, self.seed, self.intra_every,
         self.linear, self.deblocking) = [int(v) for v in self.z["meta"]]
        self.kind = str(self.z["kind"])

    def rgb(self, t):
        return synth.frame(self.w, self.h, t, self.seed, self.kind)

    def is_intra(self, t):
        return t == 0 or (self.intra_every and t % self.intra_every == 0)

    def planes(self, t, what):
        return [self.z[f"f{t}_{what}{c}"] for c in range(3)]

    def table(self, t):
        return self.z[f"f{t}_table"]

    def slice_bits(self, t):
        return self.z[f"f{t}_slice"], int(self.z[f"f{t}_slice_bits"])

    def stream(self, t):
        return self.z[f"f{t}_stream"], int(self.z[f"f{t}_stream_bits"])

    def decoded_rgb(self, t):
        return self.z[f"f{t}_rgb"]
