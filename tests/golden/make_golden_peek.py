"""Golden outputs of evx1_encoder::peek (evx1enc.cpp:170-305) from the compiled, unmodified reference
(oracle/_ref, built by `make -C oracle ref` where /root/reference exists).  Run in the build container:

    python tests/golden/make_golden_peek.py        ->  tests/golden/peek/peek_176x144.npz
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import refharness as R
from cairo_b200 import synth

W, H, Q, FRAMES = 176, 144, 16, 3
VIEWS = {"source": 0, "block_table": 2, "quant_table": 3, "spmp_table": 4, "block_variance": 5, "destination": 6}

enc = R.RefEncoder("r4")
enc.set_quality(Q)
for t in range(FRAMES):
    enc.encode(synth.frame(W, H, t, 0, "moving"))
out = {name: enc.peek(state, W, H) for name, state in VIEWS.items()}
np.savez_compressed(os.path.join(HERE, "peek", "peek_176x144.npz"), w=W, h=H, q=Q, frames=FRAMES, **out)
print({k: int(v.sum()) for k, v in out.items()})
