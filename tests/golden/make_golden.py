"""Generates tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref/libevxref_*.so,
built by `make -C oracle ref` in a container that has /root/reference).

    python tests/golden/make_golden.py

Each fixture holds, per frame: the seeded RGB input recipe (not the pixels -- synth.frame
regenerates them), the reference's block table, quantised coefficient planes, reconstruction
before and after deblocking, the slice bit string, the full public-API stream
(evx1_encoder::encode) and the reference decoder's RGB output.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import refharness as R  # noqa: E402
from cairo_b200 import synth  # noqa: E402

CASES = [
    # name,            w,   h,  frames, variant,       q,  kind,     seed, intra_every
    ("qcif_r4_q16",    176, 144, 5,     "r4",          16, "moving", 0,    0),
    ("qcif_r2_q16",    176, 144, 4,     "r2",          16, "moving", 1,    0),
    ("qcif_lin_q20",   176, 144, 3,     "r4_linear",   20, "noise",  2,    0),
    ("odd_r4_q8",      200, 120, 5,     "r4",          8,  "moving", 3,    3),
    ("dark_r4_q24",    96,  80,  4,     "r4",          24, "dark",   4,    0),
    ("noise_r4_q31",   96,  80,  3,     "r4",          31, "noise",  5,    0),
    ("static_r4_q16",  96,  80,  3,     "r4",          16, "static", 6,    0),
    ("nodb_r4_q16",    96,  80,  3,     "r4_nodeblock", 16, "moving", 7,   0),
]


def make(name, w, h, nf, variant, q, kind, seed, intra_every):
    st = R.RefStage(w, h, variant)
    enc = R.RefEncoder(variant)
    enc.set_quality(q)
    dec = R.RefDecoder(variant)
    Rn = st.R
    out = {"meta": np.array([w, h, nf, Rn, q, seed, intra_every, st.L.evxref_linear_quant(), st.L.evxref_deblocking()]),
           "kind": np.array(kind)}
    for t in range(nf):
        f = synth.frame(w, h, t, seed, kind)
        intra = t == 0 or (intra_every and t % intra_every == 0)
        if intra and t:
            enc.insert_intra()
        st.set_frame(0 if intra else 1, t, q)
        st.convert_in(f)
        for c, p in enumerate(st.planes(0)):
            out[f"f{t}_src{c}"] = p
        st.encode_slice()
        out[f"f{t}_table"] = st.block_table()
        for c, p in enumerate(st.planes(1)):
            out[f"f{t}_coef{c}"] = p
        for c, p in enumerate(st.planes(2, t % Rn)):
            out[f"f{t}_recon{c}"] = p
        d, b = st.serialize()
        out[f"f{t}_slice"] = d
        out[f"f{t}_slice_bits"] = np.array(b)
        st.deblock()
        for c, p in enumerate(st.planes(2, t % Rn)):
            out[f"f{t}_deblocked{c}"] = p
        data, bits = enc.encode(f)
        out[f"f{t}_stream"] = data
        out[f"f{t}_stream_bits"] = np.array(bits)
        out[f"f{t}_rgb"] = dec.decode(data, bits, w, h)
        assert (out[f"f{t}_rgb"] == st.convert_out()).all()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, os.path.getsize(os.path.join(HERE, name + ".npz")), "bytes")


if __name__ == "__main__":
    for case in CASES:
        make(*case)
