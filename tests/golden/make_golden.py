#!/usr/bin/env python3
"""tests/golden/make_golden.py -- regenerates tests/golden/golden_v1.npz.

Runs the REAL reference (oracle/_ref/libqcsref_seq.so = unmodified sequential
mode, libqcsref_corrected.so = D1-D3 repaired; both compiled from
/root/reference by `make -C oracle ref`) on a fixed set of gate scripts and
stores every final amplitude, scratch amplitude, measurement outcome and shot
histogram.  The .npz is committed; /root/reference is only needed to re-run
this script.  Usage:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pyoracle as po  # noqa: E402
from tests.golden.cases import CASES  # noqa: E402


def main():
    po.build_ref()
    out = {}
    for name, (n, script) in CASES.items():
        for sem, mode in (("reference", "seq"), ("corrected", "corrected")):
            ref = po.RefLib(n, mode)
            vals = po.replay(ref, script)
            key = f"{name}/{sem}"
            out[key + "/state"] = ref.state()
            out[key + "/scratch"] = ref.scratch()
            for k, (kind, v) in enumerate(vals):
                out[f"{key}/out{k}_{kind}"] = np.asarray(v)
            ref.close()
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_v1.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: {len(out)} arrays, {os.path.getsize(path)} bytes")


if __name__ == "__main__":
    main()
