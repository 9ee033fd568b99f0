#!/usr/bin/env python3
"""Golden vectors for RDOQ with frozen contexts, generated from the UNMODIFIED reference:
RdoQuant::QuantRdo (rdo_quant.cc:203-446) against the context state a picture starts from (oracle/ref_shim.cc,
xref_quant_rdo_frozen).  This is the definition a GPU kernel has to reproduce bit for bit; no kernel exists yet.

    python tests/golden/make_rdoq_golden.py        (in the build container: needs /root/reference -> oracle/_ref)
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import bindings  # noqa: E402
from xvc_b200 import workload  # noqa: E402

CASES = [   # (w, h, comp, qp, intra picture, intra CU, bitdepth)
    (8, 8, 0, 32, 0, 0, 10), (16, 16, 0, 27, 0, 0, 10), (32, 32, 0, 32, 0, 0, 10), (64, 64, 0, 37, 0, 0, 10),
    (16, 8, 0, 32, 0, 0, 10), (8, 32, 0, 22, 0, 0, 10), (64, 32, 0, 32, 0, 0, 8), (4, 4, 0, 32, 1, 1, 10),
    (8, 8, 1, 32, 0, 0, 10), (16, 16, 2, 27, 0, 0, 10), (32, 32, 1, 37, 0, 0, 12), (4, 8, 1, 32, 0, 0, 10),
    (16, 16, 0, 32, 1, 1, 10), (32, 32, 0, 27, 1, 1, 10), (8, 8, 0, 42, 0, 0, 10), (2, 2, 1, 32, 0, 0, 10),
]


def coefficients(rng, w, h, bd, scale):
    """Transform-coefficient-like blocks: magnitudes decaying with frequency, a few outliers."""
    yy, xx = np.mgrid[0:h, 0:w]
    c = rng.normal(0.0, 1.0, (h, w)) * scale * (1 << (bd - 8)) / (1.0 + 0.5 * (xx + yy))
    c[rng.random((h, w)) < 0.02] *= 6.0
    return np.clip(np.rint(c), -32768, 32767).astype(np.int16)


def main():
    ref = bindings.Ref()
    rng = np.random.default_rng(20260)
    arrays, cases = {}, []
    for k, (w, h, comp, qp, intra_pic, intra_cu, bd) in enumerate(CASES):
        for scale in (40.0, 400.0):
            coeff = coefficients(rng, w, h, bd, scale)
            lam = workload.lambda_for_qp(qp)
            lev, nz = ref.quant_rdo_frozen(w, h, bd, comp, qp, lam, intra_pic, coeff, intra_cu=intra_cu, intra_mode=1 if intra_cu else 0)
            name = "c%02d_%d" % (k, int(scale))
            arrays[name + "_in"], arrays[name + "_out"] = coeff, lev
            cases.append(dict(name=name, w=w, h=h, comp=comp, qp=qp, lam=lam, intra_pic=intra_pic, intra_cu=intra_cu, bitdepth=bd, ret=int(nz)))
    arrays["__cases__"] = np.frombuffer(json.dumps(cases).encode(), dtype=np.uint8)
    np.savez_compressed(os.path.join(HERE, "xvc_rdoq_golden.npz"), **arrays)
    print("wrote %d cases" % len(cases))


if __name__ == "__main__":
    main()
