#!/usr/bin/env python3
"""Generates tests/golden/xvc_affine_golden.npz from the UNMODIFIED reference (oracle/_ref/libxvcref.so,
built by oracle/Makefile from /root/reference): small pictures whose CUs are predicted by
InterPrediction::MotionCompensation with SetUseAffine(true) (MotionCompAffine,
inter_prediction.cc:1044-1136), C filter table (simd=0: no overshoot of 2-wide chroma sub-blocks).
Run in the development container only; the .npz is committed and replayed without the reference."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import common  # noqa: E402
from oracle import bindings  # noqa: E402
from xvc_b200 import workload  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "xvc_affine_golden.npz")


def main():
    ref = bindings.Ref()
    arrays, cases = {}, []
    for name, (width, height, bd, content, seed) in {
        "rand10": (136, 72, 10, "random", 31), "synth8": (200, 136, 8, "synth", 32), "rand12": (72, 136, 12, "random", 33),
    }.items():
        cur, r0, r1 = common.frames(width, height, bd, seed, content)
        rng = np.random.default_rng(seed)
        cus = common.mc_cus(width, height, rng, seed, min_size=8)
        aff = common.affine_cus(cus, rng)
        ses = ref.session(width, height, bd, 0, 32, workload.lambda_for_qp(32), simd=0, poc=8, sub_gop=16)
        ses.set_orig(cur)
        ses.add_ref(0, 0, 0, r0)
        ses.add_ref(1, 0, 16, r1)
        ses.set_cus(cus)
        ses.motion_compensate(threads=1)
        ses.motion_compensate_affine(aff, threads=1)
        pred = ses.get_pred()
        ses.close()
        for i in range(3):
            arrays["%s_r0_%d" % (name, i)], arrays["%s_r1_%d" % (name, i)] = r0[i], r1[i]
            arrays["%s_pred_%d" % (name, i)] = pred[i]
        arrays[name + "_cus"] = cus.view(np.uint8)
        arrays[name + "_aff"] = aff.view(np.uint8)
        cases.append(dict(name=name, bd=bd, width=width, height=height, n_aff=len(aff)))
    arrays["__cases__"] = np.frombuffer(json.dumps(cases).encode(), dtype=np.uint8)
    np.savez_compressed(OUT, **arrays)
    print("wrote %s: %s, %.1f KB" % (OUT, cases, os.path.getsize(OUT) / 1024))


if __name__ == "__main__":
    main()
