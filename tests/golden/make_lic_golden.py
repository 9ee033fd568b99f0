#!/usr/bin/env python3
"""Generates tests/golden/xvc_lic_golden.npz from the UNMODIFIED reference (oracle/_ref/libxvcref.so,
built by oracle/Makefile from /root/reference): small pictures whose CUs are predicted by
InterPrediction::MotionCompensation with SetUseLic(true) (LocalIlluminationComp / DeriveLicParams,
inter_prediction.cc:1555-1673), C filter table.  Run in the development container only; the .npz is
committed and replayed without the reference."""
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

OUT = os.path.join(ROOT, "tests", "golden", "xvc_lic_golden.npz")


def main():
    ref = bindings.Ref()
    arrays, cases = {}, []
    for name, (width, height, bd, content, seed, min_size) in {
        "synth10": (136, 72, 10, "synth", 41, 4), "rand8": (200, 136, 8, "random", 42, 8), "synth12": (72, 136, 12, "synth", 43, 4),
    }.items():
        cur, r0, r1 = common.frames(width, height, bd, seed, content)
        rng = np.random.default_rng(seed)
        cus = common.mc_cus(width, height, rng, seed, min_size=min_size)
        rec = [np.clip(p.astype(np.int32) * 7 // 8 + (3 << (bd - 8)) + rng.integers(-2, 3, size=p.shape), 0, (1 << bd) - 1).astype(np.uint16)
               for p in cur]
        lic = common.lic_cus(cus, width, height)
        ses = ref.session(width, height, bd, 0, 32, workload.lambda_for_qp(32), simd=0, poc=8, sub_gop=16)
        ses.set_orig(cur)
        ses.set_rec(rec)
        ses.add_ref(0, 0, 0, r0)
        ses.add_ref(1, 0, 16, r1)
        ses.set_cus(cus)
        assert np.array_equal(lic, ses.lic_neighbours(lic["cu"]))
        ses.motion_compensate_lic(lic)
        pred = ses.get_pred()
        ses.close()
        for i in range(3):
            arrays["%s_r0_%d" % (name, i)], arrays["%s_r1_%d" % (name, i)] = r0[i], r1[i]
            arrays["%s_rec_%d" % (name, i)], arrays["%s_pred_%d" % (name, i)] = rec[i], pred[i]
        arrays[name + "_cus"] = cus.view(np.uint8)
        arrays[name + "_lic"] = lic.view(np.uint8)
        cases.append(dict(name=name, bd=bd, width=width, height=height, n_lic=len(lic)))
    arrays["__cases__"] = np.frombuffer(json.dumps(cases).encode(), dtype=np.uint8)
    np.savez_compressed(OUT, **arrays)
    print("wrote %s: %s, %.1f KB" % (OUT, cases, os.path.getsize(OUT) / 1024))


if __name__ == "__main__":
    main()
