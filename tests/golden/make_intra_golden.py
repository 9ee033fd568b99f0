#!/usr/bin/env python3
"""Generates tests/golden/xvc_intra_golden.npz from the UNMODIFIED reference (oracle/_ref/libxvcref.so,
built by oracle/Makefile from /root/reference): for small pictures walked in coding order, the
neighbour availability, reference samples (IntraPrediction::FillReferenceState), the prediction of
all 67 modes (IntraPrediction::Predict) and the luma SATD scan (SampleMetric kSatd).
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
from xvc_b200 import abi, workload  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "xvc_intra_golden.npz")


def main():
    ref = bindings.Ref()
    arrays, cases = {}, []
    for name, (width, height, bd, content, min_size, seed) in {
        "synth10": (72, 40, 10, "synth", 4, 5), "rand8": (40, 40, 8, "random", 4, 6), "rand12": (64, 64, 12, "random", 16, 7),
    }.items():
        cur, rec, _ = common.frames(width, height, bd, seed, content)
        cus = workload.make_partition(width, height, seed=seed, min_size=min_size)
        cus["flags"] |= abi.CU_INTRA
        for comp in (0, 1):
            ses = ref.session(width, height, bd, pic_type=2)
            ses.set_orig(cur)
            ses.set_rec(rec)
            jobs, r, f, preds, satd = ses.intra_scan(cus, comp)
            ses.close()
            key = "%s_c%d" % (name, comp)
            for i in range(3):
                arrays[key + "_cur%d" % i] = cur[i]
                arrays[key + "_rec%d" % i] = rec[i]
            arrays[key + "_jobs"] = jobs.view(np.uint8)
            arrays[key + "_ref"], arrays[key + "_filt"] = r, f
            arrays[key + "_pred"] = np.concatenate([p.reshape(-1) for p in preds])
            arrays[key + "_satd"] = satd
            cases.append(dict(name=key, bd=bd, comp=comp, width=width, height=height, pred_stride=3))
    arrays["__cases__"] = np.frombuffer(json.dumps(cases).encode(), dtype=np.uint8)
    np.savez_compressed(OUT, **arrays)
    print("wrote %s: %d cases, %.1f KB" % (OUT, len(cases), os.path.getsize(OUT) / 1024))


if __name__ == "__main__":
    main()
