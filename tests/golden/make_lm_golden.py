#!/usr/bin/env python3
"""Generates tests/golden/xvc_lm_golden.npz from the UNMODIFIED reference (oracle/_ref/libxvcref.so, built by
oracle/Makefile from /root/reference): IntraPrediction::Predict(kLmChroma) (PredLmChroma,
intra_prediction.cc:560-686) of both chroma blocks of every CU of small pictures.  Run in the development
container only; the .npz is committed and replayed without the reference."""
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

OUT = os.path.join(ROOT, "tests", "golden", "xvc_lm_golden.npz")


def main():
    ref = bindings.Ref()
    arrays, cases = {}, []
    for name, (width, height, bd, content, seed) in {
        "synth10": (136, 72, 10, "synth", 81), "rand8": (72, 136, 8, "random", 82), "synth12": (200, 136, 12, "synth", 83),
    }.items():
        _, rec, _ = common.frames(width, height, bd, seed, content)
        rng = np.random.default_rng(seed)
        rec = [rec[0]] + [np.clip(p.astype(np.int32) + rng.integers(-6, 7, size=p.shape) * (1 << (bd - 8)), 0, (1 << bd) - 1).astype(np.uint16)
                          for p in rec[1:]]
        cus = workload.make_partition(width, height, seed=seed, min_size=8)
        cus["flags"] |= abi.CU_INTRA
        ses = ref.session(width, height, bd, pic_type=2)
        ses.set_rec(rec)
        blocks = ses.intra_lm_chroma(cus)
        ses.close()
        pred = [np.zeros_like(rec[1]), np.zeros_like(rec[2])]
        for cu, (bu, bv) in zip(cus, blocks):
            x, y, w, h = int(cu["x"]) // 2, int(cu["y"]) // 2, int(cu["w"]) // 2, int(cu["h"]) // 2
            pred[0][y:y + h, x:x + w], pred[1][y:y + h, x:x + w] = bu, bv
        for i in range(3):
            arrays["%s_rec_%d" % (name, i)] = rec[i]
        arrays[name + "_pred_1"], arrays[name + "_pred_2"] = pred
        arrays[name + "_cus"] = cus.view(np.uint8)
        cases.append(dict(name=name, bd=bd, width=width, height=height, n=len(cus)))
    arrays["__cases__"] = np.frombuffer(json.dumps(cases).encode(), dtype=np.uint8)
    np.savez_compressed(OUT, **arrays)
    print("wrote %s: %s, %.1f KB" % (OUT, cases, os.path.getsize(OUT) / 1024))


if __name__ == "__main__":
    main()
