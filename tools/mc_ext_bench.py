#!/usr/bin/env python3
"""Device timing of the three motion-compensation entries on one 1080p picture (the partition of
bench.py, every CU uni/bi-predicted as tests/common.mc_cus): plain, affine (every CU with w, h > 8),
LIC (every CU).  CUDA events on the context stream, L2 flushed between iterations."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import common  # noqa: E402
import torch  # noqa: E402
from xvc_b200 import lib, workload  # noqa: E402


def main():
    W, H, BD = 1920, 1080, 10
    canvas = workload.synth_canvas(W, H, 1234)
    cur, r0, r1 = [workload.synth_frame(canvas, W, H, i, BD) for i in (8, 0, 16)]
    rng = np.random.default_rng(5)
    cus = common.mc_cus(W, H, rng, 7, min_size=8)
    cus["mv"] = np.clip(cus["mv"], -1000, 1000)
    aff = common.affine_cus(cus, rng)
    lic = common.lic_cus(cus, W, H)
    ctx = lib.Context(W, H, BD, num_slots=5)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    for slot, f in ((1, r0), (2, r1)):
        ctx.upload(slot, f)
        ctx.pad_border(slot)
    ctx.upload(4, cur)
    ctx.set_cus(cus)
    refs = {(0, 0): 1, (1, 0): 2}
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    px = lambda idx: int(sum(int(cus[i]["w"]) * int(cus[i]["h"]) for i in idx))   # noqa: E731
    runs = {"motion_compensate": (lambda: ctx.motion_compensate(refs, 3), px(range(len(cus)))),
            "motion_compensate_affine": (lambda: ctx.motion_compensate_affine(aff, refs, 3), px(aff["cu"])),
            "motion_compensate_lic": (lambda: ctx.motion_compensate_lic(lic, refs, 4, 3), px(lic["cu"]))}
    for name, (fn, luma_px) in runs.items():
        times = []
        for it in range(8):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            fn()
            e1.record(stream)
            e1.synchronize()
            if it >= 3:
                times.append(e0.elapsed_time(e1))
        ms = float(np.mean(times))
        # algorithmic bytes: reference samples read once + prediction written, 1.5 samples per luma px, 2 B each
        gbs = 2 * 1.5 * luma_px * 2 / (ms * 1e-3) / 1e9
        print("%-28s %6d CUs %8.1f kpx  %.3f ms  %.0f Mpx/s  %.0f GB/s algorithmic" % (name, len(cus), luma_px / 1e3, ms, luma_px / ms / 1e3, gbs))


if __name__ == "__main__":
    main()
