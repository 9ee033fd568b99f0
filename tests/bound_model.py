#!/usr/bin/env python3
"""CPU model of the raster bound pass of tz_search_kernel on the bench picture (numpy; the oracle
gives the final best cost of each sampled job): what fraction of the raster candidates is still
below the threshold after HALF of the rows of the segment-sum bound, against the fraction that
survives the whole bound.  Decision aid for the two-stage bound of DESIGN.md section 8; thresholds use
the FINAL best cost (<= the cost the kernel holds when the raster starts) and ignore the MV rate, so
both fractions are brackets, their ratio is what matters.  Analysis tooling kept under tests/ because it uses the oracle (test infrastructure); not collected by pytest."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))   # lives under tests/: it uses the oracle (test infrastructure)
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
from oracle.bindings import Oracle, Picture  # noqa: E402
from xvc_b200 import abi  # noqa: E402


def main(n_jobs=80, seed=3):
    W, H, BD = bench.WIDTH, bench.HEIGHT, bench.BITDEPTH
    frames, cus, prm, lam = bench.picture_inputs()
    o = Oracle()
    orig = Picture(W, H, 0, frames[0])
    refs = {(0, 0): Picture(W, H, 80, frames[1]), (1, 0): Picture(W, H, 80, frames[2])}
    for r in refs.values():
        o.pad_border(r)
    rng = np.random.default_rng(seed)
    pick = rng.choice(len(cus), size=n_jobs, replace=False)
    jobs = np.zeros(n_jobs, dtype=abi.me_job_dtype)
    jobs["cu"], jobs["list"], jobs["ref_slot"] = pick, 0, 0
    jobs["search_range"] = int(prm["search_range"][0, 0, 0])
    res = o.me_search(orig, refs, BD, cus, jobs, float(prm["lambda_sqrt"][0]))
    ref = refs[(0, 0)].full[0].astype(np.int64)
    pad = 80
    # S8[y, x] = sum of 8 samples to the right
    cs = np.concatenate([np.zeros((ref.shape[0], 1), dtype=np.int64), np.cumsum(ref, axis=1)], axis=1)
    s8 = cs[:, 8:] - cs[:, :-8]
    tot = full = half = quarter = 0
    rng_r = int(jobs["search_range"][0])
    for k in range(n_jobs):
        cu = cus[pick[k]]
        x, y, w, h = int(cu["x"]), int(cu["y"]), int(cu["w"]), int(cu["h"])
        if w < 8:
            continue
        fast = h > 8
        rows = np.arange(0, h, 2 if fast else 1)
        blk = frames[0][0][y:y + h, x:x + w].astype(np.int64)
        a8 = blk.reshape(h, w // 8, 8).sum(axis=2)[rows]                     # rows x segs
        thr = int(res[k]["cost_fullpel"]) << (BD - 8)
        if fast:
            thr = (thr + 1) >> 1
        lo_x, hi_x = max(-rng_r, -(64 + 8 + x - 1)), min(rng_r, W + 8 - x - 1)
        lo_y, hi_y = max(-rng_r, -(64 + 8 + y - 1)), min(rng_r, H + 8 - y - 1)
        xs, ys = np.arange(lo_x, hi_x + 1, 5), np.arange(lo_y, hi_y + 1, 5)
        # bound per candidate, accumulated row by row
        lb = np.zeros((len(ys), len(xs)), dtype=np.int64)
        snap = {}
        for ri, r in enumerate(rows):
            for sg in range(w // 8):
                yy = pad + y + ys[:, None] + r
                xx = pad + x + xs[None, :] + 8 * sg
                lb += np.abs(s8[yy, xx] - a8[ri, sg])
            if ri + 1 in (max(1, len(rows) // 4), len(rows) // 2):
                snap[ri + 1] = (lb < thr).sum()
        n = lb.size
        tot += n
        full += (lb < thr).sum()
        half += snap.get(len(rows) // 2, 0)
        quarter += snap.get(max(1, len(rows) // 4), 0)
    print("raster candidates %d of %d sampled jobs (bench picture, list 0)" % (tot, n_jobs))
    print("below threshold after 1/4 of the rows: %.2f %%" % (100.0 * quarter / tot))
    print("below threshold after 1/2 of the rows: %.2f %%" % (100.0 * half / tot))
    print("below threshold after all rows (survivors of the bound): %.2f %%   (kernel, with rate and cost_in: 3.30 %%)" % (100.0 * full / tot))


if __name__ == "__main__":
    main()
