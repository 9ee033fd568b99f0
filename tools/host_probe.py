#!/usr/bin/env python3
"""Host-side cost of the per-picture calls (is the host keeping up with the GPU?)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
from xvc_b200 import lib

frames, cus, prm, lam = bench.picture_inputs()
ctx = lib.Context(bench.WIDTH, bench.HEIGHT, bench.BITDEPTH, num_slots=6)
prm["orig_slot"], prm["pred_slot"], prm["rec_slot"], prm["coeff_slot"] = 0, 3, 5, 4
prm["ref_slots"][0, 0, 0], prm["ref_slots"][0, 1, 0] = 1, 2
ctx.upload(0, frames[0])
for s, f in ((1, frames[1]), (2, frames[2])):
    ctx.upload(s, f); ctx.pad_border(s)
for _ in range(3):
    ctx.set_cus(cus); ctx.encode_picture(prm, want_results=False)
ctx.sync()
K = 50
t_set = t_enc = 0.0
t0 = time.perf_counter()
for _ in range(K):
    a = time.perf_counter(); ctx.set_cus(cus); b = time.perf_counter(); ctx.encode_picture(prm, want_results=False); c = time.perf_counter()
    t_set += b - a; t_enc += c - b
t1 = time.perf_counter()
ctx.sync()
t2 = time.perf_counter()
print("host per picture: set_cus %.3f ms, encode_picture (enqueue) %.3f ms, loop %.3f ms; wall incl. GPU %.3f ms per picture" %
      (1e3 * t_set / K, 1e3 * t_enc / K, 1e3 * (t1 - t0) / K, 1e3 * (t2 - t0) / K))
