#!/usr/bin/env python3
"""Where the end-to-end pipeline spends its time: variants of bench.py's e2e loop (manual probe)."""
import os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from xvc_b200 import abi, lib

frames, cus, prm, lam = bench.picture_inputs(index_offset=0)
n = len(cus)
ctx = lib.Context(bench.WIDTH, bench.HEIGHT, bench.BITDEPTH, num_slots=9, device=0)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); ctx.set_stream(stream.cuda_stream)
for slot, f in ((1, frames[1]), (2, frames[2])):
    ctx.upload(slot, f); ctx.pad_border(slot)
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
h_orig = [pin(p) for p in frames[0]]
h_rec = [[pin(np.zeros_like(p)) for p in frames[0]] for _ in range(2)]
h_lev = [[pin(np.zeros(p.shape, dtype=np.int16)) for p in frames[0]] for _ in range(2)]
h_cus = [pin(np.zeros(n, dtype=abi.cu_dtype).view(np.uint8)).view(abi.cu_dtype) for _ in range(2)]
sets = [dict(orig=0, coeff=4, rec=5), dict(orig=6, coeff=7, rec=8)]
prms = []
for st in sets:
    q = prm.copy(); q["orig_slot"], q["coeff_slot"], q["rec_slot"], q["pred_slot"] = st["orig"], st["coeff"], st["rec"], 3
    q["ref_slots"][0, 0, 0], q["ref_slots"][0, 1, 0] = 1, 2
    prms.append(q)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

def run(count, up=True, down=True, fl=True, setcus=True, cusdl=True, picdl=True):
    if up: ctx.upload_async(sets[0]["orig"], h_orig)
    host_t = 0.0
    global EV
    EV = []
    for i in range(count):
        s = i & 1
        h0 = time.perf_counter()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True); e2 = torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        if up and i + 1 < count: ctx.upload_async(sets[1 - s]["orig"], h_orig)
        if setcus: ctx.set_cus(cus)
        if fl: flush.zero_()
        e1.record(stream)
        ctx.encode_picture(prms[s], want_results=False)
        if down:
            if cusdl: ctx.get_cus_async(h_cus[s])
            if picdl: ctx.download_coeff_async(sets[s]["coeff"], h_lev[s])
            ctx.download_async(sets[s]["rec"], h_rec[s])
        e2.record(stream)
        EV.append((e0, e1, e2))
        host_t += time.perf_counter() - h0
        if down and i >= 1: ctx.wait_download(sets[1 - s]["rec"])
    if down: ctx.wait_download(sets[(count - 1) & 1]["rec"])
    torch.cuda.synchronize()
    return host_t

ctx.upload(0, frames[0]); ctx.upload(6, frames[0]); ctx.set_cus(cus)
ctx.set_profiling(True)
for name, kw in (("full", {}), ("noup", dict(up=False)), ("noup_nocusdl", dict(up=False, cusdl=False)),
                 ("noup_reconly", dict(up=False, cusdl=False, picdl=False)), ("nodown", dict(down=False)),
                 ("notransfer", dict(up=False, down=False))):
    run(4, **kw)
    t0 = time.perf_counter(); ht = run(40, **kw); dt = time.perf_counter() - t0
    pre = np.mean([a.elapsed_time(b) for a, b, c in EV[2:]]); enc = np.mean([b.elapsed_time(c) for a, b, c in EV[2:]])
    gap = np.mean([EV[i][2].elapsed_time(EV[i + 1][0]) for i in range(2, len(EV) - 1)])
    print("   setcus+flush %.3f  encode+pack %.3f  gap to next %.3f" % (pre, enc, gap))
    print("%-30s %.3f ms/picture   host enqueue %.3f ms/picture  stages %s" % (name, dt / 40 * 1e3, ht / 40 * 1e3, {k: round(v, 3) for k, v in ctx.stage_times_ms().items()}))
t0 = time.perf_counter()
for _ in range(40): flush.zero_()
torch.cuda.synchronize()
print("flush alone %.3f ms" % ((time.perf_counter() - t0) / 40 * 1e3))

# raw PCIe rates with the same pinned buffers
buf_d = torch.empty(16 << 20, dtype=torch.uint8, device="cuda")
buf_h = torch.empty(16 << 20, dtype=torch.uint8).pin_memory()
for name, fn in (("H2D", lambda: buf_d.copy_(buf_h, non_blocking=True)), ("D2H", lambda: buf_h.copy_(buf_d, non_blocking=True))):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(20): fn()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 20
    print("%s 16 MiB pinned: %.3f ms  %.1f GB/s" % (name, dt * 1e3, (16 << 20) / dt / 1e9))
# the library's own transfers, alone
for name, fn in (("upload_async", lambda: ctx.upload_async(0, h_orig)), ("download_async", lambda: ctx.download_async(5, h_rec[0]))):
    fn(); ctx.sync_copies()
    t0 = time.perf_counter()
    for _ in range(20): fn()
    host = (time.perf_counter() - t0) / 20
    ctx.sync_copies()
    dt = (time.perf_counter() - t0) / 20
    print("%s 6.2 MB: %.3f ms (host call %.3f ms)" % (name, dt * 1e3, host * 1e3))
