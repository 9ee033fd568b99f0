#!/usr/bin/env python3
"""bench.py -- the xvc hot path on B200: encoded Mpixels/s at 1080p qp32.

A step = one bi-predicted picture through the whole hot path
    partition pre-analysis on the GPU (encode workload: SAD tree over the quad tree, xvcb200_decide_partition)
    -> InterSearch::SearchMotion for every CU (TZ full-pel + sub-pel search on every reference picture of
    both lists, SearchBiIterative: FullSearch + sub-pel on the weighted original, uni / bi decision)
    -> motion compensation (uni and bi-predicted CUs)
    -> residual / forward transform / QuantFast / dequant / inverse transform / reconstruction
    -> deblocking -> border padding
of a synthetic 1920x1080 4:2:0 picture, 10-bit internal, qp 32.
--workload encode (default): panned background + objects with their own motion + camera noise; the CU partition
is decided from the content by the pre-analysis inside the timed step; xvc's default reference lists at POC 8
of a sub-GOP of 16 (two pictures per list), one bi-prediction pass, the pre-analysis vectors (scaled by POC
distance) as predictors.  The reference arm codes the same partition with the same predictors (it gets
them from the numpy statement of the pre-analysis rule, outside its timed region).
--workload raster: round 1's step (seeded random partition, one picture per list, zero predictors -> the raster scan of the
+-128 window fires for most CUs, list chosen by the sub-pel cost, noise-free pan).

  python bench.py [--gpus N] [--steps K] [--warmup W]      our arm (CUDA through the C ABI)
  python bench.py --impl reference ...                      the reference's own CPU code on the
                                                            host cores (oracle/_ref), same step

Prints ONE JSON line (rank 0).  `value` is device-resident throughput (inputs already in HBM),
`e2e` the same step through the C ABI with host buffers (H2D of the picture + CU array, D2H of
reconstruction, levels and CU decisions inside the timed region).  Beside them (--gop on, the default):
`raster_workload` (N = 1: round 1's step on the same context), `gop` (BASELINE config 5: a hierarchical-B
sequence of max(2, N) sub-GOPs encoded frame-parallel, waves across sub-GOPs, reconstructions pushed between
the GPUs and referenced after a rendezvous -- on the device through arrival tags at N = 2, a host barrier per
wave at N > 2, XVCB_GOP_RENDEZVOUS=device / host overrides) and, at N > 1, `banded` (config 4: one picture in CTU-row bands with the deblocking
halo exchange).  The line is complete before `gop` / `banded` start; they run under a watchdog
(XVCB_BENCH_EXTRAS_TIMEOUT seconds, default 180): if they do not finish, the line is printed with the reason.
"""
import argparse
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from xvc_b200 import abi, workload  # noqa: E402

WIDTH, HEIGHT, BITDEPTH, QP = 1920, 1080, 10, 32
POC, REF_POCS, SUB_GOP = 8, (0, 16), 16
METRIC = "encoded Mpixels/sec at 1080p qp32; bit-exact recon vs reference"


WORKLOAD = "encode"
# reference lists (POCs) per workload: xvc's default two pictures per list at POC 8 (list 1 repeats list 0's
# pictures in the opposite order, ReferenceListSorter), or one picture per list
LISTS = {"encode": ((0, 16), (16, 0)), "raster": ((0,), (16,))}


def picture_inputs(index_offset=0, seed=1234, partition_seed=7):
    """(current, ref POC 0, ref POC 16) planes + CU partition (raster workload; None: decided by the pre-analysis)
    + picture parameters of one step."""
    enc = WORKLOAD == "encode"
    canvas = workload.synth_canvas(WIDTH, HEIGHT, seed)
    frames = []
    for i in (POC, REF_POCS[0], REF_POCS[1]):
        f = [p.copy() for p in workload.synth_frame(canvas, WIDTH, HEIGHT, i + index_offset, BITDEPTH, frame_noise=4.0 if enc else 0.0)]
        frames.append(tuple(workload.add_objects(f, WIDTH, HEIGHT, i + index_offset, BITDEPTH)) if enc else tuple(f))
    cus = None if enc else workload.make_partition(WIDTH, HEIGHT, seed=partition_seed, min_size=8, qp=QP)
    lam = workload.lambda_for_qp(QP)
    prm = np.zeros(1, dtype=abi.picture_params_dtype)
    prm["pic_type"] = 0
    for l, pocs in enumerate(LISTS[WORKLOAD]):
        prm["num_ref"][0, l] = len(pocs)
        for r, p in enumerate(pocs):
            prm["search_range"][0, l, r] = workload.search_range_uni(POC, p, SUB_GOP)
            prm["ref_poc"][0, l, r] = p
    prm["lambda_sqrt"] = np.sqrt(lam)
    prm["chroma_offset_table"] = 1
    prm["deblock"], prm["pad"] = 1, 1
    prm["bi_iterations"], prm["bits_mode"] = (1, 1) if enc else (0, 0)
    prm["ref_slots"] = -1
    return frames, cus, prm, lam


def global_motion(index_offset=0):
    """Picture-level predictor of the pre-analysis: the content's global motion towards POC 0 (an encoder takes it
    from the previous picture's vectors), 1/16 pel."""
    return workload.true_motion(POC + index_offset, REF_POCS[0] + index_offset)


def predictors_from_partition(cus, index_offset=0):
    """Predictor per (CU, list, reference picture) from the pre-analysis vector of the CU (towards POC 0), scaled by
    POC distance like the reference's neighbour predictors (POC 16 lies as far ahead as POC 0 lies behind: -1)."""
    if WORKLOAD != "encode":
        return None
    mv0 = cus["mv"][:, 0, :].astype(np.int32)
    cols = [mv0 if p == REF_POCS[0] else -mv0 for l in LISTS[WORKLOAD] for p in l]
    return np.ascontiguousarray(np.stack(cols, axis=1).astype(np.int32))


def cpu_partition(frames, lam):
    """The reference arm's partition: the numpy statement of the pre-analysis rule (tests/partition_model.py)."""
    import partition_model
    cus, _ = partition_model.decide(frames[0][0], frames[1][0], global_motion(), float(np.sqrt(lam)), QP)
    return cus


def set_ref_slots(prm, slot_of_poc):
    for l, pocs in enumerate(LISTS[WORKLOAD]):
        for r, p in enumerate(pocs):
            prm["ref_slots"][0, l, r] = slot_of_poc[p]


def recon_digest(planes):
    h = hashlib.md5()
    for p in planes:
        h.update(np.ascontiguousarray(p).tobytes())
    return h.hexdigest()


def config_dict(n_cus, extra=None):
    cfg = {"workload": "1920x1080 synthetic YUV420 qp32, ME+transform+deblock (configs[1])" if (WIDTH, HEIGHT, QP) == (1920, 1080, 32)
           else "%dx%d synthetic YUV420 qp%d, ME+transform+deblock (context run, not the metric's configuration)" % (WIDTH, HEIGHT, QP),
           "width": WIDTH, "height": HEIGHT, "bitdepth_internal": BITDEPTH, "qp": QP,
           "picture": ("bi-predicted picture, POC %d of a sub-GOP of 16; reference lists L0 = POC %s, L1 = POC %s (xvc's default two pictures "
                       "per list; the list-1 pictures repeat list 0's and are searched once), search range %d, one SearchBiIterative pass "
                       "(FullSearch +-4 + sub-pel search on the weighted original), uni/bi decision by GetInterPredBits(fast_inter_pred_bits); "
                       "camera noise N(0,4) on every frame" if WORKLOAD == "encode" else
                       "bi-predicted picture type, one reference picture per list (POC %d, refs %s / %s), search range %d, list chosen by the "
                       "sub-pel cost (round 1's step), noise-free pan") % (
               POC, "/".join(map(str, LISTS[WORKLOAD][0])), "/".join(map(str, LISTS[WORKLOAD][1])), workload.search_range_uni(POC, REF_POCS[0], SUB_GOP)),
           "partition": ("decided from the content by the pre-analysis (SAD tree, +-8 around the picture's global motion; ours: on the GPU inside every "
                         "timed step, reference arm: the numpy statement of the same rule outside its timed region), %d CUs; one predictor per (CU, list, "
                         "reference picture) = the CU's pre-analysis vector scaled by POC distance" % n_cus) if WORKLOAD == "encode" else
                        ("seeded random quad/binary CU tree, 8..64, %d CUs; predictor mvp = 0 for every CU (the raster scan of the window fires for most jobs)" % n_cus),
           "content": "panned background + %s" % ("objects with their own motion + camera noise N(0,4)" if WORKLOAD == "encode" else "nothing else"),
           "workload_name": WORKLOAD,
           "l2": "flushed between timed iterations (256 MiB write)"}
    if extra:
        cfg.update(extra)
    return cfg


# ------------------------------------------------------------------------------------------
# reference arm: the reference's own CPU implementation of the same step
# ------------------------------------------------------------------------------------------
class CpuArm:
    """oracle/_ref (the unmodified reference compiled by oracle/Makefile) when it was built,
    else the C restatement.  Test infrastructure used as a timed baseline only."""

    def __init__(self):
        from oracle import bindings
        self.b = bindings
        self.kind = "reference" if bindings.have_ref() else "port"
        if self.kind == "reference":
            self.ref = bindings.Ref()
            self.cores = int(self.ref.L.xref_num_threads())
        else:
            self.oracle = bindings.Oracle()
            self.cores = 1

    def run(self, frames, cus, prm, lam, mvp=None):
        """One step; returns (seconds, reconstruction planes)."""
        if self.kind == "reference":
            s = self.ref.session(WIDTH, HEIGHT, BITDEPTH, 0, QP, lam, simd=1, poc=POC, sub_gop=SUB_GOP)
            s.set_orig(frames[0])
            by_poc = {REF_POCS[0]: frames[1], REF_POCS[1]: frames[2]}
            for l, pocs in enumerate(LISTS[WORKLOAD]):
                for r, p in enumerate(pocs):
                    s.add_ref(l, r, p, by_poc[p])
            t0 = time.perf_counter()
            s.encode_picture(prm, cus, threads=self.cores, mvp=mvp)
            dt = time.perf_counter() - t0
            rec = s.get_rec()
            s.close()
            return dt, rec
        if WORKLOAD != "raster":
            raise RuntimeError("the C restatement runs the raster workload only; build oracle/_ref for the encode workload")
        P = self.b.Picture
        orig = P(WIDTH, HEIGHT, 0, frames[0])
        refs = {(0, 0): P(WIDTH, HEIGHT, 80, frames[1]), (1, 0): P(WIDTH, HEIGHT, 80, frames[2])}
        for r in refs.values():
            self.oracle.pad_border(r)
        pred, rec = P(WIDTH, HEIGHT, 80), P(WIDTH, HEIGHT, 80)
        t0 = time.perf_counter()
        self.oracle.encode_picture(orig, refs, pred, rec, BITDEPTH, cus.copy(), prm)
        dt = time.perf_counter() - t0
        return dt, rec.planes()


def xvcenc_cli_sample(n_frames=9):
    """Context number (SURVEY 8d "CPU reference timing"): the reference's own encoder application
    (oracle/_ref/xvcenc, the unmodified sources compiled by oracle/Makefile) with its full RDO mode
    decision, RDOQ and CABAC on a bounded sample -- configs[0]'s 352x288 at qp 32 from the same
    synthetic generator, all host threads (frame-parallel).  A 1080p picture costs that encoder
    ~90 core-seconds, so the named resolution does not fit a bench run.  NOT the same work as the
    step timed above (fixed partition, QuantFast, no entropy coding): reported beside it, never
    used for a ratio."""
    import re
    import tempfile
    exe = os.path.join(ROOT, "oracle", "_ref", "xvcenc")
    if os.environ.get("XVCB_BENCH_SKIP_XVCENC"):
        return {"unavailable": "skipped (XVCB_BENCH_SKIP_XVCENC)"}
    if not os.path.exists(exe):
        return {"unavailable": "oracle/_ref/xvcenc not built"}
    w, h = 352, 288
    threads = min(64, os.cpu_count() or 1)
    canvas = workload.synth_canvas(w, h, 1234)
    with tempfile.TemporaryDirectory() as d:
        yuv = os.path.join(d, "in.yuv")
        with open(yuv, "wb") as f:
            for i in range(n_frames):
                for p in workload.synth_frame(canvas, w, h, i, 8):
                    f.write(p.astype(np.uint8).tobytes())
        cmd = [exe, "-input-file", yuv, "-input-width", str(w), "-input-height", str(h), "-framerate", "30", "-qp", str(QP),
               "-max-pictures", str(n_frames), "-threads", str(threads), "-output-file", os.path.join(d, "out.xvc")]
        try:
            out = subprocess.run(cmd, capture_output=True, text=True, timeout=300).stdout
        except Exception as e:  # noqa: BLE001
            return {"unavailable": repr(e)}
    m = re.search(r"Total time:\s+([0-9.]+) s", out)       # the application's own clock (encoder_app.cc:566-568)
    if not m:
        return {"unavailable": "no 'Total time' line in the xvcenc output"}
    sec = float(m.group(1))
    return {"frames_per_s": n_frames / sec, "mpixels_per_s": n_frames * w * h / sec / 1e6, "seconds": sec, "threads": threads,
            "sample": "xvcenc -qp %d -threads %d, %d synthetic 352x288 frames (configs[0]), default (slow) preset: full RDO + RDOQ + CABAC" % (QP, threads, n_frames)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    frames, cus, prm, lam = picture_inputs()
    if cus is None:
        cus = cpu_partition(frames, lam)
    mvp = predictors_from_partition(cus)
    arm = CpuArm()
    for _ in range(args.warmup):
        arm.run(frames, cus, prm, lam, mvp)
    times = [arm.run(frames, cus, prm, lam, mvp)[0] for _ in range(args.steps)]
    sec = float(np.mean(times))
    mpx = WIDTH * HEIGHT / sec / 1e6
    line = {
        "impl": "reference", "metric": METRIC, "value": mpx, "unit": "Mpixels/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u16 samples / int32 arithmetic", "data": "synthetic", "config": config_dict(len(cus)),
        "cpu_baseline": {"value": mpx, "unit": "Mpixels/s", "cores": arm.cores, "kind": arm.kind,
                         "sample": "one full 1920x1080 picture per step, %d steps, all host threads over CUs" % args.steps},
        "e2e": {"value": mpx, "unit": "Mpixels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "xvcenc_cli": xvcenc_cli_sample(),
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    def __init__(self, device):
        super().__init__(daemon=True)
        self.device, self.samples, self.stop_flag = device, [], False

    def run(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.device)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
            while not self.stop_flag:
                sm = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
                reasons = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                self.samples.append((sm, reasons))
                time.sleep(0.02)
        except Exception as e:  # noqa: BLE001
            self.error = repr(e)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable: %s" % getattr(self, "error", "no samples")]}
        names = {0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown",
                 0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}
        seen = set()
        for _, r in self.samples:
            for bit, name in names.items():
                if r & bit:
                    seen.add(name)
        return {"sm_mhz": float(np.median([s for s, _ in self.samples])), "sm_max_mhz": float(self.max_mhz),
                "reasons": sorted(seen), "samples": len(self.samples)}


def raster_measurement(torch, ctx, stream, flush, SL, args):
    """The same context, round 1's step (--workload raster): seeded partition, one picture per list, zero predictors.
    Device-timed like the main number; the CPU reference beside it on a few pictures."""
    global WORKLOAD
    saved = WORKLOAD
    WORKLOAD = "raster"
    try:
        frames, cus, prm, lam = picture_inputs()
        prm["orig_slot"], prm["pred_slot"], prm["rec_slot"], prm["coeff_slot"] = SL["orig"], SL["pred"], SL["rec"], SL["coeff"]
        set_ref_slots(prm, {REF_POCS[0]: SL["ref0"], REF_POCS[1]: SL["ref1"]})
        ctx.upload(SL["orig"], frames[0])
        for slot, f in ((SL["ref0"], frames[1]), (SL["ref1"], frames[2])):
            ctx.upload(slot, f)
            ctx.pad_border(slot)
        ms = []
        for i in range(args.warmup + args.steps):
            ctx.set_cus(cus)
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            ctx.encode_picture(prm, want_results=False)
            e1.record(stream)
            e1.synchronize()
            if i >= args.warmup:
                ms.append(e0.elapsed_time(e1))
        rec = ctx.download(SL["rec"])
        arm = CpuArm()
        t_first, rec_cpu = arm.run(frames, cus, prm, lam)
        times = [arm.run(frames, cus, prm, lam)[0] for _ in range(4)]
        cpu = WIDTH * HEIGHT / float(np.mean(times)) / 1e6
        out = {"value": WIDTH * HEIGHT / (float(np.mean(ms)) * 1e-3) / 1e6, "unit": "Mpixels/s", "ms_per_step": float(np.mean(ms)), "steps": args.steps,
               "cpu_baseline": {"value": cpu, "unit": "Mpixels/s", "cores": arm.cores, "kind": arm.kind, "sample": "4 pictures of this step"},
               "recon_bitexact_vs_cpu_baseline": recon_digest(rec_cpu) == recon_digest(rec),
               "workload": config_dict(len(cus))["picture"] + "; " + config_dict(len(cus))["partition"]}
    finally:
        WORKLOAD = saved
    return out


def gop_measurement(torch, dist, lib, sharding, rank, world, local_rank, n_sub_gops):
    """BASELINE config 5's mechanism at this run's size: a hierarchical-B sequence (key picture + n_sub_gops sub-GOPs of
    16) encoded frame-parallel -- ThreadEncoder's rule (a picture starts when its reference pictures are finished),
    waves across sub-GOPs, picture j of a wave on GPU j mod N, every finished reconstruction pushed into the slot the
    same POC has on every GPU (CUDA IPC + copy engines over NVLink) and referenced there by later waves after a
    rendezvous.  Each picture = GPU partition pre-analysis + the whole step.  Returns the dict for the bench line
    (rank 0) -- throughput, per-wave times and the fraction of GPU time spent idle."""
    from xvc_b200 import gop
    pics = gop.hierarchical_gop(n_sub_gops)
    wave_in = gop.as_wave_input(pics)
    waves = sharding.gop_waves(wave_in, done=(0,))
    n_frames = 1 + 16 * n_sub_gops
    # one reconstruction slot per POC (no ring reuse: with 56 slots POC 56 landed in the key picture's slot, which the second
    # pass then referenced as POC 0; and a reused slot's arrival tags would have two producers)
    ring = n_frames + 1
    ctx = lib.Context(WIDTH, HEIGHT, BITDEPTH, num_slots=gop.GopEngine.FIRST_RING_SLOT + ring, device=local_rank)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    peers = sharding.PeerExchange(ctx, dist, rank, world) if dist is not None else None
    canvas = workload.synth_canvas(WIDTH, HEIGHT, 1234)
    mine = set()
    for wave in waves:
        for j, poc in enumerate(wave):
            if j % world == rank:
                mine.add(poc)

    def make(poc):      # the pan (and the objects) run forward for 31 frames, then backward: no jump in long sequences
        k = gop.seq_index(poc)
        f = [p.copy() for p in workload.synth_frame(canvas, WIDTH, HEIGHT, k, BITDEPTH, frame_noise=4.0)]
        return workload.add_objects(f, WIDTH, HEIGHT, k, BITDEPTH)
    warm_pocs = (pics[0][0], pics[1][0])   # every rank warms up on the first anchor and the first B picture (local, nothing pushed)
    dev_orig = {poc: [torch.from_numpy(p.view(np.int16)).cuda() for p in make(poc)] for poc in sorted(mine | set(warm_pocs))}
    owners = {poc: j % world for wave in waves for j, poc in enumerate(wave)}
    # default: on the device where that was run on hardware in this round (N = 2), on the host (barrier per wave) at
    # N > 2 -- the device rendezvous at 4 / 8 GPUs is implemented but unverified
    device_rendezvous = peers is not None and os.environ.get("XVCB_GOP_RENDEZVOUS", "device" if world <= 2 else "host") == "device"
    eng = gop.GopEngine(ctx, peers, rank, pics, lambda poc: dev_orig[poc], QP, BITDEPTH,
                        time_events=lambda: (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)),
                        owners=owners if device_rendezvous else None, ring=ring)
    eng.load_done(0, make(0))          # the key picture is not coded here: its original stands in for its reconstruction
    # warm-up: the whole sequence once, untimed, pushes included -- the device buffers that grow with the CU count of the
    # deeper temporal layers and the first transfer into every peer slot are one-time costs (measured: 35 of the 43 ms
    # of the first 15-picture wave at N = 4) -- then the timed run from the key picture again (same pictures, same
    # slots, same results).
    def run_waves(times, host_fence):
        for wave in waves:
            t0 = time.perf_counter()
            eng.encode_many([poc for j, poc in enumerate(wave) if j % world == rank])
            t1 = time.perf_counter()
            for j, poc in enumerate(wave):
                eng.share(poc, j % world)
            t2 = time.perf_counter()
            if host_fence:
                eng.fence()
            times.append((time.perf_counter() - t0) * 1e3)
            if os.environ.get("XVCB_GOP_DEBUG"):
                print("[gop] rank %d wave of %d: enqueue %.2f ms, push calls %.2f ms, fence %.2f ms" %
                      (rank, len(wave), (t1 - t0) * 1e3, (t2 - t1) * 1e3, (time.perf_counter() - t2) * 1e3), file=sys.stderr)

    saved_owners, eng.owners = eng.owners, None      # the warm-up is local: its references are this rank's own encodes
    for p in warm_pocs:
        eng.encode(p)
    ctx.sync()
    eng.owners = saved_owners
    # (the untimed pass always meets on the host after every wave: it is the pass in which device buffers grow, and a
    # cudaFree / cudaMalloc waits for every stream of the device -- with a stream parked in a device-side wait for another
    # rank's tag two ranks can wait for each other; the timed pass repeats the same pictures: nothing grows)
    # Two untimed passes: the library double-buffers the CU blob and the predictor staging, alternating per picture, so
    # the timed (third) pass meets every buffer with the pictures it held in the first pass -- no buffer grows in it.
    for warm_pass in range(2):
        eng.pass_index = warm_pass
        run_waves([], True)
        eng.fence()
    eng.pass_index = 2
    eng.events.clear()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    wave_ms = []
    t_all = time.perf_counter()
    run_waves(wave_ms, not device_rendezvous)
    eng.fence()
    torch.cuda.synchronize()
    total_s = time.perf_counter() - t_all
    busy_ms = sum(a.elapsed_time(b) for _, a, b in eng.events)
    # every rank holds every reconstruction: digest of the last pictures' slots
    digest = recon_digest([ctx.download_padded(eng.slot_of(poc), 0) for poc in (16 * n_sub_gops, 16 * n_sub_gops - 1)])
    if dist is not None:
        t = torch.tensor([total_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_s = float(t.item())
        b = torch.tensor([busy_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(b, op=dist.ReduceOp.SUM)
        busy_ms = float(b.item())
        w = torch.tensor(wave_ms, dtype=torch.float64, device="cuda")
        dist.all_reduce(w, op=dist.ReduceOp.MAX)
        wave_ms = [float(v) for v in w.tolist()]
        digests = [None] * world
        dist.all_gather_object(digests, digest)
    else:
        digests = [digest]
    ctx.close()
    coded = 16 * n_sub_gops
    return {"frames": n_frames, "pictures_coded": coded, "sub_gops": n_sub_gops,
            "value": coded * WIDTH * HEIGHT / total_s / 1e6, "unit": "Mpixels/s", "ms_total": total_s * 1e3,
            "waves": [{"pictures": len(wv), "host_ms" if device_rendezvous else "ms": ms} for wv, ms in zip(waves, wave_ms)],
            "gpu_busy_frac": busy_ms / (world * total_s * 1e3), "gpu_idle_frac": 1.0 - busy_ms / (world * total_s * 1e3),
            "reconstructions_identical_on_all_ranks": len(set(digests)) == 1,
            "how": "ThreadEncoder's readiness rule as waves across sub-GOPs; picture j of a wave on GPU j mod N; per picture: device copy of the "
                   "original, GPU partition pre-analysis (a rank's next picture ahead of the kernels of its current one), set_cus, the whole step; finished reconstructions pushed to every GPU "
                   "(copy engines over NVLink); rendezvous: " + ("on the device -- an arrival tag written behind every pushed slot on the same copy stream, "
                   "the consumer's stream waits for the tags of its reference pictures (cuStreamWaitValue32), no host barrier between waves" if device_rendezvous else
                   "per wave on the host (own pushes done, stream idle, barrier) before they are referenced") + "; wall "
                   "clock between barriers, max over ranks; busy = CUDA-event time of the pictures' work summed over GPUs; the whole sequence runs "
                   "twice untimed (host rendezvous) before the timed pass (buffer growth, first transfer into every peer slot)",
            "rendezvous": "device" if device_rendezvous else "host",
            "search_range": "InterSearch::GetSearchRangeUniPred capped at 128 (it yields 256 for the anchor pictures; the search kernel stages +-128 windows)",
            "key_picture": "not coded: its original is uploaded as its reconstruction"}


def banded_measurement(torch, dist, lib, sharding, rank, world, local_rank, steps):
    """BASELINE config 4's mechanism: ONE picture split into CTU-row bands across the GPUs, NCCL exchange of the CU
    decisions and of the deblocking halo rows.  At N = 4 the picture is 3840x2160 (config 4 itself), else this run's
    size.  Returns the dict for the bench line (rank 0)."""
    import common
    W, H = (3840, 2160) if world == 4 else (WIDTH, HEIGHT)
    canvas = workload.synth_canvas(W, H, 1234)
    frames = [workload.synth_frame(canvas, W, H, i, BITDEPTH, frame_noise=4.0) for i in (POC, REF_POCS[0], REF_POCS[1])]
    cus = workload.make_partition(W, H, seed=7, min_size=8, qp=QP)
    workload.set_predictors(cus, POC, REF_POCS, seed=8)
    lam = workload.lambda_for_qp(QP)
    prm = common.picture_params(0, lam, ranges=(128, 128), pocs=REF_POCS, slots=dict(orig=0, ref0=1, ref1=2, pred=3, rec=4, coeff=5), pad=0)
    prm["bits_mode"], prm["bi_iterations"] = 1, 1
    ctx = lib.Context(W, H, BITDEPTH, 6, device=local_rank)
    ctx.upload(0, frames[0])
    for s_, f in ((1, frames[1]), (2, frames[2])):
        ctx.upload(s_, f)
        ctx.pad_border(s_)
    eng = sharding.GpuEngine(ctx, dict(rec=4), BITDEPTH, {(0, 0): REF_POCS[0], (1, 0): REF_POCS[1]})
    enc = sharding.BandedPictureEncoder(eng, dist, rank, world, H)
    for _ in range(2):
        enc.encode(cus, prm)
    dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        enc.encode(cus, prm)
    torch.cuda.synchronize()
    sec = time.perf_counter() - t0
    t = torch.tensor([sec], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    sec = float(t.item())
    # one GPU, the same picture, the same calls without the exchange (rank 0's device)
    single_ms = None
    if rank == 0:
        ctx.set_cus(cus)
        ctx.encode_picture(prm, want_results=False)
        ctx.sync()
        t1 = time.perf_counter()
        for _ in range(steps):
            ctx.set_cus(cus)
            ctx.encode_picture(prm, want_results=False)
        ctx.sync()
        single_ms = (time.perf_counter() - t1) / steps * 1e3
    ctx.close()
    return {"width": W, "height": H, "bands": world, "ms_per_picture": sec / steps * 1e3, "value": W * H * steps / sec / 1e6, "unit": "Mpixels/s",
            "single_gpu_ms_per_picture": single_ms, "steps": steps,
            "how": "CTU-row bands; per picture: band search / MC / T-Q on every GPU, NCCL all-gather of the CU decisions, vertical edges, halo rows "
                   "down (4 luma + 2x2 chroma), horizontal edges, modified rows back (3 + 2x1); host-synchronous protocol, wall clock, max over ranks"}


def run_extras_guarded(line, rank, limit, work):
    """work() -> dict merged into `line`.  The line is complete without it: work runs under a watchdog -- if it does not
    return within `limit` seconds (a rank stuck in a collective or in a device-side wait), rank 0 prints the line with
    the reason and the process leaves with exit code 0 (every rank runs its own watchdog)."""
    finished = threading.Event()

    def give_up():
        if finished.wait(limit):
            return
        if rank == 0:
            line["gop"] = {"unavailable": "the GOP / banded measurements did not finish within %.0f s and were abandoned" % limit}
            print(json.dumps(line))
            sys.stdout.flush()
        os._exit(0)

    threading.Thread(target=give_up, daemon=True).start()
    extra = work()
    finished.set()
    for k, v in (extra or {}).items():
        if v is not None:
            line[k] = v


def run_ours(args):
    import torch
    from xvc_b200 import lib, sharding

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- xvc_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        import datetime
        # a rank that fails must not leave the others waiting for ten minutes
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank), timeout=datetime.timedelta(seconds=180))

    # frame-parallel sharding (weak scaling): one picture per rank and step.  Per-GPU work is FIXED as N
    # grows: every rank encodes a picture with the content of the N = 1 run (the search is data
    # dependent -- pictures rank..rank+N-1 of the synthetic sequence differ by up to 18 % in search time,
    # which a max over ranks would report as a scaling loss); --distinct-pictures gives rank r picture r
    index_offset = rank if args.distinct_pictures else 0
    frames, cus, prm, lam = picture_inputs(index_offset=index_offset)
    enc = WORKLOAD == "encode"
    gm = global_motion(index_offset)
    lam_sqrt = float(prm["lambda_sqrt"][0])
    cur = {"cus": cus, "mvp": None}      # CU array + predictors of the picture coded next

    def decide_begin(orig_slot):
        """Enqueues the GPU pre-analysis of the picture in orig_slot (kernel + copy of its result)."""
        if enc:
            ctx.decide_partition_begin(orig_slot, SL["ref0"], lam_sqrt, QP, center=gm)

    def decide_end():
        """Collects it -> CU array + predictors (host work: runs while the kernels enqueued after decide_begin execute)."""
        if not enc:
            return None
        c, _ = ctx.decide_partition_end()
        return {"cus": c, "mvp": predictors_from_partition(c, index_offset)}

    def decide(orig_slot):
        decide_begin(orig_slot)
        return decide_end()

    def set_cus():
        ctx.set_cus(cur["cus"])          # also restores the flags / vectors the previous step overwrote (device copy)
        ctx.set_mv_predictors(cur["mvp"])

    # slots: 0 orig, 1/2 references, 3 prediction, 4 levels, 5.. one reconstruction slot per rank
    # (contiguous: the frame-parallel all-gather lands every rank's reconstruction in place)
    # e2e pipeline: a second set (orig, levels, reconstructions) after the first one
    ctx = lib.Context(WIDTH, HEIGHT, BITDEPTH, num_slots=5 + world + 2 + world, device=local_rank)
    # a dedicated (non-default) torch stream: the library enqueues on it, torch events time it
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    SL = dict(orig=0, ref0=1, ref1=2, pred=3, coeff=4, rec=5 + rank)
    prm["orig_slot"], prm["pred_slot"], prm["rec_slot"], prm["coeff_slot"] = SL["orig"], SL["pred"], SL["rec"], SL["coeff"]
    set_ref_slots(prm, {REF_POCS[0]: SL["ref0"], REF_POCS[1]: SL["ref1"]})
    ctx.upload(SL["orig"], frames[0])
    for slot, f in ((SL["ref0"], frames[1]), (SL["ref1"], frames[2])):
        ctx.upload(slot, f)
        ctx.pad_border(slot)
    if enc:
        cur.update(decide(SL["orig"]))
    n = len(cur["cus"])
    cus = cur["cus"]
    mvp = cur["mvp"]
    set_cus()
    ctx.sync()
    ctx.set_profiling(True)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # multi-GPU: reconstruction slots alternate between two sets, so the all-gather of picture i
    # (NCCL stream) overlaps the kernels of picture i+1; a set is reused only after its exchange
    base2 = 5 + world
    dev_prms = [prm.copy(), prm.copy()]
    dev_prms[1]["rec_slot"] = base2 + 2 + rank
    dev_first_rec = [5, base2 + 2]
    works = [None, None]
    # --exchange push (default): copy-engine pushes into the peers' arenas (CUDA IPC over NVLink), no SM
    # involved; --exchange nccl: in-place all-gather on the NCCL stream (its copy kernels compete
    # with the persistent search kernel for SMs)
    peers = None
    if dist is not None and args.exchange == "push":
        try:
            peers = sharding.PeerExchange(ctx, dist, rank, world)
        except sharding.PeerExchangeUnavailable as e:      # raised on all ranks together; stated in config.parallelism
            if rank == 0:
                print("bench.py: peer pushes unavailable (%s), exchanging over NCCL" % e, file=sys.stderr)

    def device_step(i=0, flush_l2=False, ev=None):
        """ev: list collecting (before wait, after wait, before kernels, after kernels) events."""
        s = i & 1 if dist is not None else 0
        mark = (lambda: None) if ev is None else (lambda: (ev.append(torch.cuda.Event(enable_timing=True)), ev[-1].record(stream)))
        mark()
        if works[s] is not None:
            works[s].wait()
            works[s] = None
        if peers is not None:     # the slot about to be rewritten was pushed two steps ago
            peers.wait_own(dev_first_rec[s] + rank)
        mark()
        if flush_l2:
            flush.zero_()
        mark()
        decide_begin(SL["orig"])                       # pre-analysis of the NEXT picture, ahead of this picture's kernels
        ctx.encode_picture(dev_prms[s], want_results=False)
        mark()
        if peers is not None:     # finished, padded reconstructions to every GPU that will reference them
            peers.push(dev_first_rec[s] + rank)
        elif dist is not None:
            works[s] = sharding.frame_parallel_exchange(ctx, dist, rank, world, dev_first_rec[s], async_op=True)
        nxt = decide_end()        # host work of the next picture (partition, predictors, CU classes, upload) overlaps this picture's kernels
        if nxt is not None:
            cur.update(nxt)
        set_cus()

    def drain():
        if peers is not None:
            peers.wait_own()
        for s in range(2):
            if works[s] is not None:
                works[s].wait()
                works[s] = None

    for i in range(args.warmup):
        device_step(i)
    drain()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = lib.launch_count()
    step_ms, stage_ms = [], {k: [] for k in lib.Context.STAGES}
    if dist is None:
        part_ms = []
        for _ in range(args.steps):
            if not os.environ.get("XVCB_BENCH_NOFLUSH"):      # experiments only; the default run always flushes
                flush.zero_()
            e0, ea, e1 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            e0.record(stream)
            decide_begin(SL["orig"])                           # pre-analysis of the NEXT picture, ahead of this picture's kernels
            ea.record(stream)
            ctx.encode_picture(prm, want_results=False)
            e1.record(stream)
            nxt = decide_end()                                 # host work of the next picture (partition, predictors, CU classes,
            if nxt is not None:                                # upload) overlaps this picture's kernels
                cur.update(nxt)
            set_cus()
            e1.synchronize()
            step_ms.append(e0.elapsed_time(e1))
            part_ms.append(e0.elapsed_time(ea))
            for k, v in ctx.stage_times_ms().items():
                stage_ms[k].append(v)
        stage_ms["partition"] = part_ms
        total_ms = float(np.sum(step_ms))
    else:
        # Same timed region as N = 1 (the kernels of encode_picture; the untimed set_cus + L2 flush
        # sit between steps) PLUS everything the exchange costs: it overlaps the next step's
        # kernels (any slowdown shows in that step's time), the stream stall waiting for an
        # unfinished exchange before its slot is rewritten, and the tail of the last exchanges.
        ev = []
        for i in range(args.steps):
            device_step(i, flush_l2=True, ev=ev)
        t0 = torch.cuda.Event(enable_timing=True)
        t0.record(stream)
        drain()
        t1 = torch.cuda.Event(enable_timing=True)
        t1.record(stream)
        t1.synchronize()
        for k, v in ctx.stage_times_ms().items():
            stage_ms[k].append(v)
        stall_ms = sum(ev[4 * i].elapsed_time(ev[4 * i + 1]) for i in range(args.steps))
        step_ms = [ev[4 * i + 2].elapsed_time(ev[4 * i + 3]) for i in range(args.steps)]
        tail_ms = t0.elapsed_time(t1)
        total_ms = float(np.sum(step_ms)) + stall_ms + tail_ms
        exchange_ms = {"stall_ms_per_step": stall_ms / args.steps, "tail_ms": tail_ms, "kernels_ms_per_step": float(np.mean(step_ms))}
    launches = lib.launch_count() - launches0
    barrier()
    sampler.stop_flag = True
    sampler.join(timeout=2)
    if dist is not None:
        t = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    value = world * WIDTH * HEIGHT / (ms_per_step * 1e-3) / 1e6

    # ---- end to end through the C ABI with host buffers (pinned), H2D + D2H inside the timed region.
    # Pictures alternate between two slot sets so that the transfers of one picture (the library's
    # copy stream) overlap the kernels of the next: the upload of picture i+1 is enqueued before the
    # kernels of picture i, the host consumes the results of picture i-1 while picture i runs.
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()   # noqa: E731
    h_orig = [pin(p) for p in frames[0]]
    h_rec = [[pin(np.zeros_like(p)) for p in frames[0]] for _ in range(2)]
    h_lev = [[pin(np.zeros(p.shape, dtype=np.int16)) for p in frames[0]] for _ in range(2)]
    n_cu_max = 64 * ((WIDTH + 63) // 64) * ((HEIGHT + 63) // 64)
    h_cus = [pin(np.zeros(n_cu_max, dtype=abi.cu_dtype).view(np.uint8)).view(abi.cu_dtype) for _ in range(2)]
    part_d2h = (n_cu_max * abi.cu_dtype.itemsize + 136 * (n_cu_max // 64)) if enc else 0       # what xvcb200_decide_partition brings back
    h2d = sum(p.nbytes for p in h_orig) + cus.nbytes + (mvp.nbytes if mvp is not None else 0)
    d2h = sum(p.nbytes for p in h_rec[0]) + sum(p.nbytes for p in h_lev[0]) + cus.nbytes + part_d2h
    sets = [dict(orig=SL["orig"], coeff=SL["coeff"], first_rec=5),
            dict(orig=base2, coeff=base2 + 1, first_rec=base2 + 2)]
    prms = []
    for st in sets:
        q = prm.copy()
        q["orig_slot"], q["coeff_slot"], q["rec_slot"] = st["orig"], st["coeff"], st["first_rec"] + rank
        prms.append(q)

    def e2e_run(count, flush_l2=True):
        """count pictures through the pipeline; returns the last picture's outputs (host arrays)."""
        ctx.upload_async(sets[0]["orig"], h_orig)
        if enc:
            cur.update(decide(sets[0]["orig"]))
        set_cus()
        n_of = [0, 0]
        for i in range(count):
            s = i & 1
            if i + 1 < count:
                ctx.upload_async(sets[1 - s]["orig"], h_orig)
            if peers is not None:
                peers.wait_own(sets[s]["first_rec"] + rank)
            if flush_l2:
                flush.zero_()
            # pre-analysis of picture i+1 (its upload is joined inside) ahead of the kernels of picture i
            if i + 1 < count:
                decide_begin(sets[1 - s]["orig"])
            ctx.encode_picture(prms[s], want_results=False)
            if peers is not None:
                peers.push(sets[s]["first_rec"] + rank)
            elif dist is not None:
                sharding.frame_parallel_exchange(ctx, dist, rank, world, sets[s]["first_rec"])
            n_of[s] = len(cur["cus"])
            ctx.get_cus_async(h_cus[s][:n_of[s]])
            ctx.download_coeff_async(sets[s]["coeff"], h_lev[s])
            ctx.download_async(sets[s]["first_rec"] + rank, h_rec[s])
            if i >= 1:                       # picture i-1 is complete on the host from here on
                ctx.wait_download(sets[1 - s]["first_rec"] + rank)
            if i + 1 < count:
                nxt = decide_end()
                if nxt is not None:
                    cur.update(nxt)
                set_cus()                    # host work of picture i+1 overlaps the kernels of picture i
        last = (count - 1) & 1
        ctx.wait_download(sets[last]["first_rec"] + rank)
        return h_cus[last][:n_of[last]], h_rec[last], h_lev[last]

    ctx.set_profiling(False)
    e2e_run(max(2, args.warmup))
    barrier()
    t0 = time.perf_counter()
    cus_out, rec_out, lev_out = e2e_run(args.steps)
    torch.cuda.synchronize()
    e2e_sec = (time.perf_counter() - t0)
    # the flush sits on the compute stream (the critical path) but is not part of the step:
    # time it alone and subtract
    t1 = time.perf_counter()
    for _ in range(args.steps):
        flush.zero_()
    torch.cuda.synchronize()
    e2e_sec -= (time.perf_counter() - t1)
    if dist is not None:
        t = torch.tensor([e2e_sec], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_sec = float(t.item())
    e2e_value = world * WIDTH * HEIGHT * args.steps / e2e_sec / 1e6

    # ---- candidates per step, by the reference's own count (TzSearch evaluates `num_sad` positions per search; FullSearch
    # 81 per picture of the searched list; SubpelSearch 17 per search) -- one more picture with the results brought back
    sad_candidates = None
    try:
        set_cus()           # (the CU array on the device carries the previous picture's chosen vectors)
        me, _ = ctx.encode_picture(prm, want_results=True)
        ctx.sync()
        l0, l1 = LISTS[WORKLOAD]
        J = len(l0) + len(l1)
        cols = list(range(len(l0))) + [len(l0) + r for r, p in enumerate(l1) if p not in l0 or WORKLOAD == "raster"]
        per = me["num_sad"].reshape(-1, J)[:, cols].astype(np.int64)
        n_bi = int(prm["bi_iterations"][0]) * len(per) * max(len(l0), len(l1))
        sad_candidates = {"tz_search_sad": int(per.sum()), "tz_searches": int(per.size),
                          "bi_full_search_sad": 81 * n_bi, "subpel_satd": 17 * (int(per.size) + n_bi)}
    except Exception as e:  # noqa: BLE001
        sad_candidates = {"unavailable": repr(e)}

    # ---- round 1's step beside it (context: the raster-dominated workload the round-1 numbers were quoted on)
    other = None
    if world == 1 and WORKLOAD == "encode" and not args.size and args.gop != "off":
        try:
            other = raster_measurement(torch, ctx, stream, flush, SL, args)
        except Exception as e:  # noqa: BLE001
            other = {"unavailable": repr(e)}

    slot_mb = ctx.slot_region(0)[1] / 1e6

    # ---- roofline of the dominant kernel, from the live per-stage CUDA events
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    P = WIDTH * HEIGHT
    alg_bytes = {   # SURVEY.md section 8(d), S = 2 bytes/sample, 2 reference pictures
        "tz_search": P * 2 * (1 + 2) + 16 * n * 2,
        "subpel_search": P * 2 * (1 + 2) + 32 * n * 2,
        "motion_compensate": int(1.5 * P * 2 * 2),
        "tq_reconstruct": int(4 * 1.5 * P * 2),
        "deblock": int(2 * 1.5 * P * 2) + 2 * P,
        "pad_border": int(0.2 * 1.5 * P * 2),
        "me_jobs": 60 * n,
        "partition": P * 2 * 2 + 28 * n,      # original + one reference picture once, CU array out
    }
    stage_avg = {k: float(np.mean(v)) for k, v in stage_ms.items()}
    # the dominant KERNEL: the stages me_jobs / tz_search / motion_compensate / pad_border / partition are one kernel each,
    # subpel_search (classify + 4 size classes, twice with a bi-prediction pass, + the bi search) and tq_reconstruct (one
    # launch per block shape) are several
    single = {k: stage_avg[k] for k in ("me_jobs", "tz_search", "motion_compensate", "pad_border", "partition") if k in stage_avg}
    dom = max(single, key=single.get)
    achieved = alg_bytes[dom] / (stage_avg[dom] * 1e-3) / 1e9
    traffic, traffic_src, issue = None, None, None
    import glob
    tpaths = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_traffic.json")), key=os.path.getmtime)
    tpath = tpaths[-1] if tpaths else ""      # newest ncu --set full capture of the dominant kernel
    if tpath and (WIDTH, HEIGHT) == (1920, 1080):      # the capture is of the 1080p step
        tj = json.load(open(tpath))
        if tj.get("kernel", "").startswith(dom) and tj.get("workload", "raster") == WORKLOAD:
            traffic, traffic_src = tj["dram_bytes_per_launch"], tj["source"]
            if "warp_instructions_per_launch" in tj:
                # the ceilings that actually bound the integer search: warp-instruction issue (SMs x 4 schedulers x
                # clock) and the shared-memory pipe (one wavefront per SM per clock); counts from the same ncu capture
                sm_clock = 148 * 1.965e9
                issue = {"warp_instructions_per_launch": tj["warp_instructions_per_launch"],
                         "issue_frac_of_peak": tj["warp_instructions_per_launch"] / (stage_avg[dom] * 1e-3) / (4 * sm_clock),
                         "shared_wavefronts_per_launch": tj.get("shared_wavefronts_per_launch"),
                         "shared_pipe_frac_of_peak": (tj["shared_wavefronts_per_launch"] / (stage_avg[dom] * 1e-3) / sm_clock) if tj.get("shared_wavefronts_per_launch") else None,
                         "peak": "148 SMs x 4 schedulers (issue) / x 1 wavefront (shared memory) x 1.965 GHz"}
    roofline = {"kernel": dom, "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src, "algorithmic_bytes": alg_bytes[dom], "avg_ms": stage_avg[dom],
                "note": "integer SAD search: ALU/L1-bound, not HBM-bound; see DESIGN.md for op counts",
                "compute_ceilings": issue,
                "stages_ms": stage_avg,
                "stages_frac_of_hbm": {k: alg_bytes[k] / (v * 1e-3) / 1e9 / peak for k, v in stage_avg.items() if v > 0}}

    # ---- CPU baseline beside it (bounded sample) + bit-exact reconstruction check (rank 0)
    cpu = None
    bitexact = None
    try:
        if rank != 0:
            raise RuntimeError("rank 0 only")
        arm = CpuArm()
        frames0, cus0, prm0, lam0 = picture_inputs(index_offset=0)
        if cus0 is None:      # the partition the GPU decided for this content (equal to the numpy rule: tests/test_gpu_partition.py)
            cus0 = cus if index_offset == 0 else cpu_partition(frames0, lam0)
        mvp0 = predictors_from_partition(cus0)
        t_first, rec_cpu = arm.run(frames0, cus0, prm0, lam0, mvp0)
        reps = int(min(20, max(1, 8.0 / max(t_first, 1e-3)))) if arm.kind == "reference" and world == 1 else 1
        times = [t_first] + [arm.run(frames0, cus0, prm0, lam0, mvp0)[0] for _ in range(reps - 1)]
        sec = float(np.mean(times[1:])) if len(times) > 1 else t_first
        cpu = {"value": WIDTH * HEIGHT / sec / 1e6, "unit": "Mpixels/s", "cores": arm.cores, "kind": arm.kind,
               "sample": "%d full 1920x1080 pictures of the same step (first one untimed warm-up)" % len(times)}
        bitexact = recon_digest(rec_cpu) == recon_digest(rec_out)
        if world == 1:
            cpu["xvcenc_cli"] = xvcenc_cli_sample()      # whole-encoder context, see its docstring
        else:       # the CPU baseline is an N = 1 number; at N > 1 the one CPU run above only checks bit-exactness
            cpu = None
    except Exception as e:  # noqa: BLE001
        cpu = {"value": None, "unit": "Mpixels/s", "cores": 0, "kind": "port", "sample": "unavailable: %r" % (e,)}

    line = {
        "metric": METRIC, "value": value, "unit": "Mpixels/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u16 samples / int32 arithmetic", "data": "synthetic",
        "config": config_dict(n, {"parallelism": (("frame-parallel: one picture per GPU (%s); per step every GPU's padded reconstruction (%.1f MB) goes to the %d others -- %s -- overlapped with the next picture (two reconstruction slot sets); timed: the kernels of every step (as at N=1) + stream stalls waiting for an exchange + the tail of the last exchanges" % ("rank r encodes picture r of the sequence" if args.distinct_pictures else "the same picture content on every GPU: fixed per-GPU work", slot_mb, world - 1, "copy-engine pushes into the peers' slot arenas (CUDA IPC over NVLink, no SM)" if peers is not None else "in-place NCCL all-gather on the NCCL stream"))) if world > 1 else "single GPU"}),
        "frames_per_s": value * 1e6 / (WIDTH * HEIGHT),
        "e2e": {"value": e2e_value, "unit": "Mpixels/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "pipeline": "2 pictures in flight: pinned-host H2D / D2H of neighbouring pictures on the copy stream overlap the kernels"},
        "gpu_launches": int(launches),
        "clocks": sampler.summary(),
        "roofline": roofline,
        "cpu_baseline": cpu,
        "recon_bitexact_vs_cpu_baseline": bitexact,
        "me_sad_candidates_per_step": sad_candidates,
    }
    if other is not None:
        line["raster_workload"] = other
    if dist is not None:
        line["exchange"] = exchange_ms       # rank 0's own split of its timed total

    # ---- multi-GPU mechanisms beside the weak-scaling number: the frame-parallel GOP (config 5) and the banded picture
    # (config 4).  The line above is complete without them: they run under a watchdog -- if they do not finish (a rank
    # stuck in a collective or in a device-side wait) rank 0 prints the line with the reason and every rank leaves.
    if args.gop != "off":
        def extras():
            out = {}
            try:
                ctx.close()
                out["gop"] = gop_measurement(torch, dist, lib, sharding, rank, world, local_rank,
                                             max(2, world) if args.gop_sub_gops <= 0 else args.gop_sub_gops)
            except Exception as e:  # noqa: BLE001
                out["gop"] = {"unavailable": repr(e)}
            if dist is not None:
                try:
                    out["banded"] = banded_measurement(torch, dist, lib, sharding, rank, world, local_rank, 10)
                except Exception as e:  # noqa: BLE001
                    out["banded"] = {"unavailable": repr(e)}
            return out

        run_extras_guarded(line, rank, float(os.environ.get("XVCB_BENCH_EXTRAS_TIMEOUT", "180")), extras)
    if rank == 0:
        print(json.dumps(line))
        sys.stdout.flush()
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--distinct-pictures", action="store_true", help="N>1: rank r encodes picture r of the sequence instead of picture 0")
    ap.add_argument("--exchange", default="push", choices=["push", "nccl"],
                    help="N>1: how finished reconstructions reach the other GPUs (copy-engine pushes over CUDA IPC, or NCCL all-gather)")
    ap.add_argument("--gop", default="on", choices=["on", "off"],
                    help="the frame-parallel hierarchical-B GOP measurement (and, at N > 1, the banded picture) beside the step")
    ap.add_argument("--gop-sub-gops", type=int, default=0, help="sub-GOPs of 16 pictures in the GOP measurement (default: max(2, N))")
    ap.add_argument("--workload", default="encode", choices=["encode", "raster"], help="see the module docstring")
    ap.add_argument("--size", default=None, metavar="WxH[@QP]",
                    help="context runs only (DESIGN.md table): another picture size / qp, e.g. 3840x2160@27, 7680x4320; the "
                         "default (and the only bench line the metric is quoted on) is 1920x1080@32")
    args = ap.parse_args()
    global WORKLOAD
    WORKLOAD = args.workload
    if args.size:
        global WIDTH, HEIGHT, QP
        wh, _, q = args.size.partition("@")
        WIDTH, HEIGHT = (int(v) for v in wh.lower().split("x"))
        QP = int(q) if q else QP
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
