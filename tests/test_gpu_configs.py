"""BASELINE.json configurations at FULL size on the GPU (1080p qp32, 2160p 10-bit qp27, 4320p qp32).

The oracle cannot run whole pictures of these sizes in test time, so parity is pinned through
(a) the oracle on seeded SUBSETS of the picture's CUs (search jobs incl. the picture borders,
    residual coding of individual CUs), compared with what the whole-picture pipeline produced;
(b) size-independent properties the reference's own tests use: the decoder path run on the
    encoder's output (levels + CU decisions) rebuilds the encoder's reconstruction bit for bit
    (EncodeDecodeTest / ChecksumEncDecTest, encode_decode_test.cc:118-150,
    checksum_enc_dec_test.cc:144-190), per-TU distortions equal a direct SSD of original vs
    reconstruction (transform_encoder.cc:268-283), and a second run is identical (no order
    dependence in the atomics of the parallel search).
"""
import numpy as np
import pytest

import common
from oracle.bindings import Picture
from xvc_b200 import abi, lib, workload

pytestmark = pytest.mark.gpu

CONFIGS = [
    pytest.param(1920, 1080, 10, 32, id="1080p-qp32"),
    pytest.param(3840, 2160, 10, 27, id="2160p-10bit-qp27"),
    pytest.param(7680, 4320, 10, 32, id="4320p-qp32"),
]


def _encode(ctx, cus, prm):
    ctx.set_cus(cus)
    me, tu = ctx.encode_picture(prm)
    ctx.sync()
    return me, tu


@pytest.mark.parametrize("width,height,bd,qp", CONFIGS)
def test_full_size_picture(oracle, width, height, bd, qp):
    rng = np.random.default_rng(width + qp)
    canvas = workload.synth_canvas(width, height, 1234)
    cur, r0, r1 = [workload.synth_frame(canvas, width, height, i, bd) for i in (8, 0, 16)]
    lam = workload.lambda_for_qp(qp)
    cus = workload.make_partition(width, height, seed=7, min_size=8, qp=qp)
    n = len(cus)
    ranges = tuple(workload.search_range_uni(8, p) for p in (0, 16))
    # slots: 0 orig, 1/2 references, 3 prediction, 4 reconstruction, 5 levels, 6/7 decoder side
    ctx = lib.Context(width, height, bd, num_slots=8)
    ctx.upload(0, cur)
    for s, f in ((1, r0), (2, r1)):
        ctx.upload(s, f)
        ctx.pad_border(s)
    slots = dict(orig=0, ref0=1, ref1=2, pred=3, rec=4, coeff=5)

    # ---- run A: no in-loop filter -> the reconstruction the distortions refer to
    prm_a = common.picture_params(0, lam, ranges=ranges, slots=slots, deblock=0, pad=0)
    me_a, tu_a = _encode(ctx, cus, prm_a)
    rec_a = ctx.download(4)
    pred_a = ctx.download(3)
    lev_a = ctx.download_coeff(5)
    cus_a = ctx.get_cus()
    assert np.any(lev_a[0]) and np.any(cus_a["mv"])

    # per-TU distortion = SSD(original, reconstruction) >> 2(bd-8) (sample of TUs, all components)
    for i in rng.choice(n, size=min(n, 1500), replace=False):
        c = cus[i]
        for comp in range(3):
            sh = 1 if comp else 0
            x, y, w, h = c["x"] >> sh, c["y"] >> sh, c["w"] >> sh, c["h"] >> sh
            d = cur[comp][y:y + h, x:x + w].astype(np.int64) - rec_a[comp][y:y + h, x:x + w].astype(np.int64)
            assert int(tu_a[3 * i + comp]["ssd"]) == int((d * d).sum()) >> (2 * (bd - 8)), (i, comp)
            nz = int(np.count_nonzero(lev_a[comp][y:y + h, x:x + w]))
            assert int(tu_a[3 * i + comp]["num_non_zero"]) == nz, (i, comp)
            bit = (abi.CU_CBF_Y, abi.CU_CBF_U, abi.CU_CBF_V)[comp]
            assert bool(cus_a[i]["flags"] & bit) == (nz != 0), (i, comp)

    # ---- oracle on a subset of the search jobs: picture corners / borders and random interior CUs
    order = np.lexsort((cus["x"], cus["y"]))
    border = [int(order[0]), int(order[-1])] + [int(i) for i in np.flatnonzero((cus["x"] + cus["w"] == width) | (cus["y"] + cus["h"] == height))[:6]]
    pick = sorted(set(border) | set(int(i) for i in rng.choice(n, size=10, replace=False)))
    sub = cus[pick].copy()
    jobs = np.zeros(2 * len(pick), dtype=abi.me_job_dtype)
    for k in range(len(pick)):
        for l in range(2):
            j = jobs[2 * k + l]
            j["cu"], j["list"], j["ref_slot"], j["search_range"] = k, l, 0, ranges[l]
    refs = common.oracle_refs(oracle, width, height, r0, r1)
    me_o = oracle.me_search(Picture(width, height, 0, cur), refs, bd, sub, jobs, float(np.sqrt(lam)))
    for k, i in enumerate(pick):
        for l in range(2):
            for f in ("mv_fullpel", "mv", "cost_fullpel", "dist", "cost", "num_sad"):
                assert np.array_equal(me_a[2 * i + l][f], me_o[2 * k + l][f]), (i, l, f)

    # ---- oracle residual coding of a subset of CUs on the GPU's prediction
    pick_tq = sorted(set(int(i) for i in rng.choice(n, size=24, replace=False)))
    sub = cus_a[pick_tq].copy()
    rec_o = Picture(width, height, 80)
    levels_o, tu_o = oracle.tq_reconstruct(Picture(width, height, 0, cur), Picture(width, height, 0, pred_a), rec_o, bd, sub)[:2]
    for k, i in enumerate(pick_tq):
        c = cus[i]
        for comp in range(3):
            sh = 1 if comp else 0
            x, y, w, h = c["x"] >> sh, c["y"] >> sh, c["w"] >> sh, c["h"] >> sh
            assert np.array_equal(rec_o.plane(comp)[y:y + h, x:x + w], rec_a[comp][y:y + h, x:x + w]), (i, comp)
            assert np.array_equal(levels_o[comp][y:y + h, x:x + w], lev_a[comp][y:y + h, x:x + w]), (i, comp)
            assert tu_o[3 * k + comp] == tu_a[3 * i + comp], (i, comp)

    # ---- run B: the whole step incl. deblocking and padding; decisions identical to run A
    prm_b = common.picture_params(0, lam, ranges=ranges, slots=slots, deblock=1, pad=1)
    me_b, tu_b = _encode(ctx, cus, prm_b)
    assert np.array_equal(me_a, me_b) and np.array_equal(tu_a, tu_b)
    cus_b = ctx.get_cus()
    assert np.array_equal(cus_a, cus_b)
    rec_b = [ctx.download_padded(4, c) for c in range(3)]
    lev_b = ctx.download_coeff(5)
    for c in range(3):
        assert np.array_equal(lev_a[c], lev_b[c])
    # the filter changed something, and only near block edges of the unfiltered reconstruction
    assert any(not np.array_equal(ctx.download(4)[c], rec_a[c]) for c in range(3))

    # ---- decoder path on the encoder's output: prediction from the decided vectors, dequant +
    # inverse transform + reconstruction from the levels, deblocking, padding
    ctx.upload_coeff(7, lev_b)
    ctx.set_cus(cus_b)
    ctx.motion_compensate({(0, 0): 1, (1, 0): 2}, 3)
    ctx.dequant_reconstruct(3, 6, 7)
    ctx.deblock_picture(6, 0, {(0, 0): 0, (1, 0): 16})
    ctx.pad_border(6)
    ctx.sync()
    for c in range(3):
        assert np.array_equal(ctx.download_padded(6, c), rec_b[c]), c
    ctx.close()


def test_hierarchical_gop_chain(oracle):
    """Reconstructions stay on the GPU and become reference pictures of later pictures
    (picture_encoder.cc:149-151: deblock, PadBorder, then the picture is referenced): key picture
    POC 0 -> uni-predicted POC 16 -> bi-predicted POC 8 (refs 0 / 16) -> bi-predicted POC 4
    (refs 0 / 8), in the coding order of a hierarchical-B sub-GOP.  Every picture is compared with
    the oracle running the same chain on its own reconstructions."""
    width, height, bd, qp = 264, 136, 10, 30
    canvas = workload.synth_canvas(width, height, 77)
    frame = lambda poc: workload.synth_frame(canvas, width, height, poc, bd)   # noqa: E731
    lam = workload.lambda_for_qp(qp)
    # slots: 0 orig, 1 pred, 2 levels, 3.. reconstructions by coding order
    ctx = lib.Context(width, height, bd, num_slots=7)
    ctx.upload(3, frame(0))
    ctx.pad_border(3)
    rec_o = {0: Picture(width, height, 80, frame(0))}
    oracle.pad_border(rec_o[0])
    slot_of = {0: 3}
    chain = [(16, 1, (0,)), (8, 0, (0, 16)), (4, 0, (0, 8))]       # (poc, pic_type, reference POCs L0 / L1)
    for k, (poc, pic_type, ref_pocs) in enumerate(chain):
        cur = frame(poc)
        cus = workload.make_partition(width, height, seed=40 + k, min_size=8, qp=qp)
        ranges = tuple(workload.search_range_uni(poc, p) for p in ref_pocs) + ((96,) if len(ref_pocs) == 1 else ())
        rec_slot = 4 + k
        slots = dict(orig=0, ref0=slot_of[ref_pocs[0]], ref1=slot_of[ref_pocs[1]] if len(ref_pocs) > 1 else -1, pred=1, rec=rec_slot, coeff=2)
        prm = common.picture_params(pic_type, lam, ranges=ranges[:2], pocs=(ref_pocs + (0,))[:2], slots=slots)
        ctx.upload(0, cur)
        ctx.set_cus(cus)
        me_g, tu_g = ctx.encode_picture(prm)
        ctx.sync()
        refs = {(0, 0): rec_o[ref_pocs[0]]}
        if len(ref_pocs) > 1:
            refs[(1, 0)] = rec_o[ref_pocs[1]]
        pred, rec = Picture(width, height, 80), Picture(width, height, 80)
        cus_o = cus.copy()
        levels, me_o, tu_o = oracle.encode_picture(Picture(width, height, 0, cur), refs, pred, rec, bd, cus_o, prm)
        assert np.array_equal(me_g, me_o) and np.array_equal(tu_g, tu_o), poc
        lev_g = ctx.download_coeff(2)
        for c in range(3):
            assert np.array_equal(ctx.download_padded(rec_slot, c), rec.full[c]), (poc, c)
            assert np.array_equal(lev_g[c], levels[c]), (poc, c)
        rec_o[poc] = rec
        slot_of[poc] = rec_slot
    ctx.close()


def test_peer_push_two_processes_one_gpu():
    """Frame-parallel exchange (config 5): padded reconstructions pushed between the slot arenas
    of two PROCESSES (CUDA IPC + copy engines, xvcb200_push_slot) arrive bit for bit.  Both ranks
    share cuda:0 here; tests/run_peer_push.py without `same-gpu` is the NVLink form."""
    import os
    import subprocess
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29512", os.path.join(here, "run_peer_push.py"), "same-gpu"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "True" in out.stdout


def test_frame_parallel_gop_two_processes_one_gpu():
    """BASELINE config 5's mechanism end to end: a hierarchical-B sub-GOP encoded frame-parallel by two
    processes whose reconstructions reach each other through xvcb200_push_slot and are referenced by
    the next wave == the serial encode (tests/run_frame_parallel_gop.py; without `same-gpu` it runs one
    rank per GPU over NVLink)."""
    import os
    import subprocess
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29516", os.path.join(here, "run_frame_parallel_gop.py"), "416", "240", "same-gpu"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "serial encode: True" in out.stdout


def test_ipc_open_failure_is_not_sticky():
    """A peer arena that cannot be opened (here: a handle that names nothing) is reported to the caller
    without poisoning the context: bench.py / sharding.PeerExchange then fall back to the NCCL exchange."""
    width, height, bd = 64, 64, 10
    ctx = lib.Context(width, height, bd, num_slots=1)
    with pytest.raises(lib.XvcB200Error):
        ctx.ipc_open_peer(bytes(lib.IPC_HANDLE_BYTES))        # no magic: refused before CUDA sees it
    mine = bytearray(ctx.ipc_export())
    with pytest.raises(lib.XvcB200Error):                    # a peer arena with another slot count is refused
        other = lib.Context(width, height, bd, num_slots=2)
        try:
            ctx.ipc_open_peer(other.ipc_export())
        finally:
            other.close()
    mine[0] ^= 0xff                                          # right layout, handle that names nothing
    with pytest.raises(lib.XvcB200Error):
        ctx.ipc_open_peer(bytes(mine))
    cur = common.frames(width, height, bd, 3)[0]
    ctx.upload(0, cur)
    ctx.pad_border(0)
    ctx.push_slot(0)          # no peers: a no-op
    ctx.wait_pushes()
    ctx.sync()
    assert np.array_equal(ctx.download(0)[0], cur[0])
    ctx.close()
