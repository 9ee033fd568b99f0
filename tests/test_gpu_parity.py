"""GPU parity: every entry point of libxvc_b200.so (through the C ABI) against the C oracle on
the same seeded inputs.  Bit-exact for all of them -- this path is integer arithmetic, and the
interpolation filters turned out exact as well (north_star allows +-1 LSB there; the tests
assert 0).  Nothing here reads /root/reference."""
import itertools

import numpy as np
import pytest

from oracle.bindings import Picture
from xvc_b200 import abi, lib, workload

import common
from common import CSIZES, SIZES, rnd_resi, rnd_samples

pytestmark = pytest.mark.gpu


# ------------------------------------------------------------------ table-shaped entries
@pytest.mark.parametrize("bd", [8, 10, 12])
def test_sad_ssd_table(oracle, bd):
    rng = np.random.default_rng(101)
    for w, h in itertools.product([2] + SIZES, [2] + SIZES):
        a, b = rnd_samples(rng, h, w + 3, bd), rnd_samples(rng, h, w + 5, bd)
        r, r2 = rnd_resi(rng, h, w + 1, bd), rnd_resi(rng, h, w + 7, bd)
        assert lib.sad(a, b, w, h) == oracle.sad(0, a, b, w, h)
        assert lib.sad(r, b, w, h) == oracle.sad(1, r, b, w, h)
        assert lib.ssd(a, b, w, h) == oracle.ssd(0, a, b, w, h)
        assert lib.ssd(r, b, w, h) == oracle.ssd(1, r, b, w, h)
        assert lib.ssd(r, r2, w, h) == oracle.ssd(2, r, r2, w, h)


@pytest.mark.parametrize("bd", [8, 10, 12])
def test_compare_metrics(oracle, bd):
    rng = np.random.default_rng(102)
    for w, h in itertools.product([2] + SIZES, [2] + SIZES):
        a, b = rnd_samples(rng, h, w, bd), rnd_samples(rng, h, w, bd)
        r = rnd_resi(rng, h, w, bd)
        near = np.clip(a.astype(np.int32) + rng.integers(-3, 4, size=a.shape), 0, (1 << bd) - 1).astype(np.uint16)
        for metric in (abi.METRIC_SSD, abi.METRIC_SATD, abi.METRIC_SAD, abi.METRIC_SAD_FAST):
            if metric == abi.METRIC_SAD_FAST and h < 4:
                continue
            assert lib.compare(metric, bd, a, b, w, h) == oracle.compare(metric, bd, a, b, w, h), (metric, w, h)
            assert lib.compare(metric, bd, r, b, w, h) == oracle.compare(metric, bd, r, b, w, h), (metric, w, h)
            assert lib.compare(metric, bd, a, near, w, h) == oracle.compare(metric, bd, a, near, w, h), (metric, w, h)


@pytest.mark.parametrize("bd", [8, 10, 12])
def test_filters_table(oracle, bd):
    rng = np.random.default_rng(103)
    for chroma in (0, 1):
        sizes = CSIZES if chroma else SIZES
        for w, h in itertools.product(sizes, sizes):
            taps = oracle.taps(chroma, int(rng.integers(1, 32 if chroma else 16)))
            src = rnd_samples(rng, h + 8, w + 8, bd)
            srcs = rng.integers(-8192, 8191, size=(h + 8, w + 8)).astype(np.int16)
            for kind in range(6):
                s = srcs if kind >= 4 else src
                dt = np.uint16 if kind in (0, 2, 4) else np.int16
                do, dg = np.zeros((h, w + 2), dtype=dt), np.zeros((h, w + 2), dtype=dt)
                oracle.filter(kind, chroma, w, h, bd, taps, s, (3, 3), do)
                lib.filter_block(kind, chroma, w, h, bd, taps, s, (3, 3), dg)
                assert np.array_equal(do, dg), (chroma, kind, w, h)   # incl. untouched columns: writes exactly w x h


@pytest.mark.parametrize("bd", [8, 10, 12])
def test_interp_addavg_copy(oracle, bd):
    rng = np.random.default_rng(104)
    for chroma in (0, 1):
        sizes = CSIZES if chroma else SIZES
        nfrac = 32 if chroma else 16
        for w, h in itertools.product(sizes, sizes):
            refblk = rnd_samples(rng, h + 8, w + 8, bd)
            for fx, fy in ((0, 0), (int(rng.integers(1, nfrac)), 0), (0, int(rng.integers(1, nfrac))),
                           (int(rng.integers(1, nfrac)), int(rng.integers(1, nfrac)))):
                for bip in (0, 1):
                    dt = np.int16 if bip else np.uint16
                    po, pg = np.zeros((h, 64), dtype=dt), np.zeros((h, 64), dtype=dt)
                    oracle.interp(chroma, bip, w, h, bd, fx, fy, refblk, (3, 3), po)
                    lib.interp_block(chroma, bip, w, h, bd, fx, fy, refblk, (3, 3), pg)
                    assert np.array_equal(po, pg), (chroma, w, h, fx, fy, bip)
            a = rng.integers(-8192, 8191, size=(h, 64)).astype(np.int16)
            b = rng.integers(-8192, 8191, size=(h, 64)).astype(np.int16)
            shift = max(2, 14 - bd) + 1
            offset = (1 << (shift - 1)) + 2 * 8192
            do, dg = np.zeros((h, 64), dtype=np.uint16), np.zeros((h, 64), dtype=np.uint16)
            oracle.add_avg(w, h, offset, shift, bd, a, b, do)
            lib.add_avg(w, h, offset, shift, bd, a, b, dg)
            assert np.array_equal(do, dg)
            co, cg = np.zeros((h, 64), dtype=np.int16), np.zeros((h, 64), dtype=np.int16)
            oracle.filter_copy_bipred(w, h, 8192, 14 - bd, refblk, co)
            lib.filter_copy_bipred(w, h, 8192, 14 - bd, refblk, cg)
            assert np.array_equal(co, cg)


@pytest.mark.parametrize("bd", [8, 10, 12])
def test_transforms_dct2_all_shapes(oracle, bd):
    rng = np.random.default_rng(105)
    for w, h in itertools.product([2] + SIZES, [2] + SIZES):
        for amp in (1 << bd, 8):
            resi = rng.integers(-amp + 1, amp, size=(h, w)).astype(np.int16)
            co = oracle.fwd_transform(w, h, bd, 0, 0, 0, resi)
            assert np.array_equal(lib.fwd_transform(w, h, bd, 0, 0, 0, resi), co), ("fwd", w, h)
            assert np.array_equal(lib.inv_transform(w, h, bd, 0, 0, 0, 0, co), oracle.inv_transform(w, h, bd, 0, 0, 0, 0, co))
            cf = rng.integers(-32768, 32768, size=(h, w)).astype(np.int16)
            assert np.array_equal(lib.inv_transform(w, h, bd, 0, 0, 0, 0, cf), oracle.inv_transform(w, h, bd, 0, 0, 0, 0, cf))
            dc = np.zeros((h, w), dtype=np.int16)
            dc[0, 0] = rng.integers(-2000, 2000)
            assert np.array_equal(lib.inv_transform(w, h, bd, 0, 0, 0, 1, dc), oracle.inv_transform(w, h, bd, 0, 0, 0, 1, dc))


@pytest.mark.parametrize("bd", [8, 10])
def test_transforms_all_type_pairs(oracle, bd):
    rng = np.random.default_rng(106)
    for w, h in ((4, 4), (8, 8), (16, 16), (32, 32), (64, 64), (4, 16), (32, 8), (64, 16), (16, 64)):
        resi = rnd_resi(rng, h, w, bd)
        for th, tv in itertools.product(range(6), range(6)):
            co = oracle.fwd_transform(w, h, bd, th, tv, 0, resi)
            assert np.array_equal(lib.fwd_transform(w, h, bd, th, tv, 0, resi), co), ("fwd", w, h, th, tv)
            assert np.array_equal(lib.inv_transform(w, h, bd, th, tv, 0, 0, co),
                                  oracle.inv_transform(w, h, bd, th, tv, 0, 0, co)), ("inv", w, h, th, tv)
    # round trip bound of the reference's own TransformTest.PerfectTxDcPred (transform_test.cc:153-166)
    for n in (4, 8, 16, 32):
        resi = rnd_resi(rng, n, n, bd)
        back = lib.inv_transform(n, n, bd, 0, 0, 0, 0, lib.fwd_transform(n, n, bd, 0, 0, 0, resi))
        assert np.abs(back.astype(np.int32) - resi).max() <= (1 << (bd - 8))


@pytest.mark.parametrize("bd", [8, 10])
def test_dst4x4_and_skip(oracle, bd):
    rng = np.random.default_rng(107)
    for _ in range(8):
        resi = rnd_resi(rng, 4, 4, bd)
        co = oracle.fwd_transform(4, 4, bd, 0, 0, 1, resi)
        assert np.array_equal(lib.fwd_transform(4, 4, bd, 0, 0, 1, resi), co)
        assert np.array_equal(lib.inv_transform(4, 4, bd, 0, 0, 1, 0, co), oracle.inv_transform(4, 4, bd, 0, 0, 1, 0, co))
    for w, h in ((4, 4), (2, 2), (4, 2), (2, 4), (8, 2), (2, 8)):
        resi = rnd_resi(rng, h, w, bd)
        fo = oracle.transform_skip(1, w, h, bd, resi)
        assert np.array_equal(lib.transform_skip(1, w, h, bd, resi), fo)
        assert np.array_equal(lib.transform_skip(0, w, h, bd, fo), oracle.transform_skip(0, w, h, bd, fo))


@pytest.mark.parametrize("bd", [8, 10, 12])
def test_quant_dequant(oracle, bd):
    rng = np.random.default_rng(108)
    for w, h in itertools.product([2] + SIZES, [2] + SIZES):
        for qp in (12, 27, 32, 45):
            comp = int(rng.integers(0, 3))
            qo, qg = oracle.qp(qp, bd), lib.qp_init(qp, bd)
            assert np.array_equal(qo["qp_bitdepth"], qg["qp_bitdepth"]) and np.array_equal(qo["distortion_weight"], qg["distortion_weight"])
            qbd = int(qg["qp_bitdepth"][comp])
            resi = rng.integers(-(1 << bd) // 2, (1 << bd) // 2, size=(h, w)).astype(np.int16)
            coeff = oracle.fwd_transform(w, h, bd, 0, 0, 0, resi)
            for intra_pic, scan in ((0, 0), (1, 1), (1, 2)):
                lo, nzo = oracle.quant_fast(w, h, bd, qbd, intra_pic, 1, scan, coeff)
                lg, nzg = lib.quant_fast(w, h, bd, qbd, intra_pic, 1, scan, coeff)
                assert nzo == nzg and np.array_equal(lo, lg), (w, h, qp, intra_pic, scan)
            lo2, nzo2 = oracle.quant_fast(w, h, bd, qbd, 0, 0, 0, coeff)
            lg2, nzg2 = lib.quant_fast(w, h, bd, qbd, 0, 0, 0, coeff)
            assert nzo2 == nzg2 and np.array_equal(lo2, lg2)
            big = rng.integers(-32768, 32768, size=(h, w)).astype(np.int16)
            assert np.array_equal(lib.dequant(w, h, bd, qbd, big), oracle.dequant(w, h, bd, qbd, big))


# ------------------------------------------------------------------ batched, device resident
def _ctx(width, height, bd, cur, r0, r1=None, slots=8):
    ctx = lib.Context(width, height, bd, slots)
    ctx.upload(0, cur)
    ctx.upload(1, r0)
    ctx.pad_border(1)
    if r1 is not None:
        ctx.upload(2, r1)
        ctx.pad_border(2)
    return ctx


def test_upload_download_pad_border(oracle):
    width, height, bd = 72, 40, 10
    cur, r0, r1 = common.frames(width, height, bd, 111, "random")
    ctx = _ctx(width, height, bd, cur, r0, r1)
    for c, p in enumerate(ctx.download(0)):
        assert np.array_equal(p, cur[c])
    refs = common.oracle_refs(oracle, width, height, r0, r1)
    for c in range(3):
        assert np.array_equal(ctx.download_padded(1, c), refs[(0, 0)].full[c])
        assert np.array_equal(ctx.download_padded(2, c), refs[(1, 0)].full[c])


@pytest.mark.parametrize("content,bd,min_size", [("synth", 10, 4), ("random", 10, 4), ("synth", 8, 8), ("random", 12, 8)])
def test_me_search(oracle, content, bd, min_size):
    width, height = 208, 120
    cur, r0, r1 = common.frames(width, height, bd, 112, content)
    lam = workload.lambda_for_qp(32)
    ctx = _ctx(width, height, bd, cur, r0, r1)
    rng = np.random.default_rng(113)
    cus = workload.make_partition(width, height, seed=3, min_size=min_size)
    cus["flags"][::7] |= abi.CU_FULLPEL_MV
    ctx.set_cus(cus)
    jobs = common.me_jobs(cus, rng, 2, (128, 96), 200, slots=(1, 2))
    rg = ctx.me_search(0, jobs, np.sqrt(lam))
    ojobs = jobs.copy()
    ojobs["ref_slot"] = 0
    ro = oracle.me_search(Picture(width, height, 0, cur), common.oracle_refs(oracle, width, height, r0, r1), bd, cus,
                          ojobs, np.sqrt(lam))
    for f in ("mv_fullpel", "cost_fullpel", "num_sad", "mv", "dist", "cost"):
        assert np.array_equal(rg[f], ro[f]), f
    if content == "synth":   # the search actually finds the pan (2,1) px/frame x 8 frames
        big = (cus["w"] >= 16) & (cus["h"] >= 16)
        mv0 = rg["mv"][0::2][big]
        assert np.median(mv0[:, 0]) == 16 * 16 and np.median(mv0[:, 1]) == 8 * 16


@pytest.mark.parametrize("ranges", [(1, 2), (3, 5), (8, 16), (40, 64), (200, 256)])
def test_me_search_ranges(oracle, ranges):
    """Every number of diamond rounds a pass can have (1 .. 9: radii 1, 2, 4, ... <= range), incl.
    ranges that are not powers of two and windows narrower than a column chunk of the raster bound."""
    width, height, bd = 208, 120, 10
    cur, r0, r1 = common.frames(width, height, bd, 212, "synth")
    lam = workload.lambda_for_qp(30)
    ctx = _ctx(width, height, bd, cur, r0, r1)
    rng = np.random.default_rng(214 + ranges[0])
    cus = workload.make_partition(width, height, seed=5, min_size=4)
    ctx.set_cus(cus)
    jobs = common.me_jobs(cus, rng, 2, ranges, 60, slots=(1, 2))
    rg = ctx.me_search(0, jobs, np.sqrt(lam))
    ojobs = jobs.copy()
    ojobs["ref_slot"] = 0
    ro = oracle.me_search(Picture(width, height, 0, cur), common.oracle_refs(oracle, width, height, r0, r1), bd, cus,
                          ojobs, np.sqrt(lam))
    for f in ("mv_fullpel", "cost_fullpel", "num_sad", "mv", "dist", "cost"):
        assert np.array_equal(rg[f], ro[f]), f
    bad = jobs[:1].copy()
    bad["search_range"] = 257
    with pytest.raises(lib.XvcB200Error):
        ctx.me_search(0, bad, np.sqrt(lam))


def test_me_search_far_predictors(oracle):
    """Predictors far outside the picture: ClipMv, window clipping and reads from the border."""
    width, height, bd = 136, 72, 10
    cur, r0, r1 = common.frames(width, height, bd, 114)
    lam = workload.lambda_for_qp(37)
    ctx = _ctx(width, height, bd, cur, r0, r1)
    rng = np.random.default_rng(115)
    cus = workload.make_partition(width, height, seed=5, min_size=8)
    ctx.set_cus(cus)
    jobs = common.me_jobs(cus, rng, 2, (256, 96), 5000, slots=(1, 2))
    rg = ctx.me_search(0, jobs, np.sqrt(lam))
    ojobs = jobs.copy()
    ojobs["ref_slot"] = 0
    ro = oracle.me_search(Picture(width, height, 0, cur), common.oracle_refs(oracle, width, height, r0, r1), bd, cus,
                          ojobs, np.sqrt(lam))
    for f in ("mv_fullpel", "cost_fullpel", "num_sad", "mv", "dist", "cost"):
        assert np.array_equal(rg[f], ro[f]), f


def test_me_search_recentred_raster(oracle):
    """A start point other than the predictor re-centres the raster window (DetermineMinMaxMv around the best
    start, inter_tz_search.cc:121-125): with a predictor far from the content's motion the zero vector wins the
    start, the first pass ends far out and the scan window lies outside the box staged around the predictors'
    windows -- the kernel stages a second box for those jobs (and scans from global memory what does not fit:
    the second list's predictors are scattered so that the windows of a CTU cannot share one box)."""
    width, height, bd = 448, 256, 10
    cur, r0, r1 = common.frames(width, height, bd, 118, "synth")
    lam = workload.lambda_for_qp(32)
    ctx = _ctx(width, height, bd, cur, r0, r1)
    rng = np.random.default_rng(119)
    cus = workload.make_partition(width, height, seed=9, min_size=8)
    ctx.set_cus(cus)
    jobs = common.me_jobs(cus, rng, 2, (128, 96), 0, slots=(1, 2))
    jobs["mvp"][0::2] = (-1500, -900)                                        # one far predictor for every CU
    jobs["mvp"][1::2] = rng.integers(-2400, 2401, size=(len(cus), 2))         # scattered far predictors
    jobs["prev"] = 0
    rg = ctx.me_search(0, jobs, np.sqrt(lam))
    ojobs = jobs.copy()
    ojobs["ref_slot"] = 0
    ro = oracle.me_search(Picture(width, height, 0, cur), common.oracle_refs(oracle, width, height, r0, r1), bd, cus,
                          ojobs, np.sqrt(lam))
    for f in ("mv_fullpel", "cost_fullpel", "num_sad", "mv", "dist", "cost"):
        assert np.array_equal(rg[f], ro[f]), f
    scanned = ro["num_sad"] > 1000                                           # the raster scan ran
    assert scanned[0::2].mean() > 0.3 and scanned[1::2].mean() > 0.3


def test_full_search(oracle):
    width, height, bd = 136, 72, 10
    cur, r0, r1 = common.frames(width, height, bd, 116)
    lam = workload.lambda_for_qp(32)
    ctx = _ctx(width, height, bd, cur, r0, r1)
    rng = np.random.default_rng(117)
    cus = workload.make_partition(width, height, seed=6, min_size=4)
    ctx.set_cus(cus)
    jobs = np.zeros(len(cus), dtype=abi.fullsearch_job_dtype)
    for i in range(len(cus)):
        jobs[i]["cu"], jobs[i]["ref_slot"], jobs[i]["other_pred_slot"] = i, 1, 2
        jobs[i]["mvp"] = rng.integers(-64, 65, size=2)
        jobs[i]["center"] = rng.integers(-64, 65, size=2)
        jobs[i]["range"] = 4
    rg = ctx.full_search(0, jobs, np.sqrt(lam))
    refs = common.oracle_refs(oracle, width, height, r0, r1)
    orig = Picture(width, height, 0, cur)
    lam_me = int(np.floor(65536.0 * np.sqrt(lam)))
    for i in range(len(cus)):
        mv, cost = oracle.full_search(orig, refs[(1, 0)], refs[(0, 0)], bd, cus[i:i + 1], jobs[i:i + 1], lam_me)
        assert np.array_equal(mv, rg[i]["mv_fullpel"]) and cost == rg[i]["cost_fullpel"], i


@pytest.mark.parametrize("bd", [8, 10, 12])
def test_motion_compensate(oracle, bd):
    width, height = 136, 72
    cur, r0, r1 = common.frames(width, height, bd, 118, "random")
    ctx = _ctx(width, height, bd, cur, r0, r1)
    cus = common.mc_cus(width, height, np.random.default_rng(119), 8)
    ctx.set_cus(cus)
    ctx.motion_compensate({(0, 0): 1, (1, 0): 2}, 3)
    pred = Picture(width, height, 80)
    oracle.motion_compensate(common.oracle_refs(oracle, width, height, r0, r1), bd, cus, pred)
    for c, p in enumerate(ctx.download(3)):
        assert np.array_equal(p, pred.plane(c)), c


@pytest.mark.parametrize("bd", [8, 10, 12])
def test_motion_compensate_affine(oracle, bd):
    """xvcb200_motion_compensate_affine == MotionCompAffine (oracle pinned against the reference in
    test_oracle_vs_ref.py::test_motion_compensate_affine), bit for bit."""
    width, height = 200, 136
    cur, r0, r1 = common.frames(width, height, bd, 128, "random")
    ctx = _ctx(width, height, bd, cur, r0, r1)
    rng = np.random.default_rng(129)
    cus = common.mc_cus(width, height, rng, 8)
    aff = common.affine_cus(cus, rng)
    assert len(aff) >= 10
    ctx.set_cus(cus)
    ctx.motion_compensate({(0, 0): 1, (1, 0): 2}, 3)
    ctx.motion_compensate_affine(aff, {(0, 0): 1, (1, 0): 2}, 3)
    pred = Picture(width, height, 80)
    refs = common.oracle_refs(oracle, width, height, r0, r1)
    oracle.motion_compensate(refs, bd, cus, pred)
    oracle.motion_compensate_affine(refs, bd, cus, aff, pred)
    for c, p in enumerate(ctx.download(3)):
        assert np.array_equal(p, pred.plane(c)), c
    bad = aff[:1].copy()
    bad["cu"] = len(cus)
    with pytest.raises(lib.XvcB200Error):
        ctx.motion_compensate_affine(bad, {(0, 0): 1, (1, 0): 2}, 3)


@pytest.mark.parametrize("bd,content", [(8, "synth"), (10, "random"), (12, "synth")])
def test_motion_compensate_lic(oracle, bd, content):
    """xvcb200_motion_compensate_lic == MotionCompensation with use_lic (oracle pinned against the
    reference in test_oracle_vs_ref.py::test_motion_compensate_lic), bit for bit; a subset of the CUs
    uses LIC, the others keep their translational prediction."""
    width, height = 200, 136
    cur, r0, r1 = common.frames(width, height, bd, 138, content)
    rng = np.random.default_rng(139)
    cus = common.mc_cus(width, height, rng, 12, min_size=4)
    rec_planes = [np.clip(p.astype(np.int32) * 7 // 8 + (3 << (bd - 8)) + rng.integers(-2, 3, size=p.shape), 0, (1 << bd) - 1).astype(np.uint16)
                  for p in cur]
    ctx = lib.Context(width, height, bd, num_slots=5)
    for slot, f in ((1, r0), (2, r1)):
        ctx.upload(slot, f)
        ctx.pad_border(slot)
    ctx.upload(4, rec_planes)
    ctx.set_cus(cus)
    lic = common.lic_cus(cus, width, height, indices=[i for i in range(len(cus)) if i % 4 != 3])
    assert len(lic) >= 20
    ctx.motion_compensate({(0, 0): 1, (1, 0): 2}, 3)
    ctx.motion_compensate_lic(lic, {(0, 0): 1, (1, 0): 2}, 4, 3)
    refs = common.oracle_refs(oracle, width, height, r0, r1)
    pred = Picture(width, height, 80)
    oracle.motion_compensate(refs, bd, cus, pred)
    oracle.motion_compensate_lic(refs, Picture(width, height, 80, rec_planes), bd, cus, lic, pred)
    for c, p in enumerate(ctx.download(3)):
        assert np.array_equal(p, pred.plane(c)), c
    ctx.close()


@pytest.mark.parametrize("bd,qp,min_size", [(10, 32, 4), (10, 22, 8), (8, 37, 4), (12, 27, 8)])
def test_tq_reconstruct(oracle, bd, qp, min_size):
    width, height = 200, 136
    cur, r0, r1 = common.frames(width, height, bd, 120)
    cus = workload.make_partition(width, height, seed=9, min_size=min_size, qp=qp)
    cus["qp"][::3] = qp + 3
    cus["ref_idx"][:, 0] = 0
    rng = np.random.default_rng(121)
    predp = [np.clip(p.astype(np.int32) + rng.integers(-(60 << (bd - 8)), (60 << (bd - 8)) + 1, size=p.shape), 0, (1 << bd) - 1).astype(np.uint16) for p in cur]
    ctx = lib.Context(width, height, bd, 6)
    ctx.upload(0, cur)
    ctx.upload(3, predp)
    ctx.set_cus(cus)
    tg = ctx.tq_reconstruct(0, 3, 4, 5)
    orig, pred, rec = Picture(width, height, 0, cur), Picture(width, height, 80, predp), Picture(width, height, 80)
    cus_o = cus.copy()
    levels, to = oracle.tq_reconstruct(orig, pred, rec, bd, cus_o)
    assert np.array_equal(tg["num_non_zero"], to["num_non_zero"])
    assert np.array_equal(tg["ssd"], to["ssd"])
    recg, levg = ctx.download(4), ctx.download_coeff(5)
    for c in range(3):
        assert np.array_equal(recg[c], rec.plane(c)), c
        assert np.array_equal(levg[c], levels[c]), c
    assert np.array_equal(ctx.get_cus()["flags"], cus_o["flags"])
    assert (to["num_non_zero"] > 1).sum() > len(cus) // 4          # sign hiding had something to do
    # decoder side: levels + cbf flags -> the same reconstruction (CuDecoder::DecompressComponent)
    ctx.upload(4, [np.zeros_like(p) for p in cur])
    ctx.dequant_reconstruct(3, 4, 5)
    for c, p in enumerate(ctx.download(4)):
        assert np.array_equal(p, rec.plane(c)), c


@pytest.mark.parametrize("pic_type,bd,min_size,same_poc", [(0, 10, 4, False), (0, 10, 8, True), (1, 10, 4, False), (0, 8, 4, False)])
def test_deblock(oracle, pic_type, bd, min_size, same_poc):
    width, height = 200, 104
    rng = np.random.default_rng(122 + pic_type + min_size)
    canvas = workload.synth_canvas(width, height, 5)
    cur = workload.synth_frame(canvas, width, height, 3, bd)
    cus = common.deblock_cus(width, height, rng, 10 + min_size, min_size, pic_type)
    recp = common.blocky_recon(cur, cus, rng, bd)
    poc = {(0, 0): 0, (1, 0): 0 if same_poc else 16}
    ctx = lib.Context(width, height, bd, 2)
    ctx.upload(0, recp)
    ctx.set_cus(cus)
    ctx.deblock_picture(0, pic_type, poc)
    rec = Picture(width, height, 80, recp)
    oracle.deblock_picture(rec, bd, cus, pic_type, poc)
    out = ctx.download(0)
    changed = sum(int((out[c] != recp[c]).sum()) for c in range(3))
    for c in range(3):
        assert np.array_equal(out[c], rec.plane(c)), c
    assert changed > 100


def test_deblock_is_idempotent_on_flat_picture(oracle):
    """Size-independent property: a constant picture has no edges to smooth."""
    width, height, bd = 128, 64, 10
    flat = [np.full((height, width), 512, np.uint16), np.full((height // 2, width // 2), 300, np.uint16),
            np.full((height // 2, width // 2), 700, np.uint16)]
    cus = common.deblock_cus(width, height, np.random.default_rng(5), 3, 4, 0)
    ctx = lib.Context(width, height, bd, 1)
    ctx.upload(0, flat)
    ctx.set_cus(cus)
    ctx.deblock_picture(0, 0, {(0, 0): 0, (1, 0): 16})
    for c, p in enumerate(ctx.download(0)):
        assert np.array_equal(p, flat[c])


@pytest.mark.parametrize("pic_type,bd", [(0, 10), (1, 10), (0, 8)])
def test_encode_picture(oracle, pic_type, bd):
    width, height, qp = 200, 104, 32
    cur, r0, r1 = common.frames(width, height, bd, 123)
    lam = workload.lambda_for_qp(qp)
    cus = workload.make_partition(width, height, seed=11, min_size=8, qp=qp)
    ctx = _ctx(width, height, bd, cur, r0, r1 if pic_type == 0 else None)
    ctx.set_cus(cus)
    slots = dict(orig=0, ref0=1, ref1=2 if pic_type == 0 else -1, pred=3, rec=4, coeff=5)
    prm = common.picture_params(pic_type, lam, slots=slots)
    me_g, tu_g = ctx.encode_picture(prm)
    ctx.sync()
    pred, rec = Picture(width, height, 80), Picture(width, height, 80)
    cus_o = cus.copy()
    refs = common.oracle_refs(oracle, width, height, r0, r1 if pic_type == 0 else None)
    levels, me_o, tu_o = oracle.encode_picture(Picture(width, height, 0, cur), refs, pred, rec, bd, cus_o, prm)
    for f in ("mv_fullpel", "mv", "cost_fullpel", "dist", "cost", "num_sad"):
        assert np.array_equal(me_g[f], me_o[f]), f
    assert np.array_equal(tu_g, tu_o)
    cus_g = ctx.get_cus()
    for f in ("flags", "ref_idx", "mv"):
        assert np.array_equal(cus_g[f], cus_o[f]), f
    levg = ctx.download_coeff(5)
    for c in range(3):
        assert np.array_equal(ctx.download_padded(4, c), rec.full[c]), c
        assert np.array_equal(levg[c], levels[c]), c


def test_async_transfers_pipeline(oracle):
    """Two pictures in flight on alternating slot sets (uploads / downloads on the copy stream)
    give the same levels, reconstruction and CU decisions as the synchronous calls, picture by
    picture -- the events inside the library keep every transfer ordered against the kernels."""
    width, height, bd, qp = 200, 104, 10, 32
    lam = workload.lambda_for_qp(qp)
    pics = [common.frames(width, height, bd, 500 + i) for i in range(4)]
    cus = workload.make_partition(width, height, seed=19, min_size=8, qp=qp)
    ctx = lib.Context(width, height, bd, num_slots=9)
    # slots: 1/2 references (shared), 3 prediction; set A = (0 orig, 4 rec, 5 levels), set B = (6, 7, 8)
    _, r0, r1 = pics[0]
    for s, f in ((1, r0), (2, r1)):
        ctx.upload(s, f)
        ctx.pad_border(s)
    sets = [dict(orig=0, rec=4, coeff=5), dict(orig=6, rec=7, coeff=8)]
    prms = [common.picture_params(0, lam, slots=dict(orig=st["orig"], ref0=1, ref1=2, pred=3, rec=st["rec"], coeff=st["coeff"]))
            for st in sets]
    # synchronous truth
    truth = []
    for cur, _, _ in pics:
        ctx.upload(0, cur)
        ctx.set_cus(cus)
        ctx.encode_picture(prms[0], want_results=False)
        truth.append((ctx.download(4), ctx.download_coeff(5), ctx.get_cus()))
    # pipelined
    host = [dict(rec=[np.zeros_like(p) for p in pics[0][0]], lev=[np.zeros(p.shape, dtype=np.int16) for p in pics[0][0]],
                 cus=np.zeros(len(cus), dtype=abi.cu_dtype)) for _ in range(2)]
    got = []

    def collect(i):
        s = i & 1
        ctx.wait_download(sets[s]["rec"])
        got.append(([p.copy() for p in host[s]["rec"]], [p.copy() for p in host[s]["lev"]], host[s]["cus"].copy()))

    ctx.upload_async(sets[0]["orig"], pics[0][0])
    for i in range(len(pics)):
        s = i & 1
        if i + 1 < len(pics):
            ctx.upload_async(sets[1 - s]["orig"], pics[i + 1][0])
        ctx.set_cus(cus)
        ctx.encode_picture(prms[s], want_results=False)
        ctx.get_cus_async(host[s]["cus"])
        ctx.download_coeff_async(sets[s]["coeff"], host[s]["lev"])
        ctx.download_async(sets[s]["rec"], host[s]["rec"])
        if i >= 1:
            collect(i - 1)
    collect(len(pics) - 1)
    ctx.sync_copies()
    ctx.sync()
    for i, ((rec_t, lev_t, cus_t), (rec_g, lev_g, cus_g)) in enumerate(zip(truth, got)):
        for c in range(3):
            assert np.array_equal(rec_t[c], rec_g[c]), (i, c)
            assert np.array_equal(lev_t[c], lev_g[c]), (i, c)
        for f in ("flags", "ref_idx", "mv"):
            assert np.array_equal(cus_t[f], cus_g[f]), (i, f)
    assert any(np.any(l[0]) for _, l, _ in got)      # the pictures have non-zero levels


def test_empty_inputs():
    """Edge case: a picture with no inter CU at all / empty job lists are no-ops that succeed (the
    reference's loops simply do not execute), and leave the context usable."""
    width, height, bd = 64, 64, 10
    cur, r0, r1 = common.frames(width, height, bd, 171)
    ctx = _ctx(width, height, bd, cur, r0, r1)
    ctx.set_cus(np.zeros(0, dtype=abi.cu_dtype))
    prm = common.picture_params(0, workload.lambda_for_qp(32), slots=dict(orig=0, ref0=1, ref1=2, pred=3, rec=4, coeff=5))
    me, tu = ctx.encode_picture(prm)
    assert len(me) == 0 and len(tu) == 0
    ctx.motion_compensate({(0, 0): 1, (1, 0): 2}, 3)
    ctx.motion_compensate_affine(np.zeros(0, dtype=abi.affine_cu_dtype), {(0, 0): 1, (1, 0): 2}, 3)
    ctx.motion_compensate_lic(np.zeros(0, dtype=abi.lic_cu_dtype), {(0, 0): 1, (1, 0): 2}, 4, 3)
    ctx.intra_lm_chroma(4, np.zeros(0, dtype=abi.intra_job_dtype), 3)
    ctx.sync()
    # and the context still works
    cus = workload.make_partition(width, height, seed=172, min_size=8)
    ctx.set_cus(cus)
    me, tu = ctx.encode_picture(prm)
    ctx.sync()
    assert len(me) == 2 * len(cus) and len(tu) == 3 * len(cus)
