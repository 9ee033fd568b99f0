"""Intra prediction kernels through the C ABI against the oracle and the golden vectors."""
import numpy as np
import pytest

import common
import intra_golden
from xvc_b200 import abi, lib, workload

pytestmark = pytest.mark.gpu


def _used(w, h):
    used = np.zeros(2 * abi.INTRA_REF_STRIDE, dtype=bool)
    used[:w + h + 1] = True
    used[abi.INTRA_REF_STRIDE:abi.INTRA_REF_STRIDE + w + h] = True
    return used


@pytest.mark.parametrize("bd", [8, 10, 12])
def test_intra_predict_all_modes_all_shapes(oracle, bd):
    """xvcb200_intra_predict == IntraPrediction::Predict for every mode and block shape, luma (smoothing,
    edge filters) and chroma, on full-range random reference samples."""
    rng = np.random.default_rng(900 + bd)
    for w in (4, 8, 16, 32, 64):
        for h in (4, 8, 16, 32, 64):
            ref = rng.integers(0, 1 << bd, size=2 * abi.INTRA_REF_STRIDE, dtype=np.uint16)
            filt = rng.integers(0, 1 << bd, size=2 * abi.INTRA_REF_STRIDE, dtype=np.uint16)
            for luma in (True, False):
                for mode in (range(abi.INTRA_NUM_MODES) if (w + h) % 24 == 0 or w == h else rng.choice(abi.INTRA_NUM_MODES, 12, replace=False)):
                    mode = int(mode)
                    got = lib.intra_predict(mode, w, h, bd, luma, ref, filt if luma else None)
                    want = oracle.intra_predict(mode, w, h, bd, luma, ref, filt if luma else None)
                    assert np.array_equal(got, want), (w, h, luma, mode)
    for w, h in ((2, 2), (2, 8), (8, 2), (2, 32)):          # 2-wide chroma blocks
        ref = rng.integers(0, 1 << bd, size=2 * abi.INTRA_REF_STRIDE, dtype=np.uint16)
        for mode in range(abi.INTRA_NUM_MODES):
            assert np.array_equal(lib.intra_predict(mode, w, h, bd, False, ref, None), oracle.intra_predict(mode, w, h, bd, False, ref, None)), (w, h, mode)


@pytest.mark.parametrize("bd,content,min_size", [(10, "synth", 4), (8, "random", 4), (12, "random", 8)])
def test_intra_refs_and_scan_in_coding_order(oracle, bd, content, min_size):
    """Reference samples with every availability pattern a coding-order walk produces, and the batched
    SATD scan of all 67 modes against the oracle."""
    width, height = 200, 136
    cur, rec, _ = common.frames(width, height, bd, 77 + bd, content)
    cus = workload.make_partition(width, height, seed=13 + bd, min_size=min_size)
    jobs = common.intra_jobs_in_coding_order(cus, width, height)
    refs, filts = [], []
    for i, j in enumerate(jobs):
        nb = (j["has_above_left"], j["has_above"], j["above_right"], j["has_left"], j["below_left"])
        w, h = int(j["w"]), int(j["h"])
        r_o, f_o = oracle.intra_ref_samples(w, h, bd, nb, rec[0], int(j["x"]), int(j["y"]))
        refs.append(r_o)
        filts.append(f_o)
        if i % 3 == 0:
            r_g, f_g = lib.intra_ref_samples(w, h, bd, nb, rec[0], int(j["x"]), int(j["y"]))
            assert np.array_equal(r_g[_used(w, h)], r_o[_used(w, h)]), i
            assert np.array_equal(f_g[_used(w, h)], f_o[_used(w, h)]), i
    ctx = lib.Context(width, height, bd, num_slots=2)
    ctx.upload(0, cur)
    ctx.upload(1, rec)
    got = ctx.intra_satd_scan(0, 1, jobs)
    want = np.stack([oracle.intra_satd_scan(int(j["w"]), int(j["h"]), bd, cur[0], int(j["x"]), int(j["y"]), refs[i], filts[i])
                     for i, j in enumerate(jobs)])
    assert np.array_equal(got, want)
    # chroma reference samples (half-size geometry, no smoothing requested)
    cj = common.intra_jobs_in_coding_order(cus, width, height, comp=1)
    for i in range(0, len(cj), 5):
        j = cj[i]
        nb = (j["has_above_left"], j["has_above"], j["above_right"], j["has_left"], j["below_left"])
        w, h = int(j["w"]), int(j["h"])
        r_g, _ = lib.intra_ref_samples(w, h, bd, nb, rec[1], int(j["x"]), int(j["y"]), want_filtered=False)
        r_o, _ = oracle.intra_ref_samples(w, h, bd, nb, rec[1], int(j["x"]), int(j["y"]))
        assert np.array_equal(r_g[_used(w, h)], r_o[_used(w, h)]), i
    # invalid jobs are rejected, not executed
    bad = jobs[:1].copy()
    bad["w"] = 12
    with pytest.raises(lib.XvcB200Error):
        ctx.intra_satd_scan(0, 1, bad)
    ctx.close()


def test_intra_scan_1080p(oracle):
    """Pre-analysis use at the benchmark resolution: all CUs of a 1080p partition at once, reference samples
    from the original picture with every neighbour available; a sample of CUs against the oracle."""
    width, height, bd = 1920, 1080, 10
    canvas = workload.synth_canvas(width, height, 1234)
    cur = workload.synth_frame(canvas, width, height, 8, bd)
    cus = workload.make_partition(width, height, seed=7, min_size=8)
    jobs = np.zeros(len(cus), dtype=abi.intra_job_dtype)
    jobs["x"], jobs["y"], jobs["w"], jobs["h"] = cus["x"], cus["y"], cus["w"], cus["h"]
    jobs["has_left"], jobs["has_above"] = cus["x"] > 0, cus["y"] > 0
    jobs["has_above_left"] = (cus["x"] > 0) & (cus["y"] > 0)
    jobs["above_right"] = np.where(cus["y"] > 0, np.minimum(cus["h"], width - cus["x"] - cus["w"]), 0)
    jobs["below_left"] = np.where(cus["x"] > 0, np.minimum(cus["w"], height - cus["y"] - cus["h"]), 0)
    ctx = lib.Context(width, height, bd, num_slots=1)
    ctx.upload(0, cur)
    got = ctx.intra_satd_scan(0, 0, jobs)
    rng = np.random.default_rng(5)
    for i in rng.choice(len(jobs), 40, replace=False):
        j = jobs[i]
        nb = (j["has_above_left"], j["has_above"], j["above_right"], j["has_left"], j["below_left"])
        r, f = oracle.intra_ref_samples(int(j["w"]), int(j["h"]), bd, nb, cur[0], int(j["x"]), int(j["y"]))
        want = oracle.intra_satd_scan(int(j["w"]), int(j["h"]), bd, cur[0], int(j["x"]), int(j["y"]), r, f)
        assert np.array_equal(got[i], want), i
    ctx.close()


def test_intra_golden_gpu():
    intra_golden.replay(intra_golden.GpuBackend())


def test_lm_golden_gpu():
    """Reference LM chroma outputs (tests/golden/xvc_lm_golden.npz) == xvcb200_intra_lm_chroma."""
    intra_golden.replay_lm(intra_golden.lm_gpu_backend())


@pytest.mark.parametrize("bd", [8, 10, 12])
def test_lm_chroma_gpu_vs_oracle(oracle, bd):
    """Every CU of a random partition 4..64 (chroma 2..32), picture borders included, both chroma components."""
    width, height = 264, 136
    _, rec, _ = common.frames(width, height, bd, 90 + bd, "random" if bd == 10 else "synth")
    cus = workload.make_partition(width, height, seed=90 + bd, min_size=4)
    ctx = lib.Context(width, height, bd, num_slots=2)
    ctx.upload(0, rec)
    ctx.intra_lm_chroma(0, intra_golden.lm_jobs(cus), 1)
    got = ctx.download(1)
    for cu in cus:
        x, y, w, h = int(cu["x"]), int(cu["y"]), int(cu["w"]), int(cu["h"])
        for comp in (1, 2):
            exp = oracle.intra_lm_chroma(rec, comp, x, y, w, h, bd)
            assert np.array_equal(got[comp][y // 2:(y + h) // 2, x // 2:(x + w) // 2], exp), (cu, comp)
    bad = intra_golden.lm_jobs(cus[:1])
    bad["x"] = width
    with pytest.raises(lib.XvcB200Error):
        ctx.intra_lm_chroma(0, bad, 1)
    ctx.close()
