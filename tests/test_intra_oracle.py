"""Intra prediction: the C restatement against the unmodified reference.

CUs are walked in coding order through the reference (RefSession.intra_scan): every CU sees only
the CUs before it, so corner / edge / missing above-right / missing below-left cases all occur.
Compared: the neighbour-derived reference samples (ComputeRefSamples), their smoothing
(FilterRefSamples), the prediction of all 67 modes for luma and chroma, and the luma SATD scan.
"""
import numpy as np
import pytest

import common
from xvc_b200 import abi, workload


def intra_inputs(width, height, bd, seed, min_size=4, content="synth"):
    cur, rec, _ = common.frames(width, height, bd, seed, content)
    cus = workload.make_partition(width, height, seed=seed, min_size=min_size)
    cus["flags"] |= abi.CU_INTRA
    return cur, rec, cus


@pytest.mark.parametrize("bd,content,min_size", [(10, "synth", 4), (8, "random", 4), (12, "random", 8), (10, "random", 8)])
def test_intra_oracle_vs_ref(oracle, ref, bd, content, min_size):
    width, height = 136, 104          # not multiples of 64: partial CTUs on the right and at the bottom
    cur, rec, cus = intra_inputs(width, height, bd, 31 + bd, min_size, content)
    for comp in (0, 1):
        ses = ref.session(width, height, bd, pic_type=2)
        ses.set_orig(cur)
        ses.set_rec(rec)
        jobs, ref_r, filt_r, preds_r, satd_r = ses.intra_scan(cus, comp)
        ses.close()
        assert np.array_equal(jobs, common.intra_jobs_in_coding_order(cus, width, height, comp))
        seen = set()
        for i, j in enumerate(jobs):
            nb = (j["has_above_left"], j["has_above"], j["above_right"], j["has_left"], j["below_left"])
            seen.add((bool(nb[0]), bool(nb[1]), nb[2] > 0, bool(nb[3]), nb[4] > 0))
            w, h = int(j["w"]), int(j["h"])
            ref_o, filt_o = oracle.intra_ref_samples(w, h, bd, nb, rec[comp], int(j["x"]), int(j["y"]))
            used = np.zeros(2 * abi.INTRA_REF_STRIDE, dtype=bool)
            used[:w + h + 1] = True
            used[abi.INTRA_REF_STRIDE:abi.INTRA_REF_STRIDE + w + h] = True
            assert np.array_equal(ref_o[used], ref_r[i][used]), (comp, i)
            if comp == 0:
                assert np.array_equal(filt_o[used], filt_r[i][used]), (comp, i)
            for mode in range(abi.INTRA_NUM_MODES):
                p = oracle.intra_predict(mode, w, h, bd, comp == 0, ref_r[i], filt_r[i] if comp == 0 else None)
                assert np.array_equal(p, preds_r[i][mode]), (comp, i, mode, w, h)
            if comp == 0:
                s = oracle.intra_satd_scan(w, h, bd, cur[0], int(j["x"]), int(j["y"]), ref_r[i], filt_r[i])
                assert np.array_equal(s, satd_r[i]), (i, w, h)
        # availability patterns the walk produced: nothing / left only / above only / everything, with and without the far ends
        assert (False, False, False, False, False) in seen and (True, True, True, True, True) in seen
        assert any(s[1] and not s[2] for s in seen) and any(s[3] and not s[4] for s in seen)


def test_intra_golden_oracle(oracle):
    """The committed vectors (tests/golden/xvc_intra_golden.npz, generated from the reference by
    tests/golden/make_intra_golden.py) replayed against the C restatement."""
    import intra_golden
    intra_golden.replay(intra_golden.OracleBackend(oracle))


@pytest.mark.parametrize("bd,content,seed", [(8, "synth", 71), (10, "synth", 72), (10, "random", 73), (12, "random", 74)])
def test_lm_chroma_oracle_vs_reference(oracle, ref, bd, content, seed):
    """IntraPrediction::Predict(kLmChroma) (PredLmChroma / RescaleLuma / DeriveLmParams,
    intra_prediction.cc:560-686, 873-913) for every CU of a partition -- picture borders (no / one
    neighbour side), all block shapes 8..64 (chroma 4..32), both chroma components."""
    width, height = 200, 136
    cur, rec, _ = common.frames(width, height, bd, seed, content)
    if content == "synth":      # give chroma a relation to luma that is not the generator's own
        rng = np.random.default_rng(seed)
        rec = [rec[0]] + [np.clip(p.astype(np.int32) + rng.integers(-6, 7, size=p.shape) * (1 << (bd - 8)), 0, (1 << bd) - 1).astype(np.uint16)
                          for p in rec[1:]]
    cus = workload.make_partition(width, height, seed=seed, min_size=8)
    cus["flags"] |= abi.CU_INTRA
    ses = ref.session(width, height, bd, pic_type=2)
    ses.set_orig(cur)
    ses.set_rec(rec)
    got = ses.intra_lm_chroma(cus)
    ses.close()
    for i, cu in enumerate(cus):
        for comp in (1, 2):
            p = oracle.intra_lm_chroma(rec, comp, int(cu["x"]), int(cu["y"]), int(cu["w"]), int(cu["h"]), bd)
            assert np.array_equal(p, got[i][comp - 1]), (i, comp, cu)


def test_lm_golden_oracle(oracle):
    """tests/golden/xvc_lm_golden.npz (reference LM chroma outputs) replayed against the C restatement."""
    import intra_golden
    intra_golden.replay_lm(intra_golden.lm_oracle_backend(oracle))
