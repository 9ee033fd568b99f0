"""xvcb200_encode_picture as InterSearch::SearchMotion for a whole picture: several reference pictures
per list, list-1 pictures that repeat a list-0 POC, SearchBiIterative (FullSearch + sub-pel search on
the weighted original), the fast_inter_pred_bits rate, CUs exempt from the search -- against the
UNMODIFIED reference's classes driven by oracle/ref_shim.cc (xref_encode_picture: TzSearch::Search,
InterSearch::FullSearch / SubpelSearch / MotionCompensation / GetInterPredBits, the residual chain,
DeblockingFilter, PadBorder).  Needs oracle/_ref (built from /root/reference, travels prebuilt)."""
import os

import numpy as np
import pytest

import common
from oracle import bindings
from xvc_b200 import abi, lib, workload

pytestmark = pytest.mark.gpu

POC = 8


@pytest.fixture(scope="module")
def ref():
    if not bindings.have_ref():
        pytest.skip("oracle/_ref/libxvcref.so not built (needs /root/reference)")
    lib.load()
    return bindings.Ref()


def run_both(ref, width, height, bd, qp, pic_type, pocs, bi_iterations, seed, predictors="field", mixed=False,
             max_range=64, threads=4, min_size=8, deblock=1, bits_mode=1):
    """One picture through the GPU step and through the reference's classes; returns both result sets."""
    canvas = workload.synth_canvas(width, height, seed)
    frame = lambda poc: workload.synth_frame(canvas, width, height, poc, bd, frame_noise=4.0)
    lam = workload.lambda_for_qp(qp)
    cus = workload.make_partition(width, height, seed=seed + 1, min_size=min_size, qp=qp)
    rng = np.random.default_rng(seed + 2)
    mvp = None
    if predictors == "field":            # one predictor per list: the CU's mv[list]
        workload.set_predictors(cus, POC, (pocs[0][0], pocs[1][0] if pocs[1] else None), seed=seed + 3)
    elif predictors == "per_ref":        # one per (list, reference picture): xvcb200_set_mv_predictors
        mvp = workload.mv_predictors(cus, POC, pocs, seed=seed + 3)
    elif predictors == "random":
        cus["mv"] = rng.integers(-400, 401, size=cus["mv"].shape)
    if mixed:
        for i in range(len(cus)):
            k = i % 7
            if k == 2:
                cus[i]["flags"] |= abi.CU_INTRA
            elif k == 4:                      # a merge / skip CU: vectors fixed by the host
                cus[i]["flags"] |= abi.CU_SKIP_ME
                mode = (i // 7) % (3 if pic_type == 0 else 1)
                cus[i]["ref_idx"] = [(0, -1), (-1, len(pocs[1]) - 1), (len(pocs[0]) - 1, 0)][mode]
            elif k == 5:
                cus[i]["flags"] |= abi.CU_FULLPEL_MV
    uniq = sorted({p for l in pocs for p in l})
    slot_of = {p: 1 + k for k, p in enumerate(uniq)}
    base = 1 + len(uniq)
    slots = dict(orig=0, pred=base, rec=base + 1, coeff=base + 2)
    prm = np.zeros(1, dtype=abi.picture_params_dtype)
    prm["pic_type"], prm["lambda_sqrt"], prm["chroma_offset_table"] = pic_type, np.sqrt(lam), 1
    prm["deblock"], prm["pad"] = deblock, 1
    prm["bi_iterations"], prm["bits_mode"] = bi_iterations, bits_mode
    prm["ref_slots"] = -1
    prm["orig_slot"], prm["pred_slot"], prm["rec_slot"], prm["coeff_slot"] = 0, slots["pred"], slots["rec"], slots["coeff"]
    for l in range(2):
        prm["num_ref"][0, l] = len(pocs[l])
        for r, p in enumerate(pocs[l]):
            prm["ref_slots"][0, l, r] = slot_of[p]
            prm["ref_poc"][0, l, r] = p
            prm["search_range"][0, l, r] = min(max_range, workload.search_range_uni(POC, p))
    cur = frame(POC)
    stale_pred = frame(POC + 1)               # what the prediction slot holds for CUs that are not motion compensated
    # ---- GPU
    ctx = lib.Context(width, height, bd, num_slots=base + 3)
    ctx.upload(0, cur)
    for p in uniq:
        ctx.upload(slot_of[p], frame(p))
        ctx.pad_border(slot_of[p])
    ctx.upload(slots["pred"], stale_pred)
    ctx.set_cus(cus)
    ctx.set_mv_predictors(mvp)
    me_g, tu_g = ctx.encode_picture(prm)
    ctx.sync()
    gpu = dict(me=me_g, tu=tu_g, cus=ctx.get_cus(), rec=[ctx.download_padded(slots["rec"], c) for c in range(3)],
               lev=ctx.download_coeff(slots["coeff"]))
    # ---- the reference's classes
    s = ref.session(width, height, bd, pic_type, qp, lam, simd=1, poc=POC, sub_gop=16)
    s.set_orig(cur)
    for l in range(2):
        for r, p in enumerate(pocs[l]):
            s.add_ref(l, r, p, frame(p))
    s.set_pred(stale_pred)
    me_r, tu_r, cus_r = s.encode_picture(prm, cus, threads=threads, mvp=mvp)
    cpu = dict(me=me_r, tu=tu_r, cus=cus_r, rec=[s.get_rec_padded(c) for c in range(3)], lev=s.get_coeff())
    s.close()
    return gpu, cpu, cus, prm


def assert_equal(gpu, cpu, cus_in, prm):
    J = abi.num_me_columns(prm)
    searched = (cus_in["flags"] & (abi.CU_INTRA | abi.CU_SKIP_ME)) == 0
    rows = np.repeat(searched, J)
    for f in ("mv_fullpel", "mv", "cost_fullpel", "dist", "cost"):
        assert np.array_equal(gpu["me"][f][rows], cpu["me"][f][rows]), f
    for f in ("flags", "ref_idx", "mv"):
        assert np.array_equal(gpu["cus"][f], cpu["cus"][f]), f
    assert np.array_equal(gpu["tu"], cpu["tu"])
    for c in range(3):
        assert np.array_equal(gpu["lev"][c], cpu["lev"][c]), ("levels", c)
        assert np.array_equal(gpu["rec"][c], cpu["rec"][c]), ("reconstruction", c)


CASES = {
    # name: (pic_type, (L0 POCs, L1 POCs), bi_iterations, predictors, mixed flags, bitdepth)
    "two_refs_same_pocs": (0, ((0, 16), (16, 0)), 1, "field", False, 10),      # xvc's default lists at POC 8: every L1 picture repeats an L0 one
    "two_refs_per_ref_predictors": (0, ((0, 16), (16, 0)), 1, "per_ref", True, 10),
    "three_and_one_per_ref_predictors": (0, ((4, 0, 2), (12,)), 2, "per_ref", False, 10),
    "unique_l1": (0, ((4, 0), (12, 16)), 1, "field", True, 10),
    "one_ref_four_iterations": (0, ((0,), (16,)), 4, "zero", False, 10),
    "three_and_one": (0, ((4, 0, 2), (12,)), 2, "random", True, 8),
    "partly_shared": (0, ((0, 4), (16, 4)), 2, "field", False, 10),
    "uni_two_refs": (1, ((4, 0), ()), 0, "field", True, 10),
    "bi_off": (0, ((0, 16), (16, 0)), 0, "random", False, 10),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_search_motion_pipeline(ref, name):
    pic_type, pocs, iters, predictors, mixed, bd = CASES[name]
    gpu, cpu, cus, prm = run_both(ref, 200, 104, bd, 32, pic_type, pocs, iters, seed=700 + len(name), predictors=predictors, mixed=mixed)
    assert_equal(gpu, cpu, cus, prm)
    if iters > 0:
        bi = (gpu["cus"]["ref_idx"][:, 0] >= 0) & (gpu["cus"]["ref_idx"][:, 1] >= 0) & ((cus["flags"] & abi.CU_SKIP_ME) == 0)
        assert bi.sum() > 0, "no CU chose bi-prediction: the case does not exercise SearchBiIterative"


def test_exempt_cus_keep_their_vectors(ref):
    """XVCB200_CU_INTRA / XVCB200_CU_SKIP_ME: no search, mv / ref_idx as the host set them."""
    gpu, cpu, cus, prm = run_both(ref, 200, 104, 10, 32, 0, ((0, 16), (16, 0)), 1, seed=91, predictors="random", mixed=True)
    keep = (cus["flags"] & (abi.CU_INTRA | abi.CU_SKIP_ME)) != 0
    assert keep.sum() > 10
    assert np.array_equal(gpu["cus"]["mv"][keep], cus["mv"][keep])
    assert np.array_equal(gpu["cus"]["ref_idx"][keep], cus["ref_idx"][keep])
    assert_equal(gpu, cpu, cus, prm)


FULL_SIZE = [
    pytest.param(1920, 1080, 10, 32, id="1080p-qp32"),
    pytest.param(3840, 2160, 10, 27, id="2160p-10bit-qp27"),
    pytest.param(7680, 4320, 10, 32, id="4320p-qp32"),
]


@pytest.mark.parametrize("predictors", ["per_ref", "zero"])
@pytest.mark.parametrize("width,height,bd,qp", FULL_SIZE)
def test_whole_picture_equals_reference(ref, width, height, bd, qp, predictors):
    """BASELINE.json's configurations at full size: EVERY search result, CU decision, level and sample of the
    padded, deblocked reconstruction equals what the unmodified reference's classes produce on all host
    threads (encode_decode_test.cc:106-111's whole-picture equality).  "per_ref": xvc's default reference
    lists at POC 8 of a sub-GOP of 16 (two pictures per list, the list-1 pictures repeating list 0),
    one SearchBiIterative pass, a predictor per (list, picture) near the content's motion towards that
    picture (the first diamond converges, as with neighbour-derived, POC-scaled predictors in an encoder).  "zero": round 1's step -- one picture per list, zero
    predictors (the raster scan of the +-128 window fires for most CUs), list chosen by the sub-pel cost."""
    threads = max(4, os.cpu_count() or 4)
    if predictors == "per_ref":
        gpu, cpu, cus, prm = run_both(ref, width, height, bd, qp, 0, ((0, 16), (16, 0)), 1, seed=1234, predictors="per_ref",
                                      max_range=256, threads=threads)
        bi = (gpu["cus"]["ref_idx"][:, 0] >= 0) & (gpu["cus"]["ref_idx"][:, 1] >= 0)
        assert bi.mean() > 0.05, "bi-prediction hardly ever chosen"
    else:
        gpu, cpu, cus, prm = run_both(ref, width, height, bd, qp, 0, ((0,), (16,)), 0, seed=1234, predictors="zero",
                                      max_range=256, threads=threads, bits_mode=0)
    assert_equal(gpu, cpu, cus, prm)
    assert any(np.any(l) for l in gpu["lev"]), "no residual was coded"
