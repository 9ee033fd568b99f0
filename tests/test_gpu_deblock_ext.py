"""Deblocking beyond the plain CU array: the boundary strength of affine CUs compares the
vectors at the CU corners nearest to the edge segment (deblocking_filter.cc:166-176), and intra pictures carry
a secondary (chroma) CU tree whose edges drive the chroma filter on an 8-sample grid (:57-76, 88-91) --
xvcb200_deblock_picture_ext against the UNMODIFIED reference's DeblockingFilter on a PictureData that holds
the same CUs (oracle/ref_shim.cc)."""
import numpy as np
import pytest

import common
from oracle import bindings
from xvc_b200 import abi, lib, workload

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ref():
    if not bindings.have_ref():
        pytest.skip("oracle/_ref/libxvcref.so not built (needs /root/reference)")
    lib.load()
    return bindings.Ref()


def affine_entries(cus, rng):
    """Affine CUs whose control points differ by around one integer sample: both outcomes of the >= 16 tests occur."""
    idx = [i for i in range(len(cus)) if cus[i]["w"] > 8 and cus[i]["h"] > 8 and not (cus[i]["flags"] & abi.CU_INTRA)
           and (cus[i]["ref_idx"][0] >= 0 or cus[i]["ref_idx"][1] >= 0)]
    idx = idx[::2]
    aff = np.zeros(len(idx), dtype=abi.affine_cu_dtype)
    for k, i in enumerate(idx):
        aff[k]["cu"] = i
        for l in range(2):
            base = cus[i]["mv"][l].astype(np.int64)
            aff[k]["mv"][l] = [base, base + rng.integers(-24, 25, size=2), base + rng.integers(-24, 25, size=2)]
    return aff


@pytest.mark.parametrize("pic_type,bd", [(0, 10), (1, 10), (0, 8)])
def test_deblock_affine_corner_vectors(ref, pic_type, bd):
    width, height, qp = 200, 104, 32
    rng = np.random.default_rng(410 + pic_type + bd)
    canvas = workload.synth_canvas(width, height, 6)
    cur = workload.synth_frame(canvas, width, height, 3, bd)
    cus = common.deblock_cus(width, height, rng, 44, 8, pic_type)
    cus["flags"] &= ~np.uint8(abi.CU_CBF_Y)        # cbf would make the strength 1 whatever the vectors are
    inter = (cus["flags"] & abi.CU_INTRA) == 0      # one motion for every CU: without affine corners no inter edge is filtered
    cus["ref_idx"][inter] = (0, 0) if pic_type == 0 else (0, -1)
    cus["mv"][inter] = ((40, -24), (-16, 8)) if pic_type == 0 else ((40, -24), (0, 0))
    aff = affine_entries(cus, rng)
    assert len(aff) > 5
    recp = common.blocky_recon(cur, cus, rng, bd)
    pocs = {(0, 0): 0, (1, 0): 16}
    out = {}
    for with_affine in (False, True):
        s = ref.session(width, height, bd, pic_type, qp, workload.lambda_for_qp(qp), simd=1, poc=8, sub_gop=16)
        s.add_ref(0, 0, 0, cur)
        if pic_type == 0:
            s.add_ref(1, 0, 16, cur)
        s.set_rec(recp)
        s.set_cus(cus)
        s.deblock_picture_ext(affine=aff if with_affine else None)
        want = s.get_rec()
        s.close()
        ctx = lib.Context(width, height, bd, 1)
        ctx.upload(0, recp)
        ctx.set_cus(cus)
        ctx.deblock_picture_ext(0, pic_type, pocs, affine=aff if with_affine else None)
        got = ctx.download(0)
        ctx.close()
        for c in range(3):
            assert np.array_equal(got[c], want[c]), (with_affine, c)
        out[with_affine] = want
    assert any(not np.array_equal(out[False][c], out[True][c]) for c in range(3)), "the corner vectors changed nothing: weak test"


@pytest.mark.parametrize("bd,qp", [(10, 32), (8, 27)])
def test_deblock_intra_picture_secondary_tree(ref, bd, qp):
    width, height = 200, 136
    rng = np.random.default_rng(420 + bd)
    canvas = workload.synth_canvas(width, height, 7)
    cur = workload.synth_frame(canvas, width, height, 0, bd)
    luma_cus = workload.make_partition(width, height, seed=51, min_size=4, qp=qp)
    luma_cus["flags"] |= abi.CU_INTRA
    luma_cus["qp"] = rng.integers(qp - 4, qp + 5, size=len(luma_cus))
    chroma_cus = workload.make_partition(width, height, seed=52, min_size=8, qp=qp)      # the chroma tree: its own partition
    chroma_cus["flags"] |= abi.CU_INTRA
    chroma_cus["qp"] = rng.integers(qp - 4, qp + 5, size=len(chroma_cus))
    # block edges in luma along the luma tree, in chroma along the chroma tree
    recp = common.blocky_recon(cur, luma_cus, rng, bd)
    rec_c = common.blocky_recon(cur, chroma_cus, rng, bd)
    recp = [recp[0], rec_c[1], rec_c[2]]
    s = ref.session(width, height, bd, 2, qp, workload.lambda_for_qp(qp), simd=1, poc=0, sub_gop=16)
    s.set_rec(recp)
    s.set_cus(luma_cus)
    s.deblock_picture_ext(chroma_cus=chroma_cus)
    want = s.get_rec()
    s.close()
    ctx = lib.Context(width, height, bd, 1)
    ctx.upload(0, recp)
    ctx.set_cus(luma_cus)
    ctx.deblock_picture_ext(0, 2, {}, chroma_cus=chroma_cus)
    got = ctx.download(0)
    for c in range(3):
        assert np.array_equal(got[c], want[c]), c
        assert not np.array_equal(got[c], recp[c]), ("nothing filtered", c)
    # the primary tree alone filters different chroma edges
    ctx.upload(0, recp)
    ctx.deblock_picture(0, 1, {})
    assert not np.array_equal(ctx.download(0)[1], want[1])
    ctx.close()
