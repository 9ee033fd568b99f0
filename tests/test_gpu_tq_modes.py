"""The residual-coding chain with every transform mode the reference's CodingUnit can carry (SURVEY 8 row
a14): transform-select types per direction (DCT-2 / DCT-5 / DCT-8 / DST-1 / DST-7, transform.cc:835-862,
1580-1612), transform skip of blocks of <= 16 samples (:184-215, 963-995), the 4 x 4 DST of intra luma
blocks (:217-242, 997-1017), horizontal / vertical coefficient scans of small intra CUs (:1614-1636) and
the intra-picture quantisation offset (rdo_quant.cc:168-169) -- xvcb200_tq_reconstruct /
xvcb200_dequant_reconstruct with xvcb200_set_tu_modes against the UNMODIFIED reference's
ForwardTransform / RdoQuant::QuantFast / Quantize / InverseTransform driven by oracle/ref_shim.cc."""
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


def make_modes(cus, rng, select=True, tskip=True):
    n = len(cus)
    modes = np.zeros(n, dtype=abi.tu_mode_dtype)
    intra_modes = rng.integers(0, 67, size=n).astype(np.uint8)
    intra_modes[::5] = rng.choice([18, 50, 12, 44, 56, 24], size=len(intra_modes[::5]))      # near horizontal / vertical: the adaptive scans
    if select:
        pick = rng.random(n) < 0.7
        modes["tx_ver"][pick] = rng.integers(0, 6, size=int(pick.sum()))
        modes["tx_hor"][pick] = rng.integers(0, 6, size=int(pick.sum()))
    if tskip:
        modes["tskip"] = rng.integers(0, 8, size=n)
    return modes, intra_modes


@pytest.mark.parametrize("bd,qp,min_size,pic_type,intra_frac", [(10, 32, 4, 0, 0.3), (10, 24, 4, 2, 1.0), (8, 37, 4, 1, 0.5),
                                                                 (12, 27, 8, 0, 0.2), (10, 30, 4, 0, 0.0)])
def test_tq_reconstruct_all_modes(ref, bd, qp, min_size, pic_type, intra_frac):
    width, height = 200, 136
    cur, _, _ = common.frames(width, height, bd, 320 + qp)
    lam = workload.lambda_for_qp(qp)
    cus = workload.make_partition(width, height, seed=31 + min_size, min_size=min_size, qp=qp)
    n = len(cus)
    rng = np.random.default_rng(321 + bd)
    cus["qp"][::3] = qp + 3
    intra = rng.random(n) < intra_frac
    cus["flags"][intra] |= abi.CU_INTRA
    cus["ref_idx"][~intra, 0] = 0
    predp = [np.clip(p.astype(np.int32) + rng.integers(-(60 << (bd - 8)), (60 << (bd - 8)) + 1, size=p.shape), 0, (1 << bd) - 1).astype(np.uint16) for p in cur]
    modes, intra_modes = make_modes(cus, rng)

    s = ref.session(width, height, bd, pic_type, qp, lam, simd=1, poc=8, sub_gop=16)
    s.set_orig(cur)
    s.set_pred(predp)
    s.set_cus(cus)
    s.set_tu_modes(modes, intra_modes)
    modes["scan"] = s.scan_orders(n)              # TransformHelper::DetermineScanOrder: the host's job in an encoder
    if intra_frac > 0 and min_size == 4:
        assert set(np.unique(modes["scan"])) == {0, 1, 2}
    tu_r = s.tq_reconstruct(n, threads=4)
    rec_r, lev_r = s.get_rec(), s.get_coeff()
    cus_r = s.get_cus(cus)
    s.close()

    ctx = lib.Context(width, height, bd, 6)
    ctx.upload(0, cur)
    ctx.upload(3, predp)
    ctx.set_cus(cus)
    ctx.set_tu_modes(modes)
    tu_g = ctx.tq_reconstruct(0, 3, 4, 5, intra_picture=int(pic_type == 2))
    rec_g, lev_g = ctx.download(4), ctx.download_coeff(5)
    assert np.array_equal(tu_g["num_non_zero"], tu_r["num_non_zero"])
    assert np.array_equal(tu_g["ssd"], tu_r["ssd"])
    for c in range(3):
        assert np.array_equal(lev_g[c], lev_r[c]), ("levels", c)
        assert np.array_equal(rec_g[c], rec_r[c]), ("reconstruction", c)
    assert np.array_equal(ctx.get_cus()["flags"], cus_r["flags"])
    # decoder side (CuDecoder::DecompressComponent): levels + cbf flags + the same modes -> the same reconstruction
    ctx.upload(4, [np.zeros_like(p) for p in cur])
    ctx.dequant_reconstruct(3, 4, 5)
    for c, p in enumerate(ctx.download(4)):
        assert np.array_equal(p, rec_r[c]), ("decoder", c)
    # set_cus drops the modes: the default chain again (DCT-2; 4 x 4 DST for intra luma)
    ctx.set_cus(cus)
    tu_d = ctx.tq_reconstruct(0, 3, 4, 5, intra_picture=int(pic_type == 2))
    s = ref.session(width, height, bd, pic_type, qp, lam, simd=1, poc=8, sub_gop=16)
    s.set_orig(cur)
    s.set_pred(predp)
    s.set_cus(cus)
    tu_rd = s.tq_reconstruct(n, threads=4)
    assert np.array_equal(tu_d, tu_rd)
    for c, p in enumerate(ctx.download(4)):
        assert np.array_equal(p, s.get_rec()[c]), ("default modes", c)
    s.close()
    ctx.close()
