"""The drop-in of INTEGRATION.md section 1, executed: the UNMODIFIED reference's own C++ classes
(InterSearch with TzSearch / SubpelSearch, InterPrediction::MotionCompensation, the residual chain of
TransformEncoder, SampleMetric) run with libxvc_b200's entries registered into their SIMD function
tables (oracle/ref_shim.cc, use_simd = 2: xvcb200_register_inter_prediction / _sample_metric called on
the reference's InterPrediction::SimdFunc / SampleMetric::SimdFunc objects) and must produce exactly
what they produce with the reference's C entries -- the reference's SimdTest contract
(test/xvc_test/simd_test.cc:149-176) with the CUDA library as the "SIMD" flavour.  One launch and
two copies per block: slow by construction, this is the bit-exactness vehicle, not the fast path.
Needs oracle/_ref (built here from /root/reference, travels to the GPU box prebuilt)."""
import numpy as np
import pytest

import common
from oracle import bindings
from xvc_b200 import abi, lib, workload

pytestmark = pytest.mark.gpu

W, H, BD, QP = 72, 40, 10, 32


@pytest.fixture(scope="module")
def ref():
    if not bindings.have_ref():
        pytest.skip("oracle/_ref/libxvcref.so not built (needs /root/reference)")
    lib.load()                      # fail loudly if the CUDA library is missing
    return bindings.Ref()


def _session(ref, simd, cur, r0, r1, lam):
    s = ref.session(W, H, BD, 0, QP, lam, simd=simd, poc=8, sub_gop=16)
    s.set_orig(cur)
    s.add_ref(0, 0, 0, r0)
    s.add_ref(1, 0, 16, r1)
    return s


def test_table_entries_through_reference_tables(ref):
    """Single entries, called through the reference's table objects (xref_sad / xref_ssd / xref_interp)."""
    rng = np.random.default_rng(5)
    before = lib.launch_count()
    for w, h in ((8, 8), (16, 4), (64, 32)):
        a = common.rnd_samples(rng, h, w, BD)
        b = common.rnd_samples(rng, h, w, BD)
        assert ref.sad(0, a, b, w, h, bitdepth=BD, simd=2) == ref.sad(0, a, b, w, h, bitdepth=BD, simd=0)
        assert ref.ssd(0, a, b, w, h, bitdepth=BD, simd=2) == ref.ssd(0, a, b, w, h, bitdepth=BD, simd=0)
    assert lib.launch_count() > before, "the CUDA entries did not run behind the reference's tables"


def test_reference_classes_on_cuda_tables(ref):
    cur, r0, r1 = common.frames(W, H, BD, 61)
    lam = workload.lambda_for_qp(QP)
    rng = np.random.default_rng(62)
    cus = workload.make_partition(W, H, seed=63, min_size=8, qp=QP)
    jobs = common.me_jobs(cus, rng, 2, (8, 8), 24)
    out = {}
    launches = {}
    for simd in (0, 2):
        before = lib.launch_count()
        s = _session(ref, simd, cur, r0, r1, lam)
        s.set_cus(cus)
        me = s.me_search(jobs, lam, threads=2)                       # InterSearch: TzSearch + SubpelSearch
        mc_cus = common.mc_cus(W, H, np.random.default_rng(64), 63, min_size=8)
        s.set_cus(mc_cus)
        s.motion_compensate(threads=2)                               # InterPrediction::MotionCompensation
        pred = s.get_pred()
        prm = common.picture_params(0, lam, ranges=(8, 8))
        me2, tu, cus_out = s.encode_picture(prm, cus, threads=2)     # the whole step through the reference's classes
        out[simd] = (me, pred, me2, tu, cus_out, s.get_rec(), s.get_coeff())
        launches[simd] = lib.launch_count() - before
        s.close()
    assert launches[0] == 0 and launches[2] > 100, launches
    a, b = out[0], out[2]
    assert np.array_equal(a[0], b[0]), "InterSearch results differ"
    for c in range(3):
        assert np.array_equal(a[1][c], b[1][c]), ("MotionCompensation", c)
        assert np.array_equal(a[5][c], b[5][c]), ("reconstruction", c)
        assert np.array_equal(a[6][c], b[6][c]), ("levels", c)
    assert np.array_equal(a[2], b[2]) and np.array_equal(a[3], b[3]) and np.array_equal(a[4], b[4])
