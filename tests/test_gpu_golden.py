"""GPU: libxvc_b200.so (through the C ABI) against the committed golden vectors generated from
the unmodified reference.  Bit-exact, no oracle in the loop."""
import pytest

import golden_check

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def golden():
    return golden_check.load()


@pytest.mark.parametrize("kind", ["metric", "interp", "tx", "quant", "picture"])
def test_gpu_matches_golden(golden, kind):
    z, cases = golden
    be = golden_check.GpuBackend()
    n = 0
    for c in cases:
        if c["kind"] == kind:
            golden_check.run_case(be, z, c)
            n += 1
    assert n >= 2


def test_affine_golden_gpu():
    """Reference MotionCompAffine outputs (tests/golden/xvc_affine_golden.npz) == xvcb200_motion_compensate_affine."""
    import affine_golden
    affine_golden.replay(affine_golden.gpu_backend())


def test_lic_golden_gpu():
    """Reference LocalIlluminationComp outputs (tests/golden/xvc_lic_golden.npz) == xvcb200_motion_compensate_lic."""
    import affine_golden
    affine_golden.replay_lic(affine_golden.gpu_lic_backend())
