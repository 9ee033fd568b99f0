"""CPU: the C oracle against the committed golden vectors generated from the unmodified
reference (tests/golden/make_golden.py).  Runs without /root/reference and without a GPU."""
import collections

import pytest

import golden_check


@pytest.fixture(scope="module")
def golden():
    return golden_check.load()


@pytest.mark.parametrize("kind", ["metric", "interp", "tx", "quant", "picture"])
def test_oracle_matches_golden(oracle, golden, kind):
    z, cases = golden
    be = golden_check.OracleBackend(oracle)
    n = 0
    for c in cases:
        if c["kind"] == kind:
            golden_check.run_case(be, z, c)
            n += 1
    assert n >= 2


def test_golden_covers_every_stage(golden):
    z, cases = golden
    count = collections.Counter(c["kind"] for c in cases)
    assert count["metric"] >= 20 and count["interp"] >= 10 and count["tx"] >= 100 and count["quant"] >= 30 and count["picture"] == 2


def test_affine_golden_oracle(oracle):
    """tests/golden/xvc_affine_golden.npz (reference MotionCompAffine outputs) replayed against the C
    restatement."""
    import affine_golden
    affine_golden.replay(affine_golden.oracle_backend(oracle))


def test_lic_golden_oracle(oracle):
    """tests/golden/xvc_lic_golden.npz (reference LocalIlluminationComp outputs) replayed against the C restatement."""
    import affine_golden
    affine_golden.replay_lic(affine_golden.oracle_lic_backend(oracle))
