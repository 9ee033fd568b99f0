"""xvcdec conformance of a bitstream carrying the hot path's output (see conformance.py)."""
import os

import pytest

import conformance


def _need_dec():
    if not os.path.exists(conformance.XVCDEC):
        pytest.skip("oracle/_ref/xvcdec not built (needs /root/reference)")


def test_conformance_plumbing_with_oracle(oracle, ref):
    """CPU: the C oracle as the producer -- validates the writer plumbing without a GPU."""
    _need_dec()
    size, log = conformance.run(ref, conformance.oracle_backend(oracle, 256, 128, 10), n_inter=4)
    assert size > 1000


@pytest.mark.gpu
@pytest.mark.parametrize("width,height,qp,seed,n_inter", [(256, 128, 32, 3, 1), (448, 256, 27, 11, 2), (1920, 1088, 32, 5, 1)])
def test_conformance_gpu(ref, width, height, qp, seed, n_inter):
    """The GPU pipeline's decisions, levels and reconstruction, verified by the reference decoder: one inter
    picture; a chain of two (the second one predicted from a GPU reconstruction; tests/test_gpu_partition.py runs a
    33-frame chain on the GPU-decided partition); 1080p (1920 x 1088)."""
    _need_dec()
    size, log = conformance.run(ref, conformance.gpu_backend(width, height, 10), width, height, 10, qp, seed, n_inter=n_inter)
    assert size > 1000
