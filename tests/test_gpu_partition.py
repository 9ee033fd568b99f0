"""GPU partition pre-analysis (xvcb200_decide_partition): the kernel equals the numpy
statement of its rule (tests/partition_model.py) bit for bit; the partition tiles the picture, follows the
content (moving objects are split off, uniform motion is merged), its trees are ones xvc's syntax carries --
the UNMODIFIED reference decoder decodes pictures coded on it ("Conformance verified") -- and the whole
chain pre-analysis -> search -> residual coding -> in-loop filter reproduces the decoder's output."""
import os

import numpy as np
import pytest

import conformance
import partition_model
from oracle import bindings
from xvc_b200 import abi, lib, workload

pytestmark = pytest.mark.gpu


def moving_objects(width, height, bd, seed, n_obj=6):
    """Reference / current frame pair: a panned background plus rectangles that move on their own."""
    rng = np.random.default_rng(seed)
    canvas = workload.synth_canvas(width, height, seed)
    ref = [p.copy() for p in workload.synth_frame(canvas, width, height, 4, bd, frame_noise=2.0)]
    cur = [p.copy() for p in workload.synth_frame(canvas, width, height, 5, bd, frame_noise=2.0)]
    tex = workload.synth_frame(workload.synth_canvas(width, height, seed + 1), width, height, 0, bd)[0]
    for _ in range(n_obj):
        ow, oh = int(rng.integers(16, min(72, width // 3))), int(rng.integers(16, min(72, height // 3)))
        x, y = int(rng.integers(8, width - ow - 16)), int(rng.integers(8, height - oh - 16))
        dx, dy = int(rng.integers(-6, 7)), int(rng.integers(-6, 7))
        ref[0][y:y + oh, x:x + ow] = tex[y:y + oh, x:x + ow]
        cur[0][y + dy:y + dy + oh, x + dx:x + dx + ow] = tex[y:y + oh, x:x + ow]
    return cur, ref


@pytest.mark.parametrize("width,height,bd,qp,center", [(256, 128, 10, 32, (32, 16)), (200, 104, 10, 27, (0, 0)), (448, 264, 8, 37, (-48, 35)),
                                                       (192, 72, 10, 32, (2000, -1500))])
def test_kernel_equals_model(width, height, bd, qp, center):
    cur, ref = moving_objects(width, height, bd, 900 + width)
    lam_sqrt = float(np.sqrt(workload.lambda_for_qp(qp)))
    ctx = lib.Context(width, height, bd, 2)
    ctx.upload(0, cur)
    ctx.upload(1, ref)
    ctx.pad_border(1)
    cus_g, splits_g = ctx.decide_partition(0, 1, lam_sqrt, qp, center=center)
    ctx.close()
    cus_m, splits_m = partition_model.decide(cur[0], ref[0], center, lam_sqrt, qp)
    assert workload.check_partition(cus_g, width, height)
    assert len(cus_g) == len(cus_m)
    for f in ("x", "y", "w", "h", "depth", "qp", "mv", "ref_idx", "flags"):
        assert np.array_equal(cus_g[f], cus_m[f]), f
    assert np.array_equal(splits_g, splits_m)


def test_partition_follows_the_content():
    width, height, bd, qp = 512, 256, 10, 32
    lam_sqrt = float(np.sqrt(workload.lambda_for_qp(qp)))
    canvas = workload.synth_canvas(width, height, 77)
    ref = workload.synth_frame(canvas, width, height, 4, bd)
    cur = [p.copy() for p in workload.synth_frame(canvas, width, height, 5, bd)]      # pure pan (2, 1): one motion everywhere
    ctx = lib.Context(width, height, bd, 2)
    ctx.upload(1, ref)
    ctx.pad_border(1)
    ctx.upload(0, cur)
    cus_pan, _ = ctx.decide_partition(0, 1, lam_sqrt, qp)
    assert len(cus_pan) == (width // 64) * (height // 64), "uniform motion: every CTU stays one CU"
    assert (cus_pan["mv"][:, 0, 0] == 2 * 16).all() and (cus_pan["mv"][:, 0, 1] == 1 * 16).all()
    # an object with its own motion inside one CTU: that CTU is split, the others are not
    tex = workload.synth_frame(workload.synth_canvas(width, height, 78), width, height, 0, bd)[0]
    cur[0][80:112, 144:176] = tex[80:112, 144:176]
    ctx.upload(0, cur)
    cus_obj, _ = ctx.decide_partition(0, 1, lam_sqrt, qp)
    ctx.close()
    in_ctu = (cus_obj["x"] // 64 == 2) & (cus_obj["y"] // 64 == 1)
    assert in_ctu.sum() > 1 and (~in_ctu).sum() == (width // 64) * (height // 64) - 1


@pytest.mark.parametrize("width,height,qp,n_inter", [(256, 128, 32, 2), (1920, 1088, 32, 1), (448, 256, 32, 32)])
def test_conformance_on_gpu_partition(ref, width, height, qp, n_inter):
    """The partition the GPU decided, searched / coded / filtered on the GPU, written by the reference's writer and
    decoded by the UNMODIFIED xvcdec: "Conformance verified" and decoder output == GPU reconstruction -- up to a
    33-frame low-delay chain in which every inter picture references the GPU reconstruction of the one before."""
    if not os.path.exists(conformance.XVCDEC):
        pytest.skip("oracle/_ref/xvcdec not built (needs /root/reference)")
    size, log = conformance.run(ref, conformance.gpu_backend(width, height, 10), width, height, 10, qp, 21, n_inter=n_inter,
                                partition="gpu")
    assert size > 500
